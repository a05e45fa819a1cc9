"""Texture-side plumbing around the rasterizer (host side): colour encoding of the cube texture,
its cross-layout image, the retexture operation, a lat-long view, and the reference's checkpoint tuple.

Everything here is layout work on whole tensors (slicing / stacking / one gather) done a handful of times per run,
so it is plain torch on whatever device the texture lives on — no kernels. Restated from:

* ``rgb2sh0`` / ``sh02rgb``                       ``models/texture_gaussian3d.py:16-21``
* ``cube_map`` (faces -> 3R x 4R cross image)    ``models/texture_gaussian3d.py:451-461``
* ``change_texture`` (cross image -> faces, 5 blend modes)   ``models/texture_gaussian3d.py:463-495``, used by
  ``retexture.py:48-58``
* ``sphere_map`` (lat-long image)                ``models/texture_gaussian3d.py:446-449`` +
  ``models/modules/NVDIFFREC/util.py:119-133``
* checkpoint tuple                               ``models/texture_gaussian3d.py:145-194`` (``torch.save((state_dict, iter))``,
  read back at ``retexture.py:44-45``)

Cross layout (row block, column block) of the faces, R = face resolution::

              [2]                     +y
        [1]   [4]   [0]   [5]         -x  +z  +x  -z
              [3]                     -y
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

C0 = 0.28209479177387814

# face -> (row block, column block) in the 3R x 4R cross image (models/texture_gaussian3d.py:455-460, 467-474)
CROSS_SLOT: Tuple[Tuple[int, int], ...] = ((1, 2), (1, 0), (0, 1), (2, 1), (1, 1), (1, 3))


def rgb2sh0(rgb: torch.Tensor) -> torch.Tensor:
    return (rgb - 0.5) / C0


def sh02rgb(sh0: torch.Tensor) -> torch.Tensor:
    return torch.clamp(C0 * sh0 + 0.5, 0.0, 1.0)


def cube_map(texture: torch.Tensor) -> torch.Tensor:
    """(6,R,R,3) SH-DC-encoded texture -> (3R,4R,3) RGB cross image; the two corner strips stay black."""
    if texture.dim() != 4 or texture.shape[0] != 6 or texture.shape[1] != texture.shape[2]:
        raise ValueError(f"texture must be (6,R,R,C), got {tuple(texture.shape)}")
    rgb = sh02rgb(texture)
    R = texture.shape[1]
    out = torch.zeros(3 * R, 4 * R, texture.shape[3], dtype=rgb.dtype, device=rgb.device)
    for face, (rb, cb) in enumerate(CROSS_SLOT):
        out[rb * R:(rb + 1) * R, cb * R:(cb + 1) * R] = rgb[face]
    return out


def faces_from_cross(cubemap_image: torch.Tensor) -> torch.Tensor:
    """(3R,4R,C) cross image -> (6,R,R,C) faces (inverse of the face placement of ``cube_map``)."""
    R = cubemap_image.shape[0] // 3
    if cubemap_image.dim() != 3 or cubemap_image.shape[0] != 3 * R or cubemap_image.shape[1] != 4 * R:
        raise ValueError(f"cross image must be (3R,4R,C), got {tuple(cubemap_image.shape)}")
    return torch.stack([cubemap_image[rb * R:(rb + 1) * R, cb * R:(cb + 1) * R] for rb, cb in CROSS_SLOT], dim=0)


def change_texture(texture: torch.Tensor, cubemap_image: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """The retexture operation: returns the new SH-DC-encoded (6,R,R,3) texture built from an RGB cross image.

    mode -1 replace; 0 new * luminance of (3 * old, clamped); 1 new * old; 2 old / new; 3 tint the painted region
    (new.sum > 0.01) of the old texture with twice its luminance and add (``models/texture_gaussian3d.py:479-493``)."""
    new_tex = faces_from_cross(cubemap_image).clone()
    ori_tex = sh02rgb(texture.detach())
    if ori_tex.shape != new_tex.shape:
        raise ValueError(f"cross image gives faces {tuple(new_tex.shape)}, texture is {tuple(ori_tex.shape)}")
    if mode == -1:
        pass
    elif mode == 0:
        new_tex = new_tex * (ori_tex * 3).clamp(0, 1).mean(dim=-1, keepdim=True)
    elif mode == 1:
        new_tex = new_tex * ori_tex
    elif mode == 2:
        new_tex = ori_tex / new_tex
    elif mode == 3:
        mask = new_tex.sum(-1) > 0.01
        tinted = torch.where(mask[..., None], 2 * ori_tex.mean(-1, keepdim=True) * new_tex, ori_tex)
        new_tex = new_tex + tinted
    else:
        raise ValueError(f"unknown mode {mode}")
    return rgb2sh0(new_tex)


def resize_cross(cubemap_image: torch.Tensor, face_resolution: int) -> torch.Tensor:
    """Bilinear resize of a cross image to (3R,4R) for R = ``face_resolution`` — what ``retexture.py:53`` does with
    ``cv2.resize(..., INTER_LINEAR)`` (half-pixel centres, no antialiasing)."""
    x = cubemap_image.permute(2, 0, 1)[None]
    y = F.interpolate(x, size=(3 * face_resolution, 4 * face_resolution), mode="bilinear", align_corners=False)
    return y[0].permute(1, 2, 0).contiguous()


def dir_to_face_uv(d: torch.Tensor):
    """Direction (...,3) -> (face, sx, sy), sx/sy in [-1,1]: inverse of ``cube_to_dir``
    (``models/modules/NVDIFFREC/util.py:94-101``; C twin ``renderutils/c_src/cubemap.cu:49-61``); ties x > y > z —
    the same selection the render kernels make (``cube_coord`` in ``csrc/texgs_common.cuh``)."""
    x, y, z = d.unbind(-1)
    ax, ay, az = x.abs(), y.abs(), z.abs()
    is_x = (ax >= ay) & (ax >= az)
    is_y = (~is_x) & (ay >= az)
    m = torch.where(is_x, ax, torch.where(is_y, ay, az)).clamp_min(1e-20)
    face = torch.where(is_x, torch.where(x < 0, 1, 0), torch.where(is_y, torch.where(y < 0, 3, 2), torch.where(z < 0, 5, 4)))
    sx = torch.where(is_x, torch.where(x < 0, z, -z), torch.where(is_y, x, torch.where(z < 0, -x, x))) / m
    sy = torch.where(is_x, -y, torch.where(is_y, torch.where(y < 0, -z, z), -y)) / m
    return face, sx, sy


def sample_cube(texture: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """Bilinear lookup of (6,R,R,C) at directions (...,3) with the rasterizer's filtering (texel centres at
    2(i+.5)/R-1, clamp-to-edge inside the face; spec E11)."""
    R = texture.shape[1]
    face, sx, sy = dir_to_face_uv(d)
    fx = (sx + 1.0) * (0.5 * R) - 0.5
    fy = (sy + 1.0) * (0.5 * R) - 0.5
    x0, y0 = torch.floor(fx), torch.floor(fy)
    wx, wy = (fx - x0)[..., None], (fy - y0)[..., None]
    x0, y0 = x0.long(), y0.long()
    x0c, x1c = x0.clamp(0, R - 1), (x0 + 1).clamp(0, R - 1)
    y0c, y1c = y0.clamp(0, R - 1), (y0 + 1).clamp(0, R - 1)
    top = texture[face, y0c, x0c] * (1 - wx) + texture[face, y0c, x1c] * wx
    bot = texture[face, y1c, x0c] * (1 - wx) + texture[face, y1c, x1c] * wx
    return top * (1 - wy) + bot * wy


def sphere_map(texture: torch.Tensor, resolution: Sequence[int] = (512, 1024)) -> torch.Tensor:
    """Lat-long RGB image (H,W,3) of the texture: pixel grid and direction formula of ``cubemap_to_latlong``
    (``NVDIFFREC/util.py:119-133``). The reference samples with nvdiffrast's seamless cube filtering; this uses the
    rasterizer's own per-face clamp-to-edge filtering, so the two differ only within half a texel of face edges."""
    H, W = int(resolution[0]), int(resolution[1])
    dev, dt = texture.device, texture.dtype
    gy = torch.linspace(0.0 + 1.0 / H, 1.0 - 1.0 / H, H, device=dev, dtype=dt)[:, None].expand(H, W)
    gx = torch.linspace(-1.0 + 1.0 / W, 1.0 - 1.0 / W, W, device=dev, dtype=dt)[None, :].expand(H, W)
    st, ct = torch.sin(gy * math.pi), torch.cos(gy * math.pi)
    sp, cp = torch.sin(gx * math.pi), torch.cos(gx * math.pi)
    d = torch.stack((st * sp, ct, -st * cp), dim=-1)
    return sample_cube(sh02rgb(texture), d)


# ---------------------------------------------------------------------------------------------------------------
# checkpoint tuple
# ---------------------------------------------------------------------------------------------------------------

TCNN_INPUT_PAD_VALUE = 1.0     # [EXT] what tiny-cuda-nn feeds the padded input columns of a bare ``tcnn.Network`` (see tcnn_mlp_weights)


def tcnn_mlp_weights(params: torch.Tensor, n_in: int, n_out: int, width: int = 128, n_hidden: int = 1, return_input_pad: bool = False):
    """Split the flat ``params`` vector of a tiny-cuda-nn ``FullyFusedMLP`` (``models/modules/utils.py:29-41``) into
    ``nn.Linear``-style weights ``[(width, n_in), (width, width) * (n_hidden-1), (n_out, width)]``.

    [EXT: tiny-cuda-nn is not in the reference tree] layout from its published source: row-major (out, in) matrices in
    layer order, input width padded up to a multiple of 16, output width padded up to a multiple of 16, no biases.
    The padded OUTPUT rows are dropped (their outputs are never read). The padded INPUT columns are not dead: a bare
    ``tcnn.Network`` runs behind an identity encoding that fills the columns it pads with the constant 1, so the sum of
    those columns acts as a learned bias of the first layer. ``return_input_pad=True`` also returns that ``(width,
    pad_in - n_in)`` block so the caller can fold it into a bias (``CheckpointGaussians._load_uv_net``)."""
    pad_in, pad_out = -(-n_in // 16) * 16, -(-n_out // 16) * 16
    sizes = [(width, pad_in)] + [(width, width)] * (n_hidden - 1) + [(pad_out, width)]
    need = sum(a * b for a, b in sizes)
    flat = params.detach().reshape(-1).float()
    if flat.numel() != need:
        raise ValueError(f"expected {need} tcnn parameters for {n_in}->{width}x{n_hidden}->{n_out}, got {flat.numel()}")
    out, o = [], 0
    for a, b in sizes:
        out.append(flat[o:o + a * b].reshape(a, b))
        o += a * b
    pad_cols = out[0][:, n_in:].contiguous()
    out[0] = out[0][:, :n_in].contiguous()
    out[-1] = out[-1][:n_out].contiguous()
    return (out, pad_cols) if return_input_pad else out


class CheckpointGaussians:
    """The duck-typed ``gaussians`` object ``uv_tex_render`` reads (``render/uv_tex_render.py:15,34,42-53``) built from a
    reference checkpoint's ``state_dict`` — ``params = (_xyz, _scaling, _rotation, _opacity, _shs, _texture)`` with the
    activations of ``models/texture_gaussian3d.py:25-30,196-240`` (exp / sigmoid / normalize), ``hyperparams =
    (active_sh_degree, spatial_lr_scale)``. UVs and their Jacobian come from ``uv_net`` (a ``FusedUVNet``) and the
    geometry embedding, evaluated once and cached like the reference's ``eval()`` does (``:257-262``)."""

    def __init__(self, state_dict: Dict, uv_net=None, device: Optional[torch.device] = None):
        self.active_sh_degree = int(state_dict["hyperparams"][0])
        self.spatial_lr_scale = float(state_dict["hyperparams"][1])
        names = ("_xyz", "_scaling", "_rotation", "_opacity", "_shs", "_texture")
        params = state_dict["params"]
        if len(params) != len(names):
            raise ValueError(f"params must hold {names}, got {len(params)} entries")
        for n, p in zip(names, params):
            if p is not None and device is not None:
                p = p.detach().to(device)
            setattr(self, n, p)
        self.uv_net = uv_net
        self.geo_emb = None
        net_state = state_dict.get("net_state")
        if net_state is not None and len(net_state) >= 3 and "weight" in net_state[2]:
            self.geo_emb = net_state[2]["weight"].detach().reshape(-1)
            if device is not None:
                self.geo_emb = self.geo_emb.to(device)
        if uv_net is not None and net_state is not None:
            self._load_uv_net(net_state[0])
        self._uv = None
        self._grad_uv = None

    def _load_uv_net(self, sd: Dict):
        if "pre_mlp.params" in sd and "mlp.params" in sd:            # tiny-cuda-nn networks (use_tcnn: True)
            (w1, w2), pad1 = tcnn_mlp_weights(sd["pre_mlp.params"], 3, 128, n_hidden=1, return_input_pad=True)
            w3, w4, w5 = tcnn_mlp_weights(sd["mlp.params"], 128, 3, n_hidden=2)       # 128 inputs: nothing padded
            layers = (self.uv_net.pre_mlp[0], self.uv_net.pre_mlp[2], self.uv_net.mlp[0], self.uv_net.mlp[2], self.uv_net.mlp[4])
            # the 13 padded input columns of the 3 -> 128 layer see the constant TCNN_INPUT_PAD_VALUE: a learned bias
            bias1 = TCNN_INPUT_PAD_VALUE * pad1.sum(dim=1)
            if layers[0].bias is None and float(bias1.abs().max()) > 0.0:
                raise ValueError("this tiny-cuda-nn checkpoint uses its padded input columns as a bias: build FusedUVNet(bias=True)")
            with torch.no_grad():
                for i, (lin, w) in enumerate(zip(layers, (w1, w2, w3, w4, w5))):
                    lin.weight.copy_(w.to(lin.weight.device))
                    if lin.bias is not None:
                        if i == 0:
                            lin.bias.copy_(bias1.to(lin.bias.device))
                        else:
                            lin.bias.zero_()
        else:                                                        # nn.Linear networks: same keys as FusedUVNet
            self.uv_net.load_state_dict(sd)

    get_xyz = property(lambda self: self._xyz)
    get_scaling = property(lambda self: torch.exp(self._scaling))
    get_rotation = property(lambda self: F.normalize(self._rotation))
    get_opacity = property(lambda self: torch.sigmoid(self._opacity))
    get_shs = property(lambda self: self._shs)
    get_texture = property(lambda self: self._texture)

    def set_texture(self, texture: torch.Tensor):
        self._texture = texture

    def _uv_pair(self):
        if self._uv is None:
            if self.uv_net is None or self.geo_emb is None:
                raise RuntimeError("UVs need a FusedUVNet and the checkpoint's geometry embedding")
            with torch.no_grad():
                self._uv, self._grad_uv = self.uv_net.uv_and_jacobian(self._xyz, self.geo_emb)
        return self._uv, self._grad_uv

    get_uvs = property(lambda self: self._uv_pair()[0])
    get_grad_uvs = property(lambda self: self._uv_pair()[1])


def load_checkpoint(path, uv_net=None, device=None) -> Tuple[CheckpointGaussians, int]:
    """``(state_dict, iteration) = torch.load(path)`` (``retexture.py:44``) -> ``(CheckpointGaussians, iteration)``."""
    state_dict, it = torch.load(path, map_location="cpu", weights_only=False)
    return CheckpointGaussians(state_dict, uv_net=uv_net, device=device), int(it)
