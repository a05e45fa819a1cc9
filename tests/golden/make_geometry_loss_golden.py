"""Generates tests/golden/geometry_loss.npz by running the REFERENCE's own loss code
(/root/reference/losses/pixelwise_loss.py:3-4 l1_loss, losses/norm_reg_loss.py:66-71 norm_loss,
losses/smooth_loss.py:4-27 smooth_loss — pure PyTorch, importable in this container) on seeded inputs, the way
the training step calls them (models/texture_gaussian3d.py:342-345, 354-358, 365-368):
    Lalpha = l1_loss(alpha, gt_alpha); Lnorm = norm_loss(norm, gt_norm, gt_alpha);
    Lnsm = smooth_loss(gt_image, norm, gt_alpha)
These vectors PIN oracle/loss_ref.py's geometry_losses (and through it the CUDA kernels) to the reference itself.
Run from the repo root inside the build container:  python tests/golden/make_geometry_loss_golden.py
"""
import importlib.util
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REF = Path("/root/reference/losses")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(h, w, seed, hard_mask):
    g = torch.Generator().manual_seed(seed)
    gt_image = torch.rand(3, h, w, generator=g)
    gt_image = F.avg_pool2d(gt_image[None], 3, 1, 1)[0]                  # neighbouring pixels correlate: weights not all ~0
    gt_norm = F.normalize(torch.randn(3, h, w, generator=g), dim=0)
    norm = 0.8 * gt_norm + 0.3 * torch.randn(3, h, w, generator=g)      # un-normalised, as the rasterizer blends it
    gt_alpha = torch.rand(1, h, w, generator=g)
    gt_alpha = (gt_alpha > 0.3).float() if hard_mask else gt_alpha
    alpha = (gt_alpha + 0.2 * torch.randn(1, h, w, generator=g)).clamp(0, 1)
    return alpha, norm, gt_alpha, gt_norm, gt_image


def main():
    pw, nr, sm = _load("pixelwise_loss"), _load("norm_reg_loss"), _load("smooth_loss")
    out = {}
    for tag, (h, w, seed, hard) in {"a": (23, 41, 0, True), "b": (8, 32, 1, False), "c": (5, 3, 2, True)}.items():
        alpha, norm, gt_alpha, gt_norm, gt_image = make_inputs(h, w, seed, hard)
        alpha.requires_grad_(True); norm.requires_grad_(True)
        la = pw.l1_loss(alpha, gt_alpha)
        ln = nr.norm_loss(norm, gt_norm, gt_alpha)
        ls = sm.smooth_loss(gt_image, norm, gt_alpha)
        (1.0 * la + 0.1 * ln + 0.5 * ls).backward()                      # the lambdas of configs/texture_gaussian3d.yaml:83-88
        out.update({f"{tag}_alpha": alpha.detach().numpy(), f"{tag}_norm": norm.detach().numpy(), f"{tag}_gt_alpha": gt_alpha.numpy(),
                    f"{tag}_gt_norm": gt_norm.numpy(), f"{tag}_gt_image": gt_image.numpy(),
                    f"{tag}_Lalpha": la.detach().numpy(), f"{tag}_Lnorm": ln.detach().numpy(), f"{tag}_Lnsm": ls.detach().numpy(),
                    f"{tag}_galpha": alpha.grad.numpy(), f"{tag}_gnorm": norm.grad.numpy()})
    dst = Path(__file__).resolve().parent / "geometry_loss.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: float(v) for k, v in out.items() if "_L" in k})


if __name__ == "__main__":
    main()
