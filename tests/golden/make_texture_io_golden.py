"""Generates tests/golden/texture_io.npz by running the REFERENCE's own texture helpers on a seeded texture:
``rgb2sh0`` / ``sh02rgb`` (models/texture_gaussian3d.py:16-21) and the methods ``cube_map`` (:451-461) and
``change_texture`` (:463-495) of ``TextureGaussian3D``. The model module itself cannot be imported in this container
(tinycudann, nvdiffrast, plyfile ... are absent), so the source text of exactly those definitions is cut out of the
reference file where it lies and executed against a stand-in ``self`` that only has ``_texture``. Nothing is copied
into the repo; the vectors pin ``texture_gs_b200/texture_io.py``.
Run from the repo root inside the build container:  python tests/golden/make_texture_io_golden.py
"""
import re
import textwrap
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/models/texture_gaussian3d.py")


def _cut(src: str, header: str) -> str:
    """Source of the def starting at ``header`` up to the next def at the same indentation."""
    a = src.index(header)
    indent = len(header) - len(header.lstrip())
    m = re.compile(r"\n {%d}(def |@|class )" % indent).search(src, a + len(header))
    return src[a:m.start() if m else len(src)]


def main():
    src = REF.read_text()
    ns = {"torch": torch}
    exec(_cut(src, "def rgb2sh0(rgb):"), ns)
    exec(_cut(src, "def sh02rgb(sh0):"), ns)
    for name in ("cube_map", "change_texture"):
        exec(textwrap.dedent(_cut(src, f"    def {name}(self")), ns)

    class Stand:
        pass

    R = 6
    g = torch.Generator().manual_seed(0)
    tex = (torch.rand(6, R, R, 3, generator=g) * 1.4 - 0.2 - 0.5) / 0.28209479177387814     # some values clamp in sh02rgb
    cross = torch.rand(3 * R, 4 * R, 3, generator=g)
    cross[:R, :R] = 0.0                                                                     # an unpainted region for mode 3
    cross[R:2 * R, R:2 * R][:3] = 0.0
    out = {"texture": tex.numpy(), "cross": cross.numpy()}
    s = Stand()
    s._texture = tex.clone()
    out["cube_map"] = ns["cube_map"](s).numpy()
    for mode in (-1, 0, 1, 2, 3):
        s = Stand()
        s._texture = tex.clone()
        ns["change_texture"](s, cross.clone(), mode=mode)
        out[f"changed_mode{mode}"] = s._texture.numpy()
    dst = Path(__file__).resolve().parent / "texture_io.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
