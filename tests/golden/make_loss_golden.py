"""Generates tests/golden/photometric_loss.npz by running the REFERENCE's own loss code
(/root/reference/losses/pixelwise_loss.py:3-4 l1_loss, losses/ssim_loss.py:16-54 ssim_loss — pure
PyTorch, importable in this container) on seeded images, exactly as the training loop combines them
(models/texture_gaussian3d.py:333-340):   loss = (1-l)*L1(image, gt) + l*(1 - SSIM(image, gt)).
These vectors PIN oracle/loss_ref.py (and through it the CUDA kernels) to the reference itself.
Run from the repo root inside the build container:  python tests/golden/make_loss_golden.py
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/losses")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    pw, ss = _load("pixelwise_loss"), _load("ssim_loss")
    out = {}
    for tag, (h, w, lam, seed) in {"a": (37, 45, 0.2, 0), "b": (16, 64, 0.2, 1), "c": (9, 7, 0.5, 2)}.items():
        g = torch.Generator().manual_seed(seed)
        gt = torch.rand(3, h, w, generator=g)
        img = (gt + 0.15 * torch.randn(3, h, w, generator=g)).clamp(0, 1.2).requires_grad_(True)
        l1 = pw.l1_loss(img, gt)
        ssim = ss.ssim_loss(img, gt)
        loss = (1.0 - lam) * l1 + lam * (1.0 - ssim)
        loss.backward()
        out.update({f"{tag}_img": img.detach().numpy(), f"{tag}_gt": gt.numpy(), f"{tag}_lambda": np.float32(lam),
                    f"{tag}_l1": l1.detach().numpy(), f"{tag}_ssim": ssim.detach().numpy(),
                    f"{tag}_loss": loss.detach().numpy(), f"{tag}_grad": img.grad.numpy()})
    dst = Path(__file__).resolve().parent / "photometric_loss.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: float(v) for k, v in out.items() if k.endswith("_loss")})


if __name__ == "__main__":
    main()
