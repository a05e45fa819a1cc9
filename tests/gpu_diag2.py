"""Conditioning check: CUDA fp32 and oracle fp32 both against the fp64 oracle."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from util import *

def one(n, w, h, r, deg, seed, cam_seed=None, bg=(0.2, 0.4, 0.6)):
    g = sphere_shell_scene(n, r, sh_degree=deg, seed=seed, tex_seed=seed + 1)
    cam = orbit_cameras(1, w, h, seed=seed + 2 if cam_seed is None else cam_seed)[0]
    cot = output_cotangents(h, w, seed=3)
    o64, aux64, g64 = run_oracle(g, cam, bg=bg, cot=cot, dtype=torch.float64)
    o32, aux32, g32 = run_oracle(g, cam, bg=bg, cot=cot, dtype=torch.float32)
    ocu, st, gcu = run_cuda(g, cam, bg=bg, cot=cot)
    print(f"== n={n} {w}x{h} R={r} deg={deg} seed={seed}  pairs {st.num_pairs} amb32 {int(aux32['ambiguous'].sum())}")
    for i, nm in enumerate(("image", "depth", "norm", "alpha")):
        d1 = (o32[i].double() - o64[i]).abs().max().item(); d2 = (ocu[i].double() - o64[i]).abs().max().item(); d3 = (ocu[i].double() - o32[i].double()).abs().max().item()
        print(f"  out {nm:6s} |o32-o64| {d1:.2e}  |cuda-o64| {d2:.2e}  |cuda-o32| {d3:.2e}")
    for k in g64:
        if g64[k] is None: continue
        a, b, c = g64[k], g32[k], gcu[k]
        if k == "means2D": a, b, c = a[:, :2], b[:, :2], c[:, :2]
        c = c.reshape(a.shape)
        print(f"  grad {k:9s} rel(o32,o64) {rel_err(b, a):.2e}  rel(cuda,o64) {rel_err(c, a):.2e}  rel(cuda,o32) {rel_err(c, b):.2e}  max {a.abs().max().item():.3g}")
    # where is the texture grad off?
    a, c = g64["texture"], gcu["texture"].reshape(g64["texture"].shape).double()
    d = (c - a).abs().sum(-1)
    idx = torch.nonzero(d == d.max())[0].tolist()
    print("  worst texel", idx, "ref", a[tuple(idx)].tolist(), "cuda", c[tuple(idx)].tolist(), "o32", g32["texture"][tuple(idx)].tolist())
    dsum = (c.sum((0, 1, 2)) - a.sum((0, 1, 2))).tolist()
    print("  texture-grad checksum diff per channel", dsum, "ref sums", a.sum((0, 1, 2)).tolist())

one(500, 70, 50, 16, 0, 4)
one(10000, 256, 256, 512, 3, 0, cam_seed=1, bg=(0, 0, 0))
one(2000, 128, 96, 64, 3, 0)
