"""Host-side checks of small device helpers: the function text is cut out of the .cuh, compiled as plain C with gcc
(no GPU needed) and compared with numpy. Covers arithmetic that only a cold path of the kernels exercises."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
COMMON = ROOT / "texture_gs_b200" / "csrc" / "texgs_common.cuh"


def _host_compile_eigvec(tmp_path):
    src = COMMON.read_text()
    a = src.index("#define TEXGS_JACOBI_ROT")
    b = src.index("// ---", src.index("smallest_eigvec_sym3"))
    body = src[a:b]
    c_src = ("#include <math.h>\n#define __device__\n#define __noinline__\n"
             "typedef struct { float x, y, z; } float3;\n"
             "static float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }\n"
             "static float rsqrtf(float x) { return 1.0f / sqrtf(x); }\n"
             + re.sub(r"#pragma unroll 1\n", "", body) +
             "\nvoid eig(const float* s, float* out) { float3 n = smallest_eigvec_sym3(s[0], s[1], s[2], s[3], s[4], s[5]);"
             " out[0] = n.x; out[1] = n.y; out[2] = n.z; }\n")
    cfile = tmp_path / "eig.c"
    cfile.write_text(c_src)
    so = tmp_path / "eig.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", str(cfile), "-o", str(so), "-lm"], check=True)
    lib = C.CDLL(str(so))
    lib.eig.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return lib


def test_smallest_eigenvector_of_flat_disc_covariances(tmp_path):
    """smallest_eigvec_sym3 (cov3Ds_precomp path, texgs_common.cuh): for Sigma = R diag(s)^2 R^T with one scale of
    exp(-20) (models/texture_gaussian3d.py:290-297: flat discs) and for generic anisotropic covariances the result is
    the rotation column of the smallest scale, to fp32 accuracy, at any overall magnitude."""
    lib = _host_compile_eigvec(tmp_path)
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(400):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        r, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                      [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                      [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
        if trial % 2 == 0:
            s = np.array([np.exp(rng.uniform(-7, -3)), np.exp(rng.uniform(-7, -3)), np.exp(-20.0)])
        else:
            s = np.exp(rng.uniform(-6, 0, size=3))
            s[rng.integers(3)] *= 0.3          # keep the smallest eigenvalue separated
        s = s[rng.permutation(3)] * (10.0 ** rng.integers(-3, 3))
        S = (R * s[None, :] ** 2) @ R.T
        s6 = np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]], dtype=np.float32)
        out = np.zeros(3, dtype=np.float32)
        lib.eig(s6.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
        # reference: eigenvector of the fp32-rounded matrix, in double
        S32 = np.array([[s6[0], s6[1], s6[2]], [s6[1], s6[3], s6[4]], [s6[2], s6[4], s6[5]]], dtype=np.float64)
        w, v = np.linalg.eigh(S32)
        n = v[:, 0]
        assert abs(np.linalg.norm(out) - 1.0) < 1e-5
        err = min(np.abs(out - n).max(), np.abs(out + n).max())
        gap = (w[1] - w[0]) / w[2]
        worst = max(worst, err * min(1.0, gap))
        assert err <= 2e-6 / min(1.0, gap) + 1e-6, (trial, err, gap, s)
    print("worst gap-scaled error", worst)
