// texgs_preprocess.cuh — per-Gaussian forward projection and its backward (a9).
// Spec items E1-E3, E8, E12(SH part) of SURVEY §8c; maths conventions cited in oracle/raster_ref.py.
#pragma once
#include "texgs_common.cuh"

// SH bands l>=1 evaluated at unit direction d. ``sh`` points at the first REST coefficient of this
// Gaussian (3 floats per coefficient). Sign pattern: reference utils/sh.py:69-112.
__device__ __forceinline__ float3 sh_rest_eval(int deg, const float* __restrict__ sh, float3 d) {
    float3 r = f3(0.f, 0.f, 0.f);
    if (deg <= 0) return r;
    const float x = d.x, y = d.y, z = d.z;
#define SH_ACC(k, w) { const float ww = (w); r.x += ww * sh[3 * (k)]; r.y += ww * sh[3 * (k) + 1]; r.z += ww * sh[3 * (k) + 2]; }
    SH_ACC(0, -SH_C1 * y) SH_ACC(1, SH_C1 * z) SH_ACC(2, -SH_C1 * x)
    if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        SH_ACC(3, SH_C2[0] * xy) SH_ACC(4, SH_C2[1] * yz) SH_ACC(5, SH_C2[2] * (2.f * zz - xx - yy))
        SH_ACC(6, SH_C2[3] * xz) SH_ACC(7, SH_C2[4] * (xx - yy))
        if (deg > 2) {
            SH_ACC(8, SH_C3[0] * y * (3.f * xx - yy)) SH_ACC(9, SH_C3[1] * xy * z)
            SH_ACC(10, SH_C3[2] * y * (4.f * zz - xx - yy))
            SH_ACC(11, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy))
            SH_ACC(12, SH_C3[4] * x * (4.f * zz - xx - yy)) SH_ACC(13, SH_C3[5] * z * (xx - yy))
            SH_ACC(14, SH_C3[6] * x * (xx - 3.f * yy))
        }
    }
#undef SH_ACC
    return r;
}

// Everything the forward projection of one Gaussian produces (shared by fwd and bwd kernels so
// that the backward recomputes exactly what the forward saw).
struct Proj {
    bool visible;
    float3 pv;          // view-space centre
    float4 ph;          // clip-space
    float pw;           // 1/(w+1e-7)
    float R[9];         // rotation
    float3 s;           // scales * modifier
    float tx, ty;       // clamped view x,y
    bool clampx, clampy;
    float T[6];         // 2x3: J * W
    float Sig[6];       // xx,xy,xz,yy,yz,zz
    float a, b, c, det; // 2-D covariance (dilated) and determinant
    float radius;
    float x, y;         // pixel-space mean
    int rx0, ry0, rx1, ry1;
    int kmin;           // axis of the smallest scale
    float nsign;        // +1 / -1 (flip to face the camera)
    float3 nv;          // view-space facing normal
    float3 m;           // centre - campos (world)
};

__device__ __forceinline__ void project_gaussian(const RasterParams& p, int idx, Proj& o) {
    const float3 mu = f3(p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]);
    o.visible = false;
    o.pv = xform43(p.view, mu);
    if (!(o.pv.z > TEXGS_NEAR)) return;                                     // E1
    o.ph = xform44(p.proj, mu);
    o.pw = 1.0f / (o.ph.w + 1e-7f);
    const bool covgiven = p.cov3Ds_precomp != nullptr;
    if (!covgiven) {
        const float4 q = *reinterpret_cast<const float4*>(p.rotations + 4 * idx);
        quat_to_rot(q, o.R);
        o.s = f3(p.scales[3 * idx] * p.scale_modifier, p.scales[3 * idx + 1] * p.scale_modifier,
                 p.scales[3 * idx + 2] * p.scale_modifier);
        // L = R diag(s);  Sigma = L L^T
        float L[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) { L[3 * r] = o.R[3 * r] * o.s.x; L[3 * r + 1] = o.R[3 * r + 1] * o.s.y; L[3 * r + 2] = o.R[3 * r + 2] * o.s.z; }
        o.Sig[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
        o.Sig[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
        o.Sig[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
        o.Sig[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
        o.Sig[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
        o.Sig[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
    } else {   // cov3Ds_precomp (render/render.py:52-53): used as given
#pragma unroll
        for (int k = 0; k < 6; ++k) o.Sig[k] = p.cov3Ds_precomp[(size_t)6 * idx + k];
#pragma unroll
        for (int k = 0; k < 9; ++k) o.R[k] = 0.f;
        o.s = f3(0.f, 0.f, 0.f);
    }
    // E2: EWA projection
    const float limx = 1.3f * p.tanfovx, limy = 1.3f * p.tanfovy;
    const float tz = o.pv.z;
    const float qx = o.pv.x / tz, qy = o.pv.y / tz;
    o.clampx = (qx < -limx) || (qx > limx);
    o.clampy = (qy < -limy) || (qy > limy);
    o.tx = fminf(limx, fmaxf(-limx, qx)) * tz;
    o.ty = fminf(limy, fmaxf(-limy, qy)) * tz;
    const float j00 = p.focal_x / tz, j02 = -p.focal_x * o.tx / (tz * tz);
    const float j11 = p.focal_y / tz, j12 = -p.focal_y * o.ty / (tz * tz);
    // W[c][r] = V[r][c]  (world -> view rotation);  T = J W
    const float* V = p.view.m;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o.T[r] = j00 * V[4 * r + 0] + j02 * V[4 * r + 2];
        o.T[3 + r] = j11 * V[4 * r + 1] + j12 * V[4 * r + 2];
    }
    const float* S = o.Sig;
    const float u0 = S[0] * o.T[0] + S[1] * o.T[1] + S[2] * o.T[2];
    const float u1 = S[1] * o.T[0] + S[3] * o.T[1] + S[4] * o.T[2];
    const float u2 = S[2] * o.T[0] + S[4] * o.T[1] + S[5] * o.T[2];
    const float w0 = S[0] * o.T[3] + S[1] * o.T[4] + S[2] * o.T[5];
    const float w1 = S[1] * o.T[3] + S[3] * o.T[4] + S[4] * o.T[5];
    const float w2 = S[2] * o.T[3] + S[4] * o.T[4] + S[5] * o.T[5];
    o.a = o.T[0] * u0 + o.T[1] * u1 + o.T[2] * u2 + 0.3f;
    o.b = o.T[3] * u0 + o.T[4] * u1 + o.T[5] * u2;
    o.c = o.T[3] * w0 + o.T[4] * w1 + o.T[5] * w2 + 0.3f;
    o.det = o.a * o.c - o.b * o.b;
    if (o.det == 0.0f) return;
    // E3: radius, pixel mean, tile rect
    const float mid = 0.5f * (o.a + o.c);
    const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - o.det));
    o.radius = ceilf(3.0f * sqrtf(lam));
    o.x = ((o.ph.x * o.pw + 1.0f) * (float)p.W - 1.0f) * 0.5f;
    o.y = ((o.ph.y * o.pw + 1.0f) * (float)p.H - 1.0f) * 0.5f;
    if (!(isfinite(o.x) && isfinite(o.y) && isfinite(o.radius))) return;
    const float gxf = (float)p.grid_x, gyf = (float)p.grid_y;
    o.rx0 = (int)fminf(gxf, fmaxf(0.f, floorf((o.x - o.radius) * (1.0f / TEXGS_TILE))));
    o.rx1 = (int)fminf(gxf, fmaxf(0.f, floorf((o.x + o.radius + (TEXGS_TILE - 1)) * (1.0f / TEXGS_TILE))));
    o.ry0 = (int)fminf(gyf, fmaxf(0.f, floorf((o.y - o.radius) * (1.0f / TEXGS_TILE))));
    o.ry1 = (int)fminf(gyf, fmaxf(0.f, floorf((o.y + o.radius + (TEXGS_TILE - 1)) * (1.0f / TEXGS_TILE))));
    if ((o.rx1 - o.rx0) * (o.ry1 - o.ry0) <= 0) return;
    // E8: facing disc normal
    float3 nraw;
    if (!covgiven) {
        o.kmin = argmin3(p.scales[3 * idx], p.scales[3 * idx + 1], p.scales[3 * idx + 2]);
        // column kmin of R, selected without dynamic indexing (keeps R in registers)
        nraw = (o.kmin == 0) ? f3(o.R[0], o.R[3], o.R[6]) : ((o.kmin == 1) ? f3(o.R[1], o.R[4], o.R[7]) : f3(o.R[2], o.R[5], o.R[8]));
    } else {   // the same direction when Sigma = R S^2 R^T: eigenvector of the smallest eigenvalue
        o.kmin = -1;
        nraw = smallest_eigvec_sym3(o.Sig[0], o.Sig[1], o.Sig[2], o.Sig[3], o.Sig[4], o.Sig[5]);
    }
    o.m = f3(mu.x - p.campos[0], mu.y - p.campos[1], mu.z - p.campos[2]);
    o.nsign = (dot3(nraw, o.m) > 0.f) ? -1.f : 1.f;
    o.nv = rot_w2v(p.view, f3(o.nsign * nraw.x, o.nsign * nraw.y, o.nsign * nraw.z));
    o.visible = true;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// TEXGS_PREFWD_STAGE_SH: as in the backward, the SH rows of the 32 Gaussians of a warp are one contiguous block of global
// memory (128 * 3M bytes) that a thread owning one Gaussian would walk with a 12 M-byte stride (45 scalar loads, each
// touching 32 lines per warp). One bulk copy (TMA 1-D) per warp stages the block in shared memory while the lanes
// project their Gaussians; the SH evaluation then reads its row there (odd row stride: conflict-free).
#ifndef TEXGS_PREFWD_STAGE_SH
#define TEXGS_PREFWD_STAGE_SH 1
#endif
__host__ __forceinline__ size_t prefwd_smem_bytes(int M, int mode, const void* shs) {
#if TEXGS_PREFWD_STAGE_SH
    return (shs && mode != TEXGS_MODE_PRECOMP && M > 0 && ((3 * M) & 1)) ? (size_t)8 * 32 * 3 * M * sizeof(float) + 8 * sizeof(uint64_t) : 0;
#else
    (void)M; (void)mode; (void)shs;
    return 0;
#endif
}

__global__ void __launch_bounds__(256) texgs_preprocess_fwd(const RasterParams p, int* __restrict__ radii) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const float* sh_row = (p.shs != nullptr && idx < p.P) ? p.shs + (size_t)idx * p.M * 3 : nullptr;
#if TEXGS_PREFWD_STAGE_SH
    extern __shared__ __align__(128) float prefwd_smem[];
    const int nsh = p.M * 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp-uniform: a full warp of Gaussians, odd row stride, 16-byte aligned coefficient tensor
    const bool bulk = (p.shs != nullptr) && (p.mode != TEXGS_MODE_PRECOMP) && (nsh & 1) && (idx - lane + 32 <= p.P) &&
                      (((uintptr_t)p.shs & 15) == 0);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(prefwd_smem + (size_t)8 * 32 * nsh);
    if (bulk) {
        float* const wrows = prefwd_smem + (size_t)warp * 32 * nsh;
        if (lane == 0) {
            mbar_init(&bars[warp], 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&bars[warp], 128u * (unsigned)nsh);
            bulk_g2s(wrows, p.shs + (size_t)(idx - lane) * nsh, 128u * (unsigned)nsh, &bars[warp]);
        }
        sh_row = wrows + lane * nsh;
        __syncwarp();
    }
    if (idx >= p.P) return;                    // never taken by a lane of a warp with a copy in flight (full warps only)
    Proj o;
    project_gaussian(p, idx, o);               // overlaps the copy
    if (bulk) mbar_wait(&bars[warp], 0u);      // every lane waits before it may leave: the copy must not outlive the block
#else
    if (idx >= p.P) return;
    Proj o;
    project_gaussian(p, idx, o);
#endif
    if (!o.visible) {
        radii[idx] = 0;
        p.rects[idx] = make_uint2(0u, 0u);
        return;
    }
    radii[idx] = (int)o.radius;
    p.rects[idx] = make_uint2((unsigned)o.rx0 | ((unsigned)o.rx1 << 16), (unsigned)o.ry0 | ((unsigned)o.ry1 << 16));

    const float inv_det = 1.0f / o.det;
    float3 col;
    if (p.mode == TEXGS_MODE_PRECOMP) {
        col = f3(p.colors_precomp[3 * idx], p.colors_precomp[3 * idx + 1], p.colors_precomp[3 * idx + 2]);
    } else {
        const float inv_len = rsqrtf(dot3(o.m, o.m));
        const float3 dir = f3(o.m.x * inv_len, o.m.y * inv_len, o.m.z * inv_len);
        if (p.mode == TEXGS_MODE_TEXTURE) {
            col = (sh_row != nullptr) ? sh_rest_eval(min(p.sh_degree, 3), sh_row, dir) : f3(0.f, 0.f, 0.f);
        } else {  // full SH: DC + rest
            const float* sh = sh_row;
            col = sh_rest_eval(min(p.sh_degree, 3), sh + 3, dir);
            col.x += SH_C0 * sh[0]; col.y += SH_C0 * sh[1]; col.z += SH_C0 * sh[2];
        }
        col.x += 0.5f; col.y += 0.5f; col.z += 0.5f;
        if (p.mode == TEXGS_MODE_SH) { col.x = fmaxf(col.x, 0.f); col.y = fmaxf(col.y, 0.f); col.z = fmaxf(col.z, 0.f); }
    }
    float uv[3] = {0.f, 0.f, 0.f};
    float Jp[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.mode == TEXGS_MODE_TEXTURE) {
        uv[0] = p.uvs[3 * idx]; uv[1] = p.uvs[3 * idx + 1]; uv[2] = p.uvs[3 * idx + 2];
        // J' = J * Wm^T : J'[i][c] = sum_r J[i][r] * V[r][c]   (delta_world = Wm^T delta_view)
        const float* J = p.gradient_uvs + (size_t)9 * idx;
        const float* V = p.view.m;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Jp[3 * i + c] = J[3 * i] * V[c] + J[3 * i + 1] * V[4 + c] + J[3 * i + 2] * V[8 + c];
    }
    GaussRec* r = p.recs + idx;
    r->q[0] = make_float4(o.x, o.y, o.c * inv_det, -o.b * inv_det);
    r->q[1] = make_float4(o.a * inv_det, p.opacities[idx], o.pv.z, dot3(o.nv, o.pv));
    r->q[2] = make_float4(o.nv.x, o.nv.y, o.nv.z, o.pv.x);
    r->q[3] = make_float4(o.pv.y, col.x, col.y, col.z);
    r->q[4] = make_float4(uv[0], uv[1], uv[2], Jp[0]);
    r->q[5] = make_float4(Jp[1], Jp[2], Jp[3], Jp[4]);
    r->q[6] = make_float4(Jp[5], Jp[6], Jp[7], Jp[8]);
    r->q[7] = make_float4(__int_as_float(idx), 0.f, 0.f, 0.f);

    {   // the SAME floats the scatter kernel will read back from the record
        const float4 r0 = r->q[0], r1 = r->q[1];
        for (int ty = o.ry0; ty < o.ry1; ++ty)
            for (int tx = o.rx0; tx < o.rx1; ++tx)
                if (splat_hits_tile(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tx, ty)) atomicAdd(&p.tile_count[ty * p.grid_x + tx], 1u);
    }
    atomicAdd(&p.counters->num_visible, 1u);
}

// ---------------------------------------------------------------------------------------------
// backward: accumulators (TEXGS_BWD_ACC_FLOATS per Gaussian, written by the render backward) ->
// gradients of the operator inputs.
//   acc[0..1]  dL/d(x,y) pixel units      acc[2..4]  dL/d conic (a,b,c)     acc[5] dL/d opacity
//   acc[6..8]  dL/d col                   acc[9]     dL/d depth             acc[10..12] dL/d n_v
//   acc[13..15] dL/d uv                   acc[16] S0 = sum s, acc[17..19] sum s*Delta_v (intersection path,
//                                          s = (J'^T gu . v)/(n_v . v))
// ---------------------------------------------------------------------------------------------
// accumulate mode adds ATOMICALLY (fire-and-forget RED, no read of the destination): views rendered concurrently on
// several CUDA streams accumulate into one gradient buffer (texture_gs_b200.dist.render_views_accumulate(streams=n))
__device__ __forceinline__ void put(float* p, float v, bool acc) { if (acc) atomicAdd(p, v); else *p = v; }

// ``sh`` and ``dsh`` may be the SAME row (staged path: the gradient replaces the coefficients in shared memory; each
// coefficient is read before it is overwritten), hence no __restrict__.
__device__ __forceinline__ void sh_rest_bwd(int deg, const float* sh, float3 d, float3 g,
                                            float* dsh, float3& ddir, bool accum) {
    // dsh_k = basis_k * g ;  ddir = sum_k dbasis_k/dd * (sh_k . g)
    ddir = f3(0.f, 0.f, 0.f);
    if (deg <= 0) return;
    const float x = d.x, y = d.y, z = d.z;
#define SH_B(k, w, dwx, dwy, dwz) { const float ww = (w); const float sg = sh[3 * (k)] * g.x + sh[3 * (k) + 1] * g.y + sh[3 * (k) + 2] * g.z; \
        if (dsh) { put(dsh + 3 * (k), ww * g.x, accum); put(dsh + 3 * (k) + 1, ww * g.y, accum); put(dsh + 3 * (k) + 2, ww * g.z, accum); } \
        ddir.x += (dwx) * sg; ddir.y += (dwy) * sg; ddir.z += (dwz) * sg; }
    SH_B(0, -SH_C1 * y, 0.f, -SH_C1, 0.f)
    SH_B(1, SH_C1 * z, 0.f, 0.f, SH_C1)
    SH_B(2, -SH_C1 * x, -SH_C1, 0.f, 0.f)
    if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        SH_B(3, SH_C2[0] * xy, SH_C2[0] * y, SH_C2[0] * x, 0.f)
        SH_B(4, SH_C2[1] * yz, 0.f, SH_C2[1] * z, SH_C2[1] * y)
        SH_B(5, SH_C2[2] * (2.f * zz - xx - yy), SH_C2[2] * -2.f * x, SH_C2[2] * -2.f * y, SH_C2[2] * 4.f * z)
        SH_B(6, SH_C2[3] * xz, SH_C2[3] * z, 0.f, SH_C2[3] * x)
        SH_B(7, SH_C2[4] * (xx - yy), SH_C2[4] * 2.f * x, SH_C2[4] * -2.f * y, 0.f)
        if (deg > 2) {
            SH_B(8, SH_C3[0] * y * (3.f * xx - yy), SH_C3[0] * 6.f * xy, SH_C3[0] * (3.f * xx - 3.f * yy), 0.f)
            SH_B(9, SH_C3[1] * xy * z, SH_C3[1] * yz, SH_C3[1] * xz, SH_C3[1] * xy)
            SH_B(10, SH_C3[2] * y * (4.f * zz - xx - yy), SH_C3[2] * -2.f * xy, SH_C3[2] * (4.f * zz - xx - 3.f * yy), SH_C3[2] * 8.f * yz)
            SH_B(11, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy), SH_C3[3] * -6.f * xz, SH_C3[3] * -6.f * yz, SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy))
            SH_B(12, SH_C3[4] * x * (4.f * zz - xx - yy), SH_C3[4] * (4.f * zz - 3.f * xx - yy), SH_C3[4] * -2.f * xy, SH_C3[4] * 8.f * xz)
            SH_B(13, SH_C3[5] * z * (xx - yy), SH_C3[5] * 2.f * xz, SH_C3[5] * -2.f * yz, SH_C3[5] * (xx - yy))
            SH_B(14, SH_C3[6] * x * (xx - 3.f * yy), SH_C3[6] * (3.f * xx - 3.f * yy), SH_C3[6] * -6.f * xy, 0.f)
        }
    }
#undef SH_B
}

struct BwdOut {
    float *dmeans3D, *dmeans2D, *dopacity, *dscales, *drotations, *dshs, *dcolors_precomp, *duvs, *dcov3Ds;
    unsigned acc;   // TEXGS_ACC_* bits: add into the output instead of overwriting it
};

#ifndef TEXGS_PREBWD_MIN_CTAS
#define TEXGS_PREBWD_MIN_CTAS 3
#endif
// zero gradient rows of a culled Gaussian: written in overwrite mode, nothing to add in accumulate mode
__device__ __forceinline__ void prebwd_write_zero(const RasterParams& p, const BwdOut& g, int idx, bool skip_shs) {
    const int nsh = p.M * 3;
    const bool a_m3 = g.acc & TEXGS_ACC_MEANS3D, a_m2 = g.acc & TEXGS_ACC_MEANS2D, a_op = g.acc & TEXGS_ACC_OPACITY;
    const bool a_sc = g.acc & TEXGS_ACC_SCALES, a_ro = g.acc & TEXGS_ACC_ROTATIONS, a_sh = g.acc & TEXGS_ACC_SHS;
    const bool a_cp = g.acc & TEXGS_ACC_COLORS, a_uv = g.acc & TEXGS_ACC_UVS;
    if (g.dmeans3D && !a_m3) { g.dmeans3D[3 * idx] = 0.f; g.dmeans3D[3 * idx + 1] = 0.f; g.dmeans3D[3 * idx + 2] = 0.f; }
    if (g.dmeans2D && !a_m2) { g.dmeans2D[3 * idx] = 0.f; g.dmeans2D[3 * idx + 1] = 0.f; g.dmeans2D[3 * idx + 2] = 0.f; }
    if (g.dopacity && !a_op) g.dopacity[idx] = 0.f;
    if (g.dscales && !a_sc) { g.dscales[3 * idx] = 0.f; g.dscales[3 * idx + 1] = 0.f; g.dscales[3 * idx + 2] = 0.f; }
    if (g.drotations && !a_ro) *reinterpret_cast<float4*>(g.drotations + 4 * idx) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.dshs && !a_sh && !skip_shs) for (int k = 0; k < nsh; ++k) g.dshs[(size_t)idx * nsh + k] = 0.f;
    if (g.dcolors_precomp && !a_cp) { g.dcolors_precomp[3 * idx] = 0.f; g.dcolors_precomp[3 * idx + 1] = 0.f; g.dcolors_precomp[3 * idx + 2] = 0.f; }
    if (g.duvs && !a_uv) { g.duvs[3 * idx] = 0.f; g.duvs[3 * idx + 1] = 0.f; g.duvs[3 * idx + 2] = 0.f; }
    if (g.dcov3Ds) for (int k = 0; k < 6; ++k) g.dcov3Ds[(size_t)6 * idx + k] = 0.f;
}

// gradients of one visible Gaussian. ``sh_row`` / ``dsh_row``: its SH coefficients and where their gradient goes —
// global rows (accumulate per TEXGS_ACC_SHS) or, staged, the same shared-memory row (always overwritten; the
// accumulation happens in the warp's coalesced store).
__device__ __forceinline__ void prebwd_one(const RasterParams& p, const float* __restrict__ acc_all, const BwdOut& g, int idx,
                                           const Proj& o, const float* sh_row, float* dsh_row, bool a_sh = false) {
    const int nsh = p.M * 3;
    float dmu[3] = {0.f, 0.f, 0.f};
    const bool a_m3 = g.acc & TEXGS_ACC_MEANS3D, a_m2 = g.acc & TEXGS_ACC_MEANS2D, a_op = g.acc & TEXGS_ACC_OPACITY;
    const bool a_sc = g.acc & TEXGS_ACC_SCALES, a_ro = g.acc & TEXGS_ACC_ROTATIONS;
    const bool a_cp = g.acc & TEXGS_ACC_COLORS, a_uv = g.acc & TEXGS_ACC_UVS;
    const float* acc = acc_all + (size_t)idx * TEXGS_BWD_ACC_FLOATS;
    const float* V = p.view.m;

    // ---- colour -> SH / precomp --------------------------------------------------------------
    float3 gcol = f3(acc[6], acc[7], acc[8]);
    if (p.mode == TEXGS_MODE_PRECOMP) {
        if (g.dcolors_precomp) { put(g.dcolors_precomp + 3 * idx, gcol.x, a_cp); put(g.dcolors_precomp + 3 * idx + 1, gcol.y, a_cp); put(g.dcolors_precomp + 3 * idx + 2, gcol.z, a_cp); }
    } else if (p.shs != nullptr) {
        const float len2 = dot3(o.m, o.m);
        const float inv_len = rsqrtf(len2);
        const float3 dir = f3(o.m.x * inv_len, o.m.y * inv_len, o.m.z * inv_len);
        const float* sh = sh_row;
        float* dsh = dsh_row;
        const int deg = min(p.sh_degree, 3);
        const int nrest_active = (deg + 1) * (deg + 1) - 1;
        int first_rest = 0;
        if (p.mode == TEXGS_MODE_SH) {
            // clamp mask of the per-Gaussian colour
            float3 col = sh_rest_eval(deg, sh + 3, dir);
            col.x += SH_C0 * sh[0] + 0.5f; col.y += SH_C0 * sh[1] + 0.5f; col.z += SH_C0 * sh[2] + 0.5f;
            if (col.x < 0.f) gcol.x = 0.f;
            if (col.y < 0.f) gcol.y = 0.f;
            if (col.z < 0.f) gcol.z = 0.f;
            if (dsh) { put(dsh, SH_C0 * gcol.x, a_sh); put(dsh + 1, SH_C0 * gcol.y, a_sh); put(dsh + 2, SH_C0 * gcol.z, a_sh); }
            first_rest = 1;
        }
        float3 ddir;
        sh_rest_bwd(deg, sh + 3 * first_rest, dir, gcol, dsh ? dsh + 3 * first_rest : nullptr, ddir, a_sh);
        if (dsh && !a_sh) for (int k = 3 * (first_rest + nrest_active); k < nsh; ++k) dsh[k] = 0.f;
        // dir = m/|m|
        const float dd = dot3(dir, ddir);
        dmu[0] += (ddir.x - dir.x * dd) * inv_len;
        dmu[1] += (ddir.y - dir.y * dd) * inv_len;
        dmu[2] += (ddir.z - dir.z * dd) * inv_len;
    } else if (g.dshs && !(g.acc & TEXGS_ACC_SHS)) {
        for (int k = 0; k < nsh; ++k) g.dshs[(size_t)idx * nsh + k] = 0.f;
    }

    // ---- intersection path + normal + depth: view-space grads --------------------------------
    float3 dpv = f3(0.f, 0.f, acc[9]);                    // depth = pv.z
    float3 dnv = f3(acc[10], acc[11], acc[12]);
    if (p.mode == TEXGS_MODE_TEXTURE) {
        const float3 duv = f3(acc[13], acc[14], acc[15]);
        if (g.duvs) { put(g.duvs + 3 * idx, duv.x, a_uv); put(g.duvs + 3 * idx + 1, duv.y, a_uv); put(g.duvs + 3 * idx + 2, duv.z, a_uv); }
        if (!(p.flags & TEXGS_FLAG_STOPGRAD_DELTA)) {      // E13-alt: nothing flows through Delta = t v - p_v
            const float S0 = acc[16], S1 = acc[17], S2 = acc[18], S3 = acc[19];
            // J'^T duv, J'[i][c] = sum_r J[i][r] V[r][c]
            const float* J = p.gradient_uvs + (size_t)9 * idx;
            float jtw[3];  // world: J^T duv
#pragma unroll
            for (int r = 0; r < 3; ++r) jtw[r] = J[r] * duv.x + J[3 + r] * duv.y + J[6 + r] * duv.z;
            const float3 jtv = rot_w2v(p.view, f3(jtw[0], jtw[1], jtw[2]));   // view: J'^T duv
            dpv.x += o.nv.x * S0 - jtv.x; dpv.y += o.nv.y * S0 - jtv.y; dpv.z += o.nv.z * S0 - jtv.z;
            dnv.x -= S1; dnv.y -= S2; dnv.z -= S3;      // dL/dn_v = -sum s*Delta_v
        }
    }

    // ---- 2-D mean -> clip -> world -----------------------------------------------------------
    const float dndx = acc[0] * 0.5f * (float)p.W, dndy = acc[1] * 0.5f * (float)p.H;   // NDC units
    if (g.dmeans2D) { put(g.dmeans2D + 3 * idx, dndx, a_m2); put(g.dmeans2D + 3 * idx + 1, dndy, a_m2); if (!a_m2) g.dmeans2D[3 * idx + 2] = 0.f; }
    {
        const float dphx = dndx * o.pw, dphy = dndy * o.pw;
        const float dphw = -(dndx * o.ph.x + dndy * o.ph.y) * o.pw * o.pw;
        const float* PM = p.proj.m;
#pragma unroll
        for (int r = 0; r < 3; ++r) dmu[r] += dphx * PM[4 * r + 0] + dphy * PM[4 * r + 1] + dphw * PM[4 * r + 3];
    }

    // ---- conic -> cov2D ----------------------------------------------------------------------
    const float dA = acc[2], dB = acc[3], dC = acc[4];
    const float a = o.a, b = o.b, c = o.c;
    const float id2 = 1.0f / (o.det * o.det);
    const float dLda = (-c * c * dA + b * c * dB - b * b * dC) * id2;
    const float dLdc = (-b * b * dA + a * b * dB - a * a * dC) * id2;
    const float dLdb = (2.f * b * c * dA - (a * c + b * b) * dB + 2.f * a * b * dC) * id2;
    // symmetric G_cov = [[dLda, dLdb/2],[dLdb/2, dLdc]]
    const float g00 = dLda, g01 = 0.5f * dLdb, g11 = dLdc;
    const float* T = o.T;
    const float* S = o.Sig;
    // dL/dSigma (full symmetric matrix) = T^T G T
    float GM[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            GM[3 * i + j] = T[i] * (g00 * T[j] + g01 * T[3 + j]) + T[3 + i] * (g01 * T[j] + g11 * T[3 + j]);
    if (g.dcov3Ds) {   // cov3Ds_precomp: Sigma is the input; an off-diagonal entry of the 6-vector stands for both halves
        float* d = g.dcov3Ds + (size_t)6 * idx;
        d[0] = GM[0]; d[1] = GM[1] + GM[3]; d[2] = GM[2] + GM[6]; d[3] = GM[4]; d[4] = GM[5] + GM[7]; d[5] = GM[8];
    }
    // dL/dL = 2 GM L,  L = R diag(s)
    float dR[9];
    float ds[3] = {0.f, 0.f, 0.f};
    const float sarr[3] = {o.s.x, o.s.y, o.s.z};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float dL = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) dL += GM[3 * r + j] * o.R[3 * j + k] * sarr[k];
            dL *= 2.f;
            dR[3 * r + k] = dL * sarr[k];
            ds[k] += dL * o.R[3 * r + k];
        }
    // dL/dT = 2 G T Sigma
    float TS[6];
    TS[0] = T[0] * S[0] + T[1] * S[1] + T[2] * S[2];
    TS[1] = T[0] * S[1] + T[1] * S[3] + T[2] * S[4];
    TS[2] = T[0] * S[2] + T[1] * S[4] + T[2] * S[5];
    TS[3] = T[3] * S[0] + T[4] * S[1] + T[5] * S[2];
    TS[4] = T[3] * S[1] + T[4] * S[3] + T[5] * S[4];
    TS[5] = T[3] * S[2] + T[4] * S[4] + T[5] * S[5];
    float dT[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dT[j] = 2.f * (g00 * TS[j] + g01 * TS[3 + j]);
        dT[3 + j] = 2.f * (g01 * TS[j] + g11 * TS[3 + j]);
    }
    // T = Jm W, W[c][r] = V[r][c]  ->  dJm[i][c] = sum_r dT[i][r] V[r][c]
    const float dJ00 = dT[0] * V[0] + dT[1] * V[4] + dT[2] * V[8];
    const float dJ02 = dT[0] * V[2] + dT[1] * V[6] + dT[2] * V[10];
    const float dJ11 = dT[3] * V[1] + dT[4] * V[5] + dT[5] * V[9];
    const float dJ12 = dT[3] * V[2] + dT[4] * V[6] + dT[5] * V[10];
    const float tz = o.pv.z, itz = 1.0f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float dtx = -p.focal_x * itz2 * dJ02;
    const float dty = -p.focal_y * itz2 * dJ12;
    const float dtz = -p.focal_x * itz2 * dJ00 - p.focal_y * itz2 * dJ11 + 2.f * p.focal_x * o.tx * itz3 * dJ02 +
                      2.f * p.focal_y * o.ty * itz3 * dJ12;
    // tx = clamp(x/z) * z : unclamped -> tx = x ; clamped -> tx = lim * z
    // E2-alt (TEXGS_FLAG_CLAMP_GRAD_3DGS): the 3DGS lineage passes no gradient through a clamped coordinate
    const bool drop = (p.flags & TEXGS_FLAG_CLAMP_GRAD_3DGS) != 0u;
    if (!o.clampx) dpv.x += dtx; else if (!drop) dpv.z += dtx * (o.tx * itz);
    if (!o.clampy) dpv.y += dty; else if (!drop) dpv.z += dty * (o.ty * itz);
    dpv.z += dtz;

    // ---- view -> world -----------------------------------------------------------------------
    {
        const float3 w = rot_v2w(p.view, dpv);
        dmu[0] += w.x; dmu[1] += w.y; dmu[2] += w.z;
        const float3 dn = rot_v2w(p.view, dnv);     // dL/d n (world, facing)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float on = (o.kmin == k) ? o.nsign : 0.f;
            dR[0 + k] += on * dn.x;
            dR[3 + k] += on * dn.y;
            dR[6 + k] += on * dn.z;
        }
    }

    // ---- outputs -----------------------------------------------------------------------------
    if (g.dmeans3D) { put(g.dmeans3D + 3 * idx, dmu[0], a_m3); put(g.dmeans3D + 3 * idx + 1, dmu[1], a_m3); put(g.dmeans3D + 3 * idx + 2, dmu[2], a_m3); }
    if (g.dopacity) put(g.dopacity + idx, acc[5], a_op);
    if (g.dscales) {
        put(g.dscales + 3 * idx, ds[0] * p.scale_modifier, a_sc);
        put(g.dscales + 3 * idx + 1, ds[1] * p.scale_modifier, a_sc);
        put(g.dscales + 3 * idx + 2, ds[2] * p.scale_modifier, a_sc);
    }
    if (g.drotations && p.rotations) {
        const float4 q = *reinterpret_cast<const float4*>(p.rotations + 4 * idx);
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        float4 dq;
        dq.x = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
        dq.y = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
        dq.z = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
        dq.w = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        if (a_ro) red_add_v4(g.drotations + 4 * idx, dq.x, dq.y, dq.z, dq.w);
        else *reinterpret_cast<float4*>(g.drotations + 4 * idx) = dq;
    }
}

// TEXGS_PREBWD_STAGE_SH: the SH coefficients (and their gradient) of the 32 Gaussians of a warp are one contiguous
// block of global memory, but a thread that owns one Gaussian walks it with a stride of 12 M bytes: every load /
// store instruction touches 32 sectors. Staged: the warp copies its block into shared memory with coalesced
// accesses, every lane works on its row there (row stride odd: conflict-free), the gradient replaces the
// coefficients in place and leaves with coalesced stores (read-modify-write in accumulate mode).
#ifndef TEXGS_PREBWD_STAGE_SH
#define TEXGS_PREBWD_STAGE_SH 1      // measured on B200 (500 k Gaussians, M = 15, accumulate mode): 0.236 -> 0.131 ms
#endif
// staged only when the row stride (3 M words) is odd, i.e. conflict-free in the flat layout; an even stride
// (16 coefficients in the diff_gauss SH mode) keeps the direct path
__host__ __device__ __forceinline__ bool prebwd_stageable(int M) { return M > 0 && ((3 * M) & 1) != 0; }
__host__ __forceinline__ size_t prebwd_smem_bytes(int M) {
#if TEXGS_PREBWD_STAGE_SH
    return prebwd_stageable(M) ? (size_t)8 * 32 * 3 * M * sizeof(float) + 8 * sizeof(uint64_t) : 0;     // rows + one mbarrier per warp
#else
    (void)M;
    return 0;
#endif
}

__global__ void __launch_bounds__(256, TEXGS_PREBWD_MIN_CTAS) texgs_preprocess_bwd(const RasterParams p,
                                                          const float* __restrict__ acc_all, const BwdOut g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsh = p.M * 3;
#if TEXGS_PREBWD_STAGE_SH
    extern __shared__ __align__(128) float prebwd_smem[];
    const bool staged = (p.shs != nullptr) && prebwd_stageable(p.M) && p.mode != TEXGS_MODE_PRECOMP;
    if (staged) {
        const bool a_sh = g.acc & TEXGS_ACC_SHS;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // the warp's 32 rows, in the layout they have in global memory (row stride nsh: odd -> conflict-free)
        float* const wrows = prebwd_smem + (size_t)warp * 32 * nsh;
        uint64_t* const bars = reinterpret_cast<uint64_t*>(prebwd_smem + (size_t)8 * 32 * nsh);
        float* const myrow = wrows + lane * nsh;
        const int g0 = idx - lane;                                                // first Gaussian of the warp
        const int wcnt = max(0, min(32, p.P - g0)) * nsh;                         // floats in the warp's block
        const float* __restrict__ src = p.shs + (size_t)g0 * nsh;
        // a full warp's block is 128 * nsh bytes of contiguous global memory: ONE bulk copy (TMA 1-D) brings it in and
        // one bulk copy / bulk reduction (the L2 adds, the SM never reads the destination) takes the gradient out
        const bool bulk = (wcnt == 32 * nsh) && (((uintptr_t)p.shs | (uintptr_t)g.dshs) & 15) == 0;
        if (bulk) {
            if (lane == 0) {
                mbar_init(&bars[warp], 1);
                mbar_fence_init();
                mbar_arrive_expect_tx(&bars[warp], (unsigned)wcnt * 4u);
                bulk_g2s(wrows, src, (unsigned)wcnt * 4u, &bars[warp]);
            }
        } else {
            for (int i0 = lane; i0 < wcnt; i0 += 32 * 8) {                        // 8 independent loads in flight per lane
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = (i0 + 32 * j < wcnt) ? __ldg(src + i0 + 32 * j) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (i0 + 32 * j < wcnt) wrows[i0 + 32 * j] = t[j];
            }
        }
        // every lane stays until the cooperative store at the end: ``live`` replaces the early returns
        const bool live = idx < p.P;
        bool visible = false;
        Proj o;
        if (live) { project_gaussian(p, idx, o); visible = o.visible; }       // overlaps the copy
        __syncwarp();
        if (bulk) mbar_wait(&bars[warp], 0u);
        if (live && visible) {
            prebwd_one(p, acc_all, g, idx, o, myrow, myrow);
        } else {
            for (int k = 0; k < nsh; ++k) myrow[k] = 0.f;                         // zero gradient row (adds nothing when accumulating)
            if (live) prebwd_write_zero(p, g, idx, /*skip_shs=*/true);
        }
        __syncwarp();
        if (g.dshs) {
            float* __restrict__ dst = g.dshs + (size_t)g0 * nsh;
            if (bulk) {
                fence_proxy_async_smem();          // the rows were written with ordinary stores
                __syncwarp();
                if (lane == 0) {
                    if (a_sh) bulk_reduce_add_f32(dst, wrows, (unsigned)wcnt * 4u);
                    else bulk_s2g(dst, wrows, (unsigned)wcnt * 4u);
                    bulk_commit();
                    bulk_wait<0>();                // the shared-memory source must outlive the copy
                }
            } else {
                for (int i0 = lane; i0 < wcnt; i0 += 32) {
                    if (a_sh) atomicAdd(dst + i0, wrows[i0]); else dst[i0] = wrows[i0];
                }
            }
        }
        return;
    }
#endif
    if (idx >= p.P) return;
    Proj o;
    project_gaussian(p, idx, o);
    if (!o.visible) { prebwd_write_zero(p, g, idx, false); return; }
    prebwd_one(p, acc_all, g, idx, o, p.shs ? p.shs + (size_t)idx * nsh : nullptr, g.dshs ? g.dshs + (size_t)idx * nsh : nullptr, g.acc & TEXGS_ACC_SHS);
}
