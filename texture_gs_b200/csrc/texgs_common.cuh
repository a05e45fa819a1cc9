// texgs_common.cuh — shared device helpers for the sm_100a Texture-GS rasterizer.
//
// Data layout in HBM:
//   * GaussRec: one 128-byte, 128-byte-aligned projected record per Gaussian (indexed by Gaussian
//     id, only written for visible Gaussians). One record = one L2 line = one cp.async.bulk.
//   * pairs / sorted ids: per-tile contiguous segments (tile_offset[t] .. tile_offset[t+1]).
#pragma once
#ifndef TEXGS_HOST_EMU          // tests/simt/simt_emu.h (host-side SIMT emulation of these sources, tests only) stands in
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/texgs.h"

#define TEXGS_TILE 16
#define TEXGS_ALPHA_MIN (1.0f / 255.0f)
#define TEXGS_ALPHA_MAX 0.99f
#define TEXGS_T_STOP 1e-4f
#define TEXGS_NEAR 0.2f
#define TEXGS_ND_EPS 1e-8f

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                          0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

// 128-byte projected record (8 x float4). Field map (f = float index):
//   f0  x        f1  y        f2  conic_a  f3  conic_b      | q0  \ everything the alpha test needs
//   f4  conic_c  f5  opacity  f6  depth    f7  nm = n_v.p_v | q1  / lives in the first 32-byte sector
//   f8  nv.x     f9  nv.y     f10 nv.z     f11 pv.x         | q2   view-space disc normal / centre
//   f12 pv.y     f13 col.r    f14 col.g    f15 col.b        | q3   col = SH_rest + 0.5 (pre-clamp)
//   f16 uv.x     f17 uv.y     f18 uv.z     f19 J'00         | q4   J' = J * (view->world rot)
//   f20 J'01     f21 J'02     f22 J'10     f23 J'11         | q5
//   f24 J'12     f25 J'20     f26 J'21     f27 J'22         | q6
//   f28 id(bits) f29..31 spare                              | q7
struct __align__(128) GaussRec { float4 q[8]; };
static_assert(sizeof(GaussRec) == 128, "record must be one 128 B line");

struct Mat4 { float m[16]; };   // row-major as torch stores it: element (r,c) = m[4*r+c]

// Kernel parameter block (passed by value).
struct RasterParams {
    int P, M, sh_degree, E, H, W, R, mode;
    unsigned flags;
    int grid_x, grid_y, num_tiles;
    float tanfovx, tanfovy, scale_modifier;
    float focal_x, focal_y;
    Mat4 view, proj;
    float campos[3];
    float bg[3];
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *uvs,
        *gradient_uvs, *texture, *extra_attrs, *cov3Ds_precomp;
    const float4* texture_rgba;   // optional (6,R,R,4) copy of the texture
    // workspaces
    GaussRec* recs;
    uint2* rects;            // 4 x uint16 packed: .x = x0 | x1<<16, .y = y0 | y1<<16
    TexgsCounters* counters;
    unsigned* tile_count;
    unsigned* tile_offset;
    unsigned* tile_cursor;
    uint2* pairs;            // .x = gaussian id, .y = depth bits  (as u64: depth in the high word)
    unsigned* sorted_ids;
    uint2* cull_masks;       // per (list chunk, warp): entries of the chunk that can touch the warp's left (.x) / right (.y) 4x4 block
    unsigned long long pair_capacity;
    float* final_T;
    unsigned* n_contrib;
    float* out_image_nosh;   // dual render: second image (sh_degree = 0 colour), or NULL
};

// ---------------------------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// p_view / p_hom with the reference's row-vector convention: out[c] = sum_r p[r] M[r][c] + M[3][c]
__device__ __forceinline__ float3 xform43(const Mat4& M, float3 p) {
    return f3(M.m[0] * p.x + M.m[4] * p.y + M.m[8] * p.z + M.m[12],
              M.m[1] * p.x + M.m[5] * p.y + M.m[9] * p.z + M.m[13],
              M.m[2] * p.x + M.m[6] * p.y + M.m[10] * p.z + M.m[14]);
}
__device__ __forceinline__ float4 xform44(const Mat4& M, float3 p) {
    return make_float4(M.m[0] * p.x + M.m[4] * p.y + M.m[8] * p.z + M.m[12],
                       M.m[1] * p.x + M.m[5] * p.y + M.m[9] * p.z + M.m[13],
                       M.m[2] * p.x + M.m[6] * p.y + M.m[10] * p.z + M.m[14],
                       M.m[3] * p.x + M.m[7] * p.y + M.m[11] * p.z + M.m[15]);
}
// world -> view rotation of a direction: out[c] = sum_r d[r] V[r][c]
__device__ __forceinline__ float3 rot_w2v(const Mat4& V, float3 d) {
    return f3(V.m[0] * d.x + V.m[4] * d.y + V.m[8] * d.z,
              V.m[1] * d.x + V.m[5] * d.y + V.m[9] * d.z,
              V.m[2] * d.x + V.m[6] * d.y + V.m[10] * d.z);
}
// view -> world rotation of a direction: out[r] = sum_c V[r][c] d[c]
__device__ __forceinline__ float3 rot_v2w(const Mat4& V, float3 d) {
    return f3(V.m[0] * d.x + V.m[1] * d.y + V.m[2] * d.z,
              V.m[4] * d.x + V.m[5] * d.y + V.m[6] * d.z,
              V.m[8] * d.x + V.m[9] * d.y + V.m[10] * d.z);
}

// quaternion (r,x,y,z) -> rotation, row-major R[3*r+c]  (reference utils/general.py:87-108)
__device__ __forceinline__ void quat_to_rot(float4 q, float* R) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z);       R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);       R[7] = 2.f * (y * z + r * x);       R[8] = 1.f - 2.f * (x * x + y * y);
}

// argmin of the three scales, first index wins ties (torch.argmin semantics on ties are
// "first occurrence")
__device__ __forceinline__ int argmin3(float a, float b, float c) {
    int k = 0; float m = a;
    if (b < m) { m = b; k = 1; }
    if (c < m) { k = 2; }
    return k;
}

// Unit eigenvector of the smallest eigenvalue of the symmetric 3x3 matrix S = (xx,xy,xz,yy,yz,zz): cyclic
// Jacobi rotations (high relative accuracy also for the flat discs of this path, whose smallest eigenvalue
// is ~1e-13 of the others). Cold path (cov3Ds_precomp only), kept out of line.
#define TEXGS_JACOBI_ROT(app, aqq, apq, akp, akq, v0p, v0q, v1p, v1q, v2p, v2q)                           \
    if (apq != 0.f) {                                                                                    \
        const float theta = 0.5f * (aqq - app) / apq;                                                    \
        const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));             \
        const float c = rsqrtf(t * t + 1.f), s = t * c, tau = s / (1.f + c);                             \
        const float h = t * apq;                                                                         \
        app -= h; aqq += h; apq = 0.f;                                                                   \
        float g_ = akp, h_ = akq; akp = g_ - s * (h_ + g_ * tau); akq = h_ + s * (g_ - h_ * tau);        \
        g_ = v0p; h_ = v0q; v0p = g_ - s * (h_ + g_ * tau); v0q = h_ + s * (g_ - h_ * tau);              \
        g_ = v1p; h_ = v1q; v1p = g_ - s * (h_ + g_ * tau); v1q = h_ + s * (g_ - h_ * tau);              \
        g_ = v2p; h_ = v2q; v2p = g_ - s * (h_ + g_ * tau); v2q = h_ + s * (g_ - h_ * tau);              \
    }
__device__ __noinline__ float3 smallest_eigvec_sym3(float a00, float a01, float a02, float a11, float a12, float a22) {
    // scale-invariant in exact arithmetic; normalise so squares of ~1e-9 entries do not underflow
    const float sc = fmaxf(fmaxf(fabsf(a00), fabsf(a11)), fmaxf(fabsf(a22), 1e-37f));
    const float is = 1.0f / sc;
    a00 *= is; a01 *= is; a02 *= is; a11 *= is; a12 *= is; a22 *= is;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        TEXGS_JACOBI_ROT(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21)
        TEXGS_JACOBI_ROT(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22)
        TEXGS_JACOBI_ROT(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22)
    }
    float3 n = make_float3(v00, v10, v20);
    float lam = a00;
    if (a11 < lam) { lam = a11; n = make_float3(v01, v11, v21); }
    if (a22 < lam) { n = make_float3(v02, v12, v22); }
    const float inv = rsqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    return make_float3(n.x * inv, n.y * inv, n.z * inv);
}

// ---------------------------------------------------------------------------------------------
// cube map lookup (E11). Face layout = reference NVDIFFREC/util.py:94-101 / cubemap.cu:32-61.
// ---------------------------------------------------------------------------------------------
struct CubeCoord {
    int face;        // 0..5
    bool isx, isy;   // major axis x / y (else z)
    float inv_m;     // 1/|u_major|
    float sx, sy;    // face coordinates in [-1,1]
    // sx = sgx * u[ix] * inv_m ; sy = sgy * u[iy] * inv_m ; |u_major| = sgm * u[axis]
    //   axis x: (ix,iy) = (2,1)   axis y: (0,2)   axis z: (0,1)
    float sgx, sgy, sgm;
};

// Branch-free face selection. With m = |major|:
//   x: sx = -z/x , sy = -y/m     y: sx = x/m , sy = z/y     z: sx = x/z , sy = -y/m
__device__ __forceinline__ CubeCoord cube_coord(float ux, float uy, float uz) {
    CubeCoord c;
    const float ax = fabsf(ux), ay = fabsf(uy), az = fabsf(uz);
    c.isx = (ax >= ay) && (ax >= az);
    c.isy = !c.isx && (ay >= az);
    const float maj = c.isx ? ux : (c.isy ? uy : uz);
    const bool neg = maj < 0.f;
    c.sgm = neg ? -1.f : 1.f;
    c.face = (c.isx ? 0 : (c.isy ? 2 : 4)) + (neg ? 1 : 0);
    c.inv_m = __fdividef(1.0f, fmaxf(fabsf(maj), 1e-20f));
    const float a = c.isx ? uz : ux;
    const float b = c.isy ? uz : uy;
    c.sgx = c.isx ? -c.sgm : (c.isy ? 1.f : c.sgm);
    c.sgy = c.isy ? c.sgm : -1.f;
    c.sx = c.sgx * a * c.inv_m;
    c.sy = c.sgy * b * c.inv_m;
    return c;
}

// d(sx,sy)/du applied to (dsx, dsy) (both already multiplied by inv_m): gradient w.r.t. u
__device__ __forceinline__ void cube_coord_bwd(const CubeCoord& c, float dsx, float dsy, float (&gu)[3]) {
    const float gx = c.sgx * dsx, gy = c.sgy * dsy;
    const float gm = -c.sgm * (c.sx * dsx + c.sy * dsy);
    gu[0] = c.isx ? gm : gx;
    gu[1] = c.isy ? gm : gy;
    gu[2] = c.isx ? gx : (c.isy ? gy : gm);
}

struct Bilerp {
    int i00, i01, i10, i11;   // TEXEL indices of the four taps (x3 = float offset in the rgb layout)
    float wx, wy;
};

__device__ __forceinline__ Bilerp cube_bilerp(const CubeCoord& c, int R) {
    Bilerp b;
    const float halfR = 0.5f * (float)R;
    // |minor|/|major| <= 1 up to rounding; the clamp also squashes inf/nan from degenerate inputs, so
    // floor(f) is in [-1, R-1] and only one-sided index clamps remain
    const float sx = fminf(fmaxf(c.sx, -1.0f), 1.0f), sy = fminf(fmaxf(c.sy, -1.0f), 1.0f);
    const float fx = (sx + 1.0f) * halfR - 0.5f;
    const float fy = (sy + 1.0f) * halfR - 0.5f;
    const float x0f = floorf(fx), y0f = floorf(fy);
    b.wx = fx - x0f;
    b.wy = fy - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x0c = max(x0, 0), x1c = min(x0 + 1, R - 1);
    const int y0c = max(y0, 0), y1c = min(y0 + 1, R - 1);
    const int row0 = (c.face * R + y0c) * R, row1 = (c.face * R + y1c) * R;
    b.i00 = row0 + x0c;
    b.i01 = row0 + x1c;
    b.i10 = row1 + x0c;
    b.i11 = row1 + x1c;
    return b;
}

// ---------------------------------------------------------------------------------------------
// Spec switch E11-alt (TEXGS_FLAG_SEAMLESS_CUBE): seamless filtering. A tap one texel beyond an edge of the face is the
// texel of the adjacent face that touches the same edge at the same place — what GL_TEXTURE_CUBE_MAP_SEAMLESS /
// nvdiffrast boundary_mode='cube' (the reference's own visuals, models/uv_map_gaussian3d.py:259) fetch there.
// CUBE_WRAP_TABLE[f][side] = {f', xcode, ycode}; side 0: x < 0, 1: x >= R, 2: y < 0, 3: y >= R; codes 0 -> 0, 1 -> R-1,
// 2 -> k, 3 -> R-1-k (k = position along the edge). Derived from the reference's face table by tools/gen_cube_wrap.py.
// Cold path: only the kernels' ALT instantiations reference it.
// ---------------------------------------------------------------------------------------------
__device__ __constant__ unsigned char CUBE_WRAP_TABLE[6][4][3] = {
    {{4, 1, 2}, {5, 0, 2}, {2, 1, 3}, {3, 1, 2}}, {{5, 1, 2}, {4, 0, 2}, {2, 0, 2}, {3, 0, 3}},
    {{1, 2, 0}, {0, 3, 0}, {5, 3, 0}, {4, 2, 0}}, {{1, 3, 1}, {0, 2, 1}, {4, 2, 1}, {5, 3, 1}},
    {{1, 1, 2}, {0, 0, 2}, {2, 2, 1}, {3, 2, 0}}, {{0, 1, 2}, {1, 0, 2}, {2, 3, 0}, {3, 3, 1}}};

// texel index of tap (face, x, y), x and y in [-1, R]; a tap beyond a corner is clamped in y first
__device__ __noinline__ int cube_wrap_tap(int face, int x, int y, int R) {
    const bool xo = (x < 0) || (x >= R);
    bool yo = (y < 0) || (y >= R);
    if (xo && yo) { y = min(max(y, 0), R - 1); yo = false; }
    if (xo || yo) {
        const int side = xo ? (x < 0 ? 0 : 1) : (y < 0 ? 2 : 3);
        const int k = min(max(xo ? y : x, 0), R - 1);
        const unsigned char* e = CUBE_WRAP_TABLE[face][side];
        const int vals[4] = {0, R - 1, k, R - 1 - k};
        face = e[0]; x = vals[e[1]]; y = vals[e[2]];
    }
    return (face * R + y) * R + x;
}

__device__ __forceinline__ Bilerp cube_bilerp_seamless(const CubeCoord& c, int R) {
    Bilerp b;
    const float halfR = 0.5f * (float)R;
    const float sx = fminf(fmaxf(c.sx, -1.0f), 1.0f), sy = fminf(fmaxf(c.sy, -1.0f), 1.0f);
    const float fx = (sx + 1.0f) * halfR - 0.5f;
    const float fy = (sy + 1.0f) * halfR - 0.5f;
    const float x0f = floorf(fx), y0f = floorf(fy);
    b.wx = fx - x0f;
    b.wy = fy - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;          // in [-1, R-1]
    if (x0 >= 0 && y0 >= 0 && x0 + 1 < R && y0 + 1 < R) {
        const int row0 = (c.face * R + y0) * R;
        b.i00 = row0 + x0; b.i01 = b.i00 + 1; b.i10 = b.i00 + R; b.i11 = b.i10 + 1;
    } else {
        b.i00 = cube_wrap_tap(c.face, x0, y0, R);     b.i01 = cube_wrap_tap(c.face, x0 + 1, y0, R);
        b.i10 = cube_wrap_tap(c.face, x0, y0 + 1, R); b.i11 = cube_wrap_tap(c.face, x0 + 1, y0 + 1, R);
    }
    return b;
}

// Can  q(d) = a dx^2 + 2 b dx dy + c dy^2  (d = p - mu) drop to <= tau somewhere on the rectangle
// [x0,x1] x [y0,y1]?  q is convex with its minimum at mu, so the constrained minimum is either mu
// itself (inside) or lies on an edge that FACES mu; at most two 1-D clamped minimisations.
__device__ __forceinline__ bool splat_hits_block(float mx, float my, float a, float b, float c, float opacity,
                                                 float x0, float x1, float y0, float y1) {
    const bool inx = (mx >= x0) && (mx <= x1), iny = (my >= y0) && (my <= y1);
    if (inx && iny) return true;
    float q = 3.0e38f;
    if (!inx) {
        const float dx = ((mx < x0) ? x0 : x1) - mx;
        const float dy = fminf(fmaxf(__fdividef(-b * dx, c), y0 - my), y1 - my);
        q = a * dx * dx + 2.f * b * dx * dy + c * dy * dy;
    }
    if (!iny) {
        const float dy = ((my < y0) ? y0 : y1) - my;
        const float dx = fminf(fmaxf(__fdividef(-b * dy, a), x0 - mx), x1 - mx);
        q = fminf(q, a * dx * dx + 2.f * b * dx * dy + c * dy * dy);
    }
    // alpha >= 1/255  <=>  q <= 2 ln(255 o); keep a margin far above the rounding of __expf/__logf
    const float tau = 2.0f * __logf(255.0f * opacity);
    return q <= tau * 1.001f + 0.02f;
}

// (tile, Gaussian) pair test used identically by the count (preprocess) and the scatter kernels: can the
// splat reach alpha >= 1/255 anywhere on tile (tx, ty)? Exact up to the margin above, so dropping
// the pair never changes a pixel; it only shortens the lists (spec E3's tile rect stays the outer bound).
__device__ __forceinline__ bool splat_hits_tile(float mx, float my, float a, float b, float c, float opacity, int tx, int ty) {
    const float x0 = (float)(tx * TEXGS_TILE), y0 = (float)(ty * TEXGS_TILE);
    return splat_hits_block(mx, my, a, b, c, opacity, x0, x0 + (float)(TEXGS_TILE - 1), y0, y0 + (float)(TEXGS_TILE - 1));
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D) — sm_90+/sm_100a PTX
// ---------------------------------------------------------------------------------------------
#ifndef TEXGS_HOST_EMU          // the emulator provides host versions of the PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (bytes multiple of 16, both 16-B aligned), completes on ``bar``
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy / bulk reduction (TMA 1-D, bulk_group completion). The shared-memory source was written with
// ordinary stores: fence_proxy_async_smem() first; the issuing thread commits and waits before the block may end.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// global[i] += shared[i] for bytes/4 floats, done by the L2 (no read of the destination by the SM): SASS UBLKRED
__device__ __forceinline__ void bulk_reduce_add_f32(float* gmem_dst, const float* smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif



// four bilinear taps as rgb triples, from the packed (float4) or the plain (3 floats) layout
// timing ablation only: 1 = all taps read texels 0..255 (always L1 hits) — same instructions, no gather traffic
#ifndef TEXGS_ABLATE_TAP_LOCALITY
#define TEXGS_ABLATE_TAP_LOCALITY 0
#endif
template <bool TEX4>
__device__ __forceinline__ void fetch_taps(const float* __restrict__ tex, const float4* __restrict__ tex4, const Bilerp& bl_,
                                           float (&t00)[3], float (&t01)[3], float (&t10)[3], float (&t11)[3]) {
#if TEXGS_ABLATE_TAP_LOCALITY
    Bilerp bl = bl_;
    bl.i00 &= 255; bl.i01 &= 255; bl.i10 &= 255; bl.i11 &= 255;
#else
    const Bilerp& bl = bl_;
#endif
    if (TEX4) {
        const float4 a = __ldg(tex4 + bl.i00), b = __ldg(tex4 + bl.i01), c = __ldg(tex4 + bl.i10), d = __ldg(tex4 + bl.i11);
        t00[0] = a.x; t00[1] = a.y; t00[2] = a.z;
        t01[0] = b.x; t01[1] = b.y; t01[2] = b.z;
        t10[0] = c.x; t10[1] = c.y; t10[2] = c.z;
        t11[0] = d.x; t11[1] = d.y; t11[2] = d.z;
    } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            t00[ch] = __ldg(tex + 3 * bl.i00 + ch); t01[ch] = __ldg(tex + 3 * bl.i01 + ch);
            t10[ch] = __ldg(tex + 3 * bl.i10 + ch); t11[ch] = __ldg(tex + 3 * bl.i11 + ch);
        }
    }
}

// 128-bit vector reduction (sm_90+): one L2 atomic op for a whole padded texel
#ifndef TEXGS_HOST_EMU
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

// Transposing reduction of 20 per-lane values inside each half-warp: instead of 4 shuffles per value
// (80), every step halves the number of values a lane still owns (16 -> 8 -> 4 -> 2 -> 1, then 4 -> 2 ->
// 1): 20 shuffles, and lanes 0-15 / 16-31 reduce independently in the same instructions (xor distances
// 8,4,2,1 never cross the halves). On return, inside each half,
//   outA on lane l holds the half's total of value (l & 15)                       (values 0..15)
//   outB on lane l holds the half's total of value 16 + ((l >> 2) & 3)            (values 16..19)
__device__ __forceinline__ void halfwarp_reduce20(const float (&v)[20], int lane, float& outA, float& outB) {
    const unsigned full = 0xffffffffu;
    const bool h3 = (lane & 8) != 0, h2 = (lane & 4) != 0, h1 = (lane & 2) != 0, h0 = (lane & 1) != 0;
    float a8[8], a4[4], a2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h3 ? v[i] : v[i + 8], keep = h3 ? v[i + 8] : v[i];
        a8[i] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h2 ? a8[i] : a8[i + 4], keep = h2 ? a8[i + 4] : a8[i];
        a4[i] = keep + __shfl_xor_sync(full, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h1 ? a4[i] : a4[i + 2], keep = h1 ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(full, send, 2);
    }
    {
        const float send = h0 ? a2[0] : a2[1], keep = h0 ? a2[1] : a2[0];
        outA = keep + __shfl_xor_sync(full, send, 1);      // value index 8*h3 + 4*h2 + 2*h1 + h0 = lane & 15
    }
    float b2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h3 ? v[16 + i] : v[18 + i], keep = h3 ? v[18 + i] : v[16 + i];
        b2[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const float send = h2 ? b2[0] : b2[1], keep = h2 ? b2[1] : b2[0];
        float b1 = keep + __shfl_xor_sync(full, send, 4);   // value index 16 + 2*h3 + h2
        b1 += __shfl_xor_sync(full, b1, 2);
        b1 += __shfl_xor_sync(full, b1, 1);
        outB = b1;
    }
}
