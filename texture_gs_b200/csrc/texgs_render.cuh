// texgs_render.cuh — per-tile blend kernels (SURVEY §8a rows a7 forward, a8 backward).
//
// One CTA = one 16x16 tile = 8 independent warps; warp w owns an 8x4 pixel block (one pixel per
// lane). Every warp streams the tile's depth-sorted list on its own: exact splat-vs-block culling at
// load time, one 128-byte cp.async.bulk (TMA 1-D) per surviving record into a private 2-stage ring
// tracked by mbarriers, no block barrier in the loop (see "Per-warp streaming" below).
// Spec items E5-E12.
#pragma once
#include "texgs_common.cuh"


// Warps per CTA of the two render kernels. The eight warps of a tile never talk to each other (no block barrier, private
// rings), so a tile may be split over 8 / WARPS CTAs: with half-tile CTAs a finished warp's slot returns to the scheduler
// when its three siblings are done instead of seven, and the backward fits 5 CTAs x 4 warps = 20 warps per SM at 96
// registers (28 B of spills) where 2 x 8 warps at 122 registers gave 16. Measured on B200 (profiles/r2_experiments.md):
// forward 0.776 -> 0.763 ms, backward 1.373 -> 1.323 ms; quarter-tile and single-warp CTAs lose the tile's L1 locality.
#ifndef TEXGS_FWD_WARPS
#define TEXGS_FWD_WARPS 4
#endif
#ifndef TEXGS_BWD_WARPS
#define TEXGS_BWD_WARPS 4
#endif
#ifndef TEXGS_FWD_MIN_CTAS
#define TEXGS_FWD_MIN_CTAS 6
#endif
#ifndef TEXGS_BWD_MIN_CTAS
#define TEXGS_BWD_MIN_CTAS 5
#endif
#ifndef TEXGS_FAST_EXP
#define TEXGS_FAST_EXP 1
#endif
// timing ablation only (tools/build_variants.py): 1 = skip the per-contribution texture work (intersection, UV step,
// cube lookup, taps and, in the backward, their gradients) to measure what the rest of the loop costs. Wrong images.
#ifndef TEXGS_ABLATE_HEAVY
#define TEXGS_ABLATE_HEAVY 0
#endif
#ifndef TEXGS_ABLATE_RED
#define TEXGS_ABLATE_RED 0         // timing ablation only: no texel-gradient atomics in the backward (wrong texture gradient)
#endif
#ifndef TEXGS_ABLATE_REDUCE
#define TEXGS_ABLATE_REDUCE 0      // timing ablation only: no per-Gaussian lane reduction in the backward (wrong gradients)
#endif
__device__ __forceinline__ float texgs_exp(float x) {
#if TEXGS_FAST_EXP
    return __expf(x);
#else
    return expf(x);
#endif
}

struct PixelGeom {
    int tile, warp, px, py, pix;   // warp = 0..7 inside the tile
    int bx, by;      // origin of this warp's 8x4 pixel block (= two 4x4 half-warp blocks side by side)
    bool inside;
    float vx, vy;    // view ray (vx, vy, 1)
};

// WPC = warps per CTA (8, 4, 2 or 1): a tile's eight warps never talk to each other, so a tile may be split over 8 / WPC CTAs
template <int WPC>
__device__ __forceinline__ PixelGeom pixel_geom(const RasterParams& p) {
    PixelGeom g;
    g.tile = blockIdx.x / (8 / WPC);
    const int tx = g.tile % p.grid_x, ty = g.tile / p.grid_x;
    const int warp = (blockIdx.x % (8 / WPC)) * WPC + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    g.warp = warp;
    g.bx = tx * TEXGS_TILE + (warp & 1) * 8;
    g.by = ty * TEXGS_TILE + (warp >> 1) * 4;
    // half-warp h = lane >> 4 owns the 4x4 block at (bx + 4h, by): the two halves walk their own
    // survivor lists, so one pass of the blend loop serves two splats
    g.px = g.bx + 4 * (lane >> 4) + (lane & 3);
    g.py = g.by + ((lane >> 2) & 3);
    g.inside = (g.px < p.W) && (g.py < p.H);
    g.pix = g.py * p.W + g.px;
    g.vx = ((2.0f * (float)g.px + 1.0f) / (float)p.W - 1.0f) * p.tanfovx;
    g.vy = ((2.0f * (float)g.py + 1.0f) / (float)p.H - 1.0f) * p.tanfovy;
    return g;
}

// ---------------------------------------------------------------------------------------------
// Per-warp streaming of the tile's sorted list.
//
// Each warp walks the list in chunks of 32 entries (one per lane). A lane reads its entry's id,
// then the first 32-byte sector of that Gaussian's record (centre, conic, opacity) and tests
// EXACTLY whether the splat's alpha >= 1/255 ellipse intersects the warp's 8x4 pixel block; only
// the survivors are gathered — one 128-byte cp.async.bulk each — compacted in list order into the
// warp's private stage (2 stages, an mbarrier each). While chunk c is blended, the copies of
// chunk c+1 are landing and the id / cull-sector loads of chunk c+2 are in flight. There is no
// block-level barrier in the loop and every warp stops on its own.
// ---------------------------------------------------------------------------------------------
#ifndef TEXGS_CHUNK
#define TEXGS_CHUNK 32          // list entries per chunk (<= 32: one per lane) = records per stage
#endif
#define TEXGS_STAGES 2

struct __align__(128) WarpSmem {
    GaussRec rec[TEXGS_STAGES][TEXGS_CHUNK];
    uint64_t bar[TEXGS_STAGES];
    unsigned maskL[TEXGS_STAGES];   // entries of the chunk that can touch the left  4x4 block (lanes 0-15)
    unsigned maskR[TEXGS_STAGES];   //                                   ... the right 4x4 block (lanes 16-31)
    unsigned pad[24];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "WarpSmem must keep the records 128-byte aligned");
#define TEXGS_RENDER_SMEM (8 * sizeof(WarpSmem))

__device__ __forceinline__ unsigned stream_load_id(const RasterParams& p, unsigned start, unsigned limit, int chunk, int lane,
                                                   bool& valid) {
    const unsigned e = (unsigned)chunk * TEXGS_CHUNK + (unsigned)lane;
    valid = (chunk >= 0) && (lane < TEXGS_CHUNK) && (e < limit);
    return valid ? __ldg(p.sorted_ids + start + e) : 0u;
}

// gather the entries of a chunk whose masks say they can touch one of the warp's two blocks into stage ``s``
__device__ __forceinline__ void stream_gather(const RasterParams& p, WarpSmem& ws, int s, int lane, unsigned id, unsigned mL, unsigned mR) {
    const unsigned m = mL | mR;
    if (lane == 0) {
        ws.maskL[s] = mL;
        ws.maskR[s] = mR;
        mbar_arrive_expect_tx(&ws.bar[s], (unsigned)__popc(m) * (unsigned)sizeof(GaussRec));
    }
    if ((m >> lane) & 1u) {
        const int slot = __popc(m & ((1u << lane) - 1u));
        bulk_g2s(&ws.rec[s][slot], p.recs + id, (unsigned)sizeof(GaussRec), &ws.bar[s]);
    }
}

// where the masks of (tile, chunk, warp) live: chunk c of a list that starts at ``start`` gets slot (start / CHUNK) + tile + c
// (the + tile keeps the slots of consecutive lists apart: a list of n entries has at most n / CHUNK + 1 chunks)
__device__ __forceinline__ size_t cull_mask_slot(unsigned start, const PixelGeom& g, int chunk) {
    return ((size_t)(start / TEXGS_CHUNK) + (size_t)g.tile + (size_t)chunk) * 8 + g.warp;
}

// forward: ballot the exact cull test of one chunk, keep the two masks for the backward, gather the survivors
__device__ __forceinline__ void stream_issue(const RasterParams& p, WarpSmem& ws, int s, int lane, unsigned id, bool valid,
                                             const float4& q0, const float4& q1, const PixelGeom& g, unsigned start, int chunk) {
    const float y0 = (float)g.by, y1 = (float)(g.by + 3);
    const bool passL = valid && splat_hits_block(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, (float)g.bx, (float)(g.bx + 3), y0, y1);
    const bool passR = valid && splat_hits_block(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, (float)(g.bx + 4), (float)(g.bx + 7), y0, y1);
    const unsigned mL = __ballot_sync(0xffffffffu, passL), mR = __ballot_sync(0xffffffffu, passR);
    if (lane == 0) p.cull_masks[cull_mask_slot(start, g, chunk)] = make_uint2(mL, mR);
    stream_gather(p, ws, s, lane, id, mL, mR);
}

// backward: the forward decided already which entries of a chunk reach which block (the test depends on the record and
// the block only); entries at or beyond ``limit`` (nothing blended behind the warp's last contribution) are dropped
__device__ __forceinline__ void stream_issue_saved(const RasterParams& p, WarpSmem& ws, int s, int lane, unsigned id, uint2 m,
                                                   int chunk, unsigned limit) {
    const unsigned base = (unsigned)chunk * TEXGS_CHUNK;
    const unsigned live = (limit >= base + TEXGS_CHUNK) ? 0xffffffffu : ((limit > base) ? ((1u << (limit - base)) - 1u) : 0u);
    stream_gather(p, ws, s, lane, id, m.x & live, m.y & live);
}

// u' = uv + J' (t v - p_v)   (E9/E10, evaluated in view space)
struct UvEval {
    float ux, uy, uz;
    float nd, t;
    float dx, dy, dz;   // Delta in view space
    bool safe;
};
__device__ __forceinline__ UvEval eval_uv(const float4& g1, const float4& g2, const float4& g3, const float4& g4,
                                          const float4& g5, const float4& g6, float vx, float vy) {
    UvEval e;
    e.nd = g2.x * vx + g2.y * vy + g2.z;
    e.ux = g4.x; e.uy = g4.y; e.uz = g4.z;
    e.safe = fabsf(e.nd) >= TEXGS_ND_EPS;
    e.t = 0.f;
    e.dx = e.dy = e.dz = 0.f;
    if (e.safe) {
        e.t = __fdividef(g1.w, e.nd);
        e.dx = e.t * vx - g2.w; e.dy = e.t * vy - g3.x; e.dz = e.t - g1.z;
        e.ux += g4.w * e.dx + g5.x * e.dy + g5.y * e.dz;
        e.uy += g5.z * e.dx + g5.w * e.dy + g6.x * e.dz;
        e.uz += g6.y * e.dx + g6.z * e.dy + g6.w * e.dz;
    }
    return e;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// ALT = true: cold instantiation that honours the spec switches of p.flags (TEXGS_FLAG_SEAMLESS_CUBE,
// TEXGS_FLAG_DEPTH_INTERSECTION; include/texgs.h) — the default instantiations never test them.
template <int MODE, bool TEX4, bool DUAL, bool ALT>
__global__ void __launch_bounds__(32 * TEXGS_FWD_WARPS, TEXGS_FWD_MIN_CTAS) texgs_render_fwd(const RasterParams p, float* __restrict__ out_image,
                                                      float* __restrict__ out_depth, float* __restrict__ out_norm,
                                                      float* __restrict__ out_alpha) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.counters->overflow) return;
    const PixelGeom g = pixel_geom<TEXGS_FWD_WARPS>(p);
    const int lane = threadIdx.x & 31;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const unsigned start = p.tile_offset[g.tile];
    const unsigned n = p.tile_offset[g.tile + 1] - start;
    const int nchunks = (int)((n + TEXGS_CHUNK - 1) / TEXGS_CHUNK);

    float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, D = 0.f, Nx = 0.f, Ny = 0.f, Nz = 0.f, A = 0.f;
    float Er = 0.f, Eg = 0.f, Eb = 0.f;     // dual render: colour without the SH term
    unsigned last = 0, nblend = 0;
    bool done = !g.inside;
    const float pxf = (float)g.px, pyf = (float)g.py;
    const float* __restrict__ tex = p.texture;
    const int R = p.R;
    const bool seamless = ALT && (p.flags & TEXGS_FLAG_SEAMLESS_CUBE) != 0u;
    const bool zalt = ALT && (p.flags & TEXGS_FLAG_DEPTH_INTERSECTION) != 0u;

    // a warp whose block lies completely outside the image has nothing to do
    if (nchunks > 0 && !__all_sync(0xffffffffu, done)) {
        if (lane == 0) {
            mbar_init(&ws.bar[0], 1);
            mbar_init(&ws.bar[1], 1);
            mbar_fence_init();
        }
        __syncwarp();
        // prologue: chunks 0 and 1 are gathered, the id of chunk 2 is loaded
        bool v0, v1, v2;
        const unsigned id0 = stream_load_id(p, start, n, 0, lane, v0);
        const unsigned id1 = stream_load_id(p, start, n, 1 < nchunks ? 1 : -1, lane, v1);
        unsigned id2 = stream_load_id(p, start, n, 2 < nchunks ? 2 : -1, lane, v2);
        {
            const float4* r0 = reinterpret_cast<const float4*>(p.recs + id0);
            const float4 a0 = v0 ? __ldg(r0) : make_float4(0.f, 0.f, 1.f, 0.f), a1 = v0 ? __ldg(r0 + 1) : make_float4(1.f, 0.f, 0.f, 0.f);
            stream_issue(p, ws, 0, lane, id0, v0, a0, a1, g, start, 0);
            if (1 < nchunks) {
                const float4* r1 = reinterpret_cast<const float4*>(p.recs + id1);
                const float4 b0 = v1 ? __ldg(r1) : make_float4(0.f, 0.f, 1.f, 0.f), b1 = v1 ? __ldg(r1 + 1) : make_float4(1.f, 0.f, 0.f, 0.f);
                stream_issue(p, ws, 1, lane, id1, v1, b0, b1, g, start, 1);
            }
        }
        for (int c = 0; c < nchunks; ++c) {
            const int s = c & 1;
            // loads for the chunks ahead: cull sector of chunk c+2 (its id arrived during the last
            // blend phase), id of chunk c+3
            float4 q0n = make_float4(0.f, 0.f, 1.f, 0.f), q1n = make_float4(1.f, 0.f, 0.f, 0.f);
            const unsigned idn = id2;
            const bool vn = v2;
            if (vn) {
                const float4* rn = reinterpret_cast<const float4*>(p.recs + idn);
                q0n = __ldg(rn);
                q1n = __ldg(rn + 1);
            }
            id2 = stream_load_id(p, start, n, (c + 3 < nchunks) ? c + 3 : -1, lane, v2);

            mbar_wait(&ws.bar[s], (unsigned)(c >> 1) & 1u);
            __syncwarp();
            bool warp_done = false;
            const unsigned mL = ws.maskL[s], mR = ws.maskR[s], mU = mL | mR;
            unsigned my = (lane & 16) ? mR : mL;          // this half-warp's survivors, in list order
            while (__any_sync(0xffffffffu, my != 0u)) {
                const bool has = my != 0u;
                const int l = has ? (__ffs(my) - 1) : 0;
                my &= my - 1u;
                const GaussRec& rec = ws.rec[s][__popc(mU & ((1u << l) - 1u))];   // half-uniform address
                const unsigned idx1 = (unsigned)c * TEXGS_CHUNK + (unsigned)l + 1u;
                const float4 g0 = rec.q[0], g1 = rec.q[1];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const float alpha = fminf(TEXGS_ALPHA_MAX, g1.y * texgs_exp(power));
                bool cand = has && !done && (power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN);
                if (!__any_sync(0xffffffffu, cand)) continue;
                const float test_T = T * (1.0f - alpha);
                if (cand && test_T < TEXGS_T_STOP) { done = true; cand = false; }
                if (cand) {
                    const float4 g2 = rec.q[2], g3 = rec.q[3];
                    float cr = g3.y, cg = g3.z, cb = g3.w;
                    float zc = g1.z;                                  // E7: z of the centre
                    if (MODE == TEXGS_MODE_TEXTURE && !TEXGS_ABLATE_HEAVY) {
                        const float4 g4 = rec.q[4], g5 = rec.q[5], g6 = rec.q[6];
                        const UvEval e = eval_uv(g1, g2, g3, g4, g5, g6, g.vx, g.vy);
                        if (ALT && zalt && e.safe) zc = e.t;          // E7-alt: z of the intersection (the view ray is (vx, vy, 1))
                        const CubeCoord cc = cube_coord(e.ux, e.uy, e.uz);
                        const Bilerp bl = (ALT && seamless) ? cube_bilerp_seamless(cc, R) : cube_bilerp(cc, R);
                        float tx3[3], t00[3], t01[3], t10[3], t11[3];
                        fetch_taps<TEX4>(tex, p.texture_rgba, bl, t00, t01, t10, t11);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float top = t00[ch] + bl.wx * (t01[ch] - t00[ch]);
                            const float bot = t10[ch] + bl.wx * (t11[ch] - t10[ch]);
                            tx3[ch] = top + bl.wy * (bot - top);
                        }
                        cr = fmaxf(0.f, SH_C0 * tx3[0] + cr);
                        cg = fmaxf(0.f, SH_C0 * tx3[1] + cg);
                        cb = fmaxf(0.f, SH_C0 * tx3[2] + cb);
                        if (DUAL) {
                            const float w0 = alpha * T;
                            Er += w0 * fmaxf(0.f, SH_C0 * tx3[0] + 0.5f);
                            Eg += w0 * fmaxf(0.f, SH_C0 * tx3[1] + 0.5f);
                            Eb += w0 * fmaxf(0.f, SH_C0 * tx3[2] + 0.5f);
                        }
                    }
                    const float w = alpha * T;
                    Cr += w * cr; Cg += w * cg; Cb += w * cb;
                    D += w * zc;
                    Nx += w * g2.x; Ny += w * g2.y; Nz += w * g2.z;
                    A += w;
                    T = test_T;
                    last = idx1;
                    ++nblend;
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
            __syncwarp();
            if (warp_done) {
                // never leave with a bulk copy in flight: chunk c+1 was issued one phase ago
                if (c + 1 < nchunks) mbar_wait(&ws.bar[s ^ 1], (unsigned)((c + 1) >> 1) & 1u);
                break;
            }
            // stage s is free again: gather chunk c+2 into it
            if (c + 2 < nchunks) stream_issue(p, ws, s, lane, idn, vn, q0n, q1n, g, start, c + 2);
        }
    }

    if (g.inside) {
        const int HW = p.H * p.W;
        out_image[g.pix] = Cr + T * p.bg[0];
        out_image[HW + g.pix] = Cg + T * p.bg[1];
        out_image[2 * HW + g.pix] = Cb + T * p.bg[2];
        if (DUAL) {
            p.out_image_nosh[g.pix] = Er + T * p.bg[0];
            p.out_image_nosh[HW + g.pix] = Eg + T * p.bg[1];
            p.out_image_nosh[2 * HW + g.pix] = Eb + T * p.bg[2];
        }
        out_depth[g.pix] = D;
        const float3 nw = rot_v2w(p.view, f3(Nx, Ny, Nz));
        out_norm[g.pix] = nw.x;
        out_norm[HW + g.pix] = nw.y;
        out_norm[2 * HW + g.pix] = nw.z;
        out_alpha[g.pix] = A;
        p.final_T[g.pix] = T;
        p.n_contrib[g.pix] = last;
    }
    if (p.flags & TEXGS_FLAG_DEBUG) {
        unsigned tot = nblend;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0 && tot) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&p.counters->num_blend_lo), (unsigned long long)tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct BwdIn {
    const float *dL_dimage, *dL_ddepth, *dL_dnorm, *dL_dalpha, *dL_dimage_nosh;
};

// ALT = true: cold instantiation for the spec switches of p.flags (E11-alt seamless taps, E7-alt depth of the
// intersection, E13-alt no gradient through Delta), as in the forward.
template <int MODE, bool TEX4, bool GRAD4, bool DUAL, bool ALT>
__global__ void __launch_bounds__(32 * TEXGS_BWD_WARPS, TEXGS_BWD_MIN_CTAS) texgs_render_bwd(const RasterParams p, const BwdIn in, float* __restrict__ acc,
                                                      float* __restrict__ dtex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.counters->overflow) return;
    const PixelGeom g = pixel_geom<TEXGS_BWD_WARPS>(p);
    const int lane = threadIdx.x & 31;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const unsigned start = p.tile_offset[g.tile];
    const unsigned n = p.tile_offset[g.tile + 1] - start;
    if (n == 0) return;

    const unsigned last = g.inside ? p.n_contrib[g.pix] : 0u;
    unsigned max_last = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_last = max(max_last, __shfl_xor_sync(0xffffffffu, max_last, o));
    if (max_last == 0) return;                      // uniform per warp
    // chunks are visited back to front: visit k handles chunk c = c_top - k in stage k & 1
    const int c_top = (int)((max_last - 1) / TEXGS_CHUNK);

    const int HW = p.H * p.W;
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
    float3 gnv = f3(0.f, 0.f, 0.f);
    float T_final = 1.f;
    if (g.inside) {
        if (in.dL_dimage) { gr = in.dL_dimage[g.pix]; gg = in.dL_dimage[HW + g.pix]; gb = in.dL_dimage[2 * HW + g.pix]; }
        if (in.dL_ddepth) gd = in.dL_ddepth[g.pix];
        if (in.dL_dalpha) ga = in.dL_dalpha[g.pix];
        if (in.dL_dnorm) gnv = rot_w2v(p.view, f3(in.dL_dnorm[g.pix], in.dL_dnorm[HW + g.pix], in.dL_dnorm[2 * HW + g.pix]));
        T_final = p.final_T[g.pix];
    }
    float hr = 0.f, hg = 0.f, hb = 0.f;       // dual render: cotangent of the no-SH image
    if (DUAL && g.inside && in.dL_dimage_nosh) {
        hr = in.dL_dimage_nosh[g.pix]; hg = in.dL_dimage_nosh[HW + g.pix]; hb = in.dL_dimage_nosh[2 * HW + g.pix];
    }
    const float bgdot = (gr + hr) * p.bg[0] + (gg + hg) * p.bg[1] + (gb + hb) * p.bg[2];
    const float pxf = (float)g.px, pyf = (float)g.py;
    const float* __restrict__ tex = p.texture;
    const int R = p.R;
    const float halfR = 0.5f * (float)R;
    const bool seamless = ALT && (p.flags & TEXGS_FLAG_SEAMLESS_CUBE) != 0u;
    const bool zalt = ALT && (p.flags & TEXGS_FLAG_DEPTH_INTERSECTION) != 0u;
    const bool stopgrad = ALT && (p.flags & TEXGS_FLAG_STOPGRAD_DELTA) != 0u;

    float T = T_final, acc_rec = 0.f, last_alpha = 0.f, last_X = 0.f;

    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    // prologue: visits 0 and 1 gathered, id + masks of visit 2 loaded. Which entries of a chunk reach the warp's blocks was
    // decided (exactly) by the forward and is read back — two words per chunk instead of 32 record sectors and 64 ellipse
    // tests. Entries at or beyond max_last can not contribute to this warp and are not gathered.
    bool v0, v1, v2;
    const unsigned id0 = stream_load_id(p, start, max_last, c_top, lane, v0);
    const unsigned id1 = stream_load_id(p, start, max_last, c_top - 1, lane, v1);
    unsigned id2 = stream_load_id(p, start, max_last, c_top - 2, lane, v2);
    stream_issue_saved(p, ws, 0, lane, id0, __ldg(p.cull_masks + cull_mask_slot(start, g, c_top)), c_top, max_last);
    if (c_top >= 1) stream_issue_saved(p, ws, 1, lane, id1, __ldg(p.cull_masks + cull_mask_slot(start, g, c_top - 1)), c_top - 1, max_last);
    uint2 m2 = (c_top >= 2) ? __ldg(p.cull_masks + cull_mask_slot(start, g, c_top - 2)) : make_uint2(0u, 0u);
    for (int k = 0; k <= c_top; ++k) {
        const int c = c_top - k;
        const int s = k & 1;
        const unsigned idn = id2;
        const uint2 mn = m2;
        id2 = stream_load_id(p, start, max_last, c - 3, lane, v2);
        if (c >= 3) m2 = __ldg(p.cull_masks + cull_mask_slot(start, g, c - 3));

        mbar_wait(&ws.bar[s], (unsigned)(k >> 1) & 1u);
        __syncwarp();
        const unsigned mL = ws.maskL[s], mR = ws.maskR[s], mU = mL | mR;
        unsigned my = (lane & 16) ? mR : mL;              // this half-warp's survivors, walked back to front
        while (__any_sync(0xffffffffu, my != 0u)) {
            const bool has = my != 0u;
            const int l = has ? (31 - __clz(my)) : 0;
            my &= ~(1u << l);
            const GaussRec& rec = ws.rec[s][__popc(mU & ((1u << l) - 1u))];       // half-uniform address
            const unsigned gi = (unsigned)c * TEXGS_CHUNK + (unsigned)l;   // 0-based position in the list
            const float4 g0 = rec.q[0], g1 = rec.q[1];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            const float G = texgs_exp(power);
            const float aG = g1.y * G;
            const float alpha = fminf(TEXGS_ALPHA_MAX, aG);
            const bool contrib = has && (gi < last) && (power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN);
            if (!__any_sync(0xffffffffu, contrib)) continue;

            float v[20];
#pragma unroll
            for (int q = 0; q < 20; ++q) v[q] = 0.f;
            if (contrib) {
                const float4 g2 = rec.q[2], g3 = rec.q[3];
                const float inv_1ma = __fdividef(1.0f, 1.0f - alpha);
                T = T * inv_1ma;
                const float w = alpha * T;
                float cr = g3.y, cg = g3.z, cb = g3.w;
                float mr = 1.f, mg = 1.f, mb = 1.f;
                float xdual = 0.f, kr = 0.f, kg = 0.f, kb = 0.f;
                UvEval e;
                CubeCoord cc;
                Bilerp bl;
                float t00[3], t01[3], t10[3], t11[3];
                float4 g4, g5, g6;
                if (MODE == TEXGS_MODE_TEXTURE && !TEXGS_ABLATE_HEAVY) {
                    g4 = rec.q[4]; g5 = rec.q[5]; g6 = rec.q[6];
                    e = eval_uv(g1, g2, g3, g4, g5, g6, g.vx, g.vy);
                    cc = cube_coord(e.ux, e.uy, e.uz);
                    bl = (ALT && seamless) ? cube_bilerp_seamless(cc, R) : cube_bilerp(cc, R);
                    float tx3[3];
                    fetch_taps<TEX4>(tex, p.texture_rgba, bl, t00, t01, t10, t11);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float top = t00[ch] + bl.wx * (t01[ch] - t00[ch]);
                        const float bot = t10[ch] + bl.wx * (t11[ch] - t10[ch]);
                        tx3[ch] = top + bl.wy * (bot - top);
                    }
                    cr = SH_C0 * tx3[0] + cr; cg = SH_C0 * tx3[1] + cg; cb = SH_C0 * tx3[2] + cb;
                    if (cr < 0.f) { cr = 0.f; mr = 0.f; }
                    if (cg < 0.f) { cg = 0.f; mg = 0.f; }
                    if (cb < 0.f) { cb = 0.f; mb = 0.f; }
                    if (DUAL) {   // colour of the no-SH image and its clamp mask, folded into X / the texel gradient
                        const float er = SH_C0 * tx3[0] + 0.5f, eg = SH_C0 * tx3[1] + 0.5f, eb = SH_C0 * tx3[2] + 0.5f;
                        xdual = hr * fmaxf(er, 0.f) + hg * fmaxf(eg, 0.f) + hb * fmaxf(eb, 0.f);
                        kr = (er < 0.f) ? 0.f : hr; kg = (eg < 0.f) ? 0.f : hg; kb = (eb < 0.f) ? 0.f : hb;
                    }
                }
                // E7-alt: the depth term is z of the intersection = t; its gradient then goes to t (below) instead of the centre
                const bool zint = ALT && zalt && MODE == TEXGS_MODE_TEXTURE && e.safe;
                const float X = gr * cr + gg * cg + gb * cb + gd * (zint ? e.t : g1.z) + gnv.x * g2.x + gnv.y * g2.y + gnv.z * g2.z + ga + xdual;
                acc_rec = last_alpha * last_X + (1.0f - last_alpha) * acc_rec;
                const float dL_dalpha = (X - acc_rec) * T - (T_final * inv_1ma) * bgdot;
                last_alpha = alpha;
                last_X = X;
                // alpha = min(0.99, o*G): derivative of the clamp is zero when it is active
                const float live = (aG <= TEXGS_ALPHA_MAX) ? 1.f : 0.f;
                const float dL_dG = live * g1.y * dL_dalpha;
                v[5] = live * G * dL_dalpha;
                const float GdG = G * dL_dG;
                v[0] = -GdG * (g0.z * dx + g0.w * dy);
                v[1] = -GdG * (g1.x * dy + g0.w * dx);
                v[2] = -0.5f * GdG * dx * dx;
                v[3] = -GdG * dx * dy;
                v[4] = -0.5f * GdG * dy * dy;
                const float wr = w * gr * mr, wg = w * gg * mg, wb = w * gb * mb;   // dL/d col (masked)
                v[6] = wr; v[7] = wg; v[8] = wb;
                v[9] = (zint && !stopgrad) ? 0.f : w * gd;
                v[10] = w * gnv.x; v[11] = w * gnv.y; v[12] = w * gnv.z;
                if (MODE == TEXGS_MODE_TEXTURE && !TEXGS_ABLATE_HEAVY) {
                    const float gt[3] = {SH_C0 * (wr + w * kr), SH_C0 * (wg + w * kg), SH_C0 * (wb + w * kb)};
                    float dwx = 0.f, dwy = 0.f;
                    const float w00 = (1.f - bl.wx) * (1.f - bl.wy), w01 = bl.wx * (1.f - bl.wy);
                    const float w10 = (1.f - bl.wx) * bl.wy, w11 = bl.wx * bl.wy;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float dtop = t01[ch] - t00[ch], dbot = t11[ch] - t10[ch];
                        const float top = t00[ch] + bl.wx * dtop;
                        const float bot = t10[ch] + bl.wx * dbot;
                        dwx += gt[ch] * (dtop + bl.wy * (dbot - dtop));
                        dwy += gt[ch] * (bot - top);
                        if (!GRAD4 && dtex && gt[ch] != 0.f) {
                            atomicAdd(dtex + 3 * bl.i00 + ch, gt[ch] * w00);
                            atomicAdd(dtex + 3 * bl.i01 + ch, gt[ch] * w01);
                            atomicAdd(dtex + 3 * bl.i10 + ch, gt[ch] * w10);
                            atomicAdd(dtex + 3 * bl.i11 + ch, gt[ch] * w11);
                        }
                    }
                    if (GRAD4 && dtex && !TEXGS_ABLATE_RED && (gt[0] != 0.f || gt[1] != 0.f || gt[2] != 0.f)) {
                        red_add_v4(dtex + 4 * bl.i00, gt[0] * w00, gt[1] * w00, gt[2] * w00, 0.f);
                        red_add_v4(dtex + 4 * bl.i01, gt[0] * w01, gt[1] * w01, gt[2] * w01, 0.f);
                        red_add_v4(dtex + 4 * bl.i10, gt[0] * w10, gt[1] * w10, gt[2] * w10, 0.f);
                        red_add_v4(dtex + 4 * bl.i11, gt[0] * w11, gt[1] * w11, gt[2] * w11, 0.f);
                    }
                    float gu[3];
                    cube_coord_bwd(cc, dwx * halfR * cc.inv_m, dwy * halfR * cc.inv_m, gu);
                    v[13] = gu[0]; v[14] = gu[1]; v[15] = gu[2];
                    if (e.safe && !(ALT && stopgrad)) {
                        // g_v = J'^T gu ; s = dL/dt / nd,  dL/dt = g_v . v  (+ w gd when the depth output is t: E7-alt)
                        const float gvx = g4.w * gu[0] + g5.z * gu[1] + g6.y * gu[2];
                        const float gvy = g5.x * gu[0] + g5.w * gu[1] + g6.z * gu[2];
                        const float gvz = g5.y * gu[0] + g6.x * gu[1] + g6.w * gu[2];
                        const float sden = __fdividef(gvx * g.vx + gvy * g.vy + gvz + (zint ? w * gd : 0.f), e.nd);
                        v[16] = sden;
                        v[17] = sden * e.dx;
                        v[18] = sden * e.dy;
                        v[19] = sden * e.dz;
                    }
                }
            }
            // each half-warp reduces ITS splat's 20 partials over its 16 lanes (both halves in the same
            // 20 shuffles) and adds them to that Gaussian's accumulators
            float outA, outB;
#if TEXGS_ABLATE_REDUCE
            outA = 0.f;
#pragma unroll
            for (int q = 0; q < 20; ++q) outA += v[q];
            outB = outA;
#else
            halfwarp_reduce20(v, lane, outA, outB);
#endif
            if (has) {
                float* dst = acc + (size_t)(unsigned)__float_as_int(rec.q[7].x) * TEXGS_BWD_ACC_FLOATS;
                if (outA != 0.f) atomicAdd(dst + (lane & 15), outA);
                if ((lane & 3) == 0 && outB != 0.f) atomicAdd(dst + 16 + ((lane >> 2) & 3), outB);
            }
        }
        __syncwarp();
        if (k + 2 <= c_top) stream_issue_saved(p, ws, s, lane, idn, mn, c - 2, max_last);
    }
}
