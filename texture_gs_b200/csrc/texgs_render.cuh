// texgs_render.cuh — per-tile blend kernels (SURVEY §8a rows a7 forward, a8 backward).
//
// One CTA = one 16x16 tile, 256 threads, warp w owns an 8x4 pixel block (better texel / alpha-test
// coherence than 16x2 rows). The tile's depth-sorted Gaussian list is streamed through shared
// memory in batches of TEXGS_BATCH records: each thread issues ONE 128-byte cp.async.bulk (TMA 1-D)
// per record, completion is tracked by an mbarrier per stage, two stages are in flight so the gather
// of batch b+1 overlaps the blend of batch b. Spec items E5-E12.
#pragma once
#include "texgs_common.cuh"

#define TEXGS_BATCH 128

#ifndef TEXGS_FAST_EXP
#define TEXGS_FAST_EXP 1
#endif
__device__ __forceinline__ float texgs_exp(float x) {
#if TEXGS_FAST_EXP
    return __expf(x);
#else
    return expf(x);
#endif
}

struct PixelGeom {
    int tile, px, py, pix;
    bool inside;
    float vx, vy;    // view ray (vx, vy, 1)
};

__device__ __forceinline__ PixelGeom pixel_geom(const RasterParams& p) {
    PixelGeom g;
    g.tile = blockIdx.x;
    const int tx = g.tile % p.grid_x, ty = g.tile / p.grid_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    g.px = tx * TEXGS_TILE + lx;
    g.py = ty * TEXGS_TILE + ly;
    g.inside = (g.px < p.W) && (g.py < p.H);
    g.pix = g.py * p.W + g.px;
    g.vx = ((2.0f * (float)g.px + 1.0f) / (float)p.W - 1.0f) * p.tanfovx;
    g.vy = ((2.0f * (float)g.py + 1.0f) / (float)p.H - 1.0f) * p.tanfovy;
    return g;
}

// issue the gather of batch ``b`` of this tile's list into stage b&1
__device__ __forceinline__ void issue_batch(const RasterParams& p, GaussRec (*s_rec)[TEXGS_BATCH], uint64_t* s_bar,
                                            unsigned start, unsigned n, int b) {
    const int s = b & 1;
    const unsigned cnt = min((unsigned)TEXGS_BATCH, n - (unsigned)b * TEXGS_BATCH);
    const unsigned tid = threadIdx.x;
    if (tid == 0) mbar_arrive_expect_tx(&s_bar[s], cnt * (unsigned)sizeof(GaussRec));
    if (tid < cnt) {
        const unsigned id = p.sorted_ids[start + (unsigned)b * TEXGS_BATCH + tid];
        bulk_g2s(&s_rec[s][tid], p.recs + id, (unsigned)sizeof(GaussRec), &s_bar[s]);
    }
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
        e.t = g1.w / e.nd;
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
template <int MODE>
__global__ void __launch_bounds__(256) texgs_render_fwd(const RasterParams p, float* __restrict__ out_image,
                                                      float* __restrict__ out_depth, float* __restrict__ out_norm,
                                                      float* __restrict__ out_alpha) {
    __shared__ GaussRec s_rec[2][TEXGS_BATCH];
    __shared__ __align__(8) uint64_t s_bar[2];
    if (p.counters->overflow) return;
    const PixelGeom g = pixel_geom(p);
    const unsigned start = p.tile_offset[g.tile];
    const unsigned n = p.tile_offset[g.tile + 1] - start;
    const int nb = (int)((n + TEXGS_BATCH - 1) / TEXGS_BATCH);

    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (nb > 0) issue_batch(p, s_rec, s_bar, start, n, 0);

    float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, D = 0.f, Nx = 0.f, Ny = 0.f, Nz = 0.f, A = 0.f;
    unsigned last = 0, nblend = 0;
    bool done = !g.inside;
    bool warp_done = false;          // uniform across the warp
    const float pxf = (float)g.px, pyf = (float)g.py;
    const float* __restrict__ tex = p.texture;
    const int R = p.R;

    int pending = -1;
    for (int b = 0; b < nb; ++b) {
        const int s = b & 1;
        const int cnt = (int)min((unsigned)TEXGS_BATCH, n - (unsigned)b * TEXGS_BATCH);
        mbar_wait(&s_bar[s], (unsigned)(b >> 1) & 1u);
        if (b + 1 < nb) issue_batch(p, s_rec, s_bar, start, n, b + 1);
        if (!warp_done) {
            // The loop is warp-synchronous: every lane walks the same j, per-lane state is a
            // predicate. (A per-lane continue/break loop never reconverges on sm_70+ and runs the
            // lanes one after the other.)
            for (int j = 0; j < cnt; ++j) {
                const GaussRec& rec = s_rec[s][j];
                const float4 g0 = rec.q[0], g1 = rec.q[1];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const float alpha = fminf(TEXGS_ALPHA_MAX, g1.y * texgs_exp(power));
                bool cand = !done && (power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN);
                if (!__any_sync(0xffffffffu, cand)) continue;
                const float test_T = T * (1.0f - alpha);
                if (cand && test_T < TEXGS_T_STOP) { done = true; cand = false; }
                if (cand) {
                    const float4 g2 = rec.q[2], g3 = rec.q[3];
                    float cr = g3.y, cg = g3.z, cb = g3.w;
                    if (MODE == TEXGS_MODE_TEXTURE) {
                        const float4 g4 = rec.q[4], g5 = rec.q[5], g6 = rec.q[6];
                        const UvEval e = eval_uv(g1, g2, g3, g4, g5, g6, g.vx, g.vy);
                        const CubeCoord cc = cube_coord(e.ux, e.uy, e.uz);
                        const Bilerp bl = cube_bilerp(cc, R);
                        float tx3[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float t00 = __ldg(tex + bl.i00 + ch), t01 = __ldg(tex + bl.i01 + ch);
                            const float t10 = __ldg(tex + bl.i10 + ch), t11 = __ldg(tex + bl.i11 + ch);
                            const float top = t00 + bl.wx * (t01 - t00);
                            const float bot = t10 + bl.wx * (t11 - t10);
                            tx3[ch] = top + bl.wy * (bot - top);
                        }
                        cr = fmaxf(0.f, SH_C0 * tx3[0] + cr);
                        cg = fmaxf(0.f, SH_C0 * tx3[1] + cg);
                        cb = fmaxf(0.f, SH_C0 * tx3[2] + cb);
                    }
                    const float w = alpha * T;
                    Cr += w * cr; Cg += w * cg; Cb += w * cb;
                    D += w * g1.z;
                    Nx += w * g2.x; Ny += w * g2.y; Nz += w * g2.z;
                    A += w;
                    T = test_T;
                    last = (unsigned)b * TEXGS_BATCH + (unsigned)j + 1u;
                    ++nblend;
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
        }
        if (__syncthreads_and(done ? 1 : 0)) {
            if (b + 1 < nb) pending = b + 1;
            break;
        }
    }
    if (pending >= 0) mbar_wait(&s_bar[pending & 1], (unsigned)(pending >> 1) & 1u);   // never exit with a copy in flight

    if (g.inside) {
        const int HW = p.H * p.W;
        out_image[g.pix] = Cr + T * p.bg[0];
        out_image[HW + g.pix] = Cg + T * p.bg[1];
        out_image[2 * HW + g.pix] = Cb + T * p.bg[2];
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
        if ((threadIdx.x & 31) == 0 && tot) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&p.counters->num_blend_lo), (unsigned long long)tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct BwdIn {
    const float *dL_dimage, *dL_ddepth, *dL_dnorm, *dL_dalpha;
};

template <int MODE>
__global__ void __launch_bounds__(256) texgs_render_bwd(const RasterParams p, const BwdIn in, float* __restrict__ acc,
                                                      float* __restrict__ dtex) {
    __shared__ GaussRec s_rec[2][TEXGS_BATCH];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ unsigned s_max;
    if (p.counters->overflow) return;
    const PixelGeom g = pixel_geom(p);
    const unsigned start = p.tile_offset[g.tile];
    const unsigned n = p.tile_offset[g.tile + 1] - start;
    if (n == 0) return;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); s_max = 0; }
    __syncthreads();
    const unsigned last = g.inside ? p.n_contrib[g.pix] : 0u;
    {
        unsigned m = last;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m) atomicMax(&s_max, m);
    }
    __syncthreads();
    const unsigned max_last = s_max;
    if (max_last == 0) return;
    const int nb = (int)((max_last + TEXGS_BATCH - 1) / TEXGS_BATCH);   // batches that hold contributors
    // batches are visited back to front: visit k = 0.. nb-1 handles batch b = nb-1-k, stage k&1

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
    const float bgdot = gr * p.bg[0] + gg * p.bg[1] + gb * p.bg[2];
    const float pxf = (float)g.px, pyf = (float)g.py;
    const float* __restrict__ tex = p.texture;
    const int R = p.R;
    const float halfR = 0.5f * (float)R;

    float T = T_final, acc_rec = 0.f, last_alpha = 0.f, last_X = 0.f;

    // visit 0 -> batch nb-1
    {
        const int b = nb - 1;
        const unsigned cnt = min((unsigned)TEXGS_BATCH, n - (unsigned)b * TEXGS_BATCH);
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar[0], cnt * (unsigned)sizeof(GaussRec));
        if (threadIdx.x < cnt) {
            const unsigned id = p.sorted_ids[start + (unsigned)b * TEXGS_BATCH + threadIdx.x];
            bulk_g2s(&s_rec[0][threadIdx.x], p.recs + id, (unsigned)sizeof(GaussRec), &s_bar[0]);
        }
    }
    for (int k = 0; k < nb; ++k) {
        const int b = nb - 1 - k;
        const int s = k & 1;
        const int cnt = (int)min((unsigned)TEXGS_BATCH, n - (unsigned)b * TEXGS_BATCH);
        mbar_wait(&s_bar[s], (unsigned)(k >> 1) & 1u);
        if (k + 1 < nb) {   // prefetch batch b-1 into the other stage (always a full batch)
            const int b2 = b - 1, s2 = s ^ 1;
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar[s2], TEXGS_BATCH * (unsigned)sizeof(GaussRec));
            if (threadIdx.x < TEXGS_BATCH) {
                const unsigned id = p.sorted_ids[start + (unsigned)b2 * TEXGS_BATCH + threadIdx.x];
                bulk_g2s(&s_rec[s2][threadIdx.x], p.recs + id, (unsigned)sizeof(GaussRec), &s_bar[s2]);
            }
        }
        for (int j = cnt - 1; j >= 0; --j) {
            const unsigned gi = (unsigned)b * TEXGS_BATCH + (unsigned)j;   // 0-based position in the list
            const GaussRec& rec = s_rec[s][j];
            const float4 g0 = rec.q[0], g1 = rec.q[1];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            const float G = texgs_exp(power);
            const float aG = g1.y * G;
            const float alpha = fminf(TEXGS_ALPHA_MAX, aG);
            const bool contrib = (gi < last) && (power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN);
            if (!__any_sync(0xffffffffu, contrib)) continue;

            float v[20];
#pragma unroll
            for (int q = 0; q < 20; ++q) v[q] = 0.f;
            if (contrib) {
                const float4 g2 = rec.q[2], g3 = rec.q[3];
                T = T / (1.0f - alpha);
                const float w = alpha * T;
                float cr = g3.y, cg = g3.z, cb = g3.w;
                float mr = 1.f, mg = 1.f, mb = 1.f;
                UvEval e;
                CubeCoord cc;
                Bilerp bl;
                float t00[3], t01[3], t10[3], t11[3];
                float4 g4, g5, g6;
                if (MODE == TEXGS_MODE_TEXTURE) {
                    g4 = rec.q[4]; g5 = rec.q[5]; g6 = rec.q[6];
                    e = eval_uv(g1, g2, g3, g4, g5, g6, g.vx, g.vy);
                    cc = cube_coord(e.ux, e.uy, e.uz);
                    bl = cube_bilerp(cc, R);
                    float tx3[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        t00[ch] = __ldg(tex + bl.i00 + ch); t01[ch] = __ldg(tex + bl.i01 + ch);
                        t10[ch] = __ldg(tex + bl.i10 + ch); t11[ch] = __ldg(tex + bl.i11 + ch);
                        const float top = t00[ch] + bl.wx * (t01[ch] - t00[ch]);
                        const float bot = t10[ch] + bl.wx * (t11[ch] - t10[ch]);
                        tx3[ch] = top + bl.wy * (bot - top);
                    }
                    cr = SH_C0 * tx3[0] + cr; cg = SH_C0 * tx3[1] + cg; cb = SH_C0 * tx3[2] + cb;
                    if (cr < 0.f) { cr = 0.f; mr = 0.f; }
                    if (cg < 0.f) { cg = 0.f; mg = 0.f; }
                    if (cb < 0.f) { cb = 0.f; mb = 0.f; }
                }
                const float X = gr * cr + gg * cg + gb * cb + gd * g1.z + gnv.x * g2.x + gnv.y * g2.y + gnv.z * g2.z + ga;
                acc_rec = last_alpha * last_X + (1.0f - last_alpha) * acc_rec;
                const float dL_dalpha = (X - acc_rec) * T - (T_final / (1.0f - alpha)) * bgdot;
                last_alpha = alpha;
                last_X = X;
                // alpha = min(0.99, o*G): derivative of the clamp is zero when it is active
                const float live = (aG <= TEXGS_ALPHA_MAX) ? 1.f : 0.f;
                const float dL_dG = live * g1.y * dL_dalpha;
                v[5] = live * G * dL_dalpha;
                const float gdx = -G * (g0.z * dx + g0.w * dy);   // dG/d(dx)
                const float gdy = -G * (g1.x * dy + g0.w * dx);
                v[0] = dL_dG * gdx;
                v[1] = dL_dG * gdy;
                v[2] = -0.5f * G * dx * dx * dL_dG;
                v[3] = -G * dx * dy * dL_dG;
                v[4] = -0.5f * G * dy * dy * dL_dG;
                const float wr = w * gr * mr, wg = w * gg * mg, wb = w * gb * mb;   // dL/d col (masked)
                v[6] = wr; v[7] = wg; v[8] = wb;
                v[9] = w * gd;
                v[10] = w * gnv.x; v[11] = w * gnv.y; v[12] = w * gnv.z;
                if (MODE == TEXGS_MODE_TEXTURE) {
                    const float gt[3] = {SH_C0 * wr, SH_C0 * wg, SH_C0 * wb};
                    float dwx = 0.f, dwy = 0.f;
                    const float w00 = (1.f - bl.wx) * (1.f - bl.wy), w01 = bl.wx * (1.f - bl.wy);
                    const float w10 = (1.f - bl.wx) * bl.wy, w11 = bl.wx * bl.wy;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float top = t00[ch] + bl.wx * (t01[ch] - t00[ch]);
                        const float bot = t10[ch] + bl.wx * (t11[ch] - t10[ch]);
                        dwx += gt[ch] * ((1.f - bl.wy) * (t01[ch] - t00[ch]) + bl.wy * (t11[ch] - t10[ch]));
                        dwy += gt[ch] * (bot - top);
                        if (dtex && gt[ch] != 0.f) {
                            atomicAdd(dtex + bl.i00 + ch, gt[ch] * w00);
                            atomicAdd(dtex + bl.i01 + ch, gt[ch] * w01);
                            atomicAdd(dtex + bl.i10 + ch, gt[ch] * w10);
                            atomicAdd(dtex + bl.i11 + ch, gt[ch] * w11);
                        }
                    }
                    const float dsx = dwx * halfR, dsy = dwy * halfR;
                    float gu[3] = {0.f, 0.f, 0.f};
                    const float ax = cc.sgx * dsx * cc.inv_m, ay = cc.sgy * dsy * cc.inv_m;
                    const float am = cc.sgm * (-(cc.sx * dsx + cc.sy * dsy) * cc.inv_m);
#pragma unroll
                    for (int q = 0; q < 3; ++q) gu[q] = (q == cc.ix ? ax : 0.f) + (q == cc.iy ? ay : 0.f) + (q == cc.axis ? am : 0.f);
                    v[13] = gu[0]; v[14] = gu[1]; v[15] = gu[2];
                    if (e.safe) {
                        // g_v = J'^T gu ; s = (g_v . v) / nd
                        const float gvx = g4.w * gu[0] + g5.z * gu[1] + g6.y * gu[2];
                        const float gvy = g5.x * gu[0] + g5.w * gu[1] + g6.z * gu[2];
                        const float gvz = g5.y * gu[0] + g6.x * gu[1] + g6.w * gu[2];
                        const float sden = (gvx * g.vx + gvy * g.vy + gvz) / e.nd;
                        v[16] = sden;
                        v[17] = sden * e.dx;
                        v[18] = sden * e.dy;
                        v[19] = sden * e.dz;
                    }
                }
            }
            float outA, outB;
            warp_reduce20(v, lane, outA, outB);
            {
                float* dst = acc + (size_t)(unsigned)__float_as_int(rec.q[7].x) * TEXGS_BWD_ACC_FLOATS;
                if ((lane & 1) == 0 && outA != 0.f) atomicAdd(dst + (lane >> 1), outA);
                if ((lane & 7) == 1 && outB != 0.f) atomicAdd(dst + 16 + (lane >> 3), outB);
            }
        }
        __syncthreads();   // everybody done reading stage s before it is refilled two visits later
    }
}
