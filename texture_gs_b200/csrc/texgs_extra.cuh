// texgs_extra.cuh — blending of per-Gaussian extra attribute channels: the ``extra_attrs`` kwarg of the
// operator (reference render/uv_tex_render.py:7,66 and render/render.py:8,84) and its ``extra`` output.
//
//   extra[ch](pixel) = sum_i  w_i(pixel) * extra_attrs[i][ch]          (no background term)
//
// with exactly the blend weights w_i = alpha_i T_i of the main render (spec E5). The reference tree
// always passes ``extra_attrs=None``, so this is a COLD path and it is deliberately kept out of the
// two hot render kernels (their register budget decides their occupancy): the forward re-walks each
// tile's sorted list up to the pixel's saved ``n_contrib``; the backward re-walks it back to front
// and ADDS its share of dL/d(2-D mean, conic, opacity) to the per-Gaussian accumulators the main
// render backward has already filled — the blend is linear in the blended values, so the alpha-chain
// gradient of the extra channels is an independent additive term.
//
// One CTA = one 16x16 tile, one thread = one pixel, list staged through shared memory 256 entries at
// a time (first 32-byte sector of the record only: centre, conic, opacity), channels processed
// TEXGS_EXTRA_CH at a time (one list walk per channel group).
#pragma once
#include "texgs_common.cuh"
#include "texgs_render.cuh"

#define TEXGS_EXTRA_CH 8
#define TEXGS_EXTRA_BATCH 256

struct ExtraStage {
    float4 q0[TEXGS_EXTRA_BATCH];
    float4 q1[TEXGS_EXTRA_BATCH];
    unsigned id[TEXGS_EXTRA_BATCH];
    unsigned max_last;
};

struct ExtraPixel {
    int px, py, pix;
    bool inside;
    unsigned start, last, max_last;
};

// pixel of this thread + the list bounds of its tile; ``max_last`` is uniform over the CTA
__device__ __forceinline__ ExtraPixel extra_pixel(const RasterParams& p, ExtraStage& st) {
    ExtraPixel e;
    const int tile = blockIdx.x;
    e.px = (tile % p.grid_x) * TEXGS_TILE + (int)(threadIdx.x & 15);
    e.py = (tile / p.grid_x) * TEXGS_TILE + (int)(threadIdx.x >> 4);
    e.inside = (e.px < p.W) && (e.py < p.H);
    e.pix = e.py * p.W + e.px;
    e.start = p.tile_offset[tile];
    e.last = e.inside ? p.n_contrib[e.pix] : 0u;
    if (threadIdx.x == 0) st.max_last = 0u;
    __syncthreads();
    if (e.last) atomicMax(&st.max_last, e.last);
    __syncthreads();
    e.max_last = st.max_last;
    return e;
}

// stage list entries [base, base + 256) (clipped at max_last) into shared memory
__device__ __forceinline__ void extra_stage_load(const RasterParams& p, ExtraStage& st, const ExtraPixel& e, unsigned base) {
    __syncthreads();                                   // the previous batch is consumed
    const unsigned pos = base + threadIdx.x;
    if (pos < e.max_last) {
        const unsigned id = __ldg(p.sorted_ids + e.start + pos);
        const float4* r = reinterpret_cast<const float4*>(p.recs + id);
        st.id[threadIdx.x] = id;
        st.q0[threadIdx.x] = __ldg(r);
        st.q1[threadIdx.x] = __ldg(r + 1);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) texgs_extra_fwd(const RasterParams p, float* __restrict__ out_extra) {
    __shared__ ExtraStage st;
    if (p.counters->overflow) return;
    const ExtraPixel e = extra_pixel(p, st);
    const float pxf = (float)e.px, pyf = (float)e.py;
    const size_t HW = (size_t)p.H * p.W;
    for (int c0 = 0; c0 < p.E; c0 += TEXGS_EXTRA_CH) {
        float out[TEXGS_EXTRA_CH];
#pragma unroll
        for (int k = 0; k < TEXGS_EXTRA_CH; ++k) out[k] = 0.f;
        float T = 1.0f;
        for (unsigned base = 0; base < e.max_last; base += TEXGS_EXTRA_BATCH) {
            extra_stage_load(p, st, e, base);
            const int cnt = (int)min((unsigned)TEXGS_EXTRA_BATCH, e.max_last - base);
            for (int j = 0; j < cnt; ++j) {
                if (base + (unsigned)j >= e.last) break;      // nothing at or beyond n_contrib blends into this pixel
                const float4 g0 = st.q0[j], g1 = st.q1[j];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const float alpha = fminf(TEXGS_ALPHA_MAX, g1.y * texgs_exp(power));
                if (!((power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN))) continue;
                const float w = alpha * T;
                const float* __restrict__ ea = p.extra_attrs + (size_t)st.id[j] * p.E + c0;
#pragma unroll
                for (int k = 0; k < TEXGS_EXTRA_CH; ++k)
                    if (c0 + k < p.E) out[k] += w * __ldg(ea + k);
                T = T * (1.0f - alpha);
            }
        }
        if (e.inside) {
#pragma unroll
            for (int k = 0; k < TEXGS_EXTRA_CH; ++k)
                if (c0 + k < p.E) out_extra[(size_t)(c0 + k) * HW + e.pix] = out[k];
        }
    }
}

// acc: the per-Gaussian accumulators of texgs_render_bwd (layout in texgs_preprocess.cuh); this kernel adds to
// slots 0..5 only. dextra_attrs (P,E) is pre-zeroed by the caller, or NULL when that gradient is not wanted.
__global__ void __launch_bounds__(256) texgs_extra_bwd(const RasterParams p, const float* __restrict__ dL_dextra,
                                                       float* __restrict__ acc, float* __restrict__ dextra_attrs) {
    __shared__ ExtraStage st;
    if (p.counters->overflow) return;
    const ExtraPixel e = extra_pixel(p, st);
    if (e.max_last == 0u) return;                                 // uniform over the CTA
    const int lane = threadIdx.x & 31;
    const float pxf = (float)e.px, pyf = (float)e.py;
    const size_t HW = (size_t)p.H * p.W;
    const float T_final = e.inside ? p.final_T[e.pix] : 1.0f;
    const int nbatch = (int)((e.max_last + TEXGS_EXTRA_BATCH - 1) / TEXGS_EXTRA_BATCH);
    for (int c0 = 0; c0 < p.E; c0 += TEXGS_EXTRA_CH) {
        float g[TEXGS_EXTRA_CH];
#pragma unroll
        for (int k = 0; k < TEXGS_EXTRA_CH; ++k)
            g[k] = (e.inside && c0 + k < p.E) ? dL_dextra[(size_t)(c0 + k) * HW + e.pix] : 0.f;
        float T = T_final, acc_rec = 0.f, last_alpha = 0.f, last_X = 0.f;
        for (int b = nbatch - 1; b >= 0; --b) {
            const unsigned base = (unsigned)b * TEXGS_EXTRA_BATCH;
            extra_stage_load(p, st, e, base);
            const int cnt = (int)min((unsigned)TEXGS_EXTRA_BATCH, e.max_last - base);
            for (int j = cnt - 1; j >= 0; --j) {
                const float4 g0 = st.q0[j], g1 = st.q1[j];
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const float G = texgs_exp(power);
                const float aG = g1.y * G;
                const float alpha = fminf(TEXGS_ALPHA_MAX, aG);
                const bool contrib = (base + (unsigned)j < e.last) && (power <= 0.0f) && (alpha >= TEXGS_ALPHA_MIN);
                if (!__any_sync(0xffffffffu, contrib)) continue;
                const unsigned id = st.id[j];
                float v[6 + TEXGS_EXTRA_CH];
#pragma unroll
                for (int q = 0; q < 6 + TEXGS_EXTRA_CH; ++q) v[q] = 0.f;
                if (contrib) {
                    const float inv_1ma = __fdividef(1.0f, 1.0f - alpha);
                    T = T * inv_1ma;
                    const float w = alpha * T;
                    const float* __restrict__ ea = p.extra_attrs + (size_t)id * p.E + c0;
                    float X = 0.f;
#pragma unroll
                    for (int k = 0; k < TEXGS_EXTRA_CH; ++k)
                        if (c0 + k < p.E) { X += g[k] * __ldg(ea + k); v[6 + k] = w * g[k]; }
                    acc_rec = last_alpha * last_X + (1.0f - last_alpha) * acc_rec;
                    const float dL_dalpha = (X - acc_rec) * T;                 // no background term on extra channels
                    last_alpha = alpha;
                    last_X = X;
                    const float live = (aG <= TEXGS_ALPHA_MAX) ? 1.f : 0.f;    // clamp active -> zero derivative
                    const float dL_dG = live * g1.y * dL_dalpha;
                    v[5] = live * G * dL_dalpha;
                    const float GdG = G * dL_dG;
                    v[0] = -GdG * (g0.z * dx + g0.w * dy);
                    v[1] = -GdG * (g1.x * dy + g0.w * dx);
                    v[2] = -0.5f * GdG * dx * dx;
                    v[3] = -GdG * dx * dy;
                    v[4] = -0.5f * GdG * dy * dy;
                }
#pragma unroll
                for (int q = 0; q < 6 + TEXGS_EXTRA_CH; ++q) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
                }
                if (lane == 0) {
                    float* dst = acc + (size_t)id * TEXGS_BWD_ACC_FLOATS;
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        if (v[q] != 0.f) atomicAdd(dst + q, v[q]);
                    if (dextra_attrs) {
#pragma unroll
                        for (int k = 0; k < TEXGS_EXTRA_CH; ++k)
                            if (c0 + k < p.E && v[6 + k] != 0.f) atomicAdd(dextra_attrs + (size_t)id * p.E + c0 + k, v[6 + k]);
                    }
                }
            }
        }
    }
}
