// texgs_optim.cuh — SURVEY §8f N4: the texture's optimizer step (reference models/texture_gaussian3d.py:139-143
// builds torch.optim.Adam(lr=tex_lr, eps=1e-15) over the (6,R,R,3) texture, :439-440 steps it every iteration).
// Dense Adam, same update as torch.optim.Adam (no weight decay, no amsgrad):
//     m <- m + (g - m) * (1 - b1)          v <- v * b2 + (1 - b2) * g * g
//     p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// fused with what surrounds it on this path: the gradient is read straight from the padded (6,R,R,4) buffer the
// rasterizer backward accumulates into (GradBucket), is zeroed for the next step in the same pass, and the packed
// RGBA copy of the updated texture (TexgsFwdArgs.texture_rgba) is emitted too — instead of torch's ~8 foreach
// passes + a fill + a repack. HBM-bound: 120 B per texel all options on (p, m, v read+write 72, g read 16 + zero 16,
// rgba write 16), every global access a fully coalesced 16 B per thread (texel <-> flat re-indexing goes through
// shared memory).
#pragma once
#include "texgs_common.cuh"

#define TEXGS_ADAM_THREADS 256
#define TEXGS_ADAM_TEXELS 1024                      // per CTA: 3072 parameter floats = 768 float4

struct AdamArgs {
    float* p; float* m; float* v;     // (n*3) floats each, 16-byte aligned
    const float* g3;                  // (n,3) gradient or NULL
    float* g4;                        // (n,4) padded gradient or NULL (exactly one of g3 / g4)
    float* rgba;                      // (n,4) packed copy of the updated parameter, or NULL
    unsigned long long n;             // texels (a flat tensor of k floats, k % 3 == 0, is n = k/3 "texels")
    float one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps;
    int zero_grad;
};

__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const AdamArgs& a) {
    m = m + (g - m) * a.one_minus_b1;
    v = v * a.b2 + (a.one_minus_b2 * g) * g;
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.step_size * (m / denom);
}

#ifndef TEXGS_ADAM_STREAM_HINTS
#define TEXGS_ADAM_STREAM_HINTS 0        // 1 = ld/st .cs (evict-first); measured 0.463 vs 0.446 ms without: off
#endif
#ifndef TEXGS_ADAM_MIN_CTAS
#define TEXGS_ADAM_MIN_CTAS 3
#endif

__device__ __forceinline__ float4 adam_ld(const float4* p) {
#if TEXGS_ADAM_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void adam_st(float4* p, float4 v) {
#if TEXGS_ADAM_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

__global__ void __launch_bounds__(TEXGS_ADAM_THREADS, TEXGS_ADAM_MIN_CTAS) texgs_texture_adam_kernel(const AdamArgs a) {
    __shared__ __align__(16) float sg[TEXGS_ADAM_TEXELS * 3];          // gradient, then the updated parameter
    const unsigned long long t0 = (unsigned long long)blockIdx.x * TEXGS_ADAM_TEXELS;
    const int tid = threadIdx.x;
    const bool full = t0 + TEXGS_ADAM_TEXELS <= a.n;
    constexpr int KT = TEXGS_ADAM_TEXELS / TEXGS_ADAM_THREADS;           // texel float4s per thread (4)
    constexpr int KF = 3 * TEXGS_ADAM_TEXELS / 4 / TEXGS_ADAM_THREADS;   // flat float4s per thread (3)
    if (full) {
        // 0. every global load of the tile is issued before the first store: one memory round trip per CTA
        //    (13 x 16 B in flight per thread) instead of one per unrolled iteration
        float4* p4 = reinterpret_cast<float4*>(a.p + t0 * 3);
        float4* m4 = reinterpret_cast<float4*>(a.m + t0 * 3);
        float4* v4 = reinterpret_cast<float4*>(a.v + t0 * 3);
        float4 P[KF], M[KF], V[KF], G[KT];
        if (a.g4) {
            const float4* g4 = reinterpret_cast<const float4*>(a.g4) + t0;
#pragma unroll
            for (int k = 0; k < KT; ++k) G[k] = adam_ld(g4 + tid + k * TEXGS_ADAM_THREADS);
        } else {
            const float4* g3 = reinterpret_cast<const float4*>(a.g3 + t0 * 3);
#pragma unroll
            for (int k = 0; k < KF; ++k) G[k] = adam_ld(g3 + tid + k * TEXGS_ADAM_THREADS);
        }
#pragma unroll
        for (int k = 0; k < KF; ++k) {
            const int f = tid + k * TEXGS_ADAM_THREADS;
            P[k] = adam_ld(p4 + f); M[k] = adam_ld(m4 + f); V[k] = adam_ld(v4 + f);
        }
        // 1. gradient tile -> shared memory in flat (texel*3 + channel) order; clear it in global memory
        if (a.g4) {
            float4* g4 = reinterpret_cast<float4*>(a.g4) + t0;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const int t = tid + k * TEXGS_ADAM_THREADS;
                sg[3 * t] = G[k].x; sg[3 * t + 1] = G[k].y; sg[3 * t + 2] = G[k].z;
                if (a.zero_grad) adam_st(g4 + t, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        } else {
            float4* g3 = reinterpret_cast<float4*>(const_cast<float*>(a.g3) + t0 * 3);
#pragma unroll
            for (int k = 0; k < KF; ++k) {
                const int f = tid + k * TEXGS_ADAM_THREADS;
                reinterpret_cast<float4*>(sg)[f] = G[k];
                if (a.zero_grad) adam_st(g3 + f, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        __syncthreads();
        // 2. elementwise update on flat float4s
#pragma unroll
        for (int k = 0; k < KF; ++k) {
            const int f = tid + k * TEXGS_ADAM_THREADS;
            const float4 g = reinterpret_cast<float4*>(sg)[f];
            adam_elem(P[k].x, M[k].x, V[k].x, g.x, a); adam_elem(P[k].y, M[k].y, V[k].y, g.y, a);
            adam_elem(P[k].z, M[k].z, V[k].z, g.z, a); adam_elem(P[k].w, M[k].w, V[k].w, g.w, a);
            adam_st(p4 + f, P[k]); adam_st(m4 + f, M[k]); adam_st(v4 + f, V[k]);
            reinterpret_cast<float4*>(sg)[f] = P[k];          // same thread, same slot: no hazard
        }
        // 3. packed copy of the updated texels (never given a streaming hint: the next render samples it)
        if (a.rgba) {
            __syncthreads();
            float4* o4 = reinterpret_cast<float4*>(a.rgba) + t0;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const int t = tid + k * TEXGS_ADAM_THREADS;
                o4[t] = make_float4(sg[3 * t], sg[3 * t + 1], sg[3 * t + 2], 0.f);
            }
        }
    } else {
        // tail tile (n % 1024 texels): scalar, no alignment assumptions beyond 4 bytes
        for (unsigned long long t = t0 + tid; t < a.n; t += TEXGS_ADAM_THREADS) {
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned long long e = t * 3 + c;
                const float g = a.g4 ? a.g4[t * 4 + c] : a.g3[e];
                float p = a.p[e], m = a.m[e], v = a.v[e];
                adam_elem(p, m, v, g, a);
                a.p[e] = p; a.m[e] = m; a.v[e] = v;
                o[c] = p;
                if (a.zero_grad) { if (a.g4) a.g4[t * 4 + c] = 0.f; else const_cast<float*>(a.g3)[e] = 0.f; }
            }
            if (a.zero_grad && a.g4) a.g4[t * 4 + 3] = 0.f;
            if (a.rgba) reinterpret_cast<float4*>(a.rgba)[t] = make_float4(o[0], o[1], o[2], 0.f);
        }
    }
}

// =============================================================================================================
// Data-parallel texture step: gradient reduction + Adam + parameter broadcast in ONE kernel over NVLink.
//
// With the view batch sharded over N GPUs every rank holds a dense partial gradient of the whole texture. Instead of
// all-reducing it (2 (N-1)/N x 403 MB over the links) and then running the same Adam step on every rank, rank r owns the
// texels of tile range [tile_lo, tile_hi) (1/N of the texture) and, for each of its tiles,
//   1. PULLS the N partial gradients of the tile out of the ranks' symmetric buffers and adds them
//        - multicast path: one multimem.ld_reduce per 16 bytes, the NVSwitch adds the N copies in the fabric (NVLS)
//        - peer path: N plain 128-bit loads through the peer mappings (NVLink P2P)
//   2. updates its shard of the Adam moments (kept only for the owned texels: the optimizer state is sharded)
//   3. PUSHES the updated parameter texels into every rank's copy of the texture
//        - multicast path: multimem.st (the switch replicates), peer path: N stores.
// Links carry (N-1)/N x (403 + 302) MB per rank and direction (peer path) or 403 (N-1)/N out + 302 (N-1)/N in (multicast)
// instead of 2 (N-1)/N x 403 MB each way, the optimizer arithmetic and its 24 B/texel of state traffic divide by N, and
// no gradient is written back. Callers bracket the launch with device-side barriers over the symmetric-memory signal
// pads (every rank's backward is complete before anyone pulls; nobody clears its gradient or samples the texture before
// every owner is done) — texture_gs_b200/dist.py DistTextureAdam.
// =============================================================================================================
#define TEXGS_DP_MAX_RANKS 16

struct DpAdamArgs {
    int world, rank;
    const float* grad[TEXGS_DP_MAX_RANKS];     // every rank's padded (n,4) gradient (peer mappings; [rank] is local)
    float* param[TEXGS_DP_MAX_RANKS];          // every rank's (n,3) parameter
    const float* grad_mc;                      // multicast mappings of the same two buffers, or NULL (peer path)
    float* param_mc;
    float* m; float* v;                        // moments of the OWNED texels only: ((tile_hi - tile_lo) * 1024 * 3) floats
    unsigned long long n;                      // texels of the whole texture
    unsigned long long tile_lo, tile_hi;       // owned tiles of TEXGS_ADAM_TEXELS texels
    float one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps;
};

#ifdef TEXGS_HOST_EMU      // the emulator (tests/simt) runs the PEER path of the kernel below: several "ranks" are buffers of one process
inline float4 multimem_ld_reduce_add(const float4*) { simt::fail("multimem needs an NVSwitch fabric"); return make_float4(0.f, 0.f, 0.f, 0.f); }
inline void multimem_st(float4*, float4) { simt::fail("multimem needs an NVSwitch fabric"); }
#else
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#endif

template <bool MC>
__global__ void __launch_bounds__(TEXGS_ADAM_THREADS, 2) texgs_texture_adam_dp_kernel(const DpAdamArgs a) {
    __shared__ __align__(16) float sg[TEXGS_ADAM_TEXELS * 3];
    const unsigned long long tile = a.tile_lo + blockIdx.x;
    const unsigned long long t0 = tile * TEXGS_ADAM_TEXELS;                      // first texel of the tile (global numbering)
    const unsigned long long l0 = (unsigned long long)blockIdx.x * TEXGS_ADAM_TEXELS;   // ... in the owned shard (moments)
    const int tid = threadIdx.x;
    constexpr int KT = TEXGS_ADAM_TEXELS / TEXGS_ADAM_THREADS;           // texel float4s per thread (4)
    constexpr int KF = 3 * TEXGS_ADAM_TEXELS / 4 / TEXGS_ADAM_THREADS;   // flat float4s per thread (3)
    AdamArgs e;                                                          // the elementwise update shares adam_elem()
    e.one_minus_b1 = a.one_minus_b1; e.b2 = a.b2; e.one_minus_b2 = a.one_minus_b2;
    e.step_size = a.step_size; e.inv_sqrt_bc2 = a.inv_sqrt_bc2; e.eps = a.eps;
    if (t0 + TEXGS_ADAM_TEXELS <= a.n) {
        // 1. pull + add the N partial gradients of the tile; the local parameter / moment loads fly with them
        float4 G[KT];
        if (MC) {
            const float4* g = reinterpret_cast<const float4*>(a.grad_mc) + t0;
#pragma unroll
            for (int k = 0; k < KT; ++k) G[k] = multimem_ld_reduce_add(g + tid + k * TEXGS_ADAM_THREADS);
        } else {
#pragma unroll
            for (int k = 0; k < KT; ++k) G[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < a.world; ++r) {
                const float4* g = reinterpret_cast<const float4*>(a.grad[(a.rank + r) % a.world]) + t0;   // start at home: spreads the peers
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float4 x = __ldcs(g + tid + k * TEXGS_ADAM_THREADS);
                    G[k].x += x.x; G[k].y += x.y; G[k].z += x.z;
                }
            }
        }
        float4* p4 = reinterpret_cast<float4*>(a.param[a.rank] + t0 * 3);
        float4* m4 = reinterpret_cast<float4*>(a.m + l0 * 3);
        float4* v4 = reinterpret_cast<float4*>(a.v + l0 * 3);
        float4 P[KF], M[KF], V[KF];
#pragma unroll
        for (int k = 0; k < KF; ++k) {
            const int f = tid + k * TEXGS_ADAM_THREADS;
            P[k] = p4[f]; M[k] = m4[f]; V[k] = v4[f];
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int t = tid + k * TEXGS_ADAM_THREADS;
            sg[3 * t] = G[k].x; sg[3 * t + 1] = G[k].y; sg[3 * t + 2] = G[k].z;
        }
        __syncthreads();
        // 2. Adam on flat float4s, 3. push the updated parameters to every rank
#pragma unroll
        for (int k = 0; k < KF; ++k) {
            const int f = tid + k * TEXGS_ADAM_THREADS;
            const float4 g = reinterpret_cast<float4*>(sg)[f];
            adam_elem(P[k].x, M[k].x, V[k].x, g.x, e); adam_elem(P[k].y, M[k].y, V[k].y, g.y, e);
            adam_elem(P[k].z, M[k].z, V[k].z, g.z, e); adam_elem(P[k].w, M[k].w, V[k].w, g.w, e);
            m4[f] = M[k]; v4[f] = V[k];
            if (MC) {
                multimem_st(reinterpret_cast<float4*>(a.param_mc + t0 * 3) + f, P[k]);
            } else {
                for (int r = 0; r < a.world; ++r)
                    reinterpret_cast<float4*>(a.param[(a.rank + r) % a.world] + t0 * 3)[f] = P[k];
            }
        }
    } else {
        // tail tile (n % 1024 texels): scalar, peer path on both variants (a handful of texels)
        for (unsigned long long t = t0 + tid; t < a.n; t += TEXGS_ADAM_THREADS) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float g = 0.f;
                for (int r = 0; r < a.world; ++r) g += a.grad[r][t * 4 + c];
                const unsigned long long el = (t - a.tile_lo * TEXGS_ADAM_TEXELS) * 3 + c;
                float p = a.param[a.rank][t * 3 + c], m = a.m[el], v = a.v[el];
                adam_elem(p, m, v, g, e);
                a.m[el] = m; a.v[el] = v;
                for (int r = 0; r < a.world; ++r) a.param[r][t * 3 + c] = p;
            }
        }
    }
}
