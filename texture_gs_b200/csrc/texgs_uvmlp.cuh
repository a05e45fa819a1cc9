// texgs_uvmlp.cuh — SURVEY §8f N1: the UV + Jacobian producer of the training step, fused.
//
// Reference: TextureGaussian3D.get_uvs (models/texture_gaussian3d.py:230-236) evaluates
//     uv = normalize(mlp(relu(pre_mlp(x') + emb)))         models/modules/uv_net.py:19-36
// with pre_mlp = 3 -> 128 -ReLU-> 128 and mlp = 128 -ReLU-> 128 -ReLU-> 128 -> 3
// (configs/texture_gaussian3d.yaml:18-27; tiny-cuda-nn FullyFusedMLP, fp16, no biases — or the
// nn.Linear fallback with biases, models/modules/utils.py:44-55), and get_grad_uvs (:217-227) obtains
// J = d uv / d xyz with torch.autograd.functional.jacobian: three more backward passes, every iteration.
//
// Here one kernel produces uv AND J: the value and the three forward-mode tangents (d/dx, d/dy, d/dz) of a
// point are four rows of the same GEMM, so a tile of 32 points is one M=128 operand and every hidden layer
// is one 128x128x128 tcgen05.mma chain (fp16 operands in shared memory, fp32 accumulators in TMEM).
//   rows 0-31: values, rows 32-63 / 64-95 / 96-127: tangents 0 / 1 / 2 of the same 32 points
//   -> warp s of a 4-warp group owns stream s (tcgen05.ld gives lane p of warp s TMEM lane 32 s + p);
//      ReLU masks come from the value rows: warp 0 publishes one 32-bit mask word per point and 32-column chunk.
// Two (UVMLP_GROUPS) independent 4-warp groups per CTA share the resident weights (96 KB) and interleave:
// while one group runs its epilogue (TMEM -> registers -> bias/ReLU/mask -> fp16 -> shared memory, in the
// UMMA K-major core-matrix layout so the result IS the next layer's A operand), the tensor core runs the
// other group's MMAs. The 3 -> 128 input layer is a single K = 16 MMA whose operands carry x', W1 and b1 as fp16
// hi + lo pairs (w_h x_h + w_h x_l + w_l x_h + b_h + b_l: fp32-grade, the input coordinates are NOT rounded to fp16);
// the 128 -> 3 output layer is an N = 16 MMA.
//
// Shared-memory operand layout (SWIZZLE_NONE, K-major): 16-byte chunk (row r, k-chunk kc) at
// kc * rows*16 + r*16, i.e. 8x8 core matrices are 128 contiguous bytes, SBO (next 8 rows) = 128 B,
// LBO (next 8 k) = rows*16 B. An epilogue thread owns one row: its 16-byte stores are conflict-free.
#pragma once
#include <cuda_fp16.h>
#include "texgs_common.cuh"

#define UVMLP_H 128
#define UVMLP_TILE 32
#ifndef UVMLP_GROUPS
#define UVMLP_GROUPS 3
#endif
#define UVMLP_THREADS (UVMLP_GROUPS * 128)
#define UVMLP_TMEM_COLS (UVMLP_GROUPS <= 1 ? 128 : (UVMLP_GROUPS == 2 ? 256 : 512))
#define UVMLP_MAT_BYTES (128 * 128 * 2)      // one 128x128 fp16 operand
#define UVMLP_W5_BYTES (16 * 128 * 2)        // output layer padded to N = 16

struct UvMlpParams {
    int N;
    const float* xyz;             // (N,3)
    float off[3], inv_scale[3];   // x' = (x - off) * inv_scale          (uv_net.py:22-25)
    const float* W1;              // (128,3) fp32
    const float* b1;              // (128) or NULL
    const __half* W[3];           // hidden layers 2..4, (128,128) [out][in] fp16
    const float* b[3];            // their biases or NULL
    const float* emb;             // (128): added to layer 2's output before its ReLU (uv_net.py:33)
    const __half* W5;             // (3,128) fp16
    const float* b5;              // (3) or NULL
    float* uv;                    // (N,3)
    float* J;                     // (N,9) row-major d uv_i / d x_j at 3 i + j, or NULL
    __half* stash[4];             // post-activation a1..a4 (N,128) for the backward pass, or NULL
    float* stash_inv_len;         // (N) 1 / max(|mlp output|, eps) for the backward pass, or NULL
    float* dbg;                   // debug: raw accumulators of tile 0 — [4][128][128] then [128][16]
};

struct UvMlpSmem {
    static constexpr int W = 0;                                            // 3 x 32768: W2..W4
    static constexpr int W5 = W + 3 * UVMLP_MAT_BYTES;                     // 4096: W5 padded to 16 rows
    static constexpr int W1 = W5 + UVMLP_W5_BYTES;                         // 4096: W1 / b1 as hi+lo pairs, 128 rows x K=16
    static constexpr int A = W1 + 4096;                                    // GROUPS x 32768
    static constexpr int BIAS = A + UVMLP_GROUPS * UVMLP_MAT_BYTES;        // 4 x 128 floats: 0, emb + b2, b3, b4
    static constexpr int B5 = BIAS + 4 * 128 * 4;                          // 4 floats
    static constexpr int MASK = B5 + 16;                                   // GROUPS x 2 buffers x 4 x 32 x uint4
    static constexpr int UV = MASK + UVMLP_GROUPS * 4096;                  // GROUPS x 32 x float4
    static constexpr int JB = UV + UVMLP_GROUPS * 512;                     // GROUPS x 32 x 9 floats
    static constexpr int BAR = JB + UVMLP_GROUPS * 1152;                   // GROUPS x u64
    static constexpr int TMEM = BAR + UVMLP_GROUPS * 8;                    // u32
    static constexpr int TOTAL = TMEM + 16;
};
static_assert(UvMlpSmem::TOTAL <= 227 * 1024, "shared memory budget");

// ---- PTX wrappers (CUDA 12.9, sm_100a) ---------------------------------------------------------------
__device__ __forceinline__ void uv_group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void uv_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void uv_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void uv_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void uv_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t uv_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE [61,64)
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t uv_instr_desc(int M, int Nn) {
    // cute::UMMA::InstrDescriptor: D fp32 (1 @ [4,6)), A/B fp16 (0), both K-major (0), N>>3 @ [17,23), M>>4 @ [24,29)
    return (1u << 4) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void uv_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void uv_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a descriptor mistake must end in a trap (an error code), never in a hung GPU
__device__ __forceinline__ void uv_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();
}
#define UV_TMEM_LD32(v, taddr)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
                 "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                      \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),       \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),      \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                    \
                 : "r"(taddr)                                                                                                   \
                 : "memory")
#define UV_TMEM_LD16(v, taddr)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"        \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                     \
                 : "r"(taddr)                                                                                                   \
                 : "memory")

__device__ __forceinline__ uint32_t uv_pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// fp16 hi/lo split: v ~= hi + lo with ~22 significant bits
__device__ __forceinline__ void uv_split(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t uv_h2(__half a, __half b) {
    const __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(UVMLP_THREADS, 1) texgs_uvmlp_fwd_kernel(const UvMlpParams P) {
    extern __shared__ __align__(1024) unsigned char uv_smem[];
    unsigned char* sW = uv_smem + UvMlpSmem::W;
    unsigned char* sW5 = uv_smem + UvMlpSmem::W5;
    unsigned char* sW1 = uv_smem + UvMlpSmem::W1;
    float* sBias = reinterpret_cast<float*>(uv_smem + UvMlpSmem::BIAS);
    float* sB5 = reinterpret_cast<float*>(uv_smem + UvMlpSmem::B5);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(uv_smem + UvMlpSmem::BAR);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(uv_smem + UvMlpSmem::TMEM);

    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, s = gt >> 5, lane = tid & 31;
    unsigned char* sA = uv_smem + UvMlpSmem::A + g * UVMLP_MAT_BYTES;
    uint4* sMask = reinterpret_cast<uint4*>(uv_smem + UvMlpSmem::MASK + g * 4096);      // [buf 2][k4 4][lane 32]
    float4* sUV = reinterpret_cast<float4*>(uv_smem + UvMlpSmem::UV + g * 512);
    float* sJ = reinterpret_cast<float*>(uv_smem + UvMlpSmem::JB + g * 1152);

    // ---- one-time per CTA: weights -> shared memory in the UMMA layout ---------------------------------
    for (int q = tid; q < 3 * 2048; q += UVMLP_THREADS) {
        const int l = q >> 11, r = q & 2047, n = r >> 4, kc = r & 15;
        *reinterpret_cast<uint4*>(sW + l * UVMLP_MAT_BYTES + kc * 2048 + n * 16) = reinterpret_cast<const uint4*>(P.W[l])[n * 16 + kc];
    }
    for (int q = tid; q < 256; q += UVMLP_THREADS) {
        const int n = q >> 4, kc = q & 15;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < 3) v = reinterpret_cast<const uint4*>(P.W5)[n * 16 + kc];
        *reinterpret_cast<uint4*>(sW5 + kc * 256 + n * 16) = v;
    }
    for (int c = tid; c < 128; c += UVMLP_THREADS) {
        // W1 row c, k = 0..15:  [wh0 wh1 wh2 bh | wh0 wh1 wh2 bl] [wl0 wl1 wl2 0 | 0 0 0 0]
        __half wh[3], wl[3], bh, bl;
#pragma unroll
        for (int j = 0; j < 3; ++j) uv_split(P.W1[3 * c + j], wh[j], wl[j]);
        uv_split(P.b1 ? P.b1[c] : 0.f, bh, bl);
        const __half z = __float2half_rn(0.f);
        *reinterpret_cast<uint4*>(sW1 + c * 16) = make_uint4(uv_h2(wh[0], wh[1]), uv_h2(wh[2], bh), uv_h2(wh[0], wh[1]), uv_h2(wh[2], bl));
        *reinterpret_cast<uint4*>(sW1 + 2048 + c * 16) = make_uint4(uv_h2(wl[0], wl[1]), uv_h2(wl[2], z), 0u, 0u);
        sBias[c] = 0.f;                                                   // layer 1: b1 rides in the MMA
        sBias[128 + c] = P.emb[c] + (P.b[0] ? P.b[0][c] : 0.f);
        sBias[256 + c] = P.b[1] ? P.b[1][c] : 0.f;
        sBias[384 + c] = P.b[2] ? P.b[2][c] : 0.f;
    }
    if (tid < 4) sB5[tid] = (tid < 3 && P.b5) ? P.b5[tid] : 0.f;
    if (tid == 0) {
        for (int i = 0; i < UVMLP_GROUPS; ++i) mbar_init(&sBar[i], 1);
        mbar_fence_init();
    }
    if (tid < 32) {   // warp 0 owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"((uint32_t)UVMLP_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    uv_fence_async_smem();
    uv_tc_fence_before();
    __syncthreads();
    uv_tc_fence_after();
    const uint32_t tmem_cta = *sTmem;
    const uint32_t tmem_d = tmem_cta + (uint32_t)(g * 128);                 // this group's accumulator columns
    const uint32_t tmem_row = tmem_d + ((uint32_t)(s * 32) << 16);          // this warp's 32 TMEM lanes

    const uint32_t lboA = 2048u, sboA = 128u;       // 128-row operands: next k-chunk 2048 B, next 8 rows 128 B
    const uint32_t lbo5 = 256u, sbo5 = 128u;        // 16-row W5
    const uint32_t idesc = uv_instr_desc(128, 128), idesc5 = uv_instr_desc(128, 16);
    const uint32_t aA = smem_u32(sA), aW = smem_u32(sW), aW5 = smem_u32(sW5), aW1 = smem_u32(sW1);
    const int row = s * 32 + lane;
    const int ntiles = (P.N + UVMLP_TILE - 1) / UVMLP_TILE;
    uint32_t phase = 0;
    // tangent rows: inv_scale_j (hi + lo) at slot j = s - 1
    __half t_hi = __float2half_rn(0.f), t_lo = t_hi;
    if (s > 0) uv_split((s == 1) ? P.inv_scale[0] : ((s == 2) ? P.inv_scale[1] : P.inv_scale[2]), t_hi, t_lo);

    for (int tile = blockIdx.x * UVMLP_GROUPS + g; tile < ntiles; tile += gridDim.x * UVMLP_GROUPS) {
        const int p0 = tile * UVMLP_TILE, pt = p0 + lane;
        const bool valid = pt < P.N;
        const bool dump = P.dbg != nullptr && tile == 0;
        // ---- layer 1 operand row (K = 16): value rows carry x' (hi, 1, lo, 1 | hi), tangent row j carries inv_scale_j e_j ----
        {
            const __half z = __float2half_rn(0.f), one = __float2half_rn(1.f);
            __half h0 = z, h1 = z, h2 = z, l0 = z, l1 = z, l2 = z, c3 = z;
            if (s == 0) {
                if (valid) {
                    uv_split((P.xyz[3 * pt] - P.off[0]) * P.inv_scale[0], h0, l0);
                    uv_split((P.xyz[3 * pt + 1] - P.off[1]) * P.inv_scale[1], h1, l1);
                    uv_split((P.xyz[3 * pt + 2] - P.off[2]) * P.inv_scale[2], h2, l2);
                }
                c3 = one;
            } else if (s == 1) { h0 = t_hi; l0 = t_lo; }
            else if (s == 2) { h1 = t_hi; l1 = t_lo; }
            else { h2 = t_hi; l2 = t_lo; }
            *reinterpret_cast<uint4*>(sA + row * 16) = make_uint4(uv_h2(h0, h1), uv_h2(h2, c3), uv_h2(l0, l1), uv_h2(l2, c3));
            *reinterpret_cast<uint4*>(sA + 2048 + row * 16) = make_uint4(uv_h2(h0, h1), uv_h2(h2, z), 0u, 0u);
        }
        uv_fence_async_smem();
        uv_group_bar(g);

        // ---- layers 1..4: MMA chain, then epilogue (bias, ReLU, mask, fp16) writing the next layer's A operand --------
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            if (gt == 0) {
                uv_tc_fence_after();
                if (l == 0) {
                    uv_umma(tmem_d, uv_smem_desc(aA, lboA, sboA), uv_smem_desc(aW1, lboA, sboA), idesc, 0u);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        uv_umma(tmem_d, uv_smem_desc(aA + k * 4096, lboA, sboA),
                                uv_smem_desc(aW + (l - 1) * UVMLP_MAT_BYTES + k * 4096, lboA, sboA), idesc, k > 0 ? 1u : 0u);
                }
                uv_umma_commit(&sBar[g]);
            }
            __syncwarp();
            uv_mbar_wait(&sBar[g], phase);
            phase ^= 1u;
            uv_tc_fence_after();
            const float* bias = sBias + l * 128;
            uint32_t r[4][32];
            UV_TMEM_LD32(r[0], tmem_row);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uv_tmem_wait_ld();
                if (cc < 3) UV_TMEM_LD32(r[(cc + 1) & 3], tmem_row + (uint32_t)((cc + 1) * 32));   // in flight while this chunk is processed
                if (dump && g == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) P.dbg[((size_t)l * 128 + row) * 128 + cc * 32 + i] = __uint_as_float(r[cc][i]);
                }
                uint32_t h[16];
                uint4* mbuf = sMask + (cc & 1) * 128 + lane;
                if (s == 0) {
                    const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        __half2 v = __floats2half2_rn(__uint_as_float(r[cc][2 * k]) + bias[cc * 32 + 2 * k],
                                                      __uint_as_float(r[cc][2 * k + 1]) + bias[cc * 32 + 2 * k + 1]);
                        v = __hmax2(v, zero2);
                        h[k] = *reinterpret_cast<const uint32_t*>(&v);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t m[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) m[k] = __hgt2_mask(*reinterpret_cast<const __half2*>(&h[4 * q + k]), zero2);
                        mbuf[q * 32] = make_uint4(m[0], m[1], m[2], m[3]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) h[k] = uv_pack_half2(__uint_as_float(r[cc][2 * k]), __uint_as_float(r[cc][2 * k + 1]));
                }
                uv_group_bar(g);
                if (s != 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 m = mbuf[q * 32];
                        h[4 * q] &= m.x; h[4 * q + 1] &= m.y; h[4 * q + 2] &= m.z; h[4 * q + 3] &= m.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 hv = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
                    *reinterpret_cast<uint4*>(sA + (cc * 4 + q) * 2048 + row * 16) = hv;
                    if (s == 0 && P.stash[l] && valid) reinterpret_cast<uint4*>(P.stash[l] + (size_t)pt * 128)[cc * 4 + q] = hv;
                }
            }
            uv_tc_fence_before();
            uv_fence_async_smem();
            uv_group_bar(g);
        }

        // ---- layer 5 (128 -> 3, padded to N = 16), then normalize and its Jacobian ----------------------------------
        if (gt == 0) {
            uv_tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                uv_umma(tmem_d, uv_smem_desc(aA + k * 4096, lboA, sboA), uv_smem_desc(aW5 + k * 512, lbo5, sbo5), idesc5, k > 0 ? 1u : 0u);
            uv_umma_commit(&sBar[g]);
        }
        __syncwarp();
        uv_mbar_wait(&sBar[g], phase);
        phase ^= 1u;
        uv_tc_fence_after();
        float o0, o1, o2;
        {
            uint32_t r[16];
            UV_TMEM_LD16(r, tmem_row);
            uv_tmem_wait_ld();
            uv_tc_fence_before();
            o0 = __uint_as_float(r[0]); o1 = __uint_as_float(r[1]); o2 = __uint_as_float(r[2]);
            if (dump && g == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) P.dbg[(size_t)4 * 128 * 128 + row * 16 + i] = __uint_as_float(r[i]);
            }
        }
        if (s == 0) {
            o0 += sB5[0]; o1 += sB5[1]; o2 += sB5[2];
            const float len = sqrtf(o0 * o0 + o1 * o1 + o2 * o2);
            const float inv = 1.0f / fmaxf(len, 1e-12f);                    // F.normalize(eps=1e-12)
            sUV[lane] = make_float4(o0 * inv, o1 * inv, o2 * inv, inv);
            if (P.stash_inv_len && valid) P.stash_inv_len[pt] = inv;
        }
        uv_group_bar(g);
        if (s != 0) {
            const float4 u = sUV[lane];
            const float d = u.x * o0 + u.y * o1 + u.z * o2;                  // (I - u u^T) t / |out|
            const int j = s - 1;
            sJ[lane * 9 + j] = (o0 - u.x * d) * u.w;
            sJ[lane * 9 + 3 + j] = (o1 - u.y * d) * u.w;
            sJ[lane * 9 + 6 + j] = (o2 - u.z * d) * u.w;
        }
        uv_group_bar(g);
        {
            const size_t n3 = (size_t)P.N * 3, n9 = (size_t)P.N * 9;
            if (gt < 96) {
                const size_t o = (size_t)p0 * 3 + gt;
                if (o < n3) {
                    const float4 u = sUV[gt / 3];
                    const int c = gt % 3;
                    P.uv[o] = (c == 0) ? u.x : ((c == 1) ? u.y : u.z);
                }
            }
            if (P.J) {
                for (int i = gt; i < 288; i += 128) {
                    const size_t o = (size_t)p0 * 9 + i;
                    if (o < n9) P.J[o] = sJ[i];
                }
            }
        }
        // next tile: sA is free (layer 5's MMA completed), sUV / sJ are rewritten only after later group barriers
    }

    uv_tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        uv_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_cta), "r"((uint32_t)UVMLP_TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// Backward of uv w.r.t. xyz / emb / weights. The two 128x128 products per hidden layer (delta @ W and
// delta^T @ a) run on the tensor cores in texgs_uvmlp_bwd_layer_kernel (end of this file); the K = 3 layers at both
// ends and the normalisation are memory-bound glue, two streaming kernels:
//   head : normalize backward + output layer (K = 3): d = (g - u (u.g)) / |out|;  gW5 += d^T a4, gb5 += sum d,
//          delta4 = S * (d W5) * (a4 > 0) in fp16, column sums of delta4
//   tail : input layer (K = 3): gxyz = delta1 W1 * inv_scale / S,  gW1 += delta1^T x' / S
// S is one power of two taken from max|d| on the device (mixed-precision loss scaling, no host sync).
// Thread layout everywhere: 16 lanes x 8 columns cover a row (one uint4 each), 16 rows per 256-thread pass.
// =====================================================================================================
#define UVBWD_THREADS 256
#define UVBWD_ROWS_PER_CTA 512

__device__ __forceinline__ float uv_loss_scale(float amax) { return exp2f(floorf(log2f(256.0f / fmaxf(amax, 1e-30f)))); }

__device__ __forceinline__ float3 uv_norm_bwd(const float* __restrict__ g_uv, const float* __restrict__ uv, const float* __restrict__ inv_len, int n) {
    const float g0 = g_uv[3 * n], g1 = g_uv[3 * n + 1], g2 = g_uv[3 * n + 2];
    const float u0 = uv[3 * n], u1 = uv[3 * n + 1], u2 = uv[3 * n + 2];
    const float dt = g0 * u0 + g1 * u1 + g2 * u2, il = inv_len[n];
    return make_float3((g0 - u0 * dt) * il, (g1 - u1 * dt) * il, (g2 - u2 * dt) * il);
}

__global__ void __launch_bounds__(256) texgs_uvmlp_bwd_amax_kernel(int N, const float* __restrict__ g_uv, const float* __restrict__ uv,
                                                                  const float* __restrict__ inv_len, float* __restrict__ amax) {
    float m = 0.f;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float3 d = uv_norm_bwd(g_uv, uv, inv_len, n);
        m = fmaxf(m, fmaxf(fabsf(d.x), fmaxf(fabsf(d.y), fabsf(d.z))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));   // m >= 0: int order = float order
}

// reduces per-thread column accumulators acc[NV][8] (thread = row-lane tid/16, chunk tid%16) over the CTA and adds them
// to global dst[v][128] (v < NV)
template <int NV>
__device__ __forceinline__ void uvbwd_flush(float (&acc)[NV][8], float* sred /* [NV*128] */, float* const (&dst)[NV], float mul) {
    const int tid = threadIdx.x, chunk = tid & 15;
    for (int i = tid; i < NV * 128; i += UVBWD_THREADS) sred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float x = acc[v][c];
            x += __shfl_xor_sync(0xffffffffu, x, 16);                     // the warp's two row-lanes of this chunk
            if ((tid & 31) < 16) atomicAdd(&sred[v * 128 + chunk * 8 + c], x);
        }
    __syncthreads();
    for (int i = tid; i < NV * 128; i += UVBWD_THREADS) atomicAdd(&dst[i >> 7][i & 127], sred[i] * mul);
}

__device__ __forceinline__ void uv_unpack8(const uint4& h, float (&f)[8]) {
    const __half2* p = reinterpret_cast<const __half2*>(&h);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(p[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
}

__global__ void __launch_bounds__(UVBWD_THREADS) texgs_uvmlp_bwd_head_kernel(int N, const float* __restrict__ g_uv, const float* __restrict__ uv,
                                                                           const float* __restrict__ inv_len, const __half* __restrict__ a4,
                                                                           const __half* __restrict__ W5, const float* __restrict__ amax,
                                                                           float* __restrict__ scale_out, __half* __restrict__ delta,
                                                                           float* __restrict__ gW5, float* __restrict__ gb5, float* __restrict__ colsum) {
    __shared__ float sred[4 * 128];
    const int tid = threadIdx.x, chunk = tid & 15, rl = tid >> 4;
    const float S = uv_loss_scale(*amax);
    if (blockIdx.x == 0 && tid == 0) *scale_out = S;
    float w[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i) uv_unpack8(reinterpret_cast<const uint4*>(W5 + i * 128)[chunk], w[i]);
    float acc[4][8];          // gW5 rows 0..2, column sums of delta4
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[v][c] = 0.f;
    float gb[3] = {0.f, 0.f, 0.f};
    const int r0 = blockIdx.x * UVBWD_ROWS_PER_CTA;
    for (int r = r0 + rl; r < min(N, r0 + UVBWD_ROWS_PER_CTA); r += 16) {
        const float3 d = uv_norm_bwd(g_uv, uv, inv_len, r);
        float a[8];
        uv_unpack8(reinterpret_cast<const uint4*>(a4 + (size_t)r * 128)[chunk], a);
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc[0][c] += d.x * a[c]; acc[1][c] += d.y * a[c]; acc[2][c] += d.z * a[c];
            const float pre = (d.x * w[0][c] + d.y * w[1][c] + d.z * w[2][c]) * S;
            o[c] = (a[c] > 0.f) ? pre : 0.f;
            acc[3][c] += o[c];
        }
        reinterpret_cast<uint4*>(delta + (size_t)r * 128)[chunk] =
            make_uint4(uv_pack_half2(o[0], o[1]), uv_pack_half2(o[2], o[3]), uv_pack_half2(o[4], o[5]), uv_pack_half2(o[6], o[7]));
        if (chunk == 0) { gb[0] += d.x; gb[1] += d.y; gb[2] += d.z; }
    }
    float* const dst[4] = {gW5, gW5 + 128, gW5 + 256, colsum};
    uvbwd_flush<4>(acc, sred, dst, 1.0f);
    if (chunk == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float x = gb[i];
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if ((tid & 31) == 0) atomicAdd(&gb5[i], x);
        }
    }
}

__global__ void __launch_bounds__(UVBWD_THREADS) texgs_uvmlp_bwd_tail_kernel(int N, const __half* __restrict__ delta, const float* __restrict__ xyz,
                                                                           float3 off, float3 isc, const float* __restrict__ W1,
                                                                           const float* __restrict__ scale, float* __restrict__ gxyz,
                                                                           float* __restrict__ gW1 /* (128,3) */) {
    __shared__ float sred[3 * 128];
    const int tid = threadIdx.x, chunk = tid & 15, rl = tid >> 4;
    const float invS = 1.0f / *scale;
    float w[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) w[c][j] = W1[(chunk * 8 + c) * 3 + j];
    float acc[3][8];
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[v][c] = 0.f;
    const int r0 = blockIdx.x * UVBWD_ROWS_PER_CTA;
    // every lane of a warp runs the same number of iterations (rows r and r + 1 share a warp): shuffles are safe
    for (int rb = r0; rb < min(N, r0 + UVBWD_ROWS_PER_CTA); rb += 16) {
        const int r = rb + rl;
        const bool ok = r < N;
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (ok) {
            uv_unpack8(reinterpret_cast<const uint4*>(delta + (size_t)r * 128)[chunk], d);
            x0 = (xyz[3 * r] - off.x) * isc.x; x1 = (xyz[3 * r + 1] - off.y) * isc.y; x2 = (xyz[3 * r + 2] - off.z) * isc.z;
        }
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            g0 += d[c] * w[c][0]; g1 += d[c] * w[c][1]; g2 += d[c] * w[c][2];
            acc[0][c] += d[c] * x0; acc[1][c] += d[c] * x1; acc[2][c] += d[c] * x2;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {                                  // the 16 lanes of a row
            g0 += __shfl_xor_sync(0xffffffffu, g0, o); g1 += __shfl_xor_sync(0xffffffffu, g1, o); g2 += __shfl_xor_sync(0xffffffffu, g2, o);
        }
        if (gxyz && ok && chunk == 0) {
            gxyz[3 * r] = g0 * isc.x * invS; gxyz[3 * r + 1] = g1 * isc.y * invS; gxyz[3 * r + 2] = g2 * isc.z * invS;
        }
    }
    // gW1 is (128,3): column c of accumulator j -> gW1[c*3 + j]; flush through a [3][128] staging then transpose on the add
    const int t = threadIdx.x;
    for (int i = t; i < 3 * 128; i += UVBWD_THREADS) sred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float x = acc[v][c];
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if ((t & 31) < 16) atomicAdd(&sred[v * 128 + chunk * 8 + c], x);
        }
    __syncthreads();
    for (int i = t; i < 3 * 128; i += UVBWD_THREADS) atomicAdd(&gW1[(i & 127) * 3 + (i >> 7)], sred[i] * invS);
}

// =====================================================================================================
// Backward of ONE hidden layer on the tensor cores (round 2; replaces two cuBLAS GEMMs + the mask kernel per layer):
//     gW        += delta^T @ a_prev                      (128 out x 128 in, K = points)
//     delta_out  = (delta @ W) * (a_prev > 0)  in fp16   (points x 128 in)
//     colsum    += column sums of delta_out              (bias / embedding gradient of the layer below)
// A tile is 128 points; a 4-warp group owns a tile (thread = point), two groups per CTA interleave load / MMA / epilogue,
// persistent grid. Both products are tcgen05.mma chains on fp16 operands with fp32 accumulators in TMEM:
//   * delta @ W: A = the delta tile, K-major (16-byte chunk = 8 consecutive features of a point — the rows as they lie in
//     global memory), B = W^T resident in shared memory, K-major (the host passes W transposed: [in][out]);
//   * delta^T @ a_prev: K = the tile's 128 points. A = delta^T and B = a_prev^T as MN-MAJOR operands (UMMA instruction
//     descriptor bits 15 / 16): the canonical no-swizzle MN-major layout is 8(K) x 8(MN) core matrices whose 16-byte rows are
//     8 consecutive MN elements at one K — again a 16-byte row chunk of global memory, only PLACED differently:
//     chunk (point k, feature group g) at (k / 8) * LBO + g * SBO + (k % 8) * 16. The K-major delta tile of the first product
//     (chunk (point r, feature chunk kc) at kc * 2048 + r * 16) IS that layout with LBO = 128 and SBO = 2048, so one copy of the
//     tile serves both products through two descriptors; a_prev is staged with SBO = 128, LBO = 2048. No element transposition. The accumulator stays in TMEM across all tiles of the group
//     (split-K over CTAs; one fp32 red per element at the end).
// The ReLU mask of the epilogue re-reads the a_prev chunk from the MN-major copy; column sums: transposing warp reduction
// (31 shuffles per 32 columns) + shared-memory atomics.
// =====================================================================================================
#define UVBL_GROUPS 2
#define UVBL_THREADS (UVBL_GROUPS * 128)
struct UvBwdSmem {
    static constexpr int WT = 0;                                           // 32768: W^T, K-major
    static constexpr int GRP = WT + UVMLP_MAT_BYTES;                       // per group: the delta tile, the a_prev tile
    static constexpr int GRP_BYTES = 2 * UVMLP_MAT_BYTES;
    static constexpr int COLSUM = GRP + UVBL_GROUPS * GRP_BYTES;           // 128 floats
    static constexpr int BAR = COLSUM + 512;                               // GROUPS x u64
    static constexpr int TMEM = BAR + UVBL_GROUPS * 8;                     // u32
    static constexpr int TOTAL = TMEM + 16;
};
static_assert(UvBwdSmem::TOTAL <= 227 * 1024, "shared memory budget");

// sum over the warp's 32 lanes of 32 per-lane values; lane l returns the total of value l
__device__ __forceinline__ float uv_warp_transpose_reduce32(const float (&v)[32], int lane) {
    const unsigned full = 0xffffffffu;
    float a16[16], a8[8], a4[4], a2[2];
    const bool h4 = (lane & 16) != 0, h3 = (lane & 8) != 0, h2 = (lane & 4) != 0, h1 = (lane & 2) != 0, h0 = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float send = h4 ? v[i] : v[i + 16], keep = h4 ? v[i + 16] : v[i]; a16[i] = keep + __shfl_xor_sync(full, send, 16); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float send = h3 ? a16[i] : a16[i + 8], keep = h3 ? a16[i + 8] : a16[i]; a8[i] = keep + __shfl_xor_sync(full, send, 8); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float send = h2 ? a8[i] : a8[i + 4], keep = h2 ? a8[i + 4] : a8[i]; a4[i] = keep + __shfl_xor_sync(full, send, 4); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { const float send = h1 ? a4[i] : a4[i + 2], keep = h1 ? a4[i + 2] : a4[i]; a2[i] = keep + __shfl_xor_sync(full, send, 2); }
    const float send = h0 ? a2[0] : a2[1], keep = h0 ? a2[1] : a2[0];
    return keep + __shfl_xor_sync(full, send, 1);          // value index 16 h4 + 8 h3 + 4 h2 + 2 h1 + h0 = lane
}

__global__ void __launch_bounds__(UVBL_THREADS, 1) texgs_uvmlp_bwd_layer_kernel(int N, const __half* __restrict__ delta_in,
                                                                              const __half* __restrict__ a_prev, const __half* __restrict__ Wt,
                                                                              __half* __restrict__ delta_out, float* __restrict__ gW,
                                                                              float* __restrict__ colsum) {
    extern __shared__ __align__(1024) unsigned char uvb_smem[];
    unsigned char* sWt = uvb_smem + UvBwdSmem::WT;
    float* sCol = reinterpret_cast<float*>(uvb_smem + UvBwdSmem::COLSUM);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(uvb_smem + UvBwdSmem::BAR);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(uvb_smem + UvBwdSmem::TMEM);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, s = gt >> 5, lane = tid & 31;
    unsigned char* sDK = uvb_smem + UvBwdSmem::GRP + g * UvBwdSmem::GRP_BYTES;
    unsigned char* sAM = sDK + UVMLP_MAT_BYTES;

    for (int q = tid; q < 2048; q += UVBL_THREADS) {          // W^T rows [in][out] -> K-major chunks (row n, k-chunk kc)
        const int n = q >> 4, kc = q & 15;
        *reinterpret_cast<uint4*>(sWt + kc * 2048 + n * 16) = reinterpret_cast<const uint4*>(Wt)[n * 16 + kc];
    }
    if (tid < 128) sCol[tid] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < UVBL_GROUPS; ++i) mbar_init(&sBar[i], 1);
        mbar_fence_init();
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    uv_fence_async_smem();
    uv_tc_fence_before();
    __syncthreads();
    uv_tc_fence_after();
    const uint32_t tmem_cta = *sTmem;
    const uint32_t tmem_d1 = tmem_cta + (uint32_t)(g * 256), tmem_gw = tmem_d1 + 128u;
    const uint32_t lane_off = (uint32_t)(s * 32) << 16;
    const uint32_t idesc_k = uv_instr_desc(128, 128);                                   // both operands K-major
    const uint32_t idesc_mn = uv_instr_desc(128, 128) | (1u << 15) | (1u << 16);        // both operands MN-major
    const uint32_t aDK = smem_u32(sDK), aAM = smem_u32(sAM), aWt = smem_u32(sWt);
    const int mn_off = (gt >> 3) * 2048 + (gt & 7) * 16;                                // this point's row inside an MN-major tile
    const int ntiles = (N + 127) / 128;
    uint32_t phase = 0;
    bool first = true;

    for (int tile = blockIdx.x * UVBL_GROUPS + g; tile < ntiles; tile += gridDim.x * UVBL_GROUPS) {
        const int pt = tile * 128 + gt;
        const bool valid = pt < N;
        {   // rows of delta and a_prev -> the two operand tiles (rows beyond N are zero: they add nothing). Coalesced: an
            // instruction of a warp reads two whole 256-byte rows; the 16-byte chunks are scattered into the UMMA layouts
            // (16-way bank conflicts on these stores, 0.3 us per tile — the thread-per-row loads they replace cost ten times that)
            const int chunk = lane & 15;
            uint4 d[16];
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int rl = s * 32 + 2 * it + (lane >> 4), row = tile * 128 + rl;
                d[it] = (row < N) ? __ldg(reinterpret_cast<const uint4*>(delta_in + (size_t)row * 128) + chunk) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int rl = s * 32 + 2 * it + (lane >> 4);
                *reinterpret_cast<uint4*>(sDK + chunk * 2048 + rl * 16) = d[it];
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int rl = s * 32 + 2 * it + (lane >> 4), row = tile * 128 + rl;
                d[it] = (row < N) ? __ldg(reinterpret_cast<const uint4*>(a_prev + (size_t)row * 128) + chunk) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int rl = s * 32 + 2 * it + (lane >> 4);
                *reinterpret_cast<uint4*>(sAM + (rl >> 3) * 2048 + chunk * 128 + (rl & 7) * 16) = d[it];
            }
        }
        uv_fence_async_smem();
        uv_group_bar(g);
        if (gt == 0) {
            uv_tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)        // delta @ W: K = 128 features in steps of 16
                uv_umma(tmem_d1, uv_smem_desc(aDK + k * 4096, 2048u, 128u), uv_smem_desc(aWt + k * 4096, 2048u, 128u), idesc_k, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)        // delta^T @ a_prev: K = 128 points in steps of 16 (two 8-point core-matrix rows)
                uv_umma(tmem_gw, uv_smem_desc(aDK + k * 256, 128u, 2048u), uv_smem_desc(aAM + k * 4096, 2048u, 128u), idesc_mn,
                        (first && k == 0) ? 0u : 1u);
            uv_umma_commit(&sBar[g]);
        }
        first = false;
        __syncwarp();
        uv_mbar_wait(&sBar[g], phase);
        phase ^= 1u;
        uv_tc_fence_after();
        // epilogue: this point's row of delta @ W, masked by a_prev > 0, to fp16; column sums of what is written
        uint32_t r[2][32];
        UV_TMEM_LD32(r[0], tmem_d1 + lane_off);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            uv_tmem_wait_ld();
            if (cc < 3) UV_TMEM_LD32(r[(cc + 1) & 1], tmem_d1 + lane_off + (uint32_t)((cc + 1) * 32));
            float csum[32];
            const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 av = *reinterpret_cast<const uint4*>(sAM + mn_off + (cc * 4 + q) * 128);
                const uint32_t* aw = reinterpret_cast<const uint32_t*>(&av);
                uint32_t h[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h[k] = uv_pack_half2(__uint_as_float(r[cc & 1][8 * q + 2 * k]), __uint_as_float(r[cc & 1][8 * q + 2 * k + 1]));
                    h[k] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&aw[k]), zero2);
                    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&h[k]));
                    csum[8 * q + 2 * k] = valid ? t.x : 0.f; csum[8 * q + 2 * k + 1] = valid ? t.y : 0.f;
                }
                // staged in the delta tile (its MMAs are complete) for a coalesced copy-out below; own row: conflict-free
                *reinterpret_cast<uint4*>(sDK + (cc * 4 + q) * 2048 + gt * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            }
            const float tot = uv_warp_transpose_reduce32(csum, lane);
            atomicAdd(&sCol[cc * 32 + lane], tot);
        }
        uv_tc_fence_before();
        uv_group_bar(g);
        {   // the new delta rows leave two whole rows per instruction
            const int chunk = lane & 15;
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int rl = s * 32 + 2 * it + (lane >> 4), row = tile * 128 + rl;
                const uint4 v = *reinterpret_cast<const uint4*>(sDK + chunk * 2048 + rl * 16);
                if (row < N) reinterpret_cast<uint4*>(delta_out + (size_t)row * 128)[chunk] = v;
            }
        }
        uv_group_bar(g);            // the tile's operand buffers are free again (all 16 MMAs completed before the epilogue)
    }

    // the group's weight-gradient accumulator: TMEM lane = out feature, 128 columns = in features
    if (!first) {
        uv_tc_fence_after();
        // 148 CTAs x 2 groups add into the same 64 KB: 128-bit vector reductions (a quarter of the L2 operations of scalar
        // ones) and a per-CTA rotation of the column order, so that the CTAs — which all arrive here at about the same time —
        // are not queueing on the same addresses
        const int m = gt;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
            const int cc = (c4 + (int)blockIdx.x + g) & 3;
            uint32_t r[32];
            UV_TMEM_LD32(r, tmem_gw + lane_off + (uint32_t)(cc * 32));
            uv_tmem_wait_ld();
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const int j = (i4 + (int)(blockIdx.x >> 2)) & 7;
                red_add_v4(gW + m * 128 + cc * 32 + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
    }
    uv_tc_fence_before();
    __syncthreads();
    if (tid < 128) atomicAdd(colsum + tid, sCol[tid]);
    if (tid < 32) {
        uv_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_cta), "r"(512u) : "memory");
    }
}
