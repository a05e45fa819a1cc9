// texgs_loss.cuh — SURVEY §8f N3: the photometric loss of the reference's training step fused into two
// kernels. Reference (pure PyTorch, pinned by golden vectors generated from it):
//   losses/pixelwise_loss.py:3-4   l1_loss   = mean |x - y|
//   losses/ssim_loss.py:6-54       ssim_loss = mean SSIM, 11x11 Gaussian window (sigma 1.5), zero padding 5
//   models/texture_gaussian3d.py:333-340   loss = (1-l)*L1 + l*(1 - SSIM)
// The reference spends 5 depthwise convolutions + ~10 pointwise kernels forward and their autograd
// twins backward; here one kernel produces both means plus three derivative maps, one kernel turns them
// into d loss / d image (fused-SSIM formulation: separable window, tile + halo in shared memory).
#pragma once
#include "texgs_common.cuh"

#define TEXGS_LOSS_TILE 16                                     // CTA = 16 x 16 threads
#define TEXGS_LOSS_TILE_H 32                                   // ... for a 16 x 32 pixel tile: two vertically adjacent outputs per thread
#define TEXGS_LOSS_R 5
#define TEXGS_LOSS_SPAN (TEXGS_LOSS_TILE + 2 * TEXGS_LOSS_R)     // 26 halo columns
#define TEXGS_LOSS_SPAN_H (TEXGS_LOSS_TILE_H + 2 * TEXGS_LOSS_R) // 42 halo rows
// Shared-memory row strides (words). A warp = two half-warps of 16 consecutive columns; in both passes the halves work on rows
// r and r+2 (the horizontal pass by its item order, the vertical pass because a thread owns rows 2*ly and 2*ly+1), so
// 2 * stride = 16 (mod 32) keeps them on disjoint banks: 40 for the 26-wide halo tiles, 24 for the 16-wide row-filtered ones.
// (Round 2 measured the 2-way conflict a stride of 27 gave: forward 0.158 -> 0.148 ms, backward 0.138 -> 0.107 ms at 1080p.)
#define TEXGS_LOSS_STRIDE 40
#define TEXGS_LOSS_HSTRIDE 24
#define TEXGS_SSIM_C1 0.0001f    // 0.01^2
#define TEXGS_SSIM_C2 0.0009f    // 0.03^2

// gaussian(11, 1.5) normalised in fp32 exactly as losses/ssim_loss.py:6-8 does
__device__ __constant__ float TEXGS_WIN[11] = {0.00102838012f, 0.00759875821f, 0.0360007733f, 0.109360687f, 0.213005528f,
                                               0.266011715f,   0.213005528f,   0.109360687f,  0.0360007733f, 0.00759875821f,
                                               0.00102838012f};

struct LossSums { double ssim, l1; };

// item -> (row, column) of the horizontal pass: the two half-warps of a warp take rows r and r + 2
__device__ __forceinline__ void loss_hpass_item(int i, int& r, int& c) {
    const int q = i >> 5;
    c = i & 15;
    r = ((q >> 1) << 2) + (q & 1) + ((i >> 3) & 2);
}
#define TEXGS_LOSS_HITEMS (((TEXGS_LOSS_SPAN_H + 3) / 4) * 4 * TEXGS_LOSS_TILE)     // rows rounded up to whole groups of 4

// grid (tiles_x, tiles_y, C), block 16x16
__global__ void __launch_bounds__(256) texgs_photometric_fwd_kernel(const float* __restrict__ img, const float* __restrict__ gt,
                                                                   int H, int W, float* __restrict__ d_mu1, float* __restrict__ d_e11,
                                                                   float* __restrict__ d_e12, LossSums* __restrict__ sums) {
    __shared__ float sa[TEXGS_LOSS_SPAN_H][TEXGS_LOSS_STRIDE], sb[TEXGS_LOSS_SPAN_H][TEXGS_LOSS_STRIDE];
    __shared__ float hq[5][TEXGS_LOSS_SPAN_H][TEXGS_LOSS_HSTRIDE];
    __shared__ double red[2][8];
    const int tid = threadIdx.y * TEXGS_LOSS_TILE + threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TEXGS_LOSS_TILE - TEXGS_LOSS_R, y0 = blockIdx.y * TEXGS_LOSS_TILE_H - TEXGS_LOSS_R;
    for (int i = tid; i < TEXGS_LOSS_SPAN_H * TEXGS_LOSS_SPAN; i += 256) {
        const int r = i / TEXGS_LOSS_SPAN, c = i % TEXGS_LOSS_SPAN;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;          // zero padding (conv2d padding=5)
        sa[r][c] = in ? img[plane + (size_t)gy * W + gx] : 0.f;
        sb[r][c] = in ? gt[plane + (size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < TEXGS_LOSS_HITEMS; i += 256) {                 // horizontal pass
        int r, c;
        loss_hpass_item(i, r, c);
        if (r >= TEXGS_LOSS_SPAN_H) continue;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = TEXGS_WIN[k], a = sa[r][c + k], b = sb[r][c + k];
            s0 += w * a; s1 += w * b; s2 += w * a * a; s3 += w * b * b; s4 += w * a * b;
        }
        hq[0][r][c] = s0; hq[1][r][c] = s1; hq[2][r][c] = s2; hq[3][r][c] = s3; hq[4][r][c] = s4;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = 2 * threadIdx.y;                     // this thread: pixels (ly, lx) and (ly + 1, lx) of the tile
    const int px = blockIdx.x * TEXGS_LOSS_TILE + lx, py = blockIdx.y * TEXGS_LOSS_TILE_H + ly;
    double my_ssim = 0.0, my_l1 = 0.0;
    if (px < W && py < H) {
        float m1[2] = {0.f, 0.f}, m2[2] = {0.f, 0.f}, q11[2] = {0.f, 0.f}, q22[2] = {0.f, 0.f}, q12[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 12; ++k) {                                    // vertical pass: 12 rows serve both outputs
            const float v0 = hq[0][ly + k][lx], v1 = hq[1][ly + k][lx], v2 = hq[2][ly + k][lx], v3 = hq[3][ly + k][lx], v4 = hq[4][ly + k][lx];
            if (k < 11) {
                const float w = TEXGS_WIN[k];
                m1[0] += w * v0; m2[0] += w * v1; q11[0] += w * v2; q22[0] += w * v3; q12[0] += w * v4;
            }
            if (k > 0) {
                const float w = TEXGS_WIN[k - 1];
                m1[1] += w * v0; m2[1] += w * v1; q11[1] += w * v2; q22[1] += w * v3; q12[1] += w * v4;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (py + j >= H) break;
            const float mu1 = m1[j], mu2 = m2[j], e11 = q11[j], e22 = q22[j], e12 = q12[j];
            const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
            const float A = 2.f * mu1 * mu2 + TEXGS_SSIM_C1, B = 2.f * s12 + TEXGS_SSIM_C2;
            const float Cc = mu1 * mu1 + mu2 * mu2 + TEXGS_SSIM_C1, D = s11 + s22 + TEXGS_SSIM_C2;
            const float iCD = 1.0f / (Cc * D);
            my_ssim += (double)(A * B * iCD);
            // d ssim / d (mu1, E[x^2], E[xy]) with x = image (gt is a constant)
            const float dE11 = -A * B * iCD / D;
            const float dE12 = 2.f * A * iCD;
            const float dmu1 = 2.f * mu2 * (B - A) * iCD - 2.f * mu1 * A * B * iCD / Cc - 2.f * mu1 * dE11;
            const size_t o = plane + (size_t)(py + j) * W + px;
            d_mu1[o] = dmu1; d_e11[o] = dE11; d_e12[o] = dE12;
            my_l1 += (double)fabsf(sa[ly + j + TEXGS_LOSS_R][lx + TEXGS_LOSS_R] - sb[ly + j + TEXGS_LOSS_R][lx + TEXGS_LOSS_R]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_ssim += __shfl_xor_sync(0xffffffffu, my_ssim, o);
        my_l1 += __shfl_xor_sync(0xffffffffu, my_l1, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = my_ssim; red[1][tid >> 5] = my_l1; }
    __syncthreads();
    if (tid == 0) {   // one partial per CTA, summed in a fixed order by the finalize kernel (deterministic)
        double s = 0.0, l = 0.0;
        for (int w = 0; w < 8; ++w) { s += red[0][w]; l += red[1][w]; }
        const size_t bid = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        sums[bid].ssim = s;
        sums[bid].l1 = l;
    }
}

__global__ void __launch_bounds__(1024) texgs_photometric_finalize_kernel(const LossSums* __restrict__ sums, int nparts, double inv_n,
                                                                        float lambda, float* __restrict__ out3) {
    __shared__ double red[2][32];
    double s = 0.0, l = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { s += sums[i].ssim; l += sums[i].l1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); l += __shfl_xor_sync(0xffffffffu, l, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tl = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += red[0][w]; tl += red[1][w]; }
        const float l1 = (float)(tl * inv_n);
        const float lssim = 1.0f - (float)(ts * inv_n);
        out3[0] = (1.0f - lambda) * l1 + lambda * lssim;
        out3[1] = l1;
        out3[2] = lssim;
    }
}

// d image = coef[0] * d(L1 mean)/d image + coef[1] * d(1 - SSIM mean)/d image
__global__ void __launch_bounds__(256) texgs_photometric_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
                                                                   const float* __restrict__ d_mu1, const float* __restrict__ d_e11,
                                                                   const float* __restrict__ d_e12, const float* __restrict__ coef,
                                                                   float inv_n, float* __restrict__ dimg) {
    __shared__ float sm[3][TEXGS_LOSS_SPAN_H][TEXGS_LOSS_STRIDE];
    __shared__ float hq[3][TEXGS_LOSS_SPAN_H][TEXGS_LOSS_HSTRIDE];
    const int tid = threadIdx.y * TEXGS_LOSS_TILE + threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TEXGS_LOSS_TILE - TEXGS_LOSS_R, y0 = blockIdx.y * TEXGS_LOSS_TILE_H - TEXGS_LOSS_R;
    for (int i = tid; i < TEXGS_LOSS_SPAN_H * TEXGS_LOSS_SPAN; i += 256) {
        const int r = i / TEXGS_LOSS_SPAN, c = i % TEXGS_LOSS_SPAN;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;          // derivative maps exist on image pixels only
        const size_t o = plane + (size_t)gy * W + gx;
        sm[0][r][c] = in ? d_mu1[o] : 0.f;
        sm[1][r][c] = in ? d_e11[o] : 0.f;
        sm[2][r][c] = in ? d_e12[o] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < TEXGS_LOSS_HITEMS; i += 256) {
        int r, c;
        loss_hpass_item(i, r, c);
        if (r >= TEXGS_LOSS_SPAN_H) continue;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = TEXGS_WIN[k];
            s0 += w * sm[0][r][c + k]; s1 += w * sm[1][r][c + k]; s2 += w * sm[2][r][c + k];
        }
        hq[0][r][c] = s0; hq[1][r][c] = s1; hq[2][r][c] = s2;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = 2 * threadIdx.y;
    const int px = blockIdx.x * TEXGS_LOSS_TILE + lx, py = blockIdx.y * TEXGS_LOSS_TILE_H + ly;
    if (px < W && py < H) {
        float c0[2] = {0.f, 0.f}, c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float v0 = hq[0][ly + k][lx], v1 = hq[1][ly + k][lx], v2 = hq[2][ly + k][lx];
            if (k < 11) { const float w = TEXGS_WIN[k]; c0[0] += w * v0; c1[0] += w * v1; c2[0] += w * v2; }
            if (k > 0) { const float w = TEXGS_WIN[k - 1]; c0[1] += w * v0; c1[1] += w * v1; c2[1] += w * v2; }
        }
        const float k0 = coef[0], k1 = coef[1];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (py + j >= H) break;
            const size_t o = plane + (size_t)(py + j) * W + px;
            const float a = img[o], b = gt[o];
            const float dssim = c0[j] + 2.f * a * c1[j] + b * c2[j];              // d(sum ssim)/d a
            const float d = a - b;
            const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);          // torch.abs: gradient 0 at 0
            dimg[o] = inv_n * (k0 * sgn - k1 * dssim);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Geometry losses of the same training step (models/texture_gaussian3d.py:342-368), one kernel each way:
//   Lalpha = l1_loss(alpha, gt_alpha)                      losses/pixelwise_loss.py:3-4
//   Lnorm  = norm_loss(norm, gt_norm, gt_alpha)            losses/norm_reg_loss.py:66-71 (masked branch)
//          = sum((1 - <norm, gt_norm>) * mask) / (sum(mask) + 1e-6)
//   Lnsm   = smooth_loss(gt_image, norm, gt_alpha)         losses/smooth_loss.py:4-27
//          = 1/4 * sum_d  sum|w_d * (v_a - v_b)| / (sum(w_d) + 1e-6),  four neighbour directions d
//            (right, down, down-right, up-right), w_d = exp(-sum_c|rgb_a - rgb_b| / gamma) * mask_a * mask_b
// The reference spends ~45 pointwise/reduction kernels forward on these and as many backward.
// Pair (a, b) of direction d anchored at pixel (y, x):
//   d=0: a=(y,x)   b=(y,x+1)      d=1: a=(y,x)   b=(y+1,x)
//   d=2: a=(y,x)   b=(y+1,x+1)    d=3: a=(y+1,x) b=(y,x+1)
#define TEXGS_GEO_TX 32
#define TEXGS_GEO_TY 8
#define TEXGS_GEO_NSUM 11      // alpha L1, norm num, mask sum, 4 x (w sum), 4 x (w |dv| sum)

struct GeoIn {
    const float* alpha;      // (1,H,W)
    const float* norm;       // (3,H,W)
    const float* gt_alpha;   // (1,H,W) or NULL (= ones, models/texture_gaussian3d.py:330)
    const float* gt_norm;    // (3,H,W) or NULL (Lnorm skipped)
    const float* gt_image;   // (3,H,W) or NULL (Lnsm skipped)
    int H, W;
    float inv_gamma;
};

// ws layout: [0..5] float scales {1/(HW), 1/(sum mask+eps), 1/(sum w_d + eps) x4}, then per-CTA partials
struct GeoTile {
    float rgb[3][TEXGS_GEO_TY + 2][TEXGS_GEO_TX + 2];
    float nrm[3][TEXGS_GEO_TY + 2][TEXGS_GEO_TX + 2];
    float msk[TEXGS_GEO_TY + 2][TEXGS_GEO_TX + 2];       // 0 outside the image
};

__device__ __forceinline__ void geo_load_tile(const GeoIn& g, GeoTile& t, int tid) {
    const int x0 = blockIdx.x * TEXGS_GEO_TX - 1, y0 = blockIdx.y * TEXGS_GEO_TY - 1;
    const size_t plane = (size_t)g.H * g.W;
    for (int i = tid; i < (TEXGS_GEO_TY + 2) * (TEXGS_GEO_TX + 2); i += TEXGS_GEO_TX * TEXGS_GEO_TY) {
        const int r = i / (TEXGS_GEO_TX + 2), c = i % (TEXGS_GEO_TX + 2);
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;
        const size_t o = (size_t)(in ? gy : 0) * g.W + (in ? gx : 0);
        t.msk[r][c] = in ? (g.gt_alpha ? g.gt_alpha[o] : 1.f) : 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            t.rgb[ch][r][c] = (in && g.gt_image) ? g.gt_image[ch * plane + o] : 0.f;
            t.nrm[ch][r][c] = in ? g.norm[ch * plane + o] : 0.f;
        }
    }
}

// bilateral weight of the pair (ra,ca)-(rb,cb) in tile coordinates; `valid` = both pixels inside the image
__device__ __forceinline__ float geo_weight(const GeoTile& t, int ra, int ca, int rb, int cb, bool valid, float inv_gamma) {
    if (!valid) return 0.f;
    const float d = fabsf(t.rgb[0][ra][ca] - t.rgb[0][rb][cb]) + fabsf(t.rgb[1][ra][ca] - t.rgb[1][rb][cb]) +
                    fabsf(t.rgb[2][ra][ca] - t.rgb[2][rb][cb]);
    return expf(-d * inv_gamma) * t.msk[ra][ca] * t.msk[rb][cb];
}

__global__ void __launch_bounds__(TEXGS_GEO_TX * TEXGS_GEO_TY) texgs_geometry_loss_fwd_kernel(const GeoIn g, double* __restrict__ parts) {
    __shared__ GeoTile t;
    __shared__ double red[TEXGS_GEO_NSUM][TEXGS_GEO_TX * TEXGS_GEO_TY / 32];
    const int tid = threadIdx.y * TEXGS_GEO_TX + threadIdx.x;
    geo_load_tile(g, t, tid);
    __syncthreads();
    const int px = blockIdx.x * TEXGS_GEO_TX + threadIdx.x, py = blockIdx.y * TEXGS_GEO_TY + threadIdx.y;
    const int r = threadIdx.y + 1, c = threadIdx.x + 1;
    double s[TEXGS_GEO_NSUM];
#pragma unroll
    for (int k = 0; k < TEXGS_GEO_NSUM; ++k) s[k] = 0.0;
    if (px < g.W && py < g.H) {
        const size_t o = (size_t)py * g.W + px, plane = (size_t)g.H * g.W;
        const float m = t.msk[r][c];
        s[0] = (double)fabsf(g.alpha[o] - m);
        if (g.gt_norm) {
            const float dot = t.nrm[0][r][c] * g.gt_norm[o] + t.nrm[1][r][c] * g.gt_norm[plane + o] + t.nrm[2][r][c] * g.gt_norm[2 * plane + o];
            s[1] = (double)((1.0f - dot) * m);
        }
        s[2] = (double)m;
        if (g.gt_image) {
            const bool hr = px + 1 < g.W, hd = py + 1 < g.H;
            const int ra[4] = {r, r, r, r + 1}, ca[4] = {c, c, c, c};
            const int rb[4] = {r, r + 1, r + 1, r}, cb[4] = {c + 1, c, c + 1, c + 1};
            const bool ok[4] = {hr, hd, hr && hd, hr && hd};
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float w = geo_weight(t, ra[d], ca[d], rb[d], cb[d], ok[d], g.inv_gamma);
                float a = 0.f;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) a += fabsf(w * (t.nrm[ch][ra[d]][ca[d]] - t.nrm[ch][rb[d]][cb[d]]));
                s[3 + d] = (double)w;
                s[7 + d] = (double)a;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < TEXGS_GEO_NSUM; ++k) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[k][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < TEXGS_GEO_NSUM) {
        double v = 0.0;
        for (int w = 0; w < TEXGS_GEO_TX * TEXGS_GEO_TY / 32; ++w) v += red[tid][w];
        parts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TEXGS_GEO_NSUM + tid] = v;
    }
}

// out3 = {Lalpha, Lnorm, Lnsm}; scales[6] feed the backward kernel
__global__ void __launch_bounds__(1024) texgs_geometry_loss_finalize_kernel(const double* __restrict__ parts, int nparts, double inv_hw,
                                                                          float* __restrict__ scales, float* __restrict__ out3) {
    __shared__ double red[TEXGS_GEO_NSUM][32];
    __shared__ double tot[TEXGS_GEO_NSUM];
    double s[TEXGS_GEO_NSUM];
#pragma unroll
    for (int k = 0; k < TEXGS_GEO_NSUM; ++k) s[k] = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < TEXGS_GEO_NSUM; ++k) s[k] += parts[(size_t)i * TEXGS_GEO_NSUM + k];
#pragma unroll
    for (int k = 0; k < TEXGS_GEO_NSUM; ++k) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < TEXGS_GEO_NSUM) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // fp32 arithmetic on the reduced sums, as the reference's torch.sum(...)/(torch.sum(...) + 1e-6)
        const float inv_m = 1.0f / ((float)tot[2] + 1e-6f);
        out3[0] = (float)(tot[0] * inv_hw);
        out3[1] = (float)tot[1] * inv_m;
        float nsm = 0.f;
        scales[0] = (float)inv_hw;
        scales[1] = inv_m;
        for (int d = 0; d < 4; ++d) {
            const float inv_w = 1.0f / ((float)tot[3 + d] + 1e-6f);
            scales[2 + d] = inv_w;
            nsm += (float)tot[7 + d] * inv_w;
        }
        out3[2] = nsm * 0.25f;
    }
}

__device__ __forceinline__ float geo_sign(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// dL_dalpha = coef[0] * dLalpha/dalpha ;  dL_dnorm = coef[1] * dLnorm/dnorm + coef[2] * dLnsm/dnorm
__global__ void __launch_bounds__(TEXGS_GEO_TX * TEXGS_GEO_TY) texgs_geometry_loss_bwd_kernel(const GeoIn g, const float* __restrict__ scales,
                                                                                             const float* __restrict__ coef,
                                                                                             float* __restrict__ dalpha, float* __restrict__ dnorm) {
    __shared__ GeoTile t;
    const int tid = threadIdx.y * TEXGS_GEO_TX + threadIdx.x;
    geo_load_tile(g, t, tid);
    __syncthreads();
    const int px = blockIdx.x * TEXGS_GEO_TX + threadIdx.x, py = blockIdx.y * TEXGS_GEO_TY + threadIdx.y;
    if (px >= g.W || py >= g.H) return;
    const int r = threadIdx.y + 1, c = threadIdx.x + 1;
    const size_t o = (size_t)py * g.W + px, plane = (size_t)g.H * g.W;
    const float m = t.msk[r][c];
    if (dalpha) dalpha[o] = coef[0] * scales[0] * geo_sign(g.alpha[o] - m);
    if (!dnorm) return;
    float gn[3] = {0.f, 0.f, 0.f};
    if (g.gt_norm) {
        const float k = -coef[1] * scales[1] * m;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) gn[ch] = k * g.gt_norm[ch * plane + o];
    }
    if (g.gt_image) {
        const bool hl = px > 0, hr = px + 1 < g.W, hu = py > 0, hd = py + 1 < g.H;
        // the 8 pairs this pixel belongs to: {other row, other col, direction, pair valid}; as `a` the
        // term is +w*sign(v_p - v_q), as `b` it is -w*sign(v_q - v_p) — the same expression
        const int qr[8] = {r, r + 1, r + 1, r - 1, r, r - 1, r - 1, r + 1};
        const int qc[8] = {c + 1, c, c + 1, c + 1, c - 1, c, c - 1, c - 1};
        const int dd[8] = {0, 1, 2, 3, 0, 1, 2, 3};
        const bool ok[8] = {hr, hd, hr && hd, hu && hr, hl, hu, hu && hl, hd && hl};
        const float k = 0.25f * coef[2];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float w = geo_weight(t, r, c, qr[e], qc[e], ok[e], g.inv_gamma) * scales[2 + dd[e]] * k;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) gn[ch] += w * geo_sign(t.nrm[ch][r][c] - t.nrm[ch][qr[e]][qc[e]]);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) dnorm[ch * plane + o] = gn[ch];
}
