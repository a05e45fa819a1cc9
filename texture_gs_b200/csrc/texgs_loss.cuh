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

#define TEXGS_LOSS_TILE 16
#define TEXGS_LOSS_R 5
#define TEXGS_LOSS_SPAN (TEXGS_LOSS_TILE + 2 * TEXGS_LOSS_R)   // 26
#define TEXGS_SSIM_C1 0.0001f    // 0.01^2
#define TEXGS_SSIM_C2 0.0009f    // 0.03^2

// gaussian(11, 1.5) normalised in fp32 exactly as losses/ssim_loss.py:6-8 does
__device__ __constant__ float TEXGS_WIN[11] = {0.00102838012f, 0.00759875821f, 0.0360007733f, 0.109360687f, 0.213005528f,
                                               0.266011715f,   0.213005528f,   0.109360687f,  0.0360007733f, 0.00759875821f,
                                               0.00102838012f};

struct LossSums { double ssim, l1; };

// grid (tiles_x, tiles_y, C), block 16x16
__global__ void __launch_bounds__(256) texgs_photometric_fwd_kernel(const float* __restrict__ img, const float* __restrict__ gt,
                                                                   int H, int W, float* __restrict__ d_mu1, float* __restrict__ d_e11,
                                                                   float* __restrict__ d_e12, LossSums* __restrict__ sums) {
    __shared__ float sa[TEXGS_LOSS_SPAN][TEXGS_LOSS_SPAN + 1], sb[TEXGS_LOSS_SPAN][TEXGS_LOSS_SPAN + 1];
    __shared__ float hq[5][TEXGS_LOSS_SPAN][TEXGS_LOSS_TILE];
    __shared__ double red[2][8];
    const int tid = threadIdx.y * TEXGS_LOSS_TILE + threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TEXGS_LOSS_TILE - TEXGS_LOSS_R, y0 = blockIdx.y * TEXGS_LOSS_TILE - TEXGS_LOSS_R;
    for (int i = tid; i < TEXGS_LOSS_SPAN * TEXGS_LOSS_SPAN; i += 256) {
        const int r = i / TEXGS_LOSS_SPAN, c = i % TEXGS_LOSS_SPAN;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;          // zero padding (conv2d padding=5)
        sa[r][c] = in ? img[plane + (size_t)gy * W + gx] : 0.f;
        sb[r][c] = in ? gt[plane + (size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < TEXGS_LOSS_SPAN * TEXGS_LOSS_TILE; i += 256) {   // horizontal pass
        const int r = i / TEXGS_LOSS_TILE, c = i % TEXGS_LOSS_TILE;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = TEXGS_WIN[k], a = sa[r][c + k], b = sb[r][c + k];
            s0 += w * a; s1 += w * b; s2 += w * a * a; s3 += w * b * b; s4 += w * a * b;
        }
        hq[0][r][c] = s0; hq[1][r][c] = s1; hq[2][r][c] = s2; hq[3][r][c] = s3; hq[4][r][c] = s4;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int px = blockIdx.x * TEXGS_LOSS_TILE + lx, py = blockIdx.y * TEXGS_LOSS_TILE + ly;
    double my_ssim = 0.0, my_l1 = 0.0;
    if (px < W && py < H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {                                        // vertical pass
            const float w = TEXGS_WIN[k];
            mu1 += w * hq[0][ly + k][lx]; mu2 += w * hq[1][ly + k][lx];
            e11 += w * hq[2][ly + k][lx]; e22 += w * hq[3][ly + k][lx]; e12 += w * hq[4][ly + k][lx];
        }
        const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A = 2.f * mu1 * mu2 + TEXGS_SSIM_C1, B = 2.f * s12 + TEXGS_SSIM_C2;
        const float Cc = mu1 * mu1 + mu2 * mu2 + TEXGS_SSIM_C1, D = s11 + s22 + TEXGS_SSIM_C2;
        const float iCD = 1.0f / (Cc * D);
        my_ssim = (double)(A * B * iCD);
        // d ssim / d (mu1, E[x^2], E[xy]) with x = image (gt is a constant)
        const float dE11 = -A * B * iCD / D;
        const float dE12 = 2.f * A * iCD;
        const float dmu1 = 2.f * mu2 * (B - A) * iCD - 2.f * mu1 * A * B * iCD / Cc - 2.f * mu1 * dE11;
        const size_t o = plane + (size_t)py * W + px;
        d_mu1[o] = dmu1; d_e11[o] = dE11; d_e12[o] = dE12;
        my_l1 = (double)fabsf(sa[ly + TEXGS_LOSS_R][lx + TEXGS_LOSS_R] - sb[ly + TEXGS_LOSS_R][lx + TEXGS_LOSS_R]);
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
    __shared__ float sm[3][TEXGS_LOSS_SPAN][TEXGS_LOSS_SPAN + 1];
    __shared__ float hq[3][TEXGS_LOSS_SPAN][TEXGS_LOSS_TILE];
    const int tid = threadIdx.y * TEXGS_LOSS_TILE + threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TEXGS_LOSS_TILE - TEXGS_LOSS_R, y0 = blockIdx.y * TEXGS_LOSS_TILE - TEXGS_LOSS_R;
    for (int i = tid; i < TEXGS_LOSS_SPAN * TEXGS_LOSS_SPAN; i += 256) {
        const int r = i / TEXGS_LOSS_SPAN, c = i % TEXGS_LOSS_SPAN;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;          // derivative maps exist on image pixels only
        const size_t o = plane + (size_t)gy * W + gx;
        sm[0][r][c] = in ? d_mu1[o] : 0.f;
        sm[1][r][c] = in ? d_e11[o] : 0.f;
        sm[2][r][c] = in ? d_e12[o] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < TEXGS_LOSS_SPAN * TEXGS_LOSS_TILE; i += 256) {
        const int r = i / TEXGS_LOSS_TILE, c = i % TEXGS_LOSS_TILE;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = TEXGS_WIN[k];
            s0 += w * sm[0][r][c + k]; s1 += w * sm[1][r][c + k]; s2 += w * sm[2][r][c + k];
        }
        hq[0][r][c] = s0; hq[1][r][c] = s1; hq[2][r][c] = s2;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int px = blockIdx.x * TEXGS_LOSS_TILE + lx, py = blockIdx.y * TEXGS_LOSS_TILE + ly;
    if (px < W && py < H) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = TEXGS_WIN[k];
            c0 += w * hq[0][ly + k][lx]; c1 += w * hq[1][ly + k][lx]; c2 += w * hq[2][ly + k][lx];
        }
        const size_t o = plane + (size_t)py * W + px;
        const float a = img[o], b = gt[o];
        const float dssim = c0 + 2.f * a * c1 + b * c2;                       // d(sum ssim)/d a
        const float d = a - b;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);          // torch.abs: gradient 0 at 0
        dimg[o] = inv_n * (coef[0] * sgn - coef[1] * dssim);
    }
}
