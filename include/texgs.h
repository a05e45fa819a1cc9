/*
 * texgs.h — C-ABI of the B200-native Texture-GS rasterizer (libtexgs.so).
 *
 * This is the drop-in boundary for the one hot path of slothfulxtx/Texture-GS: what the reference
 * reaches through the (un-vendored) pybind module ``diff_gauss_uv_tex._C``
 *   - ``_C.rasterize_gaussians``           <- called by GaussianRasterizer.forward, which the
 *                                              reference invokes at render/uv_tex_render.py:56-66
 *   - ``_C.rasterize_gaussians_backward``  <- autograd backward of the same call
 *                                              (loss.backward(), models/texture_gaussian3d.py:410)
 * and through ``diff_gauss._C`` for the texture-less variant (render/render.py:75-84).
 *
 * Conventions
 *   * every pointer named ``d_*`` or living in an args struct is a DEVICE pointer to contiguous
 *     fp32 / int32 memory owned by the caller (PyTorch); the library never allocates persistent
 *     memory and keeps no global mutable state (re-entrant; last error is thread-local);
 *   * ``stream`` is a ``cudaStream_t`` passed as ``void*``; all work is enqueued on it;
 *   * return value 0 = success, otherwise a non-zero code (cudaError_t value or TEXGS_E_*);
 *     ``texgs_last_error()`` returns the message for the calling thread;
 *   * no host synchronisation happens inside any call unless ``TEXGS_FLAG_DEBUG`` is set
 *     (mirrors ``raster_settings.debug``, render/uv_tex_render.py:37).
 */
#ifndef TEXGS_H_
#define TEXGS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEXGS_ABI_VERSION 4

#define TEXGS_E_INVALID   1001   /* bad argument */
#define TEXGS_E_WORKSPACE 1002   /* workspace too small */

#define TEXGS_FLAG_PREFILTERED 1u   /* GaussianRasterizationSettings.prefiltered (uv_tex_render.py:36) */
#define TEXGS_FLAG_DEBUG       2u   /* GaussianRasterizationSettings.debug       (uv_tex_render.py:37) */
/* Spec switches for the conventions the reference tree does not pin: the
 * defaults are the flags cleared. Set, the textured render takes a cold instantiation of the render kernels that
 * evaluates the alternative; needs the packed texel copy (texture_rgba) and, in the backward, dL_dtexture_rgba.      */
#define TEXGS_FLAG_SEAMLESS_CUBE      4u   /* E11-alt: bilinear taps beyond a face edge come from the adjacent face (GL seamless /
                                              nvdiffrast boundary_mode='cube', models/uv_map_gaussian3d.py:259) instead of clamp-to-edge */
#define TEXGS_FLAG_DEPTH_INTERSECTION 8u   /* E7-alt: the depth output blends z of the ray-disc intersection instead of z of the centre */
#define TEXGS_FLAG_STOPGRAD_DELTA     16u  /* E13-alt: no gradient through the intersection offset Delta (to means3D / rotations) */
#define TEXGS_FLAG_SPEC_MASK          28u  /* the three above: they select the ALT render kernels */
#define TEXGS_FLAG_CLAMP_GRAD_3DGS    32u  /* E2-alt (preprocess backward only): a clamped x/z, y/z of the EWA projection passes no gradient, as the
                                              3DGS lineage does; default = exact derivative of t.x = clamp(x/z) * z                       */

/* colour source of a splat */
#define TEXGS_MODE_TEXTURE 0   /* diff_gauss_uv_tex: C0*cube(uv + J*delta) + SH_rest + 0.5, clamped at 0 */
#define TEXGS_MODE_SH      1   /* diff_gauss:        SH(full, incl. DC) + 0.5, clamped at 0              */
#define TEXGS_MODE_PRECOMP 2   /* diff_gauss:        colors_precomp used as given                        */

/* Problem description = GaussianRasterizationSettings (render/uv_tex_render.py:25-38) + tensor
 * shapes + the kwargs of GaussianRasterizer.__call__ (render/uv_tex_render.py:56-66). */
typedef struct TexgsFwdArgs {
    int32_t P;              /* number of Gaussians                                                  */
    int32_t M;              /* SH coefficients per Gaussian in ``shs`` (row length / 3); 0 if none  */
    int32_t sh_degree;      /* active degree (raster_settings.sh_degree)                            */
    int32_t E;              /* channels of extra_attrs (0 if None)                                  */
    int32_t H, W;           /* image_height, image_width                                            */
    int32_t R;              /* cube face resolution of ``texture`` (6,R,R,3); 0 if no texture       */
    int32_t mode;           /* TEXGS_MODE_*                                                         */
    uint32_t flags;         /* TEXGS_FLAG_*                                                         */
    float tanfovx, tanfovy, scale_modifier;
    float viewmatrix[16];   /* raster_settings.viewmatrix, row-major as torch stores it (= W2C^T)   */
    float projmatrix[16];   /* raster_settings.projmatrix (= viewmatrix @ P^T)                      */
    float campos[3];
    float bg[3];
    /* inputs (device) */
    const float* means3D;        /* (P,3)                                        */
    const float* shs;            /* (P,M,3) or NULL                              */
    const float* colors_precomp; /* (P,3) or NULL (TEXGS_MODE_PRECOMP)           */
    const float* opacities;      /* (P,1)                                        */
    const float* scales;         /* (P,3)                                        */
    const float* rotations;      /* (P,4) quaternion (r,x,y,z)                   */
    const float* uvs;            /* (P,3)   or NULL                              */
    const float* gradient_uvs;   /* (P,9) row-major d uv_i / d x_j, or NULL      */
    const float* texture;        /* (6,R,R,3) or NULL                            */
    const float* extra_attrs;    /* (P,E) or NULL: blended into out_extra (E,H,W) with the weights of the main
                                    render, no background (render/uv_tex_render.py:66, render/render.py:84)    */
    /* diff_gauss only (render/render.py:52-53,83): world-space covariances given instead of scales +
     * rotations, 6 floats per Gaussian in the order xx,xy,xz,yy,yz,zz (utils/general.py:73-82), used as
     * given (scale_modifier is NOT applied: the reference applies it in get_covariance). scales and
     * rotations must then be NULL; the disc normal is the unit eigenvector of the smallest eigenvalue,
     * facing the camera, and carries no gradient. Not available in TEXGS_MODE_TEXTURE. */
    const float* cov3Ds_precomp; /* (P,6) or NULL                                */
    /* optional: the same texture repacked as (6,R,R,4) fp32 (rgb + one pad float) by
     * texgs_pack_texture; when given the render kernels fetch one 128-bit texel per tap instead
     * of three scalars. Must correspond to ``texture``. NULL = read ``texture`` directly. */
    const float* texture_rgba;
    /* optional DUAL render (texture mode only): a second (3,H,W) image blended in the same pass from
     * the colour the splats have with sh_degree = 0, max(0, C0*tex + 0.5) — what the reference obtains
     * with a second full render (models/texture_gaussian3d.py:375-389, 505-511). NULL = off. In
     * texgs_backward the forward args must carry the same pointer state (NULL / non-NULL). */
    float* out_image_nosh;
    /* optional per-kernel timing: HOST array of TEXGS_EV_COUNT cudaEvent_t (as void*), recorded on
     * the stream at the stage boundaries below; NULL = off. Forward fills slots 0..5, backward
     * (through TexgsBwdArgs.fwd) slots 6..9. */
    void* const* profile_events;
    /* optional: a cudaEvent_t (as void*) the stream waits for between binning and the render kernel (forward only). The
     * per-Gaussian stages of a view need neither the texture nor the gradient buffers, so a caller whose previous optimizer
     * step / gradient reduction is still running on another stream lets them start early and orders only the render
     * (and everything behind it on the stream) after that work (dist.render_views_accumulate(render_event=...)). NULL = off. */
    void* render_wait_event;
} TexgsFwdArgs;

#define TEXGS_EV_FWD_START      0
#define TEXGS_EV_FWD_PREPROCESS 1   /* after texgs_preprocess_fwd (+ workspace clear) */
#define TEXGS_EV_FWD_SCAN       2   /* after texgs_scan_tiles                          */
#define TEXGS_EV_FWD_SCATTER    3   /* after texgs_scatter_pairs                       */
#define TEXGS_EV_FWD_SORT       4   /* after texgs_sort_tiles                          */
#define TEXGS_EV_FWD_RENDER     5   /* after texgs_render_fwd                          */
#define TEXGS_EV_BWD_START      6
#define TEXGS_EV_BWD_CLEAR      7   /* after the accumulator / texture-grad clears     */
#define TEXGS_EV_BWD_RENDER     8   /* after texgs_render_bwd                          */
#define TEXGS_EV_BWD_PREPROCESS 9   /* after texgs_preprocess_bwd                      */
#define TEXGS_EV_COUNT          10

/* Device-written run statistics, copied to ``counters_host`` (pinned) when that pointer is given. */
typedef struct TexgsCounters {
    uint32_t num_pairs;      /* K: (tile, Gaussian) pairs this view needs                         */
    uint32_t num_visible;    /* V: Gaussians with radius > 0                                      */
    uint32_t overflow;       /* 1 if K > pair_capacity: outputs are NOT valid, retry with >= K    */
    uint32_t max_tile_len;   /* longest per-tile list                                             */
    uint32_t num_blend_lo;   /* (pixel, Gaussian) contributions blended, low / high 32 bits       */
    uint32_t num_blend_hi;   /*   (only counted when TEXGS_FLAG_DEBUG is set)                     */
    uint32_t num_long_tiles; /* tiles whose list is longer than the small-list sort handles (filled during the sort)   */
    uint32_t reserved[1];
} TexgsCounters;

/* Sizes (bytes) of the three caller-owned workspaces for a given problem and pair capacity.
 *   geom : per-Gaussian projected records (kept for backward)
 *   bin  : tile counts/offsets, unsorted + sorted (tile, Gaussian) pair lists (kept for backward)
 *   img  : per-pixel final transmittance + contributor count (kept for backward) */
int texgs_workspace_sizes(const TexgsFwdArgs* a, uint64_t pair_capacity,
                          size_t* geom_bytes, size_t* bin_bytes, size_t* img_bytes);

/* Forward: preprocess -> per-tile binning -> per-tile sort -> render.  Replaces
 * ``_C.rasterize_gaussians`` [EXT] as reached from render/uv_tex_render.py:56.
 * Outputs: image (3,H,W), depth (1,H,W), norm (3,H,W), alpha (1,H,W), radii (P,) int32,
 * extra (E,H,W) or NULL.  ``counters_host`` may be NULL; if given it must be pinned host memory and
 * is filled by an async copy on ``stream`` issued right after the tile scan, i.e. BEFORE the
 * expensive kernels; ``counters_ready_event`` (a cudaEvent_t, may be NULL) is recorded right after
 * that copy so the host can learn K / overflow without draining the stream. */
int texgs_forward(const TexgsFwdArgs* a, void* geom_ws, void* bin_ws, uint64_t pair_capacity,
                  void* img_ws, float* out_image, float* out_depth, float* out_norm,
                  float* out_alpha, int32_t* out_radii, float* out_extra,
                  TexgsCounters* counters_host, void* counters_ready_event, void* stream);

typedef struct TexgsBwdArgs {
    TexgsFwdArgs fwd;            /* the same args the forward ran with                          */
    const void* geom_ws;         /* workspaces as left by texgs_forward                         */
    const void* bin_ws;
    const void* img_ws;
    uint64_t pair_capacity;
    /* incoming cotangents (device); any may be NULL (= zero) */
    const float* dL_dimage;      /* (3,H,W) */
    const float* dL_ddepth;      /* (1,H,W) */
    const float* dL_dnorm;       /* (3,H,W) */
    const float* dL_dalpha;      /* (1,H,W) */
    const float* dL_dextra;      /* (E,H,W) */
    const float* dL_dimage_nosh; /* (3,H,W) cotangent of the dual image (or NULL)                */
    /* scratch (device): P*TEXGS_BWD_ACC_FLOATS floats, zeroed by the library */
    float* acc_ws;
    /* gradient outputs (device). Per-Gaussian ones are fully written by the library (no
     * pre-zeroing needed); dL_dtexture is ACCUMULATED into (atomics) unless
     * ``zero_texture_grad`` is non-zero, in which case the library clears it first. NULL = skip. */
    float* dL_dmeans3D;          /* (P,3)   */
    float* dL_dmeans2D;          /* (P,3)  xy slots written in NDC units, z = 0                 */
    float* dL_dopacity;          /* (P,1)   */
    float* dL_dscales;           /* (P,3)   */
    float* dL_drotations;        /* (P,4)   */
    float* dL_dshs;              /* (P,M,3) */
    float* dL_dcolors_precomp;   /* (P,3)   */
    float* dL_duvs;              /* (P,3)   */
    float* dL_dtexture;          /* (6,R,R,3) */
    float* dL_dtexture_rgba;     /* (6,R,R,4) alternative to dL_dtexture: accumulated with 128-bit vector
                                    atomics (red.global.add.v4.f32), 4th float stays 0; give exactly one */
    float* dL_dextra_attrs;      /* (P,E)   cleared and written by the library                                  */
    float* dL_dcov3Ds;           /* (P,6)   when fwd.cov3Ds_precomp is given (off-diagonal entries count both halves) */
    int32_t zero_texture_grad;
    /* TEXGS_ACC_* bits: the marked per-Gaussian outputs are ADDED to (``out += grad``, rows of culled
     * Gaussians untouched) instead of overwritten — lets the caller point them at a persistent
     * gradient bucket (the all-reduce buffer) and skip autograd's separate accumulation pass. */
    uint32_t accumulate_mask;
} TexgsBwdArgs;

#define TEXGS_ACC_MEANS3D   1u
#define TEXGS_ACC_MEANS2D   2u
#define TEXGS_ACC_OPACITY   4u
#define TEXGS_ACC_SCALES    8u
#define TEXGS_ACC_ROTATIONS 16u
#define TEXGS_ACC_SHS       32u
#define TEXGS_ACC_COLORS    64u
#define TEXGS_ACC_UVS       128u

#define TEXGS_BWD_ACC_FLOATS 24

/* Backward: per-tile back-to-front render backward -> per-Gaussian preprocess backward.  Replaces
 * ``_C.rasterize_gaussians_backward`` [EXT]. */
int texgs_backward(const TexgsBwdArgs* b, void* stream);

/* (6,R,R,3) -> (6,R,R,4) repack of the cube texture (see TexgsFwdArgs.texture_rgba). */
int texgs_pack_texture(const float* texture, int32_t R, float* texture_rgba, void* stream);

/* ---- fused photometric loss of the training step ---------------------------------
 * loss = (1-lambda)*mean|image-gt| + lambda*(1 - SSIM(image, gt))   (models/texture_gaussian3d.py:333-340,
 * losses/pixelwise_loss.py:3-4, losses/ssim_loss.py:6-54: 11x11 Gaussian window, sigma 1.5, zero padding).
 * image, gt: (C,H,W) fp32 device. ``ws`` (texgs_photometric_workspace_size bytes, 16-byte aligned) carries
 * three derivative maps from forward to backward. out3 (device) = {loss, Ll1, Lssim = 1 - SSIM}.
 * Backward: dL_dimage = coef2[0]*d(Ll1)/d(image) + coef2[1]*d(Lssim)/d(image), coef2 = 2 device floats
 * (for  d loss: {(1-lambda)*g, lambda*g}  with g the incoming scalar gradient). */
int texgs_photometric_workspace_size(int32_t C, int32_t H, int32_t W, size_t* bytes);
int texgs_photometric_forward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim,
                              void* ws, float* out3, void* stream);
int texgs_photometric_backward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, const void* ws,
                               const float* coef2, float* dL_dimage, void* stream);

/* ---- geometry losses of the same step (models/texture_gaussian3d.py:342-368) -------
 *   Lalpha = mean|alpha - gt_alpha|                                         losses/pixelwise_loss.py:3-4
 *   Lnorm  = sum((1 - <norm, gt_norm>) * gt_alpha) / (sum(gt_alpha) + 1e-6)   losses/norm_reg_loss.py:66-71
 *   Lnsm   = smooth_loss(gt_image, norm, gt_alpha, gamma)                   losses/smooth_loss.py:4-27
 * alpha (1,H,W), norm (3,H,W): rasterizer outputs. gt_alpha (1,H,W) may be NULL (= ones), gt_norm / gt_image
 * (3,H,W) may be NULL (that loss is skipped and reads 0). out3 (device) = {Lalpha, Lnorm, Lnsm}.
 * ``ws`` (texgs_geometry_loss_workspace_size bytes, 16-byte aligned) carries the normalisers to backward.
 * Backward: dL_dalpha = coef3[0]*dLalpha/dalpha, dL_dnorm = coef3[1]*dLnorm/dnorm + coef3[2]*dLnsm/dnorm
 * (coef3 = 3 device floats; either output pointer may be NULL). */
int texgs_geometry_loss_workspace_size(int32_t H, int32_t W, size_t* bytes);
int texgs_geometry_loss_forward(const float* alpha, const float* norm, const float* gt_alpha, const float* gt_norm,
                                const float* gt_image, int32_t H, int32_t W, float gamma, void* ws, float* out3, void* stream);
int texgs_geometry_loss_backward(const float* alpha, const float* norm, const float* gt_alpha, const float* gt_norm,
                                 const float* gt_image, int32_t H, int32_t W, float gamma, const void* ws, const float* coef3,
                                 float* dL_dalpha, float* dL_dnorm, void* stream);

/* ---- the texture's optimizer step -----------------------------------------------------
 * Dense Adam with torch.optim.Adam's update (models/texture_gaussian3d.py:139-143 builds it with eps=1e-15,
 * :439-440 steps it; no weight decay, no amsgrad), for a parameter of n_texels*3 floats:
 *   m += (g-m)(1-b1);  v = v*b2 + (1-b2) g^2;  p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * The gradient comes either as (n,3) ``grad3`` or as the padded (n,4) ``grad_rgba`` the rasterizer backward
 * accumulates into (exactly one non-NULL); ``zero_grad`` != 0 clears it in the same pass. ``param_rgba``
 * (optional) receives the packed (n,4) copy of the UPDATED parameter (= texgs_pack_texture of it).
 * ``step`` is the 1-based step count. All buffers 16-byte aligned device memory. */
int texgs_texture_adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* grad3, float* grad_rgba,
                            float* param_rgba, uint64_t n_texels, double lr, double beta1, double beta2, double eps,
                            int32_t step, int32_t zero_grad, void* stream);

/* ---- data-parallel texture step — gradient reduction + Adam + parameter broadcast in one kernel ----------
 * One process per GPU, ``world`` ranks, the texture gradient of every rank in a symmetric buffer (padded (n,4) layout) and the
 * texture parameter (n,3) in another. Rank ``rank`` owns the texels of tiles [tile_lo, tile_hi) (1024 texels per tile;
 * texgs_dp_shard gives the canonical split), pulls and adds the ``world`` partial gradients of its tiles over NVLink
 * (multimem.ld_reduce through ``grad_mc`` / ``param_mc``, the NVSwitch multicast mappings, when both are non-NULL; plain peer
 * loads / stores through ``grad_ptrs`` / ``param_ptrs`` otherwise), applies texgs_texture_adam_step's update to its shard
 * (``exp_avg`` / ``exp_avg_sq`` hold the owned texels only) and writes the updated texels into every rank's parameter.
 * Replaces the NCCL all-reduce of the texture gradient + ``world`` identical optimizer steps. The caller synchronises the
 * ranks before (all backward passes complete) and after (all owners done) the call; the gradients are left untouched. */
typedef struct TexgsDpAdamArgs {
    int32_t world, rank;
    const float* grad_ptrs[16];
    float* param_ptrs[16];
    const float* grad_mc;
    float* param_mc;
    float* exp_avg;
    float* exp_avg_sq;
    uint64_t n_texels, tile_lo, tile_hi;
    double lr, beta1, beta2, eps;
    int32_t step, reserved;
} TexgsDpAdamArgs;
int texgs_dp_shard(uint64_t n_texels, int32_t world, int32_t rank, uint64_t* tile_lo, uint64_t* tile_hi);
int texgs_texture_adam_dp_step(const TexgsDpAdamArgs* args, void* stream);

/* ---- UV + Jacobian producer --------------------------------------------------------------
 * uv = normalize(mlp(relu(pre_mlp((xyz - offset) / scale) + emb)))   (models/modules/uv_net.py:19-36) and
 * J[n, 3i+j] = d uv_i / d xyz_j  (TextureGaussian3D.get_grad_uvs, models/texture_gaussian3d.py:217-227) in ONE
 * kernel: value + three forward-mode tangents per point are four rows of the same tcgen05 GEMMs.
 * Network of configs/texture_gaussian3d.yaml:18-27: 3 -> 128 -ReLU-> 128 (+emb, ReLU) -> 128 -ReLU-> 128 -ReLU-> 3.
 * Hidden weights W2..W4 (128,128) and W5 (3,128) are fp16 [out][in] row-major, W1 (128,3) fp32; biases fp32 or
 * NULL (tiny-cuda-nn has none, the nn.Linear fallback models/modules/utils.py:44-55 has). Compute: fp16 operands,
 * fp32 accumulation (the reference's tcnn path is fp16 too). ``stash[k]`` (optional, fp16 (N,128)) receives the
 * post-activation inputs of layers 2..5 for a backward pass. All pointers device memory, 16-byte aligned. */
typedef struct TexgsUvMlpArgs {
    int32_t N;
    const float* xyz;
    float offset[3];
    float inv_scale[3];
    const float* W1;
    const float* b1;
    const void* W_hidden[3];      /* fp16 */
    const float* b_hidden[3];
    const float* emb;
    const void* W5;               /* fp16 */
    const float* b5;
    float* uv;                    /* (N,3) */
    float* jacobian;              /* (N,9) or NULL */
    void* stash[4];               /* fp16 (N,128) each, or NULL */
    float* stash_inv_len;         /* (N) 1/max(|mlp output|, 1e-12), or NULL */
    float* debug_accumulators;    /* NULL, or [4*128*128 + 128*16] floats: raw accumulators of the first tile */
} TexgsUvMlpArgs;
int texgs_uvmlp_forward(const TexgsUvMlpArgs* a, void* stream);

/* Backward of uv w.r.t. xyz / emb / weights: the 128x128 products per hidden layer (delta @ W, delta^T @ a) are plain
 * GEMMs the host issues through cuBLAS on the fp16 stash; these three streaming kernels are the glue around them.
 * ``delta`` tensors are fp16 (N,128) under one power-of-two loss scale S chosen on the device from max|d loss/d out|.
 *   head: from g_uv (N,3) and the stash (uv, inv_len, a4) and W5 (3,128 fp16): delta4 (out), scale (out, 1 float),
 *         gW5 (3,128) += , gb5 (3) += (both UNscaled), colsum (128) += column sums of delta4 (scaled).
 *         ``amax_scratch`` is 1 float of device scratch.
 *   mask: delta <- delta * (a > 0) in place, colsum (128) += its column sums (scaled).
 *   tail: gxyz (N,3) = delta1 @ W1 * inv_scale / S (may be NULL), gW1 (128,3) += delta1^T x' / S.
 * Accumulated outputs must be zero-initialised by the caller. */
int texgs_uvmlp_backward_head(int32_t N, const float* g_uv, const float* uv, const float* inv_len, const void* a4, const void* W5,
                              float* amax_scratch, float* scale, void* delta4, float* gW5, float* gb5, float* colsum, void* stream);
/* One hidden layer of the backward on the tensor cores (tcgen05): gW[out][in] += delta^T a_prev, delta_out = (delta W) * (a_prev > 0)
 * in fp16, colsum[in] += column sums of delta_out. ``Wt`` is the layer's weight TRANSPOSED ([in][out], fp16); ``gW`` and ``colsum``
 * are accumulated into (clear them first); delta_out may not alias delta_in. Replaces torch.mm(delta.T, a), torch.mm(delta, W) and
 * the ReLU-mask glue kernel of round 1. */
int texgs_uvmlp_backward_layer(int32_t N, const void* delta_in, const void* a_prev, const void* Wt, void* delta_out, float* gW,
                               float* colsum, void* stream);
int texgs_uvmlp_backward_tail(int32_t N, const void* delta1, const float* xyz, const float* offset3_host, const float* inv_scale3_host,
                              const float* W1, const float* scale, float* gxyz, float* gW1, void* stream);

/* Frustum test only (upstream ``GaussianRasterizer.markVisible`` [EXT]; unused in the reference
 * tree). present (P,) int32: 1 if the Gaussian passes the near-plane cull. */
int texgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix16_host,
                       const float* projmatrix16_host, int32_t* present, void* stream);

/* Debug / test introspection: byte offsets of the sub-buffers inside the workspaces. */
typedef struct TexgsLayout {
    uint64_t geom_records;     /* P x 128 B projected records                                  */
    uint64_t geom_rects;       /* P x 8 B  (4 x uint16: x0,y0,x1,y1 tile rect)                 */
    uint64_t bin_counters;     /* TexgsCounters                                                */
    uint64_t bin_tile_count;   /* T x uint32                                                   */
    uint64_t bin_tile_offset;  /* (T+1) x uint32                                               */
    uint64_t bin_tile_cursor;  /* T x uint32                                                   */
    uint64_t bin_pairs;        /* capacity x 8 B  {gaussian id, depth bits} unsorted->sorted   */
    uint64_t bin_sorted_ids;   /* capacity x uint32 gaussian ids in (tile, depth, id) order    */
    uint64_t bin_cull_masks;   /* (pair_capacity/32 + tiles + 1) * 8 uint2: per warp and list chunk, which entries can touch
                                  its left / right 4x4 block (written by the forward render, re-used by the backward)  */
    uint64_t img_final_T;      /* H*W floats                                                   */
    uint64_t img_n_contrib;    /* H*W uint32                                                   */
    uint64_t num_tiles;
    uint64_t record_bytes;
} TexgsLayout;
int texgs_workspace_layout(const TexgsFwdArgs* a, uint64_t pair_capacity, TexgsLayout* out);

int texgs_abi_version(void);
const char* texgs_last_error(void);
/* Comma-separated list of the __global__ kernels this library launches (evidence for tests). */
const char* texgs_kernel_names(void);

#ifdef __cplusplus
}
#endif
#endif /* TEXGS_H_ */
