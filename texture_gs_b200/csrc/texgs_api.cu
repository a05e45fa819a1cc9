// texgs_api.cu — extern "C" boundary of libtexgs.so (declared in include/texgs.h).
// Host side only validates arguments, carves the caller-owned workspaces and enqueues kernels on
// the caller's stream. No allocation, no global mutable state, no host sync (unless DEBUG).
#include <cstdio>
#include <cstring>
#include <string>

#ifndef TEXGS_HOST_EMU
#include <nvtx3/nvToolsExt.h>      // header-only; a no-op unless a profiler (nsys / ncu --nvtx) is attached
#define TEXGS_NVTX_PUSH(name) nvtxRangePushA(name)
#define TEXGS_NVTX_POP() nvtxRangePop()
#else
#define TEXGS_NVTX_PUSH(name) ((void)0)
#define TEXGS_NVTX_POP() ((void)0)
#endif

#include "texgs_binning.cuh"
#include "texgs_common.cuh"
#include "texgs_loss.cuh"
#include "texgs_optim.cuh"
#include "texgs_uvmlp.cuh"
#include "texgs_preprocess.cuh"
#include "texgs_render.cuh"
#include "texgs_extra.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define TEXGS_CUDA_TRY(expr)                                                                             \
    do {                                                                                                 \
        cudaError_t err__ = (expr);                                                                      \
        if (err__ != cudaSuccess)                                                                        \
            return fail((int)err__, std::string(#expr) + ": " + cudaGetErrorString(err__));              \
    } while (0)

#define TEXGS_KERNEL_CHECK(name, debug, stream)                                                          \
    do {                                                                                                 \
        cudaError_t err__ = cudaGetLastError();                                                          \
        if (err__ == cudaSuccess && (debug)) err__ = cudaStreamSynchronize(stream);                      \
        if (err__ != cudaSuccess) return fail((int)err__, std::string(name) + ": " + cudaGetErrorString(err__)); \
    } while (0)

#define TEXGS_EV(a, slot, stream)                                                                        \
    do {                                                                                                 \
        if ((a)->profile_events && (a)->profile_events[slot])                                            \
            TEXGS_CUDA_TRY(cudaEventRecord((cudaEvent_t)(a)->profile_events[slot], stream));             \
    } while (0)

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// NVTX range over the enqueue of one stage of the path (reference train.py:124-125 brackets the whole iteration with
// CUDA events; these name the pieces): texgs/forward{/preprocess,/binning,/render}, texgs/backward{/render,/preprocess}
struct NvtxRange {
    explicit NvtxRange(const char* name) { TEXGS_NVTX_PUSH(name); }
    ~NvtxRange() { TEXGS_NVTX_POP(); }
};

struct Layout {
    TexgsLayout l;
    uint64_t geom_bytes, bin_bytes, img_bytes, bin_zero_bytes;
};

int make_layout(const TexgsFwdArgs* a, uint64_t cap, Layout& L) {
    if (!a) return fail(TEXGS_E_INVALID, "args is NULL");
    if (a->P < 0 || a->H <= 0 || a->W <= 0) return fail(TEXGS_E_INVALID, "bad P/H/W");
    const uint64_t gx = (a->W + TEXGS_TILE - 1) / TEXGS_TILE, gy = (a->H + TEXGS_TILE - 1) / TEXGS_TILE;
    if (gx > 65535 || gy > 65535) return fail(TEXGS_E_INVALID, "image too large for 16-bit tile coordinates");
    const uint64_t T = gx * gy, P = (uint64_t)a->P, HW = (uint64_t)a->H * a->W;
    memset(&L, 0, sizeof(L));
    L.l.num_tiles = T;
    L.l.record_bytes = sizeof(GaussRec);
    uint64_t o = 0;
    L.l.geom_records = o; o = align_up(o + P * sizeof(GaussRec), 256);
    L.l.geom_rects = o;   o = align_up(o + P * sizeof(uint2), 256);
    L.geom_bytes = o;
    o = 0;
    L.l.bin_counters = o;    o = align_up(o + sizeof(TexgsCounters), 256);
    L.l.bin_tile_count = o;  o = align_up(o + T * 4, 256);
    L.l.bin_tile_cursor = o; o = align_up(o + T * 4, 256);
    L.bin_zero_bytes = o;    // [counters | tile_count | tile_cursor] are cleared together
    L.l.bin_tile_offset = o; o = align_up(o + (T + 1) * 4, 256);
    L.l.bin_pairs = o;       o = align_up(o + cap * 8, 256);
    L.l.bin_sorted_ids = o;  o = align_up(o + cap * 4, 256);
    L.l.bin_cull_masks = o;  o = align_up(o + (cap / TEXGS_CHUNK + T + 1) * 8 * sizeof(uint2), 256);
    L.bin_bytes = o;
    o = 0;
    L.l.img_final_T = o;   o = align_up(o + HW * 4, 256);
    L.l.img_n_contrib = o; o = align_up(o + HW * 4, 256);
    L.img_bytes = o;
    return 0;
}

int max_degree_for(int M_rest) {
    int d = 0;
    while (d < 3 && (d + 2) * (d + 2) - 1 <= M_rest) ++d;
    return d;
}

int fill_params(const TexgsFwdArgs* a, void* geom, void* bin, uint64_t cap, void* img, const Layout& L, RasterParams& p) {
    if (a->mode != TEXGS_MODE_TEXTURE && a->mode != TEXGS_MODE_SH && a->mode != TEXGS_MODE_PRECOMP)
        return fail(TEXGS_E_INVALID, "unknown mode");
    if (a->E < 0 || (a->E > 0 && a->P > 0 && !a->extra_attrs)) return fail(TEXGS_E_INVALID, "E > 0 needs extra_attrs (P,E)");
    if (a->P > 0 && (!a->means3D || !a->opacities)) return fail(TEXGS_E_INVALID, "means3D and opacities must be given");
    if (a->cov3Ds_precomp) {
        if (a->scales || a->rotations) return fail(TEXGS_E_INVALID, "give scales + rotations or cov3Ds_precomp, not both");
        if (a->mode == TEXGS_MODE_TEXTURE) return fail(TEXGS_E_INVALID, "cov3Ds_precomp is a diff_gauss argument (render/render.py:83); the textured mode needs scales + rotations");
    } else if (a->P > 0 && (!a->scales || !a->rotations)) {
        return fail(TEXGS_E_INVALID, "scales and rotations (or cov3Ds_precomp) must be given");
    }
    if (((uintptr_t)a->rotations & 15) != 0) return fail(TEXGS_E_INVALID, "rotations must be 16-byte aligned");
    if (a->mode == TEXGS_MODE_TEXTURE) {
        if ((a->P > 0 && (!a->uvs || !a->gradient_uvs)) || !a->texture || a->R <= 0)
            return fail(TEXGS_E_INVALID, "texture mode needs uvs, gradient_uvs, texture and R > 0");
        if ((uint64_t)a->R * a->R * 18 >= (1ull << 31)) return fail(TEXGS_E_INVALID, "texture too large for 32-bit texel offsets");
    } else if (a->mode == TEXGS_MODE_SH) {
        if ((a->P > 0 && !a->shs) || a->M < 1) return fail(TEXGS_E_INVALID, "SH mode needs shs with M >= 1");
    } else if (a->P > 0 && !a->colors_precomp) {
        return fail(TEXGS_E_INVALID, "precomp mode needs colors_precomp");
    }
    if ((!geom || !bin || !img)) return fail(TEXGS_E_WORKSPACE, "workspace pointer is NULL");
    if (((uintptr_t)geom & 127) || ((uintptr_t)bin & 127) || ((uintptr_t)img & 127))
        return fail(TEXGS_E_WORKSPACE, "workspaces must be 128-byte aligned");
    memset(&p, 0, sizeof(p));
    p.P = a->P; p.M = a->shs ? a->M : 0; p.E = a->extra_attrs ? a->E : 0; p.H = a->H; p.W = a->W; p.R = a->R; p.mode = a->mode;
    const int m_rest = (a->mode == TEXGS_MODE_SH) ? p.M - 1 : p.M;
    int deg = a->sh_degree < 0 ? 0 : a->sh_degree;
    const int dmax = max_degree_for(m_rest);
    p.sh_degree = deg < dmax ? deg : dmax;
    p.flags = a->flags;
    p.grid_x = (a->W + TEXGS_TILE - 1) / TEXGS_TILE;
    p.grid_y = (a->H + TEXGS_TILE - 1) / TEXGS_TILE;
    p.num_tiles = p.grid_x * p.grid_y;
    p.tanfovx = a->tanfovx; p.tanfovy = a->tanfovy; p.scale_modifier = a->scale_modifier;
    p.focal_x = (float)a->W / (2.0f * a->tanfovx);
    p.focal_y = (float)a->H / (2.0f * a->tanfovy);
    memcpy(p.view.m, a->viewmatrix, sizeof(float) * 16);
    memcpy(p.proj.m, a->projmatrix, sizeof(float) * 16);
    memcpy(p.campos, a->campos, sizeof(float) * 3);
    memcpy(p.bg, a->bg, sizeof(float) * 3);
    p.means3D = a->means3D; p.shs = a->shs; p.colors_precomp = a->colors_precomp; p.opacities = a->opacities;
    p.scales = a->scales; p.rotations = a->rotations; p.uvs = a->uvs; p.gradient_uvs = a->gradient_uvs;
    p.texture = a->texture; p.extra_attrs = a->extra_attrs; p.cov3Ds_precomp = a->cov3Ds_precomp;
    p.texture_rgba = (a->mode == TEXGS_MODE_TEXTURE) ? reinterpret_cast<const float4*>(a->texture_rgba) : nullptr;
    if (((uintptr_t)a->texture_rgba & 15) != 0) return fail(TEXGS_E_INVALID, "texture_rgba must be 16-byte aligned");
    char* g = (char*)geom; char* b = (char*)bin; char* im = (char*)img;
    p.recs = (GaussRec*)(g + L.l.geom_records);
    p.rects = (uint2*)(g + L.l.geom_rects);
    p.counters = (TexgsCounters*)(b + L.l.bin_counters);
    p.tile_count = (unsigned*)(b + L.l.bin_tile_count);
    p.tile_cursor = (unsigned*)(b + L.l.bin_tile_cursor);
    p.tile_offset = (unsigned*)(b + L.l.bin_tile_offset);
    p.pairs = (uint2*)(b + L.l.bin_pairs);
    p.sorted_ids = (unsigned*)(b + L.l.bin_sorted_ids);
    p.cull_masks = (uint2*)(b + L.l.bin_cull_masks);
    p.pair_capacity = cap;
    p.out_image_nosh = (a->mode == TEXGS_MODE_TEXTURE) ? a->out_image_nosh : nullptr;
    p.final_T = (float*)(im + L.l.img_final_T);
    p.n_contrib = (unsigned*)(im + L.l.img_n_contrib);
    return 0;
}

// The render kernels use 65 KB of dynamic shared memory (8 warps x 2 stages x 32 records): opt in
// once per device context. Idempotent, so the per-thread flag is only an optimisation.
int ensure_render_smem() {
    static thread_local int done_for_device = -1;
    int dev = 0;
    TEXGS_CUDA_TRY(cudaGetDevice(&dev));
    if (done_for_device == dev) return 0;
    const int bytes = (int)TEXGS_RENDER_SMEM;
#define TEXGS_SET_SMEM(k) TEXGS_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, true, false, false>));
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, false, false, false>));
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, true, true, false>));
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, false, true, false>));
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_SH, false, false, false>));
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, true, false, true>));      // spec-switch (ALT) instantiations
    TEXGS_SET_SMEM((texgs_render_fwd<TEXGS_MODE_TEXTURE, true, true, true>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, true, false, true>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, true, true, true>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, true, false, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, false, false, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, false, true, false, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, false, false, false, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, true, true, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, true, false, true, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, false, true, true, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_TEXTURE, false, false, true, false>));
    TEXGS_SET_SMEM((texgs_render_bwd<TEXGS_MODE_SH, false, false, false, false>));
#undef TEXGS_SET_SMEM
    done_for_device = dev;
    return 0;
}

}  // namespace

extern "C" {

int texgs_abi_version(void) { return TEXGS_ABI_VERSION; }

const char* texgs_last_error(void) { return g_last_error.c_str(); }

const char* texgs_kernel_names(void) {
    return "texgs_preprocess_fwd,texgs_scan_tiles,texgs_scatter_pairs,texgs_sort_tiles_small,texgs_sort_tiles,texgs_render_fwd,"
           "texgs_render_bwd,texgs_extra_fwd,texgs_extra_bwd,texgs_preprocess_bwd,texgs_mark_visible_kernel,texgs_pack_texture_kernel,"
           "texgs_photometric_fwd_kernel,texgs_photometric_finalize_kernel,texgs_photometric_bwd_kernel,"
           "texgs_geometry_loss_fwd_kernel,texgs_geometry_loss_finalize_kernel,texgs_geometry_loss_bwd_kernel,texgs_texture_adam_kernel,texgs_texture_adam_dp_kernel,texgs_uvmlp_fwd_kernel,"
           "texgs_uvmlp_bwd_amax_kernel,texgs_uvmlp_bwd_head_kernel,texgs_uvmlp_bwd_layer_kernel,texgs_uvmlp_bwd_tail_kernel";
}

int texgs_workspace_sizes(const TexgsFwdArgs* a, uint64_t pair_capacity, size_t* geom_bytes, size_t* bin_bytes,
                          size_t* img_bytes) {
    Layout L;
    if (int rc = make_layout(a, pair_capacity, L)) return rc;
    if (geom_bytes) *geom_bytes = (size_t)L.geom_bytes;
    if (bin_bytes) *bin_bytes = (size_t)L.bin_bytes;
    if (img_bytes) *img_bytes = (size_t)L.img_bytes;
    return 0;
}

int texgs_workspace_layout(const TexgsFwdArgs* a, uint64_t pair_capacity, TexgsLayout* out) {
    Layout L;
    if (int rc = make_layout(a, pair_capacity, L)) return rc;
    if (!out) return fail(TEXGS_E_INVALID, "out is NULL");
    *out = L.l;
    return 0;
}

int texgs_forward(const TexgsFwdArgs* a, void* geom_ws, void* bin_ws, uint64_t pair_capacity, void* img_ws,
                  float* out_image, float* out_depth, float* out_norm, float* out_alpha, int32_t* out_radii,
                  float* out_extra, TexgsCounters* counters_host, void* counters_ready_event, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (int rc = make_layout(a, pair_capacity, L)) return rc;
    RasterParams p;
    if (int rc = fill_params(a, geom_ws, bin_ws, pair_capacity, img_ws, L, p)) return rc;
    if (!out_image || !out_depth || !out_norm || !out_alpha || (!out_radii && a->P > 0)) return fail(TEXGS_E_INVALID, "output pointer is NULL");
    if (a->E > 0 && !out_extra) return fail(TEXGS_E_INVALID, "E > 0 needs out_extra (E,H,W)");
    const bool debug = (a->flags & TEXGS_FLAG_DEBUG) != 0;

    NvtxRange nvtx_fwd("texgs/forward");
    TEXGS_EV(a, TEXGS_EV_FWD_START, stream);
    TEXGS_CUDA_TRY(cudaMemsetAsync((char*)bin_ws + L.l.bin_counters, 0, L.bin_zero_bytes, stream));
    const int gblocks = (p.P + 255) / 256;
    if (p.P > 0) {
        NvtxRange r("texgs/forward/preprocess");
        const size_t smem = prefwd_smem_bytes(p.M, p.mode, p.shs);      // SH rows of every warp staged by one bulk copy each
        if (smem > 200 * 1024) return fail(TEXGS_E_INVALID, "too many SH coefficients per Gaussian for the staged forward");
        if (smem > 48 * 1024)
            TEXGS_CUDA_TRY(cudaFuncSetAttribute(texgs_preprocess_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        texgs_preprocess_fwd<<<gblocks, 256, smem, stream>>>(p, out_radii);
        TEXGS_KERNEL_CHECK("texgs_preprocess_fwd", debug, stream);
    }
    TEXGS_EV(a, TEXGS_EV_FWD_PREPROCESS, stream);
    {
    NvtxRange r("texgs/forward/binning");
    texgs_scan_tiles<<<1, TEXGS_SCAN_THREADS, 0, stream>>>(p);
    TEXGS_KERNEL_CHECK("texgs_scan_tiles", debug, stream);
    TEXGS_EV(a, TEXGS_EV_FWD_SCAN, stream);
    if (counters_host && !debug) {
        TEXGS_CUDA_TRY(cudaMemcpyAsync(counters_host, p.counters, sizeof(TexgsCounters), cudaMemcpyDeviceToHost, stream));
        if (counters_ready_event) TEXGS_CUDA_TRY(cudaEventRecord((cudaEvent_t)counters_ready_event, stream));
    }
    if (p.P > 0) {
        texgs_scatter_pairs<<<gblocks, 256, 0, stream>>>(p);
        TEXGS_KERNEL_CHECK("texgs_scatter_pairs", debug, stream);
    }
    TEXGS_EV(a, TEXGS_EV_FWD_SCATTER, stream);
    texgs_sort_tiles_small<<<p.num_tiles, TEXGS_SORT_SMALL_THREADS, 0, stream>>>(p);
    TEXGS_KERNEL_CHECK("texgs_sort_tiles_small", debug, stream);
    texgs_sort_tiles<<<TEXGS_SORT_LONG_CTAS, TEXGS_SORT_THREADS, 0, stream>>>(p);
    TEXGS_KERNEL_CHECK("texgs_sort_tiles", debug, stream);
    TEXGS_EV(a, TEXGS_EV_FWD_SORT, stream);
    }
    if (int rc = ensure_render_smem()) return rc;
    if (a->render_wait_event) TEXGS_CUDA_TRY(cudaStreamWaitEvent(stream, (cudaEvent_t)a->render_wait_event, 0));
    {
        NvtxRange r("texgs/forward/render");
        const bool t4 = p.texture_rgba != nullptr, dual = p.out_image_nosh != nullptr;
#define TEXGS_LAUNCH_FWD(M, T4, DU, AL) texgs_render_fwd<M, T4, DU, AL><<<p.num_tiles * (8 / TEXGS_FWD_WARPS), 32 * TEXGS_FWD_WARPS, TEXGS_FWD_WARPS * sizeof(WarpSmem), stream>>>(p, out_image, out_depth, out_norm, out_alpha)
        const bool alt = p.mode == TEXGS_MODE_TEXTURE && (p.flags & TEXGS_FLAG_SPEC_MASK) != 0u;
        if (alt && !t4) return fail(TEXGS_E_INVALID, "the spec-switch flags need the packed texel copy (texture_rgba)");
        if (p.mode != TEXGS_MODE_TEXTURE) TEXGS_LAUNCH_FWD(TEXGS_MODE_SH, false, false, false);
        else if (alt && dual) TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, true, true, true);
        else if (alt)         TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, true, false, true);
        else if (t4 && dual)  TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, true, true, false);
        else if (t4)          TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, true, false, false);
        else if (dual)        TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, false, true, false);
        else                  TEXGS_LAUNCH_FWD(TEXGS_MODE_TEXTURE, false, false, false);
#undef TEXGS_LAUNCH_FWD
    }
    TEXGS_KERNEL_CHECK("texgs_render_fwd", debug, stream);
    if (a->E > 0) {   // cold path (the reference tree never passes extra_attrs): separate list walk, see texgs_extra.cuh
        if (p.E > 0) texgs_extra_fwd<<<p.num_tiles, 256, 0, stream>>>(p, out_extra);
        else TEXGS_CUDA_TRY(cudaMemsetAsync(out_extra, 0, (size_t)a->E * a->H * a->W * sizeof(float), stream));   // P == 0
        TEXGS_KERNEL_CHECK("texgs_extra_fwd", debug, stream);
    }
    TEXGS_EV(a, TEXGS_EV_FWD_RENDER, stream);
    if (counters_host && debug) {   // debug: counters include the blend count, copied after the render
        TEXGS_CUDA_TRY(cudaMemcpyAsync(counters_host, p.counters, sizeof(TexgsCounters), cudaMemcpyDeviceToHost, stream));
        if (counters_ready_event) TEXGS_CUDA_TRY(cudaEventRecord((cudaEvent_t)counters_ready_event, stream));
        TEXGS_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return 0;
}

int texgs_backward(const TexgsBwdArgs* b, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!b) return fail(TEXGS_E_INVALID, "args is NULL");
    const TexgsFwdArgs* a = &b->fwd;
    Layout L;
    if (int rc = make_layout(a, b->pair_capacity, L)) return rc;
    RasterParams p;
    if (int rc = fill_params(a, (void*)b->geom_ws, (void*)b->bin_ws, b->pair_capacity, (void*)b->img_ws, L, p)) return rc;
    if (!b->acc_ws) return fail(TEXGS_E_WORKSPACE, "acc_ws is NULL");
    if ((uintptr_t)b->acc_ws & 15) return fail(TEXGS_E_WORKSPACE, "acc_ws must be 16-byte aligned");
    const bool debug = (a->flags & TEXGS_FLAG_DEBUG) != 0;

    NvtxRange nvtx_bwd("texgs/backward");
    TEXGS_EV(a, TEXGS_EV_BWD_START, stream);
    TEXGS_CUDA_TRY(cudaMemsetAsync(b->acc_ws, 0, (size_t)p.P * TEXGS_BWD_ACC_FLOATS * sizeof(float), stream));
    if (b->dL_dtexture && b->dL_dtexture_rgba) return fail(TEXGS_E_INVALID, "give dL_dtexture or dL_dtexture_rgba, not both");
    if (((uintptr_t)b->dL_drotations & 15) != 0) return fail(TEXGS_E_INVALID, "dL_drotations must be 16-byte aligned");
    if (((uintptr_t)b->dL_dtexture_rgba & 15) != 0) return fail(TEXGS_E_INVALID, "dL_dtexture_rgba must be 16-byte aligned");
    if (b->dL_dtexture && b->zero_texture_grad && p.mode == TEXGS_MODE_TEXTURE)
        TEXGS_CUDA_TRY(cudaMemsetAsync(b->dL_dtexture, 0, (size_t)6 * p.R * p.R * 3 * sizeof(float), stream));
    if (b->dL_dtexture_rgba && b->zero_texture_grad && p.mode == TEXGS_MODE_TEXTURE)
        TEXGS_CUDA_TRY(cudaMemsetAsync(b->dL_dtexture_rgba, 0, (size_t)6 * p.R * p.R * 4 * sizeof(float), stream));
    TEXGS_EV(a, TEXGS_EV_BWD_CLEAR, stream);
    BwdIn in{b->dL_dimage, b->dL_ddepth, b->dL_dnorm, b->dL_dalpha, b->dL_dimage_nosh};
    if (int rc = ensure_render_smem()) return rc;
    // variants: texel reads packed or not  x  texel-gradient writes packed or not  x  dual image
    if (p.mode == TEXGS_MODE_TEXTURE) {
        const bool rd4 = p.texture_rgba != nullptr, wr4 = b->dL_dtexture_rgba != nullptr, dual = p.out_image_nosh != nullptr;
        float* dt = wr4 ? b->dL_dtexture_rgba : b->dL_dtexture;
#define TEXGS_LAUNCH_BWD(R4, W4, DU, AL) texgs_render_bwd<TEXGS_MODE_TEXTURE, R4, W4, DU, AL><<<p.num_tiles * (8 / TEXGS_BWD_WARPS), 32 * TEXGS_BWD_WARPS, TEXGS_BWD_WARPS * sizeof(WarpSmem), stream>>>(p, in, b->acc_ws, dt)
        const bool alt = (p.flags & TEXGS_FLAG_SPEC_MASK) != 0u;
        if (alt && (!rd4 || (dt && !wr4))) return fail(TEXGS_E_INVALID, "the spec-switch flags need texture_rgba and the padded texel gradient (dL_dtexture_rgba)");
        if (alt) {
            if (dual) TEXGS_LAUNCH_BWD(true, true, true, true); else TEXGS_LAUNCH_BWD(true, true, false, true);
        } else if (dual) {
            if (rd4 && wr4) TEXGS_LAUNCH_BWD(true, true, true, false); else if (rd4) TEXGS_LAUNCH_BWD(true, false, true, false);
            else if (wr4) TEXGS_LAUNCH_BWD(false, true, true, false); else TEXGS_LAUNCH_BWD(false, false, true, false);
        } else {
            if (rd4 && wr4) TEXGS_LAUNCH_BWD(true, true, false, false); else if (rd4) TEXGS_LAUNCH_BWD(true, false, false, false);
            else if (wr4) TEXGS_LAUNCH_BWD(false, true, false, false); else TEXGS_LAUNCH_BWD(false, false, false, false);
        }
#undef TEXGS_LAUNCH_BWD
    } else {
        texgs_render_bwd<TEXGS_MODE_SH, false, false, false, false><<<p.num_tiles * (8 / TEXGS_BWD_WARPS), 32 * TEXGS_BWD_WARPS, TEXGS_BWD_WARPS * sizeof(WarpSmem), stream>>>(p, in, b->acc_ws, nullptr);
    }
    TEXGS_KERNEL_CHECK("texgs_render_bwd", debug, stream);
    if (p.E > 0 && p.P > 0) {
        if (b->dL_dextra_attrs) TEXGS_CUDA_TRY(cudaMemsetAsync(b->dL_dextra_attrs, 0, (size_t)p.P * p.E * sizeof(float), stream));
        if (b->dL_dextra) {
            texgs_extra_bwd<<<p.num_tiles, 256, 0, stream>>>(p, b->dL_dextra, b->acc_ws, b->dL_dextra_attrs);
            TEXGS_KERNEL_CHECK("texgs_extra_bwd", debug, stream);
        }
    }
    TEXGS_EV(a, TEXGS_EV_BWD_RENDER, stream);
    if (p.P > 0) {
        NvtxRange r("texgs/backward/preprocess");
        BwdOut g{b->dL_dmeans3D, b->dL_dmeans2D, b->dL_dopacity, b->dL_dscales, b->dL_drotations,
                 b->dL_dshs, b->dL_dcolors_precomp, b->dL_duvs, a->cov3Ds_precomp ? b->dL_dcov3Ds : nullptr, b->accumulate_mask};
        const size_t smem = prebwd_smem_bytes(p.M);
        if (smem > 200 * 1024) return fail(TEXGS_E_INVALID, "too many SH coefficients per Gaussian for the staged backward");
        if (smem > 48 * 1024)   // idempotent; only reached with more than 15 coefficients per Gaussian
            TEXGS_CUDA_TRY(cudaFuncSetAttribute(texgs_preprocess_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        texgs_preprocess_bwd<<<(p.P + 255) / 256, 256, smem, stream>>>(p, b->acc_ws, g);
        TEXGS_KERNEL_CHECK("texgs_preprocess_bwd", debug, stream);
    }
    TEXGS_EV(a, TEXGS_EV_BWD_PREPROCESS, stream);
    return 0;
}

__global__ void __launch_bounds__(256) texgs_pack_texture_kernel(const float* __restrict__ tex, float4* __restrict__ out, size_t ntexel) {
    // 4 texels (48 B in, 64 B out) per thread: three 128-bit loads, four 128-bit stores
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t t0 = q * 4;
    if (t0 + 3 < ntexel) {
        const float4* src = reinterpret_cast<const float4*>(tex + t0 * 3);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        out[t0] = make_float4(a.x, a.y, a.z, 0.f);
        out[t0 + 1] = make_float4(a.w, b.x, b.y, 0.f);
        out[t0 + 2] = make_float4(b.z, b.w, c.x, 0.f);
        out[t0 + 3] = make_float4(c.y, c.z, c.w, 0.f);
    } else {
        for (size_t t = t0; t < ntexel; ++t) out[t] = make_float4(tex[3 * t], tex[3 * t + 1], tex[3 * t + 2], 0.f);
    }
}

int texgs_pack_texture(const float* texture, int32_t R, float* texture_rgba, void* stream_) {
    if (!texture || !texture_rgba || R <= 0) return fail(TEXGS_E_INVALID, "bad arguments");
    if (((uintptr_t)texture & 15) || ((uintptr_t)texture_rgba & 15)) return fail(TEXGS_E_INVALID, "texture buffers must be 16-byte aligned");
    const size_t ntexel = (size_t)6 * R * R;
    const size_t nthreads = (ntexel + 3) / 4;
    texgs_pack_texture_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        texture, reinterpret_cast<float4*>(texture_rgba), ntexel);
    TEXGS_KERNEL_CHECK("texgs_pack_texture_kernel", false, (cudaStream_t)stream_);
    return 0;
}

static size_t loss_parts_bytes(int C, int H, int W) {
    const size_t nb = (size_t)((W + TEXGS_LOSS_TILE - 1) / TEXGS_LOSS_TILE) * ((H + TEXGS_LOSS_TILE_H - 1) / TEXGS_LOSS_TILE_H) * C;
    return (nb * sizeof(LossSums) + 255) / 256 * 256;
}

int texgs_photometric_workspace_size(int32_t C, int32_t H, int32_t W, size_t* bytes) {
    if (C <= 0 || H <= 0 || W <= 0 || !bytes) return fail(TEXGS_E_INVALID, "bad arguments");
    *bytes = loss_parts_bytes(C, H, W) + (size_t)3 * C * H * W * sizeof(float);     // per-CTA partial sums + 3 derivative maps
    return 0;
}

int texgs_photometric_forward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim,
                              void* ws, float* out3, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!image || !gt || !ws || !out3 || C <= 0 || H <= 0 || W <= 0 || C > 65535) return fail(TEXGS_E_INVALID, "bad arguments");
    if ((uintptr_t)ws & 15) return fail(TEXGS_E_WORKSPACE, "workspace must be 16-byte aligned");
    const size_t n = (size_t)C * H * W;
    LossSums* sums = (LossSums*)ws;
    float* maps = (float*)((char*)ws + loss_parts_bytes(C, H, W));
    const dim3 grid((W + TEXGS_LOSS_TILE - 1) / TEXGS_LOSS_TILE, (H + TEXGS_LOSS_TILE_H - 1) / TEXGS_LOSS_TILE_H, C);
    texgs_photometric_fwd_kernel<<<grid, dim3(TEXGS_LOSS_TILE, TEXGS_LOSS_TILE), 0, stream>>>(image, gt, H, W, maps, maps + n, maps + 2 * n, sums);
    TEXGS_KERNEL_CHECK("texgs_photometric_fwd_kernel", false, stream);
    texgs_photometric_finalize_kernel<<<1, 1024, 0, stream>>>(sums, (int)(grid.x * grid.y * grid.z), 1.0 / (double)n, lambda_dssim, out3);
    TEXGS_KERNEL_CHECK("texgs_photometric_finalize_kernel", false, stream);
    return 0;
}

int texgs_photometric_backward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, const void* ws,
                               const float* coef2, float* dL_dimage, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!image || !gt || !ws || !coef2 || !dL_dimage || C <= 0 || H <= 0 || W <= 0 || C > 65535) return fail(TEXGS_E_INVALID, "bad arguments");
    const size_t n = (size_t)C * H * W;
    const float* maps = (const float*)((const char*)ws + loss_parts_bytes(C, H, W));
    const dim3 grid((W + TEXGS_LOSS_TILE - 1) / TEXGS_LOSS_TILE, (H + TEXGS_LOSS_TILE_H - 1) / TEXGS_LOSS_TILE_H, C);
    texgs_photometric_bwd_kernel<<<grid, dim3(TEXGS_LOSS_TILE, TEXGS_LOSS_TILE), 0, stream>>>(image, gt, H, W, maps, maps + n, maps + 2 * n, coef2,
                                                                                             (float)(1.0 / (double)n), dL_dimage);
    TEXGS_KERNEL_CHECK("texgs_photometric_bwd_kernel", false, stream);
    return 0;
}

static size_t geo_parts(int H, int W) {
    return (size_t)((W + TEXGS_GEO_TX - 1) / TEXGS_GEO_TX) * ((H + TEXGS_GEO_TY - 1) / TEXGS_GEO_TY);
}

int texgs_geometry_loss_workspace_size(int32_t H, int32_t W, size_t* bytes) {
    if (H <= 0 || W <= 0 || !bytes) return fail(TEXGS_E_INVALID, "bad arguments");
    *bytes = 256 + geo_parts(H, W) * TEXGS_GEO_NSUM * sizeof(double);          // scales, then per-CTA partial sums
    return 0;
}

static bool geo_args(GeoIn& g, const float* alpha, const float* norm, const float* gt_alpha, const float* gt_norm, const float* gt_image,
                     int32_t H, int32_t W, float gamma) {
    if (!alpha || !norm || H <= 0 || W <= 0 || !(gamma > 0.f)) return false;
    g.alpha = alpha; g.norm = norm; g.gt_alpha = gt_alpha; g.gt_norm = gt_norm; g.gt_image = gt_image;
    g.H = H; g.W = W; g.inv_gamma = 1.0f / gamma;
    return true;
}

int texgs_geometry_loss_forward(const float* alpha, const float* norm, const float* gt_alpha, const float* gt_norm, const float* gt_image,
                                int32_t H, int32_t W, float gamma, void* ws, float* out3, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GeoIn g;
    if (!geo_args(g, alpha, norm, gt_alpha, gt_norm, gt_image, H, W, gamma) || !ws || !out3) return fail(TEXGS_E_INVALID, "bad arguments");
    if ((uintptr_t)ws & 15) return fail(TEXGS_E_WORKSPACE, "workspace must be 16-byte aligned");
    const dim3 grid((W + TEXGS_GEO_TX - 1) / TEXGS_GEO_TX, (H + TEXGS_GEO_TY - 1) / TEXGS_GEO_TY);
    double* parts = (double*)((char*)ws + 256);
    texgs_geometry_loss_fwd_kernel<<<grid, dim3(TEXGS_GEO_TX, TEXGS_GEO_TY), 0, stream>>>(g, parts);
    TEXGS_KERNEL_CHECK("texgs_geometry_loss_fwd_kernel", false, stream);
    texgs_geometry_loss_finalize_kernel<<<1, 1024, 0, stream>>>(parts, (int)geo_parts(H, W), 1.0 / ((double)H * W), (float*)ws, out3);
    TEXGS_KERNEL_CHECK("texgs_geometry_loss_finalize_kernel", false, stream);
    return 0;
}

int texgs_geometry_loss_backward(const float* alpha, const float* norm, const float* gt_alpha, const float* gt_norm, const float* gt_image,
                                 int32_t H, int32_t W, float gamma, const void* ws, const float* coef3, float* dL_dalpha, float* dL_dnorm,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GeoIn g;
    if (!geo_args(g, alpha, norm, gt_alpha, gt_norm, gt_image, H, W, gamma) || !ws || !coef3 || (!dL_dalpha && !dL_dnorm))
        return fail(TEXGS_E_INVALID, "bad arguments");
    const dim3 grid((W + TEXGS_GEO_TX - 1) / TEXGS_GEO_TX, (H + TEXGS_GEO_TY - 1) / TEXGS_GEO_TY);
    texgs_geometry_loss_bwd_kernel<<<grid, dim3(TEXGS_GEO_TX, TEXGS_GEO_TY), 0, stream>>>(g, (const float*)ws, coef3, dL_dalpha, dL_dnorm);
    TEXGS_KERNEL_CHECK("texgs_geometry_loss_bwd_kernel", false, stream);
    return 0;
}

int texgs_texture_adam_step(float* param, float* exp_avg, float* exp_avg_sq, const float* grad3, float* grad_rgba, float* param_rgba,
                            uint64_t n_texels, double lr, double beta1, double beta2, double eps, int32_t step, int32_t zero_grad,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!param || !exp_avg || !exp_avg_sq || (grad3 == nullptr) == (grad_rgba == nullptr) || step < 1 ||
        !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0))
        return fail(TEXGS_E_INVALID, "bad arguments (exactly one of grad3 / grad_rgba, step >= 1, betas in [0,1))");
    if (n_texels == 0) return 0;
    if (((uintptr_t)param | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)grad3 | (uintptr_t)grad_rgba | (uintptr_t)param_rgba) & 15)
        return fail(TEXGS_E_INVALID, "all buffers must be 16-byte aligned");
    AdamArgs a;
    a.p = param; a.m = exp_avg; a.v = exp_avg_sq; a.g3 = grad3; a.g4 = grad_rgba; a.rgba = param_rgba; a.n = n_texels;
    // scalars formed in double on the host and rounded once, as torch.optim.Adam's scalar path does
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    a.one_minus_b1 = (float)(1.0 - beta1); a.b2 = (float)beta2; a.one_minus_b2 = (float)(1.0 - beta2);
    a.step_size = (float)(lr / bc1); a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2)); a.eps = (float)eps;
    a.zero_grad = zero_grad;
    const uint64_t ctas = (n_texels + TEXGS_ADAM_TEXELS - 1) / TEXGS_ADAM_TEXELS;
    if (ctas > 0x7fffffffull) return fail(TEXGS_E_INVALID, "tensor too large");
    texgs_texture_adam_kernel<<<(unsigned)ctas, TEXGS_ADAM_THREADS, 0, stream>>>(a);
    TEXGS_KERNEL_CHECK("texgs_texture_adam_kernel", false, stream);
    return 0;
}

int texgs_dp_shard(uint64_t n_texels, int32_t world, int32_t rank, uint64_t* tile_lo, uint64_t* tile_hi) {
    if (world < 1 || rank < 0 || rank >= world || !tile_lo || !tile_hi) return fail(TEXGS_E_INVALID, "bad arguments");
    const uint64_t tiles = (n_texels + TEXGS_ADAM_TEXELS - 1) / TEXGS_ADAM_TEXELS;
    *tile_lo = tiles * (uint64_t)rank / (uint64_t)world;
    *tile_hi = tiles * (uint64_t)(rank + 1) / (uint64_t)world;
    return 0;
}

int texgs_texture_adam_dp_step(const TexgsDpAdamArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || a->world < 1 || a->world > TEXGS_DP_MAX_RANKS || a->rank < 0 || a->rank >= a->world || a->step < 1 || !a->exp_avg || !a->exp_avg_sq ||
        !(a->beta1 >= 0.0 && a->beta1 < 1.0) || !(a->beta2 >= 0.0 && a->beta2 < 1.0) || a->tile_hi < a->tile_lo ||
        a->tile_hi * TEXGS_ADAM_TEXELS >= a->n_texels + TEXGS_ADAM_TEXELS)
        return fail(TEXGS_E_INVALID, "bad arguments (1 <= world <= 16, step >= 1, betas in [0,1), tile range inside the texture)");
    uintptr_t al = (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq | (uintptr_t)a->grad_mc | (uintptr_t)a->param_mc;
    for (int r = 0; r < a->world; ++r) {
        if (!a->grad_ptrs[r] || !a->param_ptrs[r]) return fail(TEXGS_E_INVALID, "a peer pointer is NULL");
        al |= (uintptr_t)a->grad_ptrs[r] | (uintptr_t)a->param_ptrs[r];
    }
    if (al & 15) return fail(TEXGS_E_INVALID, "all buffers must be 16-byte aligned");
    if ((a->grad_mc == nullptr) != (a->param_mc == nullptr)) return fail(TEXGS_E_INVALID, "give both multicast mappings or neither");
#ifdef TEXGS_HOST_EMU
    if (a->grad_mc) return fail(TEXGS_E_INVALID, "multimem needs an NVSwitch fabric (the host emulation runs the peer path only)");
#endif
    if (a->tile_hi == a->tile_lo) return 0;
    DpAdamArgs k;
    k.world = a->world; k.rank = a->rank;
    for (int r = 0; r < TEXGS_DP_MAX_RANKS; ++r) { k.grad[r] = r < a->world ? a->grad_ptrs[r] : nullptr; k.param[r] = r < a->world ? a->param_ptrs[r] : nullptr; }
    k.grad_mc = a->grad_mc; k.param_mc = a->param_mc; k.m = a->exp_avg; k.v = a->exp_avg_sq;
    k.n = a->n_texels; k.tile_lo = a->tile_lo; k.tile_hi = a->tile_hi;
    const double bc1 = 1.0 - pow(a->beta1, (double)a->step), bc2 = 1.0 - pow(a->beta2, (double)a->step);
    k.one_minus_b1 = (float)(1.0 - a->beta1); k.b2 = (float)a->beta2; k.one_minus_b2 = (float)(1.0 - a->beta2);
    k.step_size = (float)(a->lr / bc1); k.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2)); k.eps = (float)a->eps;
    const unsigned ctas = (unsigned)(a->tile_hi - a->tile_lo);
    if (a->grad_mc) texgs_texture_adam_dp_kernel<true><<<ctas, TEXGS_ADAM_THREADS, 0, stream>>>(k);
    else texgs_texture_adam_dp_kernel<false><<<ctas, TEXGS_ADAM_THREADS, 0, stream>>>(k);
    TEXGS_KERNEL_CHECK("texgs_texture_adam_dp_kernel", false, stream);
    return 0;
}

int texgs_uvmlp_forward(const TexgsUvMlpArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || a->N < 0) return fail(TEXGS_E_INVALID, "bad arguments");
    if (a->N == 0) return 0;
    if (!a->xyz || !a->W1 || !a->W_hidden[0] || !a->W_hidden[1] || !a->W_hidden[2] || !a->emb || !a->W5 || !a->uv)
        return fail(TEXGS_E_INVALID, "xyz, W1, W_hidden[0..2], emb, W5 and uv are required");
    uintptr_t al = (uintptr_t)a->W_hidden[0] | (uintptr_t)a->W_hidden[1] | (uintptr_t)a->W_hidden[2] | (uintptr_t)a->W5;
    for (int i = 0; i < 4; ++i) al |= (uintptr_t)a->stash[i];
    if (al & 15) return fail(TEXGS_E_INVALID, "fp16 weight / stash buffers must be 16-byte aligned");
    UvMlpParams P;
    P.N = a->N; P.xyz = a->xyz;
    for (int i = 0; i < 3; ++i) {
        P.off[i] = a->offset[i]; P.inv_scale[i] = a->inv_scale[i];
        P.W[i] = (const __half*)a->W_hidden[i]; P.b[i] = a->b_hidden[i];
    }
    P.W1 = a->W1; P.b1 = a->b1; P.emb = a->emb; P.W5 = (const __half*)a->W5; P.b5 = a->b5;
    P.uv = a->uv; P.J = a->jacobian;
    for (int i = 0; i < 4; ++i) P.stash[i] = (__half*)a->stash[i];
    P.stash_inv_len = a->stash_inv_len;
    P.dbg = a->debug_accumulators;
    static thread_local int attr_set_for_device = -1;
    int dev = 0, sms = 0;
    TEXGS_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_set_for_device != dev) {
        TEXGS_CUDA_TRY(cudaFuncSetAttribute(texgs_uvmlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UvMlpSmem::TOTAL));
        attr_set_for_device = dev;
    }
    TEXGS_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ntiles = (a->N + UVMLP_TILE - 1) / UVMLP_TILE;
    const int ctas = std::max(1, std::min(sms, (ntiles + UVMLP_GROUPS - 1) / UVMLP_GROUPS));
    texgs_uvmlp_fwd_kernel<<<ctas, UVMLP_THREADS, UvMlpSmem::TOTAL, stream>>>(P);
    TEXGS_KERNEL_CHECK("texgs_uvmlp_fwd_kernel", false, stream);
    return 0;
}

static inline unsigned uvbwd_ctas(int N) { return (unsigned)((N + UVBWD_ROWS_PER_CTA - 1) / UVBWD_ROWS_PER_CTA); }

int texgs_uvmlp_backward_head(int32_t N, const float* g_uv, const float* uv, const float* inv_len, const void* a4, const void* W5,
                              float* amax_scratch, float* scale, void* delta4, float* gW5, float* gb5, float* colsum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N <= 0 || !g_uv || !uv || !inv_len || !a4 || !W5 || !amax_scratch || !scale || !delta4 || !gW5 || !gb5 || !colsum)
        return fail(TEXGS_E_INVALID, "bad arguments");
    if (((uintptr_t)a4 | (uintptr_t)W5 | (uintptr_t)delta4) & 15) return fail(TEXGS_E_INVALID, "fp16 buffers must be 16-byte aligned");
    TEXGS_CUDA_TRY(cudaMemsetAsync(amax_scratch, 0, sizeof(float), stream));
    texgs_uvmlp_bwd_amax_kernel<<<std::min(1184, (N + 255) / 256), 256, 0, stream>>>(N, g_uv, uv, inv_len, amax_scratch);
    TEXGS_KERNEL_CHECK("texgs_uvmlp_bwd_amax_kernel", false, stream);
    texgs_uvmlp_bwd_head_kernel<<<uvbwd_ctas(N), UVBWD_THREADS, 0, stream>>>(N, g_uv, uv, inv_len, (const __half*)a4, (const __half*)W5, amax_scratch,
                                                                             scale, (__half*)delta4, gW5, gb5, colsum);
    TEXGS_KERNEL_CHECK("texgs_uvmlp_bwd_head_kernel", false, stream);
    return 0;
}

int texgs_uvmlp_backward_layer(int32_t N, const void* delta_in, const void* a_prev, const void* Wt, void* delta_out, float* gW,
                               float* colsum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N <= 0 || !delta_in || !a_prev || !Wt || !delta_out || !gW || !colsum || delta_in == delta_out) return fail(TEXGS_E_INVALID, "bad arguments");
    if (((uintptr_t)delta_in | (uintptr_t)a_prev | (uintptr_t)Wt | (uintptr_t)delta_out) & 15) return fail(TEXGS_E_INVALID, "fp16 buffers must be 16-byte aligned");
#ifdef TEXGS_HOST_EMU
    return fail(TEXGS_E_INVALID, "tcgen05 kernels have no host emulation");
#else
    static thread_local int attr_set_for_device = -1;
    int dev = 0, sms = 0;
    TEXGS_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_set_for_device != dev) {
        TEXGS_CUDA_TRY(cudaFuncSetAttribute(texgs_uvmlp_bwd_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UvBwdSmem::TOTAL));
        attr_set_for_device = dev;
    }
    TEXGS_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ntiles = (N + 127) / 128;
    const int ctas = std::max(1, std::min(sms, (ntiles + UVBL_GROUPS - 1) / UVBL_GROUPS));
    texgs_uvmlp_bwd_layer_kernel<<<ctas, UVBL_THREADS, UvBwdSmem::TOTAL, stream>>>(N, (const __half*)delta_in, (const __half*)a_prev, (const __half*)Wt,
                                                                                 (__half*)delta_out, gW, colsum);
    TEXGS_KERNEL_CHECK("texgs_uvmlp_bwd_layer_kernel", false, stream);
    return 0;
#endif
}

int texgs_uvmlp_backward_tail(int32_t N, const void* delta1, const float* xyz, const float* offset3_host, const float* inv_scale3_host,
                              const float* W1, const float* scale, float* gxyz, float* gW1, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N <= 0 || !delta1 || !xyz || !offset3_host || !inv_scale3_host || !W1 || !scale || !gW1) return fail(TEXGS_E_INVALID, "bad arguments");
    if ((uintptr_t)delta1 & 15) return fail(TEXGS_E_INVALID, "fp16 buffers must be 16-byte aligned");
    texgs_uvmlp_bwd_tail_kernel<<<uvbwd_ctas(N), UVBWD_THREADS, 0, stream>>>(
        N, (const __half*)delta1, xyz, make_float3(offset3_host[0], offset3_host[1], offset3_host[2]),
        make_float3(inv_scale3_host[0], inv_scale3_host[1], inv_scale3_host[2]), W1, scale, gxyz, gW1);
    TEXGS_KERNEL_CHECK("texgs_uvmlp_bwd_tail_kernel", false, stream);
    return 0;
}

__global__ void texgs_mark_visible_kernel(int P, const float* __restrict__ means3D, Mat4 view, int* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 pv = xform43(view, f3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]));
    present[idx] = (pv.z > TEXGS_NEAR) ? 1 : 0;
}

int texgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix16_host, const float* projmatrix16_host,
                       int32_t* present, void* stream_) {
    (void)projmatrix16_host;
    if (P < 0 || (P > 0 && (!means3D || !present)) || !viewmatrix16_host) return fail(TEXGS_E_INVALID, "bad arguments");
    Mat4 v;
    memcpy(v.m, viewmatrix16_host, sizeof(float) * 16);
    if (P > 0) {
        texgs_mark_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(P, means3D, v, present);
        TEXGS_KERNEL_CHECK("texgs_mark_visible_kernel", false, (cudaStream_t)stream_);
    }
    return 0;
}

}  // extern "C"
