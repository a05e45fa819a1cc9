/* ORACLE (test infrastructure, NOT product code) — scalar C restatement of the Texture-GS textured rasterizer,
 * forward and backward, spec items E1-E13 of SURVEY.md §8c.
 *
 *     *** PARITY UNPINNED *** — same status and same sources as oracle/raster_ref.py (read its header): the
 *     reference's arithmetic for this path lives in the un-vendored pip-git dependency diff_gauss_uv_tex
 *     (reference requirements.txt:15, imported at render/uv_tex_render.py:4); what the reference tree does pin
 *     (camera / pixel-centre / quaternion / SH / cube-map conventions, output semantics) is cited at each step below.
 *
 * Why a second oracle: oracle/raster_ref.py (torch, autograd) is the definition; this file is an independent,
 * hand-differentiated restatement in plain loops that (a) is checked against it to rounding in float64
 * (tests/test_oracle_c.py), (b) is fast enough (OpenMP over tiles) to check the CUDA kernels DIRECTLY at
 * BASELINE.json's full sizes instead of through properties only, and (c) is the CPU arm of bench.py.
 * Only tests/, __graft_entry__ and bench.py's CPU legs may load it (oracle/raster_c.py); the product never does.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -ffp-contract=off -DREAL=double|float raster_c.c -lm   (oracle/raster_c.py)
 *
 * Conventions (paths relative to /root/reference):
 *   row-vector camera maths p_view = [p,1] @ world_view_transform, p_clip = [p,1] @ full_proj_transform
 *       utils/cameras.py:62-65, utils/graphics.py:38-71            (matrices row-major as torch stores them)
 *   pixel centres ndc = (2 pix + 1)/S - 1                           losses/norm_reg_loss.py:25-30
 *   view ray of a pixel (ndc_x tanfovx, ndc_y tanfovy, 1)           losses/norm_reg_loss.py:30
 *   quaternion (r,x,y,z) -> R, Sigma = (R S)(R S)^T                 utils/general.py:87-119
 *   SH basis, constants, signs                                     utils/sh.py:26-112
 *   texture value -> rgb  C0 t + 0.5                                models/texture_gaussian3d.py:16-21
 *   Jacobian layout J[3i+j] = d uv_i / d x_j                        models/texture_gaussian3d.py:223-227
 *   cube faces / texel centres / index order [face,row,col,rgb]     models/modules/NVDIFFREC/util.py:94-116,
 *                                                                   NVDIFFREC/renderutils/c_src/cubemap.cu:32-61
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

#define TILE 16
#define SH_C0 ((real)0.28209479177387814)
#define SH_C1 ((real)0.4886025119029199)
static const real SH_C2[5] = {(real)1.0925484305920792, (real)-1.0925484305920792, (real)0.31539156525252005,
                              (real)-1.0925484305920792, (real)0.5462742152960396};
static const real SH_C3[7] = {(real)-0.5900435899266435, (real)2.890611442640554, (real)-0.4570457994644658,
                              (real)0.3731763325901154, (real)-0.4570457994644658, (real)1.445305721320277,
                              (real)-0.5900435899266435};
#define ALPHA_MIN ((real)1.0 / (real)255.0)
#define ALPHA_MAX ((real)0.99)
#define T_STOP ((real)1e-4)
#define NEAR_CULL ((real)0.2)
#define ND_EPS ((real)1e-8)
/* test-side conditioning flags (never change a rendered value), same constants as oracle/raster_ref.py */
#define FACE_TIE_REL ((real)1e-4)
#define TEXEL_TIE ((real)2e-6)
#define DEPTH_TIE_REL ((real)1e-6)
#define GRAZING_COS ((real)0.05)
#define FLAG_THRESHOLD 1
#define FLAG_GRAZING 2
#define FLAG_FACE_TIE 4
#define FLAG_TEXEL_TIE 8
#define FLAG_DEPTH_TIE 16
#define FLAG_RECT 32          /* a splat whose integer radius / tile rect (E3) is decided by rounding reaches this pixel */
#ifndef THRESH_ULPS
#define THRESH_ULPS 256      /* alpha / T threshold proximity, in units of rel (4e-6 in float32): see DESIGN.md section 7 */
#endif
#define FLAG_CLAMP_TIE 64      /* gradient tests only: a colour channel within rounding of the clamp max(0, .) — the value is
                                 continuous there, the clamp mask (and with it d/dSH, d/dtexture, d/duv) is not */
#define CLAMP_TIE ((real)4e-6)
#define RADIUS_TIE_REL ((real)1e-5)
#define RECT_TIE_PX ((real)1e-3)

typedef struct {
    int32_t P, M, sh_degree, H, W, R;
    int32_t threads, reserved;
    real tanfovx, tanfovy, scale_modifier;
    real view[16], proj[16], campos[3], bg[3];
    /* inputs (P,3) (P,M,3) (P) (P,3) (P,4) (P,3) (P,9) (6,R,R,3) */
    const real *xyz, *shs, *opacity, *scaling, *rotation, *uvs, *grad_uvs, *texture;
    /* outputs: (3,H,W) (H,W) (3,H,W) (H,W); radii (P); final_T (H,W); n_contrib (H,W); flags (H,W) */
    real *image, *depth, *norm, *alpha;
    int32_t* radii;
    real* final_T;
    int32_t* n_contrib;
    uint8_t* flags;
    int64_t* counters; /* [0] pairs (spec E3 tile rects), [1] visible, [2] blended contributions, [3] longest list */
    /* backward (all NULL = forward only): cotangents of the four outputs, gradients of the inputs */
    const real *g_image, *g_depth, *g_norm, *g_alpha;
    real *d_xyz, *d_means2D, *d_shs, *d_opacity, *d_scaling, *d_rotation, *d_uvs, *d_texture;
} OracleArgs;

typedef struct {
    int visible, radius, rx0, ry0, rx1, ry1, kmin;
    real xy[2], conic[3], opacity, depth, normal[3], m[3], csh[3], flip;
    real r3;            /* 3 sqrt(lambda_max) before the ceil; 0 if the Gaussian was culled before E3 */
} Proj;

typedef struct { real depth; int32_t id; } Entry;

static int cmp_entry(const void* a, const void* b) {
    const Entry *x = (const Entry*)a, *y = (const Entry*)b;
    if (x->depth < y->depth) return -1;
    if (x->depth > y->depth) return 1;
    return (x->id > y->id) - (x->id < y->id);          /* E4: ties by Gaussian index */
}

static void quat_to_rot(const real* q, real* R) {       /* utils/general.py:87-108, quaternion used as given */
    const real r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z);     R[2] = 2 * (x * z + r * y);
    R[3] = 2 * (x * y + r * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (x * z - r * y);     R[7] = 2 * (y * z + r * x);     R[8] = 1 - 2 * (x * x + y * y);
}

/* SH bands l >= 1 (utils/sh.py:57-112): basis values and their derivatives w.r.t. the direction */
static int sh_basis(int deg, real x, real y, real z, real* b, real (*d)[3]) {
    int n = 0;
    memset(b, 0, 15 * sizeof(real));
    memset(d, 0, 15 * 3 * sizeof(real));
    if (deg < 1) return 0;
    b[0] = -SH_C1 * y; d[0][1] = -SH_C1;
    b[1] = SH_C1 * z;  d[1][2] = SH_C1;
    b[2] = -SH_C1 * x; d[2][0] = -SH_C1;
    n = 3;
    if (deg > 1) {
        const real xx = x * x, yy = y * y, zz = z * z;
        b[3] = SH_C2[0] * x * y;               d[3][0] = SH_C2[0] * y;  d[3][1] = SH_C2[0] * x;
        b[4] = SH_C2[1] * y * z;               d[4][1] = SH_C2[1] * z;  d[4][2] = SH_C2[1] * y;
        b[5] = SH_C2[2] * (2 * zz - xx - yy);  d[5][0] = -2 * SH_C2[2] * x; d[5][1] = -2 * SH_C2[2] * y; d[5][2] = 4 * SH_C2[2] * z;
        b[6] = SH_C2[3] * x * z;               d[6][0] = SH_C2[3] * z;  d[6][2] = SH_C2[3] * x;
        b[7] = SH_C2[4] * (xx - yy);           d[7][0] = 2 * SH_C2[4] * x; d[7][1] = -2 * SH_C2[4] * y;
        n = 8;
        if (deg > 2) {
            b[8] = SH_C3[0] * y * (3 * xx - yy);            d[8][0] = SH_C3[0] * 6 * x * y; d[8][1] = SH_C3[0] * (3 * xx - 3 * yy);
            b[9] = SH_C3[1] * x * y * z;                    d[9][0] = SH_C3[1] * y * z; d[9][1] = SH_C3[1] * x * z; d[9][2] = SH_C3[1] * x * y;
            b[10] = SH_C3[2] * y * (4 * zz - xx - yy);      d[10][0] = -2 * SH_C3[2] * x * y; d[10][1] = SH_C3[2] * (4 * zz - xx - 3 * yy); d[10][2] = 8 * SH_C3[2] * y * z;
            b[11] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy); d[11][0] = -6 * SH_C3[3] * x * z; d[11][1] = -6 * SH_C3[3] * y * z; d[11][2] = SH_C3[3] * (6 * zz - 3 * xx - 3 * yy);
            b[12] = SH_C3[4] * x * (4 * zz - xx - yy);      d[12][0] = SH_C3[4] * (4 * zz - 3 * xx - yy); d[12][1] = -2 * SH_C3[4] * x * y; d[12][2] = 8 * SH_C3[4] * x * z;
            b[13] = SH_C3[5] * z * (xx - yy);               d[13][0] = 2 * SH_C3[5] * x * z; d[13][1] = -2 * SH_C3[5] * y * z; d[13][2] = SH_C3[5] * (xx - yy);
            b[14] = SH_C3[6] * x * (xx - 3 * yy);           d[14][0] = SH_C3[6] * (3 * xx - 3 * yy); d[14][1] = -6 * SH_C3[6] * x * y;
            n = 15;
        }
    }
    return n;
}

/* E1-E3, E8, per-Gaussian part of E12 */
static void project(const OracleArgs* a, int i, Proj* o) {
    const real* V = a->view; const real* PM = a->proj;
    const real* p = a->xyz + 3 * i;
    memset(o, 0, sizeof(*o));
    real pv[3], ph[4];
    for (int c = 0; c < 3; ++c) pv[c] = p[0] * V[c] + p[1] * V[4 + c] + p[2] * V[8 + c] + V[12 + c];
    for (int c = 0; c < 4; ++c) ph[c] = p[0] * PM[c] + p[1] * PM[4 + c] + p[2] * PM[8 + c] + PM[12 + c];
    const real pw = (real)1.0 / (ph[3] + (real)1e-7);
    o->depth = pv[2];
    const int in_front = pv[2] > NEAR_CULL;                                       /* E1 */
    real R[9];
    quat_to_rot(a->rotation + 4 * i, R);
    const real* s = a->scaling + 3 * i;
    real L[9], Sg[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) L[3 * r + c] = R[3 * r + c] * (s[c] * a->scale_modifier);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Sg[3 * r + c] = L[3 * r] * L[3 * c] + L[3 * r + 1] * L[3 * c + 1] + L[3 * r + 2] * L[3 * c + 2];
    /* E2: EWA projection */
    const real limx = (real)1.3 * a->tanfovx, limy = (real)1.3 * a->tanfovy;
    const real tz = in_front ? pv[2] : (real)1.0;
    real txr = pv[0] / tz, tyr = pv[1] / tz;
    txr = txr < -limx ? -limx : (txr > limx ? limx : txr);
    tyr = tyr < -limy ? -limy : (tyr > limy ? limy : tyr);
    const real tx = txr * tz, ty = tyr * tz;
    const real fx = (real)a->W / ((real)2.0 * a->tanfovx), fy = (real)a->H / ((real)2.0 * a->tanfovy);
    const real J[6] = {fx / tz, 0, -fx * tx / (tz * tz), 0, fy / tz, -fy * ty / (tz * tz)};
    real T[6];   /* T = J @ Wm, Wm[k][c] = V[c][k] */
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) T[3 * r + c] = J[3 * r] * V[4 * c] + J[3 * r + 1] * V[4 * c + 1] + J[3 * r + 2] * V[4 * c + 2];
    real TS[6];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) TS[3 * r + c] = T[3 * r] * Sg[c] + T[3 * r + 1] * Sg[3 + c] + T[3 * r + 2] * Sg[6 + c];
    const real c00 = TS[0] * T[0] + TS[1] * T[1] + TS[2] * T[2];
    const real c01 = TS[0] * T[3] + TS[1] * T[4] + TS[2] * T[5];
    const real c11 = TS[3] * T[3] + TS[4] * T[4] + TS[5] * T[5];
    const real ca = c00 + (real)0.3, cb = c01, cc = c11 + (real)0.3;
    const real det = ca * cc - cb * cb;
    const int det_ok = det != 0;
    const real ds = det_ok ? det : (real)1.0;
    o->conic[0] = cc / ds; o->conic[1] = -cb / ds; o->conic[2] = ca / ds;
    /* E3: radius and tile rect */
    const real mid = (real)0.5 * (ca + cc);
    real disc = mid * mid - det; if (disc < (real)0.1) disc = (real)0.1;
    const real r3 = (real)3.0 * sqrt(mid + sqrt(disc));
    const real radius = ceil(r3);
    o->xy[0] = ((ph[0] * pw + (real)1.0) * a->W - (real)1.0) * (real)0.5;
    o->xy[1] = ((ph[1] * pw + (real)1.0) * a->H - (real)1.0) * (real)0.5;
    const int gx = (a->W + TILE - 1) / TILE, gy = (a->H + TILE - 1) / TILE;
    const int fin = isfinite(o->xy[0]) && isfinite(o->xy[1]) && isfinite(radius);
    if (fin) {
        real v;
        v = floor((o->xy[0] - radius) / TILE); o->rx0 = (int)(v < 0 ? 0 : (v > gx ? gx : v));
        v = floor((o->xy[0] + radius + (TILE - 1)) / TILE); o->rx1 = (int)(v < 0 ? 0 : (v > gx ? gx : v));
        v = floor((o->xy[1] - radius) / TILE); o->ry0 = (int)(v < 0 ? 0 : (v > gy ? gy : v));
        v = floor((o->xy[1] + radius + (TILE - 1)) / TILE); o->ry1 = (int)(v < 0 ? 0 : (v > gy ? gy : v));
    }
    o->visible = in_front && det_ok && fin && (o->rx1 - o->rx0) * (o->ry1 - o->ry0) > 0;
    o->r3 = (in_front && det_ok && fin) ? r3 : 0;
    o->radius = o->visible ? (int)radius : 0;
    o->opacity = a->opacity[i];
    /* E8: disc normal = rotation column of the smallest scale (first index wins ties), facing the camera */
    int k = 0; if (s[1] < s[k]) k = 1; if (s[2] < s[k]) k = 2;
    o->kmin = k;
    for (int c = 0; c < 3; ++c) o->m[c] = p[c] - a->campos[c];
    const real nm = R[k] * o->m[0] + R[3 + k] * o->m[1] + R[6 + k] * o->m[2];
    o->flip = nm > 0 ? (real)-1.0 : (real)1.0;
    for (int c = 0; c < 3; ++c) o->normal[c] = o->flip * R[3 * c + k];
    /* E12 (per-Gaussian part): SH-rest colour in the direction camera -> Gaussian */
    const real len = sqrt(o->m[0] * o->m[0] + o->m[1] * o->m[1] + o->m[2] * o->m[2]);
    real b[15], d[15][3];
    const int nb = (a->shs && a->M > 0) ? sh_basis(a->sh_degree, o->m[0] / len, o->m[1] / len, o->m[2] / len, b, d) : 0;
    for (int kk = 0; kk < nb && kk < a->M; ++kk)
        for (int c = 0; c < 3; ++c) o->csh[c] += b[kk] * a->shs[((size_t)i * a->M + kk) * 3 + c];
}

/* cube lookup (E11): face, face coordinates, and what the backward needs */
typedef struct {
    int face, axis;          /* axis: 0 x, 1 y, 2 z */
    real inv, sx, sy, sgn_major;
    int ix, iy;              /* component feeding sx / sy */
    real sgx, sgy;           /* sx = sgx * u[ix] * inv, sy = sgy * u[iy] * inv */
} Cube;

static void cube_coords(const real* u, Cube* c) {
    const real ax = fabs(u[0]), ay = fabs(u[1]), az = fabs(u[2]);
    const int is_x = (ax >= ay) && (ax >= az), is_y = !is_x && (ay >= az);
    real m = is_x ? ax : (is_y ? ay : az);
    if (m < (real)1e-20) m = (real)1e-20;
    c->inv = (real)1.0 / m;
    if (is_x)      { c->axis = 0; c->face = u[0] < 0 ? 1 : 0; c->ix = 2; c->sgx = u[0] < 0 ? (real)1 : (real)-1; c->iy = 1; c->sgy = -1; c->sgn_major = u[0] < 0 ? (real)-1 : (real)1; }
    else if (is_y) { c->axis = 1; c->face = u[1] < 0 ? 3 : 2; c->ix = 0; c->sgx = 1; c->iy = 2; c->sgy = u[1] < 0 ? (real)-1 : (real)1; c->sgn_major = u[1] < 0 ? (real)-1 : (real)1; }
    else           { c->axis = 2; c->face = u[2] < 0 ? 5 : 4; c->ix = 0; c->sgx = u[2] < 0 ? (real)-1 : (real)1; c->iy = 1; c->sgy = -1; c->sgn_major = u[2] < 0 ? (real)-1 : (real)1; }
    c->sx = c->sgx * u[c->ix] * c->inv;
    c->sy = c->sgy * u[c->iy] * c->inv;
}

typedef struct {
    int32_t g, pos;                  /* Gaussian id, 0-based position in the tile's list */
    real alpha, T, G, dx, dy, live;  /* live: derivative of min(0.99, o G) */
    real col[3], mask[3];            /* clamped colour and the clamp mask of max(0, .) */
    /* texture path */
    int safe;
    real dw[3], nd, nmv, delta[3], u[3];
    Cube cube;
    int i00, i01, i10, i11;          /* texel indices (x3 = float offset) */
    real wx, wy, taps[4][3];
} Contrib;

static inline void atomic_add(real* p, real v) {
#pragma omp atomic
    *p += v;
}

int texgs_oracle_real_bytes(void) { return (int)sizeof(real); }

int texgs_oracle_run(const OracleArgs* a) {
    const int P = a->P, H = a->H, W = a->W, Rr = a->R, M = a->M;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
    const int backward = a->g_image || a->g_depth || a->g_norm || a->g_alpha;
    const real rel = sizeof(real) == 4 ? (real)4e-6 : (real)1e-12;
#ifdef _OPENMP
    if (a->threads > 0) omp_set_num_threads(a->threads);
#endif
    Proj* pr = (Proj*)malloc((size_t)(P > 0 ? P : 1) * sizeof(Proj));
    if (!pr) return 1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) project(a, i, &pr[i]);
    int64_t nvis = 0;
    for (int i = 0; i < P; ++i) { a->radii[i] = pr[i].radius; nvis += pr[i].visible; }

    /* E4: per-tile lists sorted by (depth, index) */
    int64_t* offs = (int64_t*)calloc((size_t)ntiles + 1, sizeof(int64_t));
    for (int i = 0; i < P; ++i)
        if (pr[i].visible)
            for (int ty = pr[i].ry0; ty < pr[i].ry1; ++ty)
                for (int tx = pr[i].rx0; tx < pr[i].rx1; ++tx) offs[ty * gx + tx + 1]++;
    int64_t longest = 0;
    for (int t = 0; t < ntiles; ++t) { if (offs[t + 1] > longest) longest = offs[t + 1]; offs[t + 1] += offs[t]; }
    const int64_t K = offs[ntiles];
    Entry* list = (Entry*)malloc((size_t)(K > 0 ? K : 1) * sizeof(Entry));
    int64_t* cur = (int64_t*)malloc((size_t)ntiles * sizeof(int64_t));
    memcpy(cur, offs, (size_t)ntiles * sizeof(int64_t));
    for (int i = 0; i < P; ++i)
        if (pr[i].visible)
            for (int ty = pr[i].ry0; ty < pr[i].ry1; ++ty)
                for (int tx = pr[i].rx0; tx < pr[i].rx1; ++tx) {
                    Entry* e = &list[cur[ty * gx + tx]++];
                    e->depth = pr[i].depth; e->id = i;
                }
#pragma omp parallel for schedule(dynamic, 8)
    for (int t = 0; t < ntiles; ++t) qsort(list + offs[t], (size_t)(offs[t + 1] - offs[t]), sizeof(Entry), cmp_entry);

    /* per-Gaussian accumulators of the backward: xy(2) conic(3) opacity(1) csh(3) depth(1) normal(3) uv(3) m(3) */
    enum { A_XY = 0, A_CON = 2, A_OP = 5, A_CSH = 6, A_Z = 9, A_N = 10, A_UV = 13, A_M = 16, NACC = 19 };
    real* acc = backward ? (real*)calloc((size_t)(P > 0 ? P : 1) * NACC, sizeof(real)) : NULL;
    if (backward && a->d_texture) memset(a->d_texture, 0, (size_t)6 * Rr * Rr * 3 * sizeof(real));
    int64_t nblend = 0;
    const real* V = a->view;

#pragma omp parallel reduction(+ : nblend)
    {
        Contrib* cs = (Contrib*)malloc((size_t)(longest > 0 ? longest : 1) * sizeof(Contrib));
        real* tacc = backward ? (real*)malloc((size_t)(longest > 0 ? longest : 1) * NACC * sizeof(real)) : NULL;
#pragma omp for schedule(dynamic, 2)
        for (int t = 0; t < ntiles; ++t) {
            const Entry* L = list + offs[t];
            const int n = (int)(offs[t + 1] - offs[t]);
            if (tacc) memset(tacc, 0, (size_t)n * NACC * sizeof(real));
            const int tx0 = (t % gx) * TILE, ty0 = (t / gx) * TILE;
            for (int py = ty0; py < ty0 + TILE && py < H; ++py)
                for (int px = tx0; px < tx0 + TILE && px < W; ++px) {
                    const int pix = py * W + px;
                    const real ndcx = ((real)2.0 * px + (real)1.0) / W - (real)1.0, ndcy = ((real)2.0 * py + (real)1.0) / H - (real)1.0;
                    const real vr[3] = {ndcx * a->tanfovx, ndcy * a->tanfovy, (real)1.0};
                    real dw[3];                                   /* world direction of the pixel ray */
                    for (int r = 0; r < 3; ++r) dw[r] = vr[0] * V[4 * r] + vr[1] * V[4 * r + 1] + vr[2] * V[4 * r + 2];
                    const real dwlen = sqrt(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
                    real T = 1, C[3] = {0, 0, 0}, D = 0, Nn[3] = {0, 0, 0}, A = 0;
                    int nc = 0, last = 0;
                    uint8_t flag = 0;
                    int prev_pos = -2;
                    for (int l = 0; l < n; ++l) {                 /* E5: front to back */
                        const int g = L[l].id;
                        const Proj* q = &pr[g];
                        const real dx = q->xy[0] - (real)px, dy = q->xy[1] - (real)py;
                        const real power = (real)-0.5 * (q->conic[0] * dx * dx + q->conic[2] * dy * dy) - q->conic[1] * dx * dy;
                        const real G = exp(power);
                        const real oG = q->opacity * G;
                        const real alpha = oG > ALPHA_MAX ? ALPHA_MAX : oG;
                        if (fabs(alpha - ALPHA_MIN) < rel * THRESH_ULPS * ALPHA_MIN) flag |= FLAG_THRESHOLD;
                        if (!(power <= 0) || !(alpha >= ALPHA_MIN)) continue;
                        const real Tn = T * ((real)1.0 - alpha);
                        if (fabs(Tn - T_STOP) < rel * THRESH_ULPS * T_STOP) flag |= FLAG_THRESHOLD;
                        if (Tn < T_STOP) break;                   /* stop BEFORE adding */
                        Contrib* c = &cs[nc++];
                        c->g = g; c->pos = l; c->alpha = alpha; c->T = T; c->G = G; c->dx = dx; c->dy = dy;
                        c->live = oG <= ALPHA_MAX ? (real)1.0 : (real)0.0;
                        /* the previous CONTRIBUTION of this pixel (not necessarily the previous list entry) at the same depth */
                        if (prev_pos >= 0 && fabs(L[l].depth - L[prev_pos].depth) <= DEPTH_TIE_REL * fabs(L[l].depth)) flag |= FLAG_DEPTH_TIE;
                        prev_pos = l;
                        /* E9: ray - disc plane intersection, E10: first-order UV step */
                        c->nd = q->normal[0] * dw[0] + q->normal[1] * dw[1] + q->normal[2] * dw[2];
                        c->nmv = q->normal[0] * q->m[0] + q->normal[1] * q->m[1] + q->normal[2] * q->m[2];
                        c->safe = fabs(c->nd) >= ND_EPS;
                        const real nlen = sqrt(q->normal[0] * q->normal[0] + q->normal[1] * q->normal[1] + q->normal[2] * q->normal[2]);
                        const real cosang = fabs(c->nd) / (dwlen * (nlen > (real)1e-30 ? nlen : (real)1e-30));
                        if (cosang < GRAZING_COS) flag |= FLAG_GRAZING;
                        const real tp = c->safe ? c->nmv / c->nd : 0;
                        const real* Jg = a->grad_uvs + (size_t)9 * g;
                        for (int k = 0; k < 3; ++k) { c->dw[k] = dw[k]; c->delta[k] = c->safe ? tp * dw[k] - q->m[k] : 0; }
                        for (int k = 0; k < 3; ++k)
                            c->u[k] = a->uvs[3 * g + k] + Jg[3 * k] * c->delta[0] + Jg[3 * k + 1] * c->delta[1] + Jg[3 * k + 2] * c->delta[2];
                        {
                            real u0 = fabs(c->u[0]), u1 = fabs(c->u[1]), u2 = fabs(c->u[2]), hi, mid2;
                            hi = u0 > u1 ? (u0 > u2 ? u0 : u2) : (u1 > u2 ? u1 : u2);
                            mid2 = u0 + u1 + u2 - hi - (u0 < u1 ? (u0 < u2 ? u0 : u2) : (u1 < u2 ? u1 : u2));
                            if (hi - mid2 < FACE_TIE_REL * hi) flag |= FLAG_FACE_TIE;
                        }
                        /* E11: cube lookup, bilinear, clamp-to-edge inside the face */
                        cube_coords(c->u, &c->cube);
                        const real fxx = (c->cube.sx + (real)1.0) * ((real)0.5 * Rr) - (real)0.5, fyy = (c->cube.sy + (real)1.0) * ((real)0.5 * Rr) - (real)0.5;
                        {
                            const real ddx = fabs(fxx - floor(fxx + (real)0.5)), ddy = fabs(fyy - floor(fyy + (real)0.5));
                            /* the intersection amplifies rounding by 1 / cos(angle(n, d)): so does the margin */
                            if ((ddx < ddy ? ddx : ddy) * (cosang > GRAZING_COS ? cosang : GRAZING_COS) < TEXEL_TIE * Rr) flag |= FLAG_TEXEL_TIE;
                        }
                        const real x0f = floor(fxx), y0f = floor(fyy);
                        c->wx = fxx - x0f; c->wy = fyy - y0f;
                        long x0 = (long)x0f, y0 = (long)y0f;
                        const long x0c = x0 < 0 ? 0 : (x0 > Rr - 1 ? Rr - 1 : x0), x1c = x0 + 1 < 0 ? 0 : (x0 + 1 > Rr - 1 ? Rr - 1 : x0 + 1);
                        const long y0c = y0 < 0 ? 0 : (y0 > Rr - 1 ? Rr - 1 : y0), y1c = y0 + 1 < 0 ? 0 : (y0 + 1 > Rr - 1 ? Rr - 1 : y0 + 1);
                        c->i00 = (int)(((long)c->cube.face * Rr + y0c) * Rr + x0c); c->i01 = (int)(((long)c->cube.face * Rr + y0c) * Rr + x1c);
                        c->i10 = (int)(((long)c->cube.face * Rr + y1c) * Rr + x0c); c->i11 = (int)(((long)c->cube.face * Rr + y1c) * Rr + x1c);
                        const int idx[4] = {c->i00, c->i01, c->i10, c->i11};
                        for (int k = 0; k < 4; ++k) for (int ch = 0; ch < 3; ++ch) c->taps[k][ch] = a->texture[(size_t)3 * idx[k] + ch];
                        const real w = alpha * T;
                        for (int ch = 0; ch < 3; ++ch) {          /* E12 */
                            const real top = c->taps[0][ch] + c->wx * (c->taps[1][ch] - c->taps[0][ch]);
                            const real bot = c->taps[2][ch] + c->wx * (c->taps[3][ch] - c->taps[2][ch]);
                            const real pre = SH_C0 * (top + c->wy * (bot - top)) + q->csh[ch] + (real)0.5;
                            if (fabs(pre) < CLAMP_TIE) flag |= FLAG_CLAMP_TIE;
                            c->mask[ch] = pre >= 0 ? (real)1.0 : (real)0.0;
                            c->col[ch] = pre >= 0 ? pre : 0;
                            C[ch] += w * c->col[ch];
                            Nn[ch] += w * q->normal[ch];          /* E8 */
                        }
                        D += w * q->depth;                        /* E7: z of the centre */
                        A += w;
                        T = Tn;
                        last = l + 1;
                    }
                    nblend += nc;
                    for (int ch = 0; ch < 3; ++ch) {              /* E6: background on the colour only */
                        a->image[(size_t)ch * H * W + pix] = C[ch] + T * a->bg[ch];
                        a->norm[(size_t)ch * H * W + pix] = Nn[ch];
                    }
                    a->depth[pix] = D; a->alpha[pix] = A;
                    if (a->final_T) a->final_T[pix] = T;
                    if (a->n_contrib) a->n_contrib[pix] = last;
                    if (a->flags) a->flags[pix] = flag;
                    if (!backward || nc == 0) continue;

                    /* E13: exact derivative of the forward above, back to front */
                    real gC[3] = {0, 0, 0}, gN[3] = {0, 0, 0}, gD = 0, gA = 0;
                    for (int ch = 0; ch < 3; ++ch) {
                        if (a->g_image) gC[ch] = a->g_image[(size_t)ch * H * W + pix];
                        if (a->g_norm) gN[ch] = a->g_norm[(size_t)ch * H * W + pix];
                    }
                    if (a->g_depth) gD = a->g_depth[pix];
                    if (a->g_alpha) gA = a->g_alpha[pix];
                    const real bgdot = gC[0] * a->bg[0] + gC[1] * a->bg[1] + gC[2] * a->bg[2];
                    real suffix = 0;                              /* sum over later contributions of w_j X_j */
                    for (int k = nc - 1; k >= 0; --k) {
                        Contrib* c = &cs[k];
                        const Proj* q = &pr[c->g];
                        real* ta = tacc + (size_t)c->pos * NACC;
                        const real w = c->alpha * c->T;
                        const real X = gC[0] * c->col[0] + gC[1] * c->col[1] + gC[2] * c->col[2] + gD * q->depth +
                                       gN[0] * q->normal[0] + gN[1] * q->normal[1] + gN[2] * q->normal[2] + gA;
                        const real inv1 = (real)1.0 / ((real)1.0 - c->alpha);
                        const real dalpha = c->T * X - suffix * inv1 - T * bgdot * inv1;      /* T here = final T of the pixel */
                        suffix += w * X;
                        const real dG = c->live * q->opacity * dalpha;
                        ta[A_OP] += c->live * c->G * dalpha;
                        const real dpow = c->G * dG;
                        ta[A_CON + 0] += (real)-0.5 * c->dx * c->dx * dpow;
                        ta[A_CON + 1] += -c->dx * c->dy * dpow;
                        ta[A_CON + 2] += (real)-0.5 * c->dy * c->dy * dpow;
                        ta[A_XY + 0] += -(q->conic[0] * c->dx + q->conic[1] * c->dy) * dpow;
                        ta[A_XY + 1] += -(q->conic[2] * c->dy + q->conic[1] * c->dx) * dpow;
                        ta[A_Z] += w * gD;
                        real gt[3];
                        for (int ch = 0; ch < 3; ++ch) {
                            ta[A_N + ch] += w * gN[ch];
                            const real gcol = w * gC[ch] * c->mask[ch];
                            ta[A_CSH + ch] += gcol;
                            gt[ch] = SH_C0 * gcol;
                        }
                        /* texture and its coordinates */
                        const real w00 = (1 - c->wx) * (1 - c->wy), w01 = c->wx * (1 - c->wy), w10 = (1 - c->wx) * c->wy, w11 = c->wx * c->wy;
                        real gwx = 0, gwy = 0;
                        for (int ch = 0; ch < 3; ++ch) {
                            const real dtop = c->taps[1][ch] - c->taps[0][ch], dbot = c->taps[3][ch] - c->taps[2][ch];
                            const real top = c->taps[0][ch] + c->wx * dtop, bot = c->taps[2][ch] + c->wx * dbot;
                            gwx += gt[ch] * (dtop + c->wy * (dbot - dtop));
                            gwy += gt[ch] * (bot - top);
                            if (a->d_texture && gt[ch] != 0) {
                                atomic_add(a->d_texture + (size_t)3 * c->i00 + ch, gt[ch] * w00);
                                atomic_add(a->d_texture + (size_t)3 * c->i01 + ch, gt[ch] * w01);
                                atomic_add(a->d_texture + (size_t)3 * c->i10 + ch, gt[ch] * w10);
                                atomic_add(a->d_texture + (size_t)3 * c->i11 + ch, gt[ch] * w11);
                            }
                        }
                        const real gsx = gwx * ((real)0.5 * Rr), gsy = gwy * ((real)0.5 * Rr);
                        real gu[3] = {0, 0, 0};
                        gu[c->cube.ix] += c->cube.sgx * c->cube.inv * gsx;
                        gu[c->cube.iy] += c->cube.sgy * c->cube.inv * gsy;
                        gu[c->cube.axis] += -(c->cube.sx * gsx + c->cube.sy * gsy) * c->cube.inv * c->cube.sgn_major;
                        for (int ch = 0; ch < 3; ++ch) ta[A_UV + ch] += gu[ch];
                        if (c->safe) {
                            const real* Jg = a->grad_uvs + (size_t)9 * c->g;
                            real gdl[3];
                            for (int j = 0; j < 3; ++j) gdl[j] = Jg[j] * gu[0] + Jg[3 + j] * gu[1] + Jg[6 + j] * gu[2];
                            const real gt_ = gdl[0] * c->dw[0] + gdl[1] * c->dw[1] + gdl[2] * c->dw[2];
                            const real gnm = gt_ / c->nd, gnd = -gt_ * c->nmv / (c->nd * c->nd);
                            for (int j = 0; j < 3; ++j) {
                                ta[A_M + j] += -gdl[j] + gnm * q->normal[j];
                                ta[A_N + j] += gnm * q->m[j] + gnd * c->dw[j];
                            }
                        }
                    }
                }
            if (tacc)
                for (int l = 0; l < n; ++l) {
                    const real* ta = tacc + (size_t)l * NACC;
                    real* ga = acc + (size_t)L[l].id * NACC;
                    for (int k = 0; k < NACC; ++k) if (ta[k] != 0) atomic_add(ga + k, ta[k]);
                }
        }
        free(cs);
        free(tacc);
    }
    if (a->counters) { a->counters[0] = K; a->counters[1] = nvis; a->counters[2] = nblend; a->counters[3] = longest; }

    /* conditioning flag for E3: radius = ceil(3 sqrt(lambda)) and the tile rect are INTEGER decisions; an implementation
     * whose fp32 rounding lands on the other side lists the splat in one more / one fewer row or column of tiles. Every
     * pixel of such a tile that the splat reaches with alpha >= 1/255 is flagged (a handful of splats per view). */
    if (a->flags)
        for (int i = 0; i < P; ++i) {
            const Proj* o = &pr[i];
            if (!(o->r3 > 0)) continue;
            const real near_int = fabs(o->r3 - floor(o->r3 + (real)0.5)) < RADIUS_TIE_REL * (o->r3 > 1 ? o->r3 : 1);
            const real rad = ceil(o->r3);
            const real r_lo = (near_int ? rad - 1 : rad) - RECT_TIE_PX, r_hi = (near_int ? rad + 1 : rad) + RECT_TIE_PX;
            int lo[4], hi[4];     /* x0 x1 y0 y1 of the smallest / largest rect within the ambiguity */
            real v;
#define CLAMPI(val, top) ((int)((val) < 0 ? 0 : ((val) > (top) ? (top) : (val))))
            v = floor((o->xy[0] - r_lo) / TILE); lo[0] = CLAMPI(v, gx); v = floor((o->xy[0] + r_lo + (TILE - 1)) / TILE); lo[1] = CLAMPI(v, gx);
            v = floor((o->xy[1] - r_lo) / TILE); lo[2] = CLAMPI(v, gy); v = floor((o->xy[1] + r_lo + (TILE - 1)) / TILE); lo[3] = CLAMPI(v, gy);
            v = floor((o->xy[0] - r_hi) / TILE); hi[0] = CLAMPI(v, gx); v = floor((o->xy[0] + r_hi + (TILE - 1)) / TILE); hi[1] = CLAMPI(v, gx);
            v = floor((o->xy[1] - r_hi) / TILE); hi[2] = CLAMPI(v, gy); v = floor((o->xy[1] + r_hi + (TILE - 1)) / TILE); hi[3] = CLAMPI(v, gy);
#undef CLAMPI
            if (lo[0] == hi[0] && lo[1] == hi[1] && lo[2] == hi[2] && lo[3] == hi[3]) continue;
            for (int ty = hi[2]; ty < hi[3]; ++ty)
                for (int tx = hi[0]; tx < hi[1]; ++tx) {
                    if (tx >= lo[0] && tx < lo[1] && ty >= lo[2] && ty < lo[3]) continue;      /* listed under every rounding */
                    for (int py = ty * TILE; py < ty * TILE + TILE && py < H; ++py)
                        for (int px = tx * TILE; px < tx * TILE + TILE && px < W; ++px) {
                            const real dx = o->xy[0] - (real)px, dy = o->xy[1] - (real)py;
                            const real power = (real)-0.5 * (o->conic[0] * dx * dx + o->conic[2] * dy * dy) - o->conic[1] * dx * dy;
                            if (power <= 0 && o->opacity * exp(power) >= (real)0.5 * ALPHA_MIN) a->flags[py * W + px] |= FLAG_RECT;
                        }
                }
        }

    if (backward) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < P; ++i) {
            real gp[3] = {0, 0, 0}, gq[4] = {0, 0, 0, 0}, gs[3] = {0, 0, 0}, gm2[3] = {0, 0, 0};
            const real* ga = acc + (size_t)i * NACC;
            const Proj* o = &pr[i];
            if (a->d_shs) memset(a->d_shs + (size_t)i * M * 3, 0, (size_t)M * 3 * sizeof(real));
            if (o->visible) {
                const real* p = a->xyz + 3 * i; const real* s = a->scaling + 3 * i; const real* PM = a->proj;
                real R[9]; quat_to_rot(a->rotation + 4 * i, R);
                real gR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gpv[3] = {0, 0, 0}, gmm[3] = {ga[A_M], ga[A_M + 1], ga[A_M + 2]};
                /* colour -> SH coefficients and view direction */
                const real len = sqrt(o->m[0] * o->m[0] + o->m[1] * o->m[1] + o->m[2] * o->m[2]);
                const real dir[3] = {o->m[0] / len, o->m[1] / len, o->m[2] / len};
                real b[15], d[15][3], gdir[3] = {0, 0, 0};
                const int nb = (a->shs && M > 0) ? sh_basis(a->sh_degree, dir[0], dir[1], dir[2], b, d) : 0;
                for (int k = 0; k < nb && k < M; ++k) {
                    real sg = 0;
                    for (int c = 0; c < 3; ++c) {
                        if (a->d_shs) a->d_shs[((size_t)i * M + k) * 3 + c] = b[k] * ga[A_CSH + c];
                        sg += a->shs[((size_t)i * M + k) * 3 + c] * ga[A_CSH + c];
                    }
                    for (int c = 0; c < 3; ++c) gdir[c] += d[k][c] * sg;
                }
                const real dd = dir[0] * gdir[0] + dir[1] * gdir[1] + dir[2] * gdir[2];
                for (int c = 0; c < 3; ++c) gmm[c] += (gdir[c] - dir[c] * dd) / len;
                /* normal -> rotation column kmin */
                for (int c = 0; c < 3; ++c) gR[3 * c + o->kmin] += o->flip * ga[A_N + c];
                gpv[2] += ga[A_Z];
                /* pixel position -> clip coordinates -> xyz */
                real ph[4];
                for (int c = 0; c < 4; ++c) ph[c] = p[0] * PM[c] + p[1] * PM[4 + c] + p[2] * PM[8 + c] + PM[12 + c];
                const real pw = (real)1.0 / (ph[3] + (real)1e-7);
                const real gnx = ga[A_XY] * (real)0.5 * W, gny = ga[A_XY + 1] * (real)0.5 * H;
                gm2[0] = gnx; gm2[1] = gny;
                const real gph[4] = {gnx * pw, gny * pw, 0, -pw * pw * (gnx * ph[0] + gny * ph[1])};
                for (int r = 0; r < 3; ++r) gp[r] += PM[4 * r] * gph[0] + PM[4 * r + 1] * gph[1] + PM[4 * r + 3] * gph[3];
                /* conic -> 2-D covariance */
                real pv[3];
                for (int c = 0; c < 3; ++c) pv[c] = p[0] * V[c] + p[1] * V[4 + c] + p[2] * V[8 + c] + V[12 + c];
                real Lm[9], Sg[9];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Lm[3 * r + c] = R[3 * r + c] * (s[c] * a->scale_modifier);
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Sg[3 * r + c] = Lm[3 * r] * Lm[3 * c] + Lm[3 * r + 1] * Lm[3 * c + 1] + Lm[3 * r + 2] * Lm[3 * c + 2];
                const real limx = (real)1.3 * a->tanfovx, limy = (real)1.3 * a->tanfovy;
                const real tz = pv[2];
                real txr = pv[0] / tz, tyr = pv[1] / tz;
                const int cxl = txr < -limx, cxh = txr > limx, cyl = tyr < -limy, cyh = tyr > limy;
                txr = cxl ? -limx : (cxh ? limx : txr); tyr = cyl ? -limy : (cyh ? limy : tyr);
                const real tx = txr * tz, ty = tyr * tz;
                const real fx = (real)W / ((real)2.0 * a->tanfovx), fy = (real)H / ((real)2.0 * a->tanfovy);
                const real J[6] = {fx / tz, 0, -fx * tx / (tz * tz), 0, fy / tz, -fy * ty / (tz * tz)};
                real T[6];
                for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) T[3 * r + c] = J[3 * r] * V[4 * c] + J[3 * r + 1] * V[4 * c + 1] + J[3 * r + 2] * V[4 * c + 2];
                real TS[6];
                for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) TS[3 * r + c] = T[3 * r] * Sg[c] + T[3 * r + 1] * Sg[3 + c] + T[3 * r + 2] * Sg[6 + c];
                const real ca = TS[0] * T[0] + TS[1] * T[1] + TS[2] * T[2] + (real)0.3;
                const real cb = TS[0] * T[3] + TS[1] * T[4] + TS[2] * T[5];
                const real cc = TS[3] * T[3] + TS[4] * T[4] + TS[5] * T[5] + (real)0.3;
                const real det = ca * cc - cb * cb;
                const real g0 = ga[A_CON], g1 = ga[A_CON + 1], g2 = ga[A_CON + 2];
                const real Sx = (cc * g0 - cb * g1 + ca * g2) / (det * det);
                const real g_a = g2 / det - Sx * cc, g_c = g0 / det - Sx * ca, g_b = -g1 / det + 2 * cb * Sx;
                /* cov = T Sigma T^T with d cov = [[g_a, g_b], [0, g_c]] */
                const real Gs[4] = {2 * g_a, g_b, g_b, 2 * g_c};                 /* G + G^T */
                real gT[6];
                for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) gT[3 * r + c] = Gs[2 * r] * TS[c] + Gs[2 * r + 1] * TS[3 + c];
                real gSg[9];                                                      /* T^T G T */
                const real Gm[4] = {g_a, g_b, 0, g_c};
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
                    gSg[3 * r + c] = T[r] * (Gm[0] * T[c] + Gm[1] * T[3 + c]) + T[3 + r] * (Gm[2] * T[c] + Gm[3] * T[3 + c]);
                real gJ[6];
                for (int r = 0; r < 2; ++r) for (int k = 0; k < 3; ++k) gJ[3 * r + k] = gT[3 * r] * V[k] + gT[3 * r + 1] * V[4 + k] + gT[3 * r + 2] * V[8 + k];
                real gtz = -fx / (tz * tz) * gJ[0] + 2 * fx * tx / (tz * tz * tz) * gJ[2] - fy / (tz * tz) * gJ[4] + 2 * fy * ty / (tz * tz * tz) * gJ[5];
                const real gtx = -fx / (tz * tz) * gJ[2], gty = -fy / (tz * tz) * gJ[5];
                if (cxl || cxh) gtz += txr * gtx; else gpv[0] += gtx;
                if (cyl || cyh) gtz += tyr * gty; else gpv[1] += gty;
                gpv[2] += gtz;
                for (int r = 0; r < 3; ++r) gp[r] += V[4 * r] * gpv[0] + V[4 * r + 1] * gpv[1] + V[4 * r + 2] * gpv[2];
                /* Sigma = L L^T, L = R diag(s mod) */
                real gL[9];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
                    gL[3 * r + c] = (gSg[3 * r] + gSg[r]) * Lm[c] + (gSg[3 * r + 1] + gSg[3 + r]) * Lm[3 + c] + (gSg[3 * r + 2] + gSg[6 + r]) * Lm[6 + c];
                for (int c = 0; c < 3; ++c) {
                    real acc_s = 0;
                    for (int r = 0; r < 3; ++r) { gR[3 * r + c] += gL[3 * r + c] * s[c] * a->scale_modifier; acc_s += gL[3 * r + c] * R[3 * r + c]; }
                    gs[c] = acc_s * a->scale_modifier;
                }
                const real* qq = a->rotation + 4 * i;
                const real r_ = qq[0], x = qq[1], y = qq[2], z = qq[3];
                gq[0] = 2 * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
                gq[1] = 2 * (y * gR[1] + z * gR[2] + y * gR[3] - 2 * x * gR[4] - r_ * gR[5] + z * gR[6] + r_ * gR[7] - 2 * x * gR[8]);
                gq[2] = 2 * (-2 * y * gR[0] + x * gR[1] + r_ * gR[2] + x * gR[3] + z * gR[5] - r_ * gR[6] + z * gR[7] - 2 * y * gR[8]);
                gq[3] = 2 * (-2 * z * gR[0] - r_ * gR[1] + x * gR[2] + r_ * gR[3] - 2 * z * gR[4] + y * gR[5] + x * gR[6] + y * gR[7]);
                for (int c = 0; c < 3; ++c) gp[c] += gmm[c];
            }
            for (int c = 0; c < 3; ++c) {
                if (a->d_xyz) a->d_xyz[3 * i + c] = gp[c];
                if (a->d_means2D) a->d_means2D[3 * i + c] = gm2[c];
                if (a->d_scaling) a->d_scaling[3 * i + c] = gs[c];
                if (a->d_uvs) a->d_uvs[3 * i + c] = o->visible ? ga[A_UV + c] : 0;
            }
            if (a->d_rotation) for (int c = 0; c < 4; ++c) a->d_rotation[4 * i + c] = gq[c];
            if (a->d_opacity) a->d_opacity[i] = o->visible ? ga[A_OP] : 0;
        }
    }
    free(acc); free(cur); free(list); free(offs); free(pr);
    return 0;
}
