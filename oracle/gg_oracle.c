/* TEST INFRASTRUCTURE -- plain-C restatement of the Gaussian-splat rasterizer hot path.
 *
 * PARITY UNPINNED: the reference's rasterizer is the third-party, un-vendored and un-pinned
 * CUDA extension `diff_gaussian_rasterization_depth_alpha`
 * (/root/reference/setup.sh:26-29; imported at /root/reference/gaussian_renderer/__init__.py:16).
 * Its sources are not under /root/reference, so this file restates the published 3DGS
 * rasterization algorithm (+ depth and accumulated-alpha channels), anchored on the
 * reference's call sites (gaussian_renderer/__init__.py:39-54 settings, :103-111 call and
 * 4-tuple return) and on the in-tree statements of the same math:
 *   SH basis / constants            utils/sh_utils.py:25-42,56-111   (+0.5, clamp: gaussian_renderer/__init__.py:84-85)
 *   Sigma3D = (R S)(R S)^T, packing scene/gaussian_model.py:27-31, utils/general_utils.py:74-120
 *   view / projection matrices      scene/cameras.py:53-62, utils/graphics_utils.py:38-81
 *   means2D-gradient consumer       scene/gaussian_model.py:410-412
 * It is validated against oracle/torch_oracle.py (autograd) and the golden fixtures in
 * tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product path never does.
 *
 * Forward arithmetic is fp32 (same operation order as the CUDA kernels where that is cheap);
 * per-Gaussian gradient sums are accumulated in double so that this is the more accurate side
 * of every comparison.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define NEAR_Z 0.2f
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_STOP 0.0001f
#define BLUR 0.3f

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    int32_t N, M, D, W, H;
    float tanfovx, tanfovy, scale_modifier;
    float bg[3];
    float viewmatrix[16]; /* flat row-major of the torch tensor world_view_transform (= W2C^T) */
    float projmatrix[16]; /* flat row-major of full_proj_transform (= (P W2C)^T) */
    float campos[3];
    int32_t prefiltered;
} ggo_params;

typedef struct {
    ggo_params p;
    int gx, gy, T;
    int64_t K;
    /* per Gaussian */
    float *xy, *depth, *conic_o, *rgb, *cov6;
    uint8_t* clamped;
    int32_t* radii;
    int32_t* rect; /* x0,y0,x1,y1 */
    /* inputs kept by pointer value copies */
    float *means3D, *shs, *colors, *opac, *scales, *rots, *cov_pre;
    /* binning */
    int64_t* tile_off; /* T+1 */
    uint32_t* inst;    /* K sorted gaussian ids */
    /* per pixel */
    int32_t* n_contrib;
    float* final_T;
} ggo_state;

static void* dupf(const float* src, size_t n) {
    if (!src || n == 0) return NULL;
    float* d = (float*)malloc(n * sizeof(float));
    memcpy(d, src, n * sizeof(float));
    return d;
}

void ggo_free(ggo_state* s) {
    if (!s) return;
    free(s->xy); free(s->depth); free(s->conic_o); free(s->rgb); free(s->cov6); free(s->clamped);
    free(s->radii); free(s->rect); free(s->means3D); free(s->shs); free(s->colors); free(s->opac);
    free(s->scales); free(s->rots); free(s->cov_pre); free(s->tile_off); free(s->inst);
    free(s->n_contrib); free(s->final_T);
    free(s);
}

int64_t ggo_num_rendered(const ggo_state* s) { return s ? s->K : -1; }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* c6) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                  2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                  2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
    float M[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i * 3 + j] = R[i * 3 + j] * (mod * s[j]);
    float S[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            S[i * 3 + j] = M[i * 3 + 0] * M[j * 3 + 0] + M[i * 3 + 1] * M[j * 3 + 1] + M[i * 3 + 2] * M[j * 3 + 2];
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

/* SH -> RGB (before +0.5 / clamp) for one Gaussian; shs is [M][3] */
static void sh_eval(int deg, const float* sh, float x, float y, float z, float* out) {
    for (int c = 0; c < 3; c++) {
        float r = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.f * zz - xx - yy) * sh[13 * 3 + c] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.f * yy) * sh[15 * 3 + c];
                }
            }
        }
        out[c] = r;
    }
}

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* ------------------------------------------------------------------------------------------- */
int ggo_forward(const ggo_params* pp, const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, const float* rotations,
                const float* cov3D_precomp, const float* means2D_offset /* [N,3] or NULL */,
                float* out_color, float* out_depth, float* out_alpha, int32_t* out_radii,
                uint8_t* out_fragile /* [H*W] or NULL */, float fragile_eps, ggo_state** out_state) {
    const ggo_params p = *pp;
    const int N = p.N, W = p.W, H = p.H;
    ggo_state* s = (ggo_state*)calloc(1, sizeof(ggo_state));
    s->p = p;
    s->gx = (W + TILE - 1) / TILE;
    s->gy = (H + TILE - 1) / TILE;
    s->T = s->gx * s->gy;
    const int gx = s->gx, gy = s->gy, T = s->T;
    const size_t P = (size_t)W * H;
    memset(out_color, 0, 3 * P * sizeof(float));
    memset(out_depth, 0, P * sizeof(float));
    memset(out_alpha, 0, P * sizeof(float));
    if (out_fragile) memset(out_fragile, 0, P);
    s->n_contrib = (int32_t*)calloc(P ? P : 1, sizeof(int32_t));
    s->final_T = (float*)malloc((P ? P : 1) * sizeof(float));
    for (size_t i = 0; i < P; i++) s->final_T[i] = 1.f;
    s->tile_off = (int64_t*)calloc((size_t)T + 1, sizeof(int64_t));
    if (N == 0) { /* upstream returns an all-zero image when there are no Gaussians */
        s->K = 0;
        *out_state = s;
        return 0;
    }
    s->xy = (float*)calloc((size_t)N * 2, 4);
    s->depth = (float*)calloc(N, 4);
    s->conic_o = (float*)calloc((size_t)N * 4, 4);
    s->rgb = (float*)calloc((size_t)N * 3, 4);
    s->cov6 = (float*)calloc((size_t)N * 6, 4);
    s->clamped = (uint8_t*)calloc((size_t)N * 3, 1);
    s->radii = (int32_t*)calloc(N, 4);
    s->rect = (int32_t*)calloc((size_t)N * 4, 4);
    s->means3D = dupf(means3D, (size_t)N * 3);
    s->shs = dupf(shs, shs ? (size_t)N * p.M * 3 : 0);
    s->colors = dupf(colors_precomp, colors_precomp ? (size_t)N * 3 : 0);
    s->opac = dupf(opacities, N);
    s->scales = dupf(scales, scales ? (size_t)N * 3 : 0);
    s->rots = dupf(rotations, rotations ? (size_t)N * 4 : 0);
    s->cov_pre = dupf(cov3D_precomp, cov3D_precomp ? (size_t)N * 6 : 0);

    const float* V = p.viewmatrix;
    const float* Pm = p.projmatrix;
    const float focal_x = W / (2.0f * p.tanfovx), focal_y = H / (2.0f * p.tanfovy);

    /* ---- a5: per-Gaussian preprocess ---- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        s->radii[i] = 0;
        const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
        float tvx = V[0] * x + V[4] * y + V[8] * z + V[12];
        float tvy = V[1] * x + V[5] * y + V[9] * z + V[13];
        float tvz = fmaf(V[10], z, fmaf(V[6], y, fmaf(V[2], x, V[14]))); /* sort key: same fused chain as the CUDA kernel */
        if (!(tvz > NEAR_Z)) continue;
        float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
        float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
        float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
        float pw = 1.0f / (hw + 0.0000001f);
        float ndx = hx * pw, ndy = hy * pw;
        float* c6 = s->cov6 + 6 * (size_t)i;
        if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * (size_t)i, 24);
        else cov3d_from_scale_rot(scales + 3 * (size_t)i, p.scale_modifier, rotations + 4 * (size_t)i, c6);
        /* EWA projection */
        const float limx = 1.3f * p.tanfovx, limy = 1.3f * p.tanfovy;
        float txtz = tvx / tvz, tytz = tvy / tvz;
        float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz;
        float ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
        float J00 = focal_x / tvz, J02 = -(focal_x * tx) / (tvz * tvz);
        float J11 = focal_y / tvz, J12 = -(focal_y * ty) / (tvz * tvz);
        /* W3[r][c] = V[c*4+r];  Tm = J * W3 (2x3) */
        float T00 = J00 * V[0] + J02 * V[2], T01 = J00 * V[4] + J02 * V[6], T02 = J00 * V[8] + J02 * V[10];
        float T10 = J11 * V[1] + J12 * V[2], T11 = J11 * V[5] + J12 * V[6], T12 = J11 * V[9] + J12 * V[10];
        float S00 = c6[0], S01 = c6[1], S02 = c6[2], S11 = c6[3], S12 = c6[4], S22 = c6[5];
        float u0 = T00 * S00 + T01 * S01 + T02 * S02, u1 = T00 * S01 + T01 * S11 + T02 * S12,
              u2 = T00 * S02 + T01 * S12 + T02 * S22;
        float v0 = T10 * S00 + T11 * S01 + T12 * S02, v1 = T10 * S01 + T11 * S11 + T12 * S12,
              v2 = T10 * S02 + T11 * S12 + T12 * S22;
        float a = u0 * T00 + u1 * T01 + u2 * T02 + BLUR;
        float b = u0 * T10 + u1 * T11 + u2 * T12;
        float c = v0 * T10 + v1 * T11 + v2 * T12 + BLUR;
        float det = a * c - b * b;
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float mid = 0.5f * (a + c);
        float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float rad = ceilf(3.f * sqrtf(lam));
        float px = ((ndx + 1.0f) * W - 1.0f) * 0.5f, py = ((ndy + 1.0f) * H - 1.0f) * 0.5f;
        if (!(isfinite(rad) && isfinite(px) && isfinite(py))) continue;
        if (means2D_offset) { /* means2D is a zero-valued gradient sink in the reference; kept for oracle symmetry */
            px += means2D_offset[3 * (size_t)i] * 0.5f * W;
            py += means2D_offset[3 * (size_t)i + 1] * 0.5f * H;
        }
        int x0 = (int)((px - rad) / (float)TILE), y0 = (int)((py - rad) / (float)TILE);
        int x1 = (int)((px + rad + (float)(TILE - 1)) / (float)TILE), y1 = (int)((py + rad + (float)(TILE - 1)) / (float)TILE);
        x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0); x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
        y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0); y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        /* colour */
        float* rgb = s->rgb + 3 * (size_t)i;
        if (colors_precomp) {
            rgb[0] = colors_precomp[3 * (size_t)i]; rgb[1] = colors_precomp[3 * (size_t)i + 1]; rgb[2] = colors_precomp[3 * (size_t)i + 2];
        } else {
            float dx = x - p.campos[0], dy = y - p.campos[1], dz = z - p.campos[2];
            float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float res[3];
            sh_eval(p.D, shs + (size_t)i * p.M * 3, dx * inv, dy * inv, dz * inv, res);
            for (int ch = 0; ch < 3; ch++) {
                float r = res[ch] + 0.5f;
                s->clamped[3 * (size_t)i + ch] = (r < 0.f);
                rgb[ch] = fmaxf(r, 0.f);
            }
        }
        s->depth[i] = tvz;
        s->radii[i] = (int)rad;
        s->xy[2 * (size_t)i] = px; s->xy[2 * (size_t)i + 1] = py;
        s->conic_o[4 * (size_t)i] = c * det_inv; s->conic_o[4 * (size_t)i + 1] = -b * det_inv;
        s->conic_o[4 * (size_t)i + 2] = a * det_inv; s->conic_o[4 * (size_t)i + 3] = opacities[i];
        s->rect[4 * (size_t)i] = x0; s->rect[4 * (size_t)i + 1] = y0; s->rect[4 * (size_t)i + 2] = x1; s->rect[4 * (size_t)i + 3] = y1;
    }
    memcpy(out_radii, s->radii, (size_t)N * 4);

    /* ---- a6-a8: binning.  Bucket by tile (Gaussian order), then order each tile by
     * (depth bits, Gaussian index) -- identical to a stable sort of (tile<<32 | depth) keys. ---- */
    int64_t* cnt = s->tile_off; /* use as counts first (shifted by one) */
    for (int i = 0; i < N; i++) {
        if (s->radii[i] <= 0) continue;
        const int32_t* r = s->rect + 4 * (size_t)i;
        for (int ty = r[1]; ty < r[3]; ty++)
            for (int tx = r[0]; tx < r[2]; tx++) cnt[(size_t)ty * gx + tx + 1]++;
    }
    for (int t = 0; t < T; t++) cnt[t + 1] += cnt[t];
    const int64_t K = cnt[T];
    s->K = K;
    uint64_t* keys = (uint64_t*)malloc((K ? K : 1) * sizeof(uint64_t));
    int64_t* fill = (int64_t*)calloc((size_t)T, sizeof(int64_t));
    for (int i = 0; i < N; i++) {
        if (s->radii[i] <= 0) continue;
        const int32_t* r = s->rect + 4 * (size_t)i;
        uint64_t key = ((uint64_t)f2u(s->depth[i]) << 32) | (uint32_t)i;
        for (int ty = r[1]; ty < r[3]; ty++)
            for (int tx = r[0]; tx < r[2]; tx++) {
                size_t t = (size_t)ty * gx + tx;
                keys[s->tile_off[t] + fill[t]++] = key;
            }
    }
    free(fill);
    s->inst = (uint32_t*)malloc((K ? K : 1) * sizeof(uint32_t));
#pragma omp parallel for schedule(dynamic, 8)
    for (int t = 0; t < T; t++) {
        int64_t o = s->tile_off[t], n = s->tile_off[t + 1] - o;
        if (n > 1) qsort(keys + o, (size_t)n, sizeof(uint64_t), cmp_u64);
        for (int64_t j = 0; j < n; j++) s->inst[o + j] = (uint32_t)(keys[o + j] & 0xffffffffu);
    }
    free(keys);

    /* ---- a9: per-tile front-to-back blend ---- */
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < T; t++) {
        const int tyi = t / gx, txi = t % gx;
        const int64_t o = s->tile_off[t], n = s->tile_off[t + 1] - o;
        for (int ly = 0; ly < TILE; ly++) {
            const int py = tyi * TILE + ly;
            if (py >= H) break;
            for (int lx = 0; lx < TILE; lx++) {
                const int px = txi * TILE + lx;
                if (px >= W) break;
                const size_t pid = (size_t)py * W + px;
                float Tr = 1.f, C0 = 0, C1 = 0, C2 = 0, Dp = 0, Ac = 0;
                int last = 0;
                uint8_t frag = 0;
                for (int64_t j = 0; j < n; j++) {
                    const uint32_t g = s->inst[o + j];
                    const float dx = s->xy[2 * (size_t)g] - (float)px, dy = s->xy[2 * (size_t)g + 1] - (float)py;
                    const float* co = s->conic_o + 4 * (size_t)g;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.f) continue;
                    const float alpha = fminf(ALPHA_MAX, co[3] * expf(power));
                    if (out_fragile && fabsf(alpha * 255.f - 1.f) < fragile_eps) frag = 1;
                    if (alpha < ALPHA_MIN) continue;
                    const float test_T = Tr * (1.f - alpha);
                    if (out_fragile && fabsf(test_T / T_STOP - 1.f) < fragile_eps) frag = 1;
                    if (test_T < T_STOP) break;
                    const float w = alpha * Tr;
                    const float* rgb = s->rgb + 3 * (size_t)g;
                    C0 += rgb[0] * w; C1 += rgb[1] * w; C2 += rgb[2] * w;
                    Dp += s->depth[g] * w;
                    Ac += w;
                    Tr = test_T;
                    last = (int)j + 1;
                }
                s->final_T[pid] = Tr;
                s->n_contrib[pid] = last;
                out_color[pid] = C0 + Tr * p.bg[0];
                out_color[P + pid] = C1 + Tr * p.bg[1];
                out_color[2 * P + pid] = C2 + Tr * p.bg[2];
                out_depth[pid] = Dp;
                out_alpha[pid] = Ac;
                if (out_fragile) out_fragile[pid] = frag;
            }
        }
    }
    *out_state = s;
    return 0;
}

/* expose intermediates for stage-wise parity tests */
int ggo_get_geom(const ggo_state* s, float* xy, float* depth, float* conic_o, float* rgb, int32_t* rect) {
    const size_t N = (size_t)s->p.N;
    if (N == 0) return 0;
    if (xy) memcpy(xy, s->xy, N * 8);
    if (depth) memcpy(depth, s->depth, N * 4);
    if (conic_o) memcpy(conic_o, s->conic_o, N * 16);
    if (rgb) memcpy(rgb, s->rgb, N * 12);
    if (rect) memcpy(rect, s->rect, N * 16);
    return 0;
}
int ggo_get_binning(const ggo_state* s, int64_t* tile_off, uint32_t* inst, int32_t* n_contrib, float* final_T) {
    if (tile_off) memcpy(tile_off, s->tile_off, ((size_t)s->T + 1) * 8);
    if (inst && s->K) memcpy(inst, s->inst, (size_t)s->K * 4);
    const size_t P = (size_t)s->p.W * s->p.H;
    if (n_contrib) memcpy(n_contrib, s->n_contrib, P * 4);
    if (final_T) memcpy(final_T, s->final_T, P * 4);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* a10 + a11.  Any output pointer may be NULL. dL_dmeans2D is [N,3] (z = 0). */
static int backward_impl(const ggo_state* s, int forward_order, const float* dL_dcolor, const float* dL_ddepth,
                         const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                         float* dL_dcolors_precomp, float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                         float* dL_dcov3D) {
    const ggo_params p = s->p;
    const int N = p.N, W = p.W, H = p.H, gx = s->gx, T = s->T;
    const size_t P = (size_t)W * H;
    if (N == 0) return 0;
    /* per-Gaussian accumulators: mean2D(2) conic(3) opacity(1) rgb(3) depth(1) */
    double* acc = (double*)calloc((size_t)N * 10, sizeof(double));

#pragma omp parallel
    {
        double* loc = NULL;
        size_t loc_cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int t = 0; t < T; t++) {
            const int tyi = t / gx, txi = t % gx;
            const int64_t o = s->tile_off[t], n = s->tile_off[t + 1] - o;
            if (n == 0) continue;
            if ((size_t)n * 10 > loc_cap) {
                loc_cap = (size_t)n * 10;
                loc = (double*)realloc(loc, loc_cap * sizeof(double));
            }
            memset(loc, 0, (size_t)n * 10 * sizeof(double));
            for (int ly = 0; ly < TILE; ly++) {
                const int py = tyi * TILE + ly;
                if (py >= H) break;
                for (int lx = 0; lx < TILE; lx++) {
                    const int px = txi * TILE + lx;
                    if (px >= W) break;
                    const size_t pid = (size_t)py * W + px;
                    const float T_final = s->final_T[pid];
                    float Tr = T_final;
                    const float gC[3] = {dL_dcolor ? dL_dcolor[pid] : 0.f, dL_dcolor ? dL_dcolor[P + pid] : 0.f,
                                         dL_dcolor ? dL_dcolor[2 * P + pid] : 0.f};
                    const float gD = dL_ddepth ? dL_ddepth[pid] : 0.f, gA = dL_dalpha ? dL_dalpha[pid] : 0.f;
                    const float bg_dot = p.bg[0] * gC[0] + p.bg[1] * gC[1] + p.bg[2] * gC[2];
                    float accum_c[3] = {0, 0, 0}, accum_d = 0, accum_a = 0;
                    float last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
                    if (forward_order) {
                        /* Front-to-back formulation (what a per-Gaussian-parallel backward needs: no "colour behind"
                         * recurrence, only running prefixes and the pixel's totals).  With v_i = (c_i, z_i, 1),
                         * upstream g = (gC, gD, gA), w_i = alpha_i T_i, U_i = sum_{j<=i} w_j (g.v_j) and
                         * TOT = sum_j w_j (g.v_j) + T_final (gC.bg):
                         *     dL/dalpha_i = T_i (g.v_i) - (TOT - U_i) / (1 - alpha_i)                          */
                        double tot = (double)T_final * bg_dot;
                        {
                            float Tt = 1.f;
                            for (int64_t j = 0; j < s->n_contrib[pid]; j++) {
                                const uint32_t g = s->inst[o + j];
                                const float dx = s->xy[2 * (size_t)g] - (float)px, dy = s->xy[2 * (size_t)g + 1] - (float)py;
                                const float* co = s->conic_o + 4 * (size_t)g;
                                const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                                if (power > 0.f) continue;
                                const float alpha = fminf(ALPHA_MAX, co[3] * expf(power));
                                if (alpha < ALPHA_MIN) continue;
                                const float* rgb = s->rgb + 3 * (size_t)g;
                                const float gv = rgb[0] * gC[0] + rgb[1] * gC[1] + rgb[2] * gC[2] + s->depth[g] * gD + gA;
                                tot += (double)(alpha * Tt) * gv;
                                Tt *= (1.f - alpha);
                            }
                        }
                        float Tt = 1.f;
                        double upto = 0.0;
                        for (int64_t j = 0; j < s->n_contrib[pid]; j++) {
                            const uint32_t g = s->inst[o + j];
                            const float dx = s->xy[2 * (size_t)g] - (float)px, dy = s->xy[2 * (size_t)g + 1] - (float)py;
                            const float* co = s->conic_o + 4 * (size_t)g;
                            const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                            if (power > 0.f) continue;
                            const float G = expf(power);
                            const float alpha = fminf(ALPHA_MAX, co[3] * G);
                            if (alpha < ALPHA_MIN) continue;
                            const float w = alpha * Tt;
                            const float* rgb = s->rgb + 3 * (size_t)g;
                            const float gv = rgb[0] * gC[0] + rgb[1] * gC[1] + rgb[2] * gC[2] + s->depth[g] * gD + gA;
                            upto += (double)w * gv;
                            const float dL_dalpha_ = Tt * gv - (float)((tot - upto) / (double)(1.f - alpha));
                            double* L = loc + (size_t)j * 10;
                            L[6] += (double)(w * gC[0]); L[7] += (double)(w * gC[1]); L[8] += (double)(w * gC[2]);
                            L[9] += (double)(w * gD);
                            const float dL_dG = co[3] * dL_dalpha_;
                            const float gdx = G * dx, gdy = G * dy;
                            L[0] += (double)(dL_dG * (-gdx * co[0] - gdy * co[1]) * 0.5f * W);
                            L[1] += (double)(dL_dG * (-gdy * co[2] - gdx * co[1]) * 0.5f * H);
                            L[2] += (double)(-0.5f * gdx * dx * dL_dG);
                            L[3] += (double)(-gdx * dy * dL_dG);
                            L[4] += (double)(-0.5f * gdy * dy * dL_dG);
                            L[5] += (double)(G * dL_dalpha_);
                            Tt *= (1.f - alpha);
                        }
                        continue;
                    }
                    for (int64_t j = s->n_contrib[pid] - 1; j >= 0; j--) {
                        const uint32_t g = s->inst[o + j];
                        const float dx = s->xy[2 * (size_t)g] - (float)px, dy = s->xy[2 * (size_t)g + 1] - (float)py;
                        const float* co = s->conic_o + 4 * (size_t)g;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0.f) continue;
                        const float G = expf(power);
                        const float alpha = fminf(ALPHA_MAX, co[3] * G);
                        if (alpha < ALPHA_MIN) continue;
                        Tr = Tr / (1.f - alpha);
                        const float w = alpha * Tr;
                        double* L = loc + (size_t)j * 10;
                        float dL_dalpha_ = 0.f;
                        const float* rgb = s->rgb + 3 * (size_t)g;
                        for (int ch = 0; ch < 3; ch++) {
                            accum_c[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum_c[ch];
                            last_c[ch] = rgb[ch];
                            dL_dalpha_ += (rgb[ch] - accum_c[ch]) * gC[ch];
                            L[6 + ch] += (double)(w * gC[ch]);
                        }
                        const float dep = s->depth[g];
                        accum_d = last_alpha * last_d + (1.f - last_alpha) * accum_d;
                        last_d = dep;
                        dL_dalpha_ += (dep - accum_d) * gD;
                        L[9] += (double)(w * gD);
                        accum_a = last_alpha * 1.f + (1.f - last_alpha) * accum_a;
                        dL_dalpha_ += (1.f - accum_a) * gA;
                        dL_dalpha_ *= Tr;
                        last_alpha = alpha;
                        dL_dalpha_ += (-T_final / (1.f - alpha)) * bg_dot;
                        const float dL_dG = co[3] * dL_dalpha_;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                        const float dG_ddely = -gdy * co[2] - gdx * co[1];
                        L[0] += (double)(dL_dG * dG_ddelx * 0.5f * W);
                        L[1] += (double)(dL_dG * dG_ddely * 0.5f * H);
                        L[2] += (double)(-0.5f * gdx * dx * dL_dG);
                        L[3] += (double)(-gdx * dy * dL_dG); /* full d/dB */
                        L[4] += (double)(-0.5f * gdy * dy * dL_dG);
                        L[5] += (double)(G * dL_dalpha_);
                    }
                }
            }
            for (int64_t j = 0; j < n; j++) {
                const uint32_t g = s->inst[o + j];
                for (int k = 0; k < 10; k++) {
                    const double v = loc[(size_t)j * 10 + k];
                    if (v != 0.0) {
#pragma omp atomic
                        acc[(size_t)g * 10 + k] += v;
                    }
                }
            }
        }
        free(loc);
    }

    const float* V = p.viewmatrix;
    const float* Pm = p.projmatrix;
    const float focal_x = W / (2.0f * p.tanfovx), focal_y = H / (2.0f * p.tanfovy);
    const int Mc = p.M;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        float gm[3] = {0, 0, 0};
        if (dL_dmeans2D) { dL_dmeans2D[3 * (size_t)i] = 0; dL_dmeans2D[3 * (size_t)i + 1] = 0; dL_dmeans2D[3 * (size_t)i + 2] = 0; }
        if (dL_dshs) memset(dL_dshs + (size_t)i * Mc * 3, 0, (size_t)Mc * 12);
        if (dL_dcolors_precomp) memset(dL_dcolors_precomp + 3 * (size_t)i, 0, 12);
        if (dL_dopacities) dL_dopacities[i] = 0;
        if (dL_dscales) memset(dL_dscales + 3 * (size_t)i, 0, 12);
        if (dL_drotations) memset(dL_drotations + 4 * (size_t)i, 0, 16);
        if (dL_dcov3D) memset(dL_dcov3D + 6 * (size_t)i, 0, 24);
        if (dL_dmeans3D) memset(dL_dmeans3D + 3 * (size_t)i, 0, 12);
        if (s->radii[i] <= 0) continue;
        const double* A = acc + (size_t)i * 10;
        const float g2x = (float)A[0], g2y = (float)A[1];
        const float gcA = (float)A[2], gcB = (float)A[3], gcC = (float)A[4];
        const float g_op = (float)A[5];
        const float g_rgb[3] = {(float)A[6], (float)A[7], (float)A[8]};
        const float g_dep = (float)A[9];
        const float x = s->means3D[3 * (size_t)i], y = s->means3D[3 * (size_t)i + 1], z = s->means3D[3 * (size_t)i + 2];
        if (dL_dmeans2D) { dL_dmeans2D[3 * (size_t)i] = g2x; dL_dmeans2D[3 * (size_t)i + 1] = g2y; }
        if (dL_dopacities) dL_dopacities[i] = g_op;

        /* ---- colour ---- */
        if (s->colors) {
            if (dL_dcolors_precomp) for (int ch = 0; ch < 3; ch++) dL_dcolors_precomp[3 * (size_t)i + ch] = g_rgb[ch];
        } else {
            float vx = x - p.campos[0], vy = y - p.campos[1], vz = z - p.campos[2];
            float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
            float dxn = vx * inv, dyn = vy * inv, dzn = vz * inv;
            float gr[3];
            for (int ch = 0; ch < 3; ch++) gr[ch] = s->clamped[3 * (size_t)i + ch] ? 0.f : g_rgb[ch];
            const float* sh = s->shs + (size_t)i * Mc * 3;
            float* gsh = dL_dshs ? dL_dshs + (size_t)i * Mc * 3 : NULL;
            float dRx[3] = {0, 0, 0}, dRy[3] = {0, 0, 0}, dRz[3] = {0, 0, 0};
            float basis[16];
            int nb = (p.D + 1) * (p.D + 1);
            {
                float xx = dxn * dxn, yy = dyn * dyn, zz = dzn * dzn, xy = dxn * dyn, yz = dyn * dzn, xz = dxn * dzn;
                basis[0] = SH_C0;
                basis[1] = -SH_C1 * dyn; basis[2] = SH_C1 * dzn; basis[3] = -SH_C1 * dxn;
                basis[4] = SH_C2[0] * xy; basis[5] = SH_C2[1] * yz; basis[6] = SH_C2[2] * (2.f * zz - xx - yy);
                basis[7] = SH_C2[3] * xz; basis[8] = SH_C2[4] * (xx - yy);
                basis[9] = SH_C3[0] * dyn * (3.f * xx - yy); basis[10] = SH_C3[1] * xy * dzn;
                basis[11] = SH_C3[2] * dyn * (4.f * zz - xx - yy);
                basis[12] = SH_C3[3] * dzn * (2.f * zz - 3.f * xx - 3.f * yy);
                basis[13] = SH_C3[4] * dxn * (4.f * zz - xx - yy); basis[14] = SH_C3[5] * dzn * (xx - yy);
                basis[15] = SH_C3[6] * dxn * (xx - 3.f * yy);
                if (gsh) for (int k = 0; k < nb; k++) for (int ch = 0; ch < 3; ch++) gsh[k * 3 + ch] = basis[k] * gr[ch];
                for (int ch = 0; ch < 3; ch++) {
                    const float* h = sh + ch; /* h[k*3] */
                    float ddx = 0, ddy = 0, ddz = 0;
                    if (p.D > 0) {
                        ddx += -SH_C1 * h[3 * 3]; ddy += -SH_C1 * h[1 * 3]; ddz += SH_C1 * h[2 * 3];
                        if (p.D > 1) {
                            ddx += SH_C2[0] * dyn * h[4 * 3] + SH_C2[2] * (-2.f * dxn) * h[6 * 3] + SH_C2[3] * dzn * h[7 * 3] + SH_C2[4] * 2.f * dxn * h[8 * 3];
                            ddy += SH_C2[0] * dxn * h[4 * 3] + SH_C2[1] * dzn * h[5 * 3] + SH_C2[2] * (-2.f * dyn) * h[6 * 3] + SH_C2[4] * (-2.f * dyn) * h[8 * 3];
                            ddz += SH_C2[1] * dyn * h[5 * 3] + SH_C2[2] * 4.f * dzn * h[6 * 3] + SH_C2[3] * dxn * h[7 * 3];
                            if (p.D > 2) {
                                ddx += SH_C3[0] * h[9 * 3] * 6.f * xy + SH_C3[1] * h[10 * 3] * yz + SH_C3[2] * h[11 * 3] * (-2.f * xy) +
                                       SH_C3[3] * h[12 * 3] * (-6.f * xz) + SH_C3[4] * h[13 * 3] * (4.f * zz - 3.f * xx - yy) +
                                       SH_C3[5] * h[14 * 3] * 2.f * xz + SH_C3[6] * h[15 * 3] * (3.f * xx - 3.f * yy);
                                ddy += SH_C3[0] * h[9 * 3] * (3.f * xx - 3.f * yy) + SH_C3[1] * h[10 * 3] * xz +
                                       SH_C3[2] * h[11 * 3] * (4.f * zz - xx - 3.f * yy) + SH_C3[3] * h[12 * 3] * (-6.f * yz) +
                                       SH_C3[4] * h[13 * 3] * (-2.f * xy) + SH_C3[5] * h[14 * 3] * (-2.f * yz) + SH_C3[6] * h[15 * 3] * (-6.f * xy);
                                ddz += SH_C3[1] * h[10 * 3] * xy + SH_C3[2] * h[11 * 3] * 8.f * yz +
                                       SH_C3[3] * h[12 * 3] * (6.f * zz - 3.f * xx - 3.f * yy) + SH_C3[4] * h[13 * 3] * 8.f * xz +
                                       SH_C3[5] * h[14 * 3] * (xx - yy);
                            }
                        }
                    }
                    dRx[ch] = ddx; dRy[ch] = ddy; dRz[ch] = ddz;
                }
            }
            float gdx = dRx[0] * gr[0] + dRx[1] * gr[1] + dRx[2] * gr[2];
            float gdy = dRy[0] * gr[0] + dRy[1] * gr[1] + dRy[2] * gr[2];
            float gdz = dRz[0] * gr[0] + dRz[1] * gr[1] + dRz[2] * gr[2];
            float dot = dxn * gdx + dyn * gdy + dzn * gdz;
            gm[0] += (gdx - dxn * dot) * inv; gm[1] += (gdy - dyn * dot) * inv; gm[2] += (gdz - dzn * dot) * inv;
        }

        /* ---- conic -> Sigma2D ---- */
        float tvx = V[0] * x + V[4] * y + V[8] * z + V[12];
        float tvy = V[1] * x + V[5] * y + V[9] * z + V[13];
        float tvz = fmaf(V[10], z, fmaf(V[6], y, fmaf(V[2], x, V[14]))); /* sort key: same fused chain as the CUDA kernel */
        const float limx = 1.3f * p.tanfovx, limy = 1.3f * p.tanfovy;
        float txtz = tvx / tvz, tytz = tvy / tvz;
        float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz;
        float ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
        const float gate_x = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float gate_y = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        float J00 = focal_x / tvz, J02 = -(focal_x * tx) / (tvz * tvz);
        float J11 = focal_y / tvz, J12 = -(focal_y * ty) / (tvz * tvz);
        float T00 = J00 * V[0] + J02 * V[2], T01 = J00 * V[4] + J02 * V[6], T02 = J00 * V[8] + J02 * V[10];
        float T10 = J11 * V[1] + J12 * V[2], T11 = J11 * V[5] + J12 * V[6], T12 = J11 * V[9] + J12 * V[10];
        const float* c6 = s->cov6 + 6 * (size_t)i;
        float S00 = c6[0], S01 = c6[1], S02 = c6[2], S11 = c6[3], S12 = c6[4], S22 = c6[5];
        float u0 = T00 * S00 + T01 * S01 + T02 * S02, u1 = T00 * S01 + T01 * S11 + T02 * S12, u2 = T00 * S02 + T01 * S12 + T02 * S22;
        float v0 = T10 * S00 + T11 * S01 + T12 * S02, v1 = T10 * S01 + T11 * S11 + T12 * S12, v2 = T10 * S02 + T11 * S12 + T12 * S22;
        float a = u0 * T00 + u1 * T01 + u2 * T02 + BLUR;
        float b = u0 * T10 + u1 * T11 + u2 * T12;
        float c = v0 * T10 + v1 * T11 + v2 * T12 + BLUR;
        float det = a * c - b * b;
        float d2 = 1.f / (det * det + 0.0000001f); /* upstream's guard; det >= 0.09 with the 0.3 px^2 blur, so <= 1.3e-5 relative */
        float dLa = 0, dLb = 0, dLc = 0;
        if (det != 0.f) {
            dLa = d2 * (-c * c * gcA + b * c * gcB - b * b * gcC);
            dLc = d2 * (-b * b * gcA + a * b * gcB - a * a * gcC);
            dLb = d2 * (2.f * b * c * gcA - (det + 2.f * b * b) * gcB + 2.f * a * b * gcC);
        }
        /* dSigma3D (full symmetric) = Tm^T Ghat Tm, Ghat = [[dLa, dLb/2],[dLb/2, dLc]] */
        float hb = 0.5f * dLb;
        float dS00 = T00 * T00 * dLa + 2.f * T00 * T10 * hb + T10 * T10 * dLc;
        float dS11 = T01 * T01 * dLa + 2.f * T01 * T11 * hb + T11 * T11 * dLc;
        float dS22 = T02 * T02 * dLa + 2.f * T02 * T12 * hb + T12 * T12 * dLc;
        float dS01 = T00 * T01 * dLa + (T00 * T11 + T01 * T10) * hb + T10 * T11 * dLc;
        float dS02 = T00 * T02 * dLa + (T00 * T12 + T02 * T10) * hb + T10 * T12 * dLc;
        float dS12 = T01 * T02 * dLa + (T01 * T12 + T02 * T11) * hb + T11 * T12 * dLc;
        if (dL_dcov3D && s->cov_pre) {
            float* o6 = dL_dcov3D + 6 * (size_t)i;
            o6[0] = dS00; o6[1] = 2.f * dS01; o6[2] = 2.f * dS02; o6[3] = dS11; o6[4] = 2.f * dS12; o6[5] = dS22;
        }
        /* dL/dTm = 2 Ghat Tm Sigma ; (Tm Sigma) rows are u*, v* */
        float dT00 = 2.f * (dLa * u0 + hb * v0), dT01 = 2.f * (dLa * u1 + hb * v1), dT02 = 2.f * (dLa * u2 + hb * v2);
        float dT10 = 2.f * (hb * u0 + dLc * v0), dT11 = 2.f * (hb * u1 + dLc * v1), dT12 = 2.f * (hb * u2 + dLc * v2);
        /* dL/dJ = dL/dTm * W3^T ; W3[r][c] = V[c*4+r] */
        float dJ00 = dT00 * V[0] + dT01 * V[4] + dT02 * V[8];
        float dJ02 = dT00 * V[2] + dT01 * V[6] + dT02 * V[10];
        float dJ11 = dT10 * V[1] + dT11 * V[5] + dT12 * V[9];
        float dJ12 = dT10 * V[2] + dT11 * V[6] + dT12 * V[10];
        float tz1 = 1.f / tvz, tz2 = tz1 * tz1, tz3 = tz2 * tz1;
        float dtx = gate_x * (-focal_x * tz2) * dJ02;
        float dty = gate_y * (-focal_y * tz2) * dJ12;
        float dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 + (2.f * focal_x * tx) * tz3 * dJ02 + (2.f * focal_y * ty) * tz3 * dJ12;
        /* dL/dmean += W3^T dL/dt ; t_r = sum_c W3[r][c] mean_c, W3[r][c]=V[c*4+r] */
        gm[0] += V[0] * dtx + V[1] * dty + V[2] * dtz;
        gm[1] += V[4] * dtx + V[5] * dty + V[6] * dtz;
        gm[2] += V[8] * dtx + V[9] * dty + V[10] * dtz;

        /* ---- pixel mean -> mean3D (stored g2 = dL/dndc) and depth ---- */
        float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
        float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
        float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
        float mw = 1.0f / (hw + 0.0000001f);
        float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        gm[0] += (Pm[0] * mw - Pm[3] * mul1) * g2x + (Pm[1] * mw - Pm[3] * mul2) * g2y + V[2] * g_dep;
        gm[1] += (Pm[4] * mw - Pm[7] * mul1) * g2x + (Pm[5] * mw - Pm[7] * mul2) * g2y + V[6] * g_dep;
        gm[2] += (Pm[8] * mw - Pm[11] * mul1) * g2x + (Pm[9] * mw - Pm[11] * mul2) * g2y + V[10] * g_dep;
        if (dL_dmeans3D) { dL_dmeans3D[3 * (size_t)i] = gm[0]; dL_dmeans3D[3 * (size_t)i + 1] = gm[1]; dL_dmeans3D[3 * (size_t)i + 2] = gm[2]; }

        /* ---- Sigma3D -> scale, rotation ---- */
        if (!s->cov_pre && (dL_dscales || dL_drotations)) {
            const float* q = s->rots + 4 * (size_t)i;
            const float* sc = s->scales + 3 * (size_t)i;
            float r = q[0], qx = q[1], qy = q[2], qz = q[3];
            float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - r * qz), 2.f * (qx * qz + r * qy),
                          2.f * (qx * qy + r * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - r * qx),
                          2.f * (qx * qz - r * qy), 2.f * (qy * qz + r * qx), 1.f - 2.f * (qx * qx + qy * qy)};
            float sm[3] = {p.scale_modifier * sc[0], p.scale_modifier * sc[1], p.scale_modifier * sc[2]};
            float dS[9] = {dS00, dS01, dS02, dS01, dS11, dS12, dS02, dS12, dS22};
            float Mm[9], dM[9];
            for (int ii = 0; ii < 3; ii++) for (int jj = 0; jj < 3; jj++) Mm[ii * 3 + jj] = R[ii * 3 + jj] * sm[jj];
            for (int ii = 0; ii < 3; ii++) for (int jj = 0; jj < 3; jj++)
                dM[ii * 3 + jj] = 2.f * (dS[ii * 3 + 0] * Mm[0 * 3 + jj] + dS[ii * 3 + 1] * Mm[1 * 3 + jj] + dS[ii * 3 + 2] * Mm[2 * 3 + jj]);
            if (dL_dscales) for (int jj = 0; jj < 3; jj++)
                dL_dscales[3 * (size_t)i + jj] = p.scale_modifier * (R[0 * 3 + jj] * dM[0 * 3 + jj] + R[1 * 3 + jj] * dM[1 * 3 + jj] + R[2 * 3 + jj] * dM[2 * 3 + jj]);
            float G[9];
            for (int ii = 0; ii < 3; ii++) for (int jj = 0; jj < 3; jj++) G[ii * 3 + jj] = dM[ii * 3 + jj] * sm[jj];
            if (dL_drotations) {
                float* o = dL_drotations + 4 * (size_t)i;
                o[0] = 2.f * (qz * (G[3] - G[1]) + qy * (G[2] - G[6]) + qx * (G[7] - G[5]));
                o[1] = 2.f * (qy * (G[1] + G[3]) + qz * (G[2] + G[6]) + r * (G[7] - G[5])) - 4.f * qx * (G[4] + G[8]);
                o[2] = 2.f * (qx * (G[1] + G[3]) + r * (G[2] - G[6]) + qz * (G[5] + G[7])) - 4.f * qy * (G[0] + G[8]);
                o[3] = 2.f * (r * (G[3] - G[1]) + qx * (G[2] + G[6]) + qy * (G[5] + G[7])) - 4.f * qz * (G[0] + G[4]);
            }
        }
    }
    free(acc);
    return 0;
}

int ggo_backward(const ggo_state* s, const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp,
                 float* dL_dopacities, float* dL_dscales, float* dL_drotations, float* dL_dcov3D) {
    return backward_impl(s, 0, dL_dcolor, dL_ddepth, dL_dalpha, dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors_precomp,
                         dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D);
}

/* Same gradients through the front-to-back formulation (prefix sums + pixel totals instead of the back-to-front
 * "colour behind" recurrence).  Kept as a checked statement of the math a per-Gaussian-parallel backward kernel
 * would use; tests/test_oracle_cpu.py compares it with ggo_backward. */
int ggo_backward_forward_order(const ggo_state* s, const float* dL_dcolor, const float* dL_ddepth,
                               const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs,
                               float* dL_dcolors_precomp, float* dL_dopacities, float* dL_dscales,
                               float* dL_drotations, float* dL_dcov3D) {
    return backward_impl(s, 1, dL_dcolor, dL_ddepth, dL_dalpha, dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors_precomp,
                         dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D);
}

int ggo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ggo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
