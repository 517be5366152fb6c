// preprocess_bwd.cu -- per-Gaussian backward: conic -> Sigma2D
// -> Sigma3D / view-space position -> mean3D, scale, quaternion; pixel mean and depth -> mean3D;
// SH backward (dL/dsh and, through the view direction, dL/dmean3D).
//
// Pure streaming kernel: nothing is kept from the forward pass except `radii`; Sigma3D, the EWA
// terms, the SH basis and the colour clamp mask are recomputed from the inputs (cheaper than
// 40+ B/Gaussian of saved state through HBM).  M == 16: the 192 B SH rows are staged in with
// coalesced 16-byte cp.async into padded shared memory, and dL/dsh leaves through the same padded
// rows so the 192 B/Gaussian write is coalesced too.
#include "common.cuh"

namespace gg {

constexpr int PB_BLOCK = 128;
constexpr int PB_ROW_U = 13;

struct PBArgs {
    int N, M, D, W, H;
    float tanfovx, tanfovy, mod;
    const float *means3D, *shs, *colors, *scales, *rots, *cov_pre, *view, *proj, *campos;
    const int32_t* radii;
    const float4 *a0, *a1;
    const float2* a2;
    float *g_means3D, *g_means2D, *g_shs, *g_colors, *g_opac, *g_scales, *g_rots, *g_cov;
};

// geometry part shared by both kernels; returns gm (dL/dmean3D without the SH term)
__device__ __forceinline__ void geom_backward(const PBArgs& A, const float* V, const float* Pm, int i, float4 q0,
                                              float4 q1, float2 q2, float* gm) {
    const float x = A.means3D[3 * (size_t)i], y = A.means3D[3 * (size_t)i + 1], z = A.means3D[3 * (size_t)i + 2];
    const float g2x = q0.x, g2y = q0.y, gcA = q0.z, gcB = q0.w, gcC = q1.x, g_op = q1.y, g_dep = q2.y;
    if (A.g_means2D) {
        A.g_means2D[3 * (size_t)i] = g2x;
        A.g_means2D[3 * (size_t)i + 1] = g2y;
        A.g_means2D[3 * (size_t)i + 2] = 0.f;
    }
    if (A.g_opac) A.g_opac[i] = g_op;

    float c6[6];
    float sc[3] = {0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (A.cov_pre) {
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = A.cov_pre[6 * (size_t)i + k];
    } else {
        sc[0] = A.scales[3 * (size_t)i]; sc[1] = A.scales[3 * (size_t)i + 1]; sc[2] = A.scales[3 * (size_t)i + 2];
        const float4 q4 = reinterpret_cast<const float4*>(A.rots)[i];
        q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
        cov3d_from_scale_rot(sc, A.mod, q, c6);
    }
    const float focal_x = A.W / (2.0f * A.tanfovx), focal_y = A.H / (2.0f * A.tanfovy);
    Ewa e;
    ewa_project(V, x, y, z, c6, focal_x, focal_y, A.tanfovx, A.tanfovy, e);

    // conic (A,B,C) = (c,-b,a)/det  ->  (a,b,c)
    const float a = e.a, b = e.b, c = e.c, det = e.det;
    float dLa = 0.f, dLb = 0.f, dLc = 0.f;
    if (det != 0.f) {
        const float d2 = 1.f / (det * det + 0.0000001f);      // upstream keeps the denominator finite the same way
        dLa = d2 * (-c * c * gcA + b * c * gcB - b * b * gcC);
        dLc = d2 * (-b * b * gcA + a * b * gcB - a * a * gcC);
        dLb = d2 * (2.f * b * c * gcA - (det + 2.f * b * b) * gcB + 2.f * a * b * gcC);
    }
    const float hb = 0.5f * dLb;
    // dSigma3D (full symmetric) = Tm^T Ghat Tm
    const float dS00 = e.T00 * e.T00 * dLa + 2.f * e.T00 * e.T10 * hb + e.T10 * e.T10 * dLc;
    const float dS11 = e.T01 * e.T01 * dLa + 2.f * e.T01 * e.T11 * hb + e.T11 * e.T11 * dLc;
    const float dS22 = e.T02 * e.T02 * dLa + 2.f * e.T02 * e.T12 * hb + e.T12 * e.T12 * dLc;
    const float dS01 = e.T00 * e.T01 * dLa + (e.T00 * e.T11 + e.T01 * e.T10) * hb + e.T10 * e.T11 * dLc;
    const float dS02 = e.T00 * e.T02 * dLa + (e.T00 * e.T12 + e.T02 * e.T10) * hb + e.T10 * e.T12 * dLc;
    const float dS12 = e.T01 * e.T02 * dLa + (e.T01 * e.T12 + e.T02 * e.T11) * hb + e.T11 * e.T12 * dLc;
    if (A.g_cov && A.cov_pre) {
        float* o6 = A.g_cov + 6 * (size_t)i;
        o6[0] = dS00; o6[1] = 2.f * dS01; o6[2] = 2.f * dS02; o6[3] = dS11; o6[4] = 2.f * dS12; o6[5] = dS22;
    }
    // dL/dTm = 2 Ghat (Tm Sigma)
    const float dT00 = 2.f * (dLa * e.u0 + hb * e.v0), dT01 = 2.f * (dLa * e.u1 + hb * e.v1), dT02 = 2.f * (dLa * e.u2 + hb * e.v2);
    const float dT10 = 2.f * (hb * e.u0 + dLc * e.v0), dT11 = 2.f * (hb * e.u1 + dLc * e.v1), dT12 = 2.f * (hb * e.u2 + dLc * e.v2);
    const float dJ00 = dT00 * V[0] + dT01 * V[4] + dT02 * V[8];
    const float dJ02 = dT00 * V[2] + dT01 * V[6] + dT02 * V[10];
    const float dJ11 = dT10 * V[1] + dT11 * V[5] + dT12 * V[9];
    const float dJ12 = dT10 * V[2] + dT11 * V[6] + dT12 * V[10];
    const float tz1 = 1.f / e.tvz, tz2 = tz1 * tz1, tz3 = tz2 * tz1;
    const float dtx = e.gate_x * (-focal_x * tz2) * dJ02;
    const float dty = e.gate_y * (-focal_y * tz2) * dJ12;
    const float dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 + (2.f * focal_x * e.tx) * tz3 * dJ02 +
                      (2.f * focal_y * e.ty) * tz3 * dJ12;
    gm[0] += V[0] * dtx + V[1] * dty + V[2] * dtz;
    gm[1] += V[4] * dtx + V[5] * dty + V[6] * dtz;
    gm[2] += V[8] * dtx + V[9] * dty + V[10] * dtz;

    // pixel mean (stored as dL/dndc) and depth
    const float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
    const float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
    const float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
    const float mw = 1.0f / (hw + 0.0000001f);
    const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
    gm[0] += (Pm[0] * mw - Pm[3] * mul1) * g2x + (Pm[1] * mw - Pm[3] * mul2) * g2y + V[2] * g_dep;
    gm[1] += (Pm[4] * mw - Pm[7] * mul1) * g2x + (Pm[5] * mw - Pm[7] * mul2) * g2y + V[6] * g_dep;
    gm[2] += (Pm[8] * mw - Pm[11] * mul1) * g2x + (Pm[9] * mw - Pm[11] * mul2) * g2y + V[10] * g_dep;

    // Sigma3D -> scale, quaternion
    if (!A.cov_pre && (A.g_scales || A.g_rots)) {
        float R[9];
        quat_to_rot(q[0], q[1], q[2], q[3], R);
        const float sm[3] = {A.mod * sc[0], A.mod * sc[1], A.mod * sc[2]};
        const float dS[9] = {dS00, dS01, dS02, dS01, dS11, dS12, dS02, dS12, dS22};
        float Mm[9], dM[9], G[9];
#pragma unroll
        for (int ii = 0; ii < 3; ii++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++) Mm[ii * 3 + jj] = R[ii * 3 + jj] * sm[jj];
#pragma unroll
        for (int ii = 0; ii < 3; ii++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
                dM[ii * 3 + jj] = 2.f * (dS[ii * 3] * Mm[jj] + dS[ii * 3 + 1] * Mm[3 + jj] + dS[ii * 3 + 2] * Mm[6 + jj]);
        if (A.g_scales) {
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
                A.g_scales[3 * (size_t)i + jj] = A.mod * (R[jj] * dM[jj] + R[3 + jj] * dM[3 + jj] + R[6 + jj] * dM[6 + jj]);
        }
#pragma unroll
        for (int ii = 0; ii < 3; ii++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++) G[ii * 3 + jj] = dM[ii * 3 + jj] * sm[jj];
        if (A.g_rots) {
            const float r = q[0], qx = q[1], qy = q[2], qz = q[3];
            float4 o;
            o.x = 2.f * (qz * (G[3] - G[1]) + qy * (G[2] - G[6]) + qx * (G[7] - G[5]));
            o.y = 2.f * (qy * (G[1] + G[3]) + qz * (G[2] + G[6]) + r * (G[7] - G[5])) - 4.f * qx * (G[4] + G[8]);
            o.z = 2.f * (qx * (G[1] + G[3]) + r * (G[2] - G[6]) + qz * (G[5] + G[7])) - 4.f * qy * (G[0] + G[8]);
            o.w = 2.f * (r * (G[3] - G[1]) + qx * (G[2] + G[6]) + qy * (G[5] + G[7])) - 4.f * qz * (G[0] + G[4]);
            reinterpret_cast<float4*>(A.g_rots)[i] = o;
        }
    }
}

// d(colour)/d(view direction) -> mean3D, given h_k = sum_ch sh[k][ch] * gr[ch] (zero above the active degree)
__device__ __forceinline__ void sh_dir_backward(int deg, const float* h, float dxn, float dyn, float dzn, float inv, float* gm) {
    const float X = dxn, Y = dyn, Z = dzn;
    float gx_ = -GG_SH_C1 * h[3], gy_ = -GG_SH_C1 * h[1], gz_ = GG_SH_C1 * h[2];
    if (deg > 1) {
        const float xx = X * X, yy = Y * Y, zz = Z * Z, xy = X * Y, yz = Y * Z, xz = X * Z;
        gx_ += GG_SH_C2_0 * Y * h[4] + GG_SH_C2_2 * (-2.f * X) * h[6] + GG_SH_C2_3 * Z * h[7] + GG_SH_C2_4 * 2.f * X * h[8];
        gy_ += GG_SH_C2_0 * X * h[4] + GG_SH_C2_1 * Z * h[5] + GG_SH_C2_2 * (-2.f * Y) * h[6] + GG_SH_C2_4 * (-2.f * Y) * h[8];
        gz_ += GG_SH_C2_1 * Y * h[5] + GG_SH_C2_2 * 4.f * Z * h[6] + GG_SH_C2_3 * X * h[7];
        if (deg > 2) {
            gx_ += GG_SH_C3_0 * h[9] * 6.f * xy + GG_SH_C3_1 * h[10] * yz + GG_SH_C3_2 * h[11] * (-2.f * xy) +
                   GG_SH_C3_3 * h[12] * (-6.f * xz) + GG_SH_C3_4 * h[13] * (4.f * zz - 3.f * xx - yy) +
                   GG_SH_C3_5 * h[14] * 2.f * xz + GG_SH_C3_6 * h[15] * (3.f * xx - 3.f * yy);
            gy_ += GG_SH_C3_0 * h[9] * (3.f * xx - 3.f * yy) + GG_SH_C3_1 * h[10] * xz +
                   GG_SH_C3_2 * h[11] * (4.f * zz - xx - 3.f * yy) + GG_SH_C3_3 * h[12] * (-6.f * yz) +
                   GG_SH_C3_4 * h[13] * (-2.f * xy) + GG_SH_C3_5 * h[14] * (-2.f * yz) + GG_SH_C3_6 * h[15] * (-6.f * xy);
            gz_ += GG_SH_C3_1 * h[10] * xy + GG_SH_C3_2 * h[11] * 8.f * yz +
                   GG_SH_C3_3 * h[12] * (6.f * zz - 3.f * xx - 3.f * yy) + GG_SH_C3_4 * h[13] * 8.f * xz +
                   GG_SH_C3_5 * h[14] * (xx - yy);
        }
    }
    const float dot = dxn * gx_ + dyn * gy_ + dzn * gz_;
    gm[0] += (gx_ - dxn * dot) * inv;
    gm[1] += (gy_ - dyn * dot) * inv;
    gm[2] += (gz_ - dzn * dot) * inv;
}

// SH backward for one Gaussian held in registers: sh[3k+ch] in, gsh[3k+ch] out (all 3*M entries
// written, zeros above the active degree); adds the view-direction term to gm.
__device__ __forceinline__ void sh_backward(int deg, const float* sh, const float* campos, float x, float y, float z,
                                            const float* g_rgb, float* gsh, float* gm) {
    const float vx = x - campos[0], vy = y - campos[1], vz = z - campos[2];
    const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
    const float dxn = vx * inv, dyn = vy * inv, dzn = vz * inv;
    float b[16];
    sh_basis(deg, dxn, dyn, dzn, b);
    const int nb = (deg + 1) * (deg + 1);
    float res[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < nb) {
            res[0] += b[k] * sh[3 * k]; res[1] += b[k] * sh[3 * k + 1]; res[2] += b[k] * sh[3 * k + 2];
        }
    float gr[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) gr[ch] = (res[ch] + 0.5f < 0.f) ? 0.f : g_rgb[ch];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float bk = (k < nb) ? b[k] : 0.f;
        gsh[3 * k] = bk * gr[0]; gsh[3 * k + 1] = bk * gr[1]; gsh[3 * k + 2] = bk * gr[2];
    }
    if (deg == 0) return;
    // h_k = sum_ch sh[k][ch] * gr[ch]  (direction derivative only needs this contraction)
    float h[16];
#pragma unroll
    for (int k = 0; k < 16; k++) h[k] = (k < nb) ? sh[3 * k] * gr[0] + sh[3 * k + 1] * gr[1] + sh[3 * k + 2] * gr[2] : 0.f;
    sh_dir_backward(deg, h, dxn, dyn, dzn, inv, gm);
}

__device__ __forceinline__ void write_zero_small(const PBArgs& A, int i) {
    if (A.g_means3D) { A.g_means3D[3 * (size_t)i] = 0.f; A.g_means3D[3 * (size_t)i + 1] = 0.f; A.g_means3D[3 * (size_t)i + 2] = 0.f; }
    if (A.g_means2D) { A.g_means2D[3 * (size_t)i] = 0.f; A.g_means2D[3 * (size_t)i + 1] = 0.f; A.g_means2D[3 * (size_t)i + 2] = 0.f; }
    if (A.g_opac) A.g_opac[i] = 0.f;
    if (A.g_scales) { A.g_scales[3 * (size_t)i] = 0.f; A.g_scales[3 * (size_t)i + 1] = 0.f; A.g_scales[3 * (size_t)i + 2] = 0.f; }
    if (A.g_rots) reinterpret_cast<float4*>(A.g_rots)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A.g_cov) for (int k = 0; k < 6; k++) A.g_cov[6 * (size_t)i + k] = 0.f;
    if (A.g_colors) { A.g_colors[3 * (size_t)i] = 0.f; A.g_colors[3 * (size_t)i + 1] = 0.f; A.g_colors[3 * (size_t)i + 2] = 0.f; }
}

// ---- M == 16 fast path ------------------------------------------------------------------------
__global__ void __launch_bounds__(PB_BLOCK, 8) preprocess_bwd16_kernel(PBArgs A) {   // <= 64 registers: 8 CTAs (1024 threads) per SM
    __shared__ __align__(16) float4 rows[PB_BLOCK * PB_ROW_U];
    __shared__ float cam[35];
    if (threadIdx.x < 16) cam[threadIdx.x] = A.view[threadIdx.x];
    else if (threadIdx.x < 32) cam[threadIdx.x] = A.proj[threadIdx.x - 16];
    else if (threadIdx.x < 35) cam[threadIdx.x] = A.campos[threadIdx.x - 32];
    const int base = blockIdx.x * PB_BLOCK;
    const int i = base + threadIdx.x;
    const int nG = min(PB_BLOCK, A.N - base);
    const bool live = (i < A.N) && (A.radii[i] > 0);
    const bool any_live = __syncthreads_or(live);
    const bool want_sh = A.g_shs != nullptr;
    if (any_live) {
        const float4* src = reinterpret_cast<const float4*>(A.shs) + (size_t)base * 12;
        for (int u = threadIdx.x; u < nG * 12; u += PB_BLOCK) {
            const int gI = u / 12, j = u - gI * 12;
            cp_async16(&rows[gI * PB_ROW_U + j], src + u);
        }
        cp_async_wait_all();
        __syncthreads();
    }
    const int tid_row = threadIdx.x * PB_ROW_U;
    if (live) {
        const float4 q0 = A.a0[i], q1 = A.a1[i];
        const float2 q2 = A.a2[i];
        float gm[3] = {0.f, 0.f, 0.f};
        geom_backward(A, cam, cam + 16, i, q0, q1, q2, gm);
        // SH backward straight on the padded shared-memory row: the 48 coefficients and their 48 gradients never sit in
        // registers together (86 -> ~56 registers, 29 % -> 44 % occupancy for this streaming kernel)
        const float* campos = cam + 32;
        const float vx = A.means3D[3 * (size_t)i] - campos[0], vy = A.means3D[3 * (size_t)i + 1] - campos[1],
                    vz = A.means3D[3 * (size_t)i + 2] - campos[2];
        const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
        const float dxn = vx * inv, dyn = vy * inv, dzn = vz * inv;
        float b[16];
#pragma unroll
        for (int k = 0; k < 16; k++) b[k] = 0.f;
        sh_basis(A.D, dxn, dyn, dzn, b);
        const int nb = (A.D + 1) * (A.D + 1);
        float res[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 12; j++) {                       // pass 1: the colour, for the clamp mask
            const float4 q = rows[tid_row + j];
            const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int f = 4 * j + e;                      // flat index 3 k + channel (compile-time)
                if (f / 3 < nb) res[f % 3] += b[f / 3] * v[e];
            }
        }
        const float g_rgb[3] = {q1.z, q1.w, q2.x};
        float gr[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) gr[ch] = (res[ch] + 0.5f < 0.f) ? 0.f : g_rgb[ch];
        float h[16];
#pragma unroll
        for (int k = 0; k < 16; k++) h[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 12; j++) {                       // pass 2: gradients replace the coefficients in place
            const float4 q = rows[tid_row + j];
            const float v[4] = {q.x, q.y, q.z, q.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int f = 4 * j + e, k = f / 3, ch = f % 3;
                const bool on = k < nb;
                o[e] = on ? b[k] * gr[ch] : 0.f;
                if (on) h[k] += v[e] * gr[ch];
            }
            if (want_sh) rows[tid_row + j] = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (A.D > 0) sh_dir_backward(A.D, h, dxn, dyn, dzn, inv, gm);
        if (A.g_means3D) {
            A.g_means3D[3 * (size_t)i] = gm[0]; A.g_means3D[3 * (size_t)i + 1] = gm[1]; A.g_means3D[3 * (size_t)i + 2] = gm[2];
        }
    } else if (i < A.N) {
        write_zero_small(A, i);
        if (want_sh && any_live) {
#pragma unroll
            for (int j = 0; j < 12; j++) rows[tid_row + j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (!want_sh) return;
    float4* dst = reinterpret_cast<float4*>(A.g_shs) + (size_t)base * 12;
    if (any_live) {
        // rows are private per thread until here; now stream the whole slab out coalesced
        __syncthreads();
        for (int u = threadIdx.x; u < nG * 12; u += PB_BLOCK) {
            const int gI = u / 12, j = u - gI * 12;
            dst[u] = rows[gI * PB_ROW_U + j];
        }
    } else {
        for (int u = threadIdx.x; u < nG * 12; u += PB_BLOCK) dst[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- generic path (M != 16, colours precomputed, or no SH gradient layout constraints) ---------
__global__ void __launch_bounds__(256) preprocess_bwd_generic_kernel(PBArgs A) {
    __shared__ float cam[35];
    if (threadIdx.x < 16) cam[threadIdx.x] = A.view[threadIdx.x];
    else if (threadIdx.x < 32) cam[threadIdx.x] = A.proj[threadIdx.x - 16];
    else if (threadIdx.x < 35) cam[threadIdx.x] = A.campos[threadIdx.x - 32];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.N) return;
    const int M = A.M;
    if (A.radii[i] <= 0) {
        write_zero_small(A, i);
        if (A.g_shs) for (int k = 0; k < 3 * M; k++) A.g_shs[(size_t)i * M * 3 + k] = 0.f;
        return;
    }
    const float4 q0 = A.a0[i], q1 = A.a1[i];
    const float2 q2 = A.a2[i];
    float gm[3] = {0.f, 0.f, 0.f};
    geom_backward(A, cam, cam + 16, i, q0, q1, q2, gm);
    const float g_rgb[3] = {q1.z, q1.w, q2.x};
    if (A.colors) {
        if (A.g_colors) {
            A.g_colors[3 * (size_t)i] = g_rgb[0]; A.g_colors[3 * (size_t)i + 1] = g_rgb[1]; A.g_colors[3 * (size_t)i + 2] = g_rgb[2];
        }
    } else {
        float sh[48], gsh[48];
        const int nb = (A.D + 1) * (A.D + 1);
#pragma unroll
        for (int k = 0; k < 48; k++) sh[k] = (k < 3 * nb) ? A.shs[(size_t)i * M * 3 + k] : 0.f;
        sh_backward(A.D, sh, cam + 32, A.means3D[3 * (size_t)i], A.means3D[3 * (size_t)i + 1],
                    A.means3D[3 * (size_t)i + 2], g_rgb, gsh, gm);
        if (A.g_shs) {
#pragma unroll
            for (int k = 0; k < 48; k++)
                if (k < 3 * M) A.g_shs[(size_t)i * M * 3 + k] = gsh[k];
            for (int k = 48; k < 3 * M; k++) A.g_shs[(size_t)i * M * 3 + k] = 0.f;
        }
    }
    if (A.g_means3D) {
        A.g_means3D[3 * (size_t)i] = gm[0]; A.g_means3D[3 * (size_t)i + 1] = gm[1]; A.g_means3D[3 * (size_t)i + 2] = gm[2];
    }
}

int launch_preprocess_bwd(const gg_view& v, const gg_inputs& in, const int32_t* radii, const AccumWS& acc,
                          float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors,
                          float* dL_dopacities, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                          cudaStream_t s) {
    const int N = v.num_gaussians;
    if (N == 0) return 0;
    PBArgs A;
    A.N = N; A.M = v.sh_coeffs; A.D = v.sh_degree; A.W = v.image_width; A.H = v.image_height;
    A.tanfovx = v.tanfovx; A.tanfovy = v.tanfovy; A.mod = v.scale_modifier;
    A.means3D = in.means3D; A.shs = in.shs; A.colors = in.colors_precomp; A.scales = in.scales;
    A.rots = in.rotations; A.cov_pre = in.cov3D_precomp; A.view = in.viewmatrix; A.proj = in.projmatrix;
    A.campos = in.campos; A.radii = radii; A.a0 = acc.a0; A.a1 = acc.a1; A.a2 = acc.a2;
    A.g_means3D = dL_dmeans3D; A.g_means2D = dL_dmeans2D; A.g_shs = dL_dshs; A.g_colors = dL_dcolors;
    A.g_opac = dL_dopacities; A.g_scales = dL_dscales; A.g_rots = dL_drotations; A.g_cov = dL_dcov3D;
    if (!in.colors_precomp && v.sh_coeffs == 16)
        preprocess_bwd16_kernel<<<(N + PB_BLOCK - 1) / PB_BLOCK, PB_BLOCK, 0, s>>>(A);
    else
        preprocess_bwd_generic_kernel<<<(N + 255) / 256, 256, 0, s>>>(A);
    return 1;
}

}  // namespace gg
