// mesh_binding_math.h -- per-face frame and per-Gaussian binding transform, forward and backward.
// Pure inline math shared by the CUDA kernels (mesh_binding.cu) and a host test harness
// (tests/native/mesh_math_host.cpp), so the hand-derived backward is checked against autograd on CPU.
//
// Restates the arithmetic of
//   /root/reference/scene/mesh_gaussian_model.py:90-95    update_face_coor
//   /root/reference/scene/mesh_gaussian_model.py:105-128  get_scaling / get_rotation / get_xyz
//   /root/reference/utils/graphics_utils.py:100-137       length / safe_normalize / compute_face_orientation
// and of roma.rotmat_to_unitquat / quat_product (xyzw; largest-of-(m00,m11,m22,trace) branch).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GG_HD __host__ __device__ __forceinline__
#else
#define GG_HD inline
#endif

namespace ggmb {

constexpr float EPS_LEN2 = 1e-20f;   // graphics_utils.length: sqrt(clamp(dot, min=1e-20))

struct V3 { float x, y, z; };
GG_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GG_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GG_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GG_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
GG_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GG_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// y = x / sqrt(max(x.x, eps)); returns length used and whether the clamp was active
GG_HD V3 safe_normalize(V3 x, float& len, bool& clamped) {
    const float d = dot(x, x);
    clamped = d < EPS_LEN2;
    len = sqrtf(clamped ? EPS_LEN2 : d);
    return x * (1.0f / len);
}
// backward of safe_normalize: gy -> gx
GG_HD V3 safe_normalize_bwd(V3 y, float len, bool clamped, V3 gy) {
    if (clamped) return gy * (1.0f / len);
    return (gy - y * dot(y, gy)) * (1.0f / len);
}

// Per-face frame (17 floats): R columns a0|a1|a2 row-major R[r*3+c], scale, centre, quaternion wxyz
struct FaceFrame {
    float R[9];
    float scale;
    V3 center;
    float q[4];       // wxyz, unit
};
struct FaceAux {      // intermediates the backward needs (recomputed, never stored in HBM)
    V3 e1, e2, a0, a1, a2, n, m;
    float len_e1, len_n, len_m, s1_sign;
    bool c_e1, c_n, c_m;
    int branch;       // quaternion branch: 0,1,2 = diagonal entry largest, 3 = trace largest
    float qraw[4];    // xyzw before normalisation
    float qraw_len;
};

GG_HD void quat_from_rotmat_xyzw(const float* M, float* qraw, int& branch) {
    const float m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[3], m11 = M[4], m12 = M[5], m20 = M[6], m21 = M[7], m22 = M[8];
    const float tr = m00 + m11 + m22;
    branch = 0;
    float best = m00;
    if (m11 > best) { best = m11; branch = 1; }
    if (m22 > best) { best = m22; branch = 2; }
    if (tr > best) { branch = 3; }
    if (branch == 3) { qraw[0] = m21 - m12; qraw[1] = m02 - m20; qraw[2] = m10 - m01; qraw[3] = 1.f + tr; }
    else if (branch == 0) { qraw[0] = 1.f - tr + 2.f * m00; qraw[1] = m10 + m01; qraw[2] = m20 + m02; qraw[3] = m21 - m12; }
    else if (branch == 1) { qraw[0] = m10 + m01; qraw[1] = 1.f - tr + 2.f * m11; qraw[2] = m21 + m12; qraw[3] = m02 - m20; }
    else { qraw[0] = m20 + m02; qraw[1] = m21 + m12; qraw[2] = 1.f - tr + 2.f * m22; qraw[3] = m10 - m01; }
}
// d qraw (xyzw) -> dM (accumulates)
GG_HD void quat_from_rotmat_xyzw_bwd(int branch, const float* g, float* dM) {
    // dM index: m00=0 m01=1 m02=2 m10=3 m11=4 m12=5 m20=6 m21=7 m22=8 ; tr = m00+m11+m22
    if (branch == 3) {
        dM[7] += g[0]; dM[5] -= g[0];
        dM[2] += g[1]; dM[6] -= g[1];
        dM[3] += g[2]; dM[1] -= g[2];
        dM[0] += g[3]; dM[4] += g[3]; dM[8] += g[3];
    } else if (branch == 0) {
        dM[0] += g[0]; dM[4] -= g[0]; dM[8] -= g[0];          // 1 - tr + 2 m00 = 1 + m00 - m11 - m22
        dM[3] += g[1]; dM[1] += g[1];
        dM[6] += g[2]; dM[2] += g[2];
        dM[7] += g[3]; dM[5] -= g[3];
    } else if (branch == 1) {
        dM[3] += g[0]; dM[1] += g[0];
        dM[4] += g[1]; dM[0] -= g[1]; dM[8] -= g[1];
        dM[7] += g[2]; dM[5] += g[2];
        dM[2] += g[3]; dM[6] -= g[3];
    } else {
        dM[6] += g[0]; dM[2] += g[0];
        dM[7] += g[1]; dM[5] += g[1];
        dM[8] += g[2]; dM[0] -= g[2]; dM[4] -= g[2];
        dM[3] += g[3]; dM[1] -= g[3];
    }
}

GG_HD void face_frame_fwd(V3 v0, V3 v1, V3 v2, FaceFrame& f, FaceAux& x) {
    x.e1 = v1 - v0;
    x.e2 = v2 - v0;
    x.a0 = safe_normalize(x.e1, x.len_e1, x.c_e1);
    x.n = cross(x.a0, x.e2);
    x.a1 = safe_normalize(x.n, x.len_n, x.c_n);
    x.m = cross(x.a1, x.a0);
    V3 mh = safe_normalize(x.m, x.len_m, x.c_m);
    x.a2 = mh * -1.0f;
    f.R[0] = x.a0.x; f.R[1] = x.a1.x; f.R[2] = x.a2.x;
    f.R[3] = x.a0.y; f.R[4] = x.a1.y; f.R[5] = x.a2.y;
    f.R[6] = x.a0.z; f.R[7] = x.a1.z; f.R[8] = x.a2.z;
    const float s0 = x.len_e1;                       // length(v1 - v0) (same clamp)
    const float d = dot(x.a2, x.e2);
    x.s1_sign = d >= 0.f ? 1.f : -1.f;
    f.scale = 0.5f * (s0 + fabsf(d));
    f.center = (v0 + v1 + v2) * (1.0f / 3.0f);
    quat_from_rotmat_xyzw(f.R, x.qraw, x.branch);
    x.qraw_len = sqrtf(x.qraw[0] * x.qraw[0] + x.qraw[1] * x.qraw[1] + x.qraw[2] * x.qraw[2] + x.qraw[3] * x.qraw[3]);
    const float inv = 1.0f / x.qraw_len;
    f.q[0] = x.qraw[3] * inv; f.q[1] = x.qraw[0] * inv; f.q[2] = x.qraw[1] * inv; f.q[3] = x.qraw[2] * inv;   // xyzw -> wxyz
}

// gradients wrt the frame outputs -> gradients wrt the three vertices
GG_HD void face_frame_bwd(const FaceFrame& f, const FaceAux& x, const float* gR, float g_scale, V3 g_center,
                          const float* g_q /*wxyz*/, V3& gv0, V3& gv1, V3& gv2) {
    float dM[9];
    for (int k = 0; k < 9; k++) dM[k] = gR[k];
    // quaternion: q = qraw / |qraw| (then wxyz reorder)
    {
        const float gq_xyzw[4] = {g_q[1], g_q[2], g_q[3], g_q[0]};
        const float qn[4] = {f.q[1], f.q[2], f.q[3], f.q[0]};
        const float dd = qn[0] * gq_xyzw[0] + qn[1] * gq_xyzw[1] + qn[2] * gq_xyzw[2] + qn[3] * gq_xyzw[3];
        float graw[4];
        for (int k = 0; k < 4; k++) graw[k] = (gq_xyzw[k] - qn[k] * dd) / x.qraw_len;
        quat_from_rotmat_xyzw_bwd(x.branch, graw, dM);
    }
    V3 ga0 = v3(dM[0], dM[3], dM[6]), ga1 = v3(dM[1], dM[4], dM[7]), ga2 = v3(dM[2], dM[5], dM[8]);
    V3 ge1 = v3(0, 0, 0), ge2 = v3(0, 0, 0);
    // scale = (s0 + |a2 . e2|) / 2
    const float gs = 0.5f * g_scale;
    if (!x.c_e1) ge1 = ge1 + x.e1 * (gs / x.len_e1);                 // d sqrt(e1.e1) = e1 / len
    ga2 = ga2 + x.e2 * (gs * x.s1_sign);
    ge2 = ge2 + x.a2 * (gs * x.s1_sign);
    // a2 = -normalize(m), m = a1 x a0
    {
        V3 mh = x.a2 * -1.0f;
        V3 gm = safe_normalize_bwd(mh, x.len_m, x.c_m, ga2 * -1.0f);
        ga1 = ga1 + cross(x.a0, gm);          // c = a x b: dL/da = b x gc
        ga0 = ga0 + cross(gm, x.a1);          //            dL/db = gc x a
    }
    // a1 = normalize(n), n = a0 x e2
    {
        V3 gn = safe_normalize_bwd(x.a1, x.len_n, x.c_n, ga1);
        ga0 = ga0 + cross(x.e2, gn);
        ge2 = ge2 + cross(gn, x.a0);
    }
    // a0 = normalize(e1)
    ge1 = ge1 + safe_normalize_bwd(x.a0, x.len_e1, x.c_e1, ga0);
    const V3 gc3 = g_center * (1.0f / 3.0f);
    gv1 = ge1 + gc3;
    gv2 = ge2 + gc3;
    gv0 = gc3 - ge1 - ge2;
}

// ---- per-Gaussian binding ---------------------------------------------------------------------
// Hamilton product in wxyz (identical to roma.quat_product on the xyzw-reordered operands)
GG_HD void qmul(const float* p, const float* q, float* r) {
    r[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
    r[1] = p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2];
    r[2] = p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1];
    r[3] = p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0];
}
// r = p*q: given gr, gp = gr * conj(q), gq = conj(p) * gr   (bilinear)
GG_HD void qmul_bwd(const float* p, const float* q, const float* gr, float* gp, float* gq) {
    const float qc[4] = {q[0], -q[1], -q[2], -q[3]};
    const float pc[4] = {p[0], -p[1], -p[2], -p[3]};
    qmul(gr, qc, gp);
    qmul(pc, gr, gq);
}
// torch.nn.functional.normalize (eps 1e-12): y = x / max(|x|, eps)
GG_HD float qnormalize(const float* x, float* y) {
    float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    n = n > 1e-12f ? n : 1e-12f;
    const float inv = 1.0f / n;
    for (int k = 0; k < 4; k++) y[k] = x[k] * inv;
    return n;
}
GG_HD void qnormalize_bwd(const float* y, float n, const float* gy, float* gx) {
    const float d = y[0] * gy[0] + y[1] * gy[1] + y[2] * gy[2] + y[3] * gy[3];
    for (int k = 0; k < 4; k++) gx[k] = (gy[k] - y[k] * d) / n;
}

struct BindOut { V3 xyz; V3 scaling; float rot[4]; };

// `anchor`: where the local frame sits -- the face centre (MeshGaussianModel.get_xyz, scene/mesh_gaussian_model.py:124-128)
// or the barycentric point a v0 + b v1 + c v2 (AvatarGaussianModel.get_xyz / get_final_xyz / get_barycentric_3d,
// scene/avatar_gaussian_model.py:140-159).  `scal_scale`: the face scale get_scaling multiplies with -- the current one,
// or the frozen `face_scaling_remembered` (scene/mesh_gaussian_model.py:98-110), which carries no gradient.
GG_HD void bind_fwd_ex(const FaceFrame& f, V3 anchor, float scal_scale, V3 lxyz, V3 lscal, const float* lrot, BindOut& o) {
    const V3 r = v3(f.R[0] * lxyz.x + f.R[1] * lxyz.y + f.R[2] * lxyz.z, f.R[3] * lxyz.x + f.R[4] * lxyz.y + f.R[5] * lxyz.z,
                    f.R[6] * lxyz.x + f.R[7] * lxyz.y + f.R[8] * lxyz.z);
    o.xyz = r * f.scale + anchor;
    o.scaling = v3(expf(lscal.x) * scal_scale, expf(lscal.y) * scal_scale, expf(lscal.z) * scal_scale);
    float rl[4], fq[4], w[4];
    qnormalize(lrot, rl);
    qnormalize(f.q, fq);
    qmul(fq, rl, w);
    qnormalize(w, o.rot);
}
GG_HD void bind_fwd(const FaceFrame& f, V3 lxyz, V3 lscal, const float* lrot, BindOut& o) {
    bind_fwd_ex(f, f.center, f.scale, lxyz, lscal, lrot, o);
}

// gradients wrt outputs -> gradients wrt local params and wrt the face frame (17 floats, to be accumulated).
// gF[10..12] is the gradient of the ANCHOR (= centre slots for the face-centre variant; the barycentric variant
// scatters it to the three vertices itself).  scale_is_live = false: get_scaling used a frozen face scale.
GG_HD void bind_bwd_ex(const FaceFrame& f, float scal_scale, bool scale_is_live, V3 lxyz, V3 lscal, const float* lrot,
                       V3 g_xyz, V3 g_scal, const float* g_rot, V3& gl_xyz, V3& gl_scal, float* gl_rot,
                       float* gF /*[17]: R9, scale, anchor3, q4*/) {
    const V3 r = v3(f.R[0] * lxyz.x + f.R[1] * lxyz.y + f.R[2] * lxyz.z, f.R[3] * lxyz.x + f.R[4] * lxyz.y + f.R[5] * lxyz.z,
                    f.R[6] * lxyz.x + f.R[7] * lxyz.y + f.R[8] * lxyz.z);
    const V3 gs = g_xyz * f.scale;                          // gradient wrt r
    gl_xyz = v3(f.R[0] * gs.x + f.R[3] * gs.y + f.R[6] * gs.z, f.R[1] * gs.x + f.R[4] * gs.y + f.R[7] * gs.z,
                f.R[2] * gs.x + f.R[5] * gs.y + f.R[8] * gs.z);
    gF[0] = gs.x * lxyz.x; gF[1] = gs.x * lxyz.y; gF[2] = gs.x * lxyz.z;
    gF[3] = gs.y * lxyz.x; gF[4] = gs.y * lxyz.y; gF[5] = gs.y * lxyz.z;
    gF[6] = gs.z * lxyz.x; gF[7] = gs.z * lxyz.y; gF[8] = gs.z * lxyz.z;
    const V3 ex = v3(expf(lscal.x), expf(lscal.y), expf(lscal.z));
    gF[9] = dot(g_xyz, r) + (scale_is_live ? g_scal.x * ex.x + g_scal.y * ex.y + g_scal.z * ex.z : 0.f);
    gF[10] = g_xyz.x; gF[11] = g_xyz.y; gF[12] = g_xyz.z;
    gl_scal = v3(g_scal.x * ex.x * scal_scale, g_scal.y * ex.y * scal_scale, g_scal.z * ex.z * scal_scale);
    float rl[4], fq[4], w[4], out[4];
    const float n_rl = qnormalize(lrot, rl);
    const float n_fq = qnormalize(f.q, fq);
    qmul(fq, rl, w);
    const float n_w = qnormalize(w, out);
    float gw[4], gfq[4], grl[4];
    qnormalize_bwd(out, n_w, g_rot, gw);
    qmul_bwd(fq, rl, gw, gfq, grl);
    qnormalize_bwd(rl, n_rl, grl, gl_rot);
    qnormalize_bwd(fq, n_fq, gfq, gF + 13);
}
GG_HD void bind_bwd(const FaceFrame& f, V3 lxyz, V3 lscal, const float* lrot, V3 g_xyz, V3 g_scal, const float* g_rot,
                    V3& gl_xyz, V3& gl_scal, float* gl_rot, float* gF) {
    bind_bwd_ex(f, f.scale, true, lxyz, lscal, lrot, g_xyz, g_scal, g_rot, gl_xyz, gl_scal, gl_rot, gF);
}

}  // namespace ggmb
