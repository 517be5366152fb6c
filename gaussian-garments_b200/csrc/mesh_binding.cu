// mesh_binding.cu -- fused mesh-binding transform, forward and backward.
//
// Replaces, per optimisation iteration of the reference, ~15 small torch kernels + bmm + the roma
// quaternion ops of  MeshGaussianModel.update_face_coor / get_xyz / get_scaling / get_rotation
// (/root/reference/scene/mesh_gaussian_model.py:90-128, utils/graphics_utils.py:118-137) and their
// autograd backward by four streaming kernels:
//   face_frame_kernel       per face:      3 vertices -> frame (R, scale, centre, quaternion; 17 floats)
//   bind_fwd_kernel         per Gaussian:  local (xyz, log-scale, quaternion) -> world (xyz, scale, quaternion)
//   bind_bwd_kernel         per Gaussian:  world grads -> local grads + 17 red.add into its face's frame gradient
//   face_frame_bwd_kernel   per face:      frame gradient -> 9 red.add into the vertex gradient (mesh.v)
// All arithmetic is in mesh_binding_math.h (also compiled for the host by tests/native/mesh_math_host.cpp).
#include "common.cuh"
#include "mesh_binding_math.h"

namespace gg {
using namespace ggmb;

__device__ __forceinline__ V3 ld3(const float* __restrict__ p, size_t i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(float* __restrict__ p, size_t i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

__device__ __forceinline__ void load_frame(const float* __restrict__ frames, int f, FaceFrame& fr) {
    const float* s = frames + 17 * (size_t)f;
#pragma unroll
    for (int k = 0; k < 9; k++) fr.R[k] = s[k];
    fr.scale = s[9];
    fr.center = v3(s[10], s[11], s[12]);
#pragma unroll
    for (int k = 0; k < 4; k++) fr.q[k] = s[13 + k];
}

__global__ void __launch_bounds__(256)
face_frame_kernel(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces, float* __restrict__ frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    FaceFrame fr;
    FaceAux ax;
    face_frame_fwd(ld3(verts, faces[3 * f]), ld3(verts, faces[3 * f + 1]), ld3(verts, faces[3 * f + 2]), fr, ax);
    float* o = frames + 17 * (size_t)f;
#pragma unroll
    for (int k = 0; k < 9; k++) o[k] = fr.R[k];
    o[9] = fr.scale; o[10] = fr.center.x; o[11] = fr.center.y; o[12] = fr.center.z;
#pragma unroll
    for (int k = 0; k < 4; k++) o[13 + k] = fr.q[k];
}

// bary != NULL: AvatarGaussianModel anchor (barycentric point of the bound face, scene/avatar_gaussian_model.py:151-154);
// scale_rem != NULL: get_scaling with the frozen face_scaling_remembered (scene/mesh_gaussian_model.py:98-110).
__global__ void __launch_bounds__(256)
bind_fwd_kernel(int N, const float* __restrict__ frames, const int32_t* __restrict__ binding,
                const float* __restrict__ lxyz, const float* __restrict__ lscal, const float* __restrict__ lrot,
                const float* __restrict__ verts, const int32_t* __restrict__ faces, const float* __restrict__ bary,
                const float* __restrict__ scale_rem, float* __restrict__ o_xyz, float* __restrict__ o_scal,
                float* __restrict__ o_rot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int f = binding[i];
    FaceFrame fr;
    load_frame(frames, f, fr);
    const float4 q = reinterpret_cast<const float4*>(lrot)[i];
    const float lr[4] = {q.x, q.y, q.z, q.w};
    V3 anchor = fr.center;
    if (bary) {
        const V3 bc = ld3(bary, i);
        anchor = ld3(verts, faces[3 * f]) * bc.x + ld3(verts, faces[3 * f + 1]) * bc.y + ld3(verts, faces[3 * f + 2]) * bc.z;
    }
    BindOut o;
    bind_fwd_ex(fr, anchor, scale_rem ? scale_rem[f] : fr.scale, ld3(lxyz, i), ld3(lscal, i), lr, o);
    st3(o_xyz, i, o.xyz);
    st3(o_scal, i, o.scaling);
    reinterpret_cast<float4*>(o_rot)[i] = make_float4(o.rot[0], o.rot[1], o.rot[2], o.rot[3]);
}

__global__ void __launch_bounds__(256)
bind_bwd_kernel(int N, const float* __restrict__ frames, const int32_t* __restrict__ binding,
                const float* __restrict__ lxyz, const float* __restrict__ lscal, const float* __restrict__ lrot,
                const float* __restrict__ g_xyz, const float* __restrict__ g_scal, const float* __restrict__ g_rot,
                const int32_t* __restrict__ faces, const float* __restrict__ bary, const float* __restrict__ scale_rem,
                float* __restrict__ gl_xyz, float* __restrict__ gl_scal, float* __restrict__ gl_rot,
                float* __restrict__ gF, float* __restrict__ g_verts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int f = binding[i];
    FaceFrame fr;
    load_frame(frames, f, fr);
    const float4 q = reinterpret_cast<const float4*>(lrot)[i];
    const float lr[4] = {q.x, q.y, q.z, q.w};
    const V3 gx = g_xyz ? ld3(g_xyz, i) : v3(0, 0, 0), gs = g_scal ? ld3(g_scal, i) : v3(0, 0, 0);
    float gr[4] = {0.f, 0.f, 0.f, 0.f};
    if (g_rot) {
        const float4 g4 = reinterpret_cast<const float4*>(g_rot)[i];
        gr[0] = g4.x; gr[1] = g4.y; gr[2] = g4.z; gr[3] = g4.w;
    }
    V3 a, b;
    float gq[4], g17[17];
    bind_bwd_ex(fr, scale_rem ? scale_rem[f] : fr.scale, scale_rem == nullptr, ld3(lxyz, i), ld3(lscal, i), lr, gx, gs, gr,
                a, b, gq, g17);
    if (bary && g_verts) {       // barycentric anchor: its gradient goes straight to the face's three vertices
        const V3 bc = ld3(bary, i);
        const float w3[3] = {bc.x, bc.y, bc.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float* dst = g_verts + 3 * (size_t)faces[3 * f + k];
            atomicAdd(dst, w3[k] * gx.x); atomicAdd(dst + 1, w3[k] * gx.y); atomicAdd(dst + 2, w3[k] * gx.z);
        }
        g17[10] = g17[11] = g17[12] = 0.f;
    }
    if (gl_xyz) st3(gl_xyz, i, a);
    if (gl_scal) st3(gl_scal, i, b);
    if (gl_rot) reinterpret_cast<float4*>(gl_rot)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
    if (gF) {
        float* dst = gF + 17 * (size_t)f;
#pragma unroll
        for (int k = 0; k < 17; k++) atomicAdd(dst + k, g17[k]);
    }
}

__global__ void __launch_bounds__(256)
face_frame_bwd_kernel(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                      const float* __restrict__ gF, float* __restrict__ g_verts) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    FaceFrame fr;
    FaceAux ax;
    face_frame_fwd(ld3(verts, i0), ld3(verts, i1), ld3(verts, i2), fr, ax);     // recompute: cheaper than 100+ B/face of saved state
    float g[17];
#pragma unroll
    for (int k = 0; k < 17; k++) g[k] = gF[17 * (size_t)f + k];
    V3 g0, g1, g2;
    face_frame_bwd(fr, ax, g, g[9], v3(g[10], g[11], g[12]), g + 13, g0, g1, g2);
    atomicAdd(g_verts + 3 * (size_t)i0, g0.x); atomicAdd(g_verts + 3 * (size_t)i0 + 1, g0.y); atomicAdd(g_verts + 3 * (size_t)i0 + 2, g0.z);
    atomicAdd(g_verts + 3 * (size_t)i1, g1.x); atomicAdd(g_verts + 3 * (size_t)i1 + 1, g1.y); atomicAdd(g_verts + 3 * (size_t)i1 + 2, g1.z);
    atomicAdd(g_verts + 3 * (size_t)i2, g2.x); atomicAdd(g_verts + 3 * (size_t)i2 + 1, g2.y); atomicAdd(g_verts + 3 * (size_t)i2 + 2, g2.z);
}

int launch_mesh_bind_forward(int F, int N, const float* verts, const int32_t* faces, const int32_t* binding,
                             const float* lxyz, const float* lscal, const float* lrot, const float* bary,
                             const float* scale_rem, float* frames, float* o_xyz, float* o_scal, float* o_rot,
                             cudaStream_t s) {
    int n = 0;
    if (F > 0) { face_frame_kernel<<<(F + 255) / 256, 256, 0, s>>>(F, verts, faces, frames); n++; }
    if (N > 0) {
        bind_fwd_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, frames, binding, lxyz, lscal, lrot, verts, faces, bary, scale_rem,
                                                        o_xyz, o_scal, o_rot);
        n++;
    }
    return n;
}

int launch_mesh_bind_backward(int F, int N, const float* verts, const int32_t* faces, const int32_t* binding,
                              const float* lxyz, const float* lscal, const float* lrot, const float* bary,
                              const float* scale_rem, const float* frames, const float* g_xyz, const float* g_scal,
                              const float* g_rot, float* gF, float* g_verts, float* gl_xyz, float* gl_scal, float* gl_rot,
                              cudaStream_t s) {
    int n = 0;
    if (N > 0) {
        bind_bwd_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, frames, binding, lxyz, lscal, lrot, g_xyz, g_scal, g_rot, faces,
                                                        bary, scale_rem, gl_xyz, gl_scal, gl_rot,
                                                        g_verts ? gF : nullptr, g_verts);
        n++;
    }
    if (F > 0 && g_verts) { face_frame_bwd_kernel<<<(F + 255) / 256, 256, 0, s>>>(F, verts, faces, gF, g_verts); n++; }
    return n;
}

}  // namespace gg
