// Host harness: runs the product's mesh-binding math header (gaussian-garments_b200/csrc/mesh_binding_math.h)
// on the CPU so tests can compare its hand-derived backward with autograd without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../gaussian-garments_b200/csrc/mesh_binding_math.h"
using namespace ggmb;

static V3 ld(const float* p, int64_t i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

static V3 anchor_of(const FaceFrame& fr, const float* verts, const int64_t* faces, int64_t f, const float* bary, int64_t i) {
    if (!bary) return fr.center;
    return ld(verts, faces[3 * f]) * bary[3 * i] + ld(verts, faces[3 * f + 1]) * bary[3 * i + 1] + ld(verts, faces[3 * f + 2]) * bary[3 * i + 2];
}

extern "C" void mb_forward(int F, const float* verts, const int64_t* faces, int N, const int64_t* binding,
                           const float* lxyz, const float* lscal, const float* lrot, float* o_xyz, float* o_scal,
                           float* o_rot, float* o_frames /*[F,17] or null*/, const float* bary /*[N,3] or null*/,
                           const float* scale_rem /*[F] or null*/) {
    std::vector<FaceFrame> fr(F);
    for (int f = 0; f < F; f++) {
        FaceAux x;
        face_frame_fwd(ld(verts, faces[3 * f]), ld(verts, faces[3 * f + 1]), ld(verts, faces[3 * f + 2]), fr[f], x);
        if (o_frames) {
            float* o = o_frames + 17 * (size_t)f;
            memcpy(o, fr[f].R, 36); o[9] = fr[f].scale; o[10] = fr[f].center.x; o[11] = fr[f].center.y; o[12] = fr[f].center.z;
            memcpy(o + 13, fr[f].q, 16);
        }
    }
    for (int i = 0; i < N; i++) {
        BindOut o;
        const int64_t f = binding[i];
        bind_fwd_ex(fr[f], anchor_of(fr[f], verts, faces, f, bary, i), scale_rem ? scale_rem[f] : fr[f].scale, ld(lxyz, i),
                    ld(lscal, i), lrot + 4 * (size_t)i, o);
        o_xyz[3 * i] = o.xyz.x; o_xyz[3 * i + 1] = o.xyz.y; o_xyz[3 * i + 2] = o.xyz.z;
        o_scal[3 * i] = o.scaling.x; o_scal[3 * i + 1] = o.scaling.y; o_scal[3 * i + 2] = o.scaling.z;
        memcpy(o_rot + 4 * (size_t)i, o.rot, 16);
    }
}

extern "C" void mb_backward(int V, int F, const float* verts, const int64_t* faces, int N, const int64_t* binding,
                            const float* lxyz, const float* lscal, const float* lrot, const float* g_xyz,
                            const float* g_scal, const float* g_rot, float* gv, float* gl_xyz, float* gl_scal,
                            float* gl_rot, const float* bary, const float* scale_rem) {
    std::vector<FaceFrame> fr(F);
    std::vector<FaceAux> ax(F);
    std::vector<float> gF((size_t)F * 17, 0.f);
    for (int f = 0; f < F; f++)
        face_frame_fwd(ld(verts, faces[3 * f]), ld(verts, faces[3 * f + 1]), ld(verts, faces[3 * f + 2]), fr[f], ax[f]);
    memset(gv, 0, sizeof(float) * 3 * (size_t)V);
    for (int i = 0; i < N; i++) {
        V3 a, b;
        float gq[4], g17[17];
        const int64_t f = binding[i];
        bind_bwd_ex(fr[f], scale_rem ? scale_rem[f] : fr[f].scale, scale_rem == nullptr, ld(lxyz, i), ld(lscal, i),
                    lrot + 4 * (size_t)i, ld(g_xyz, i), ld(g_scal, i), g_rot + 4 * (size_t)i, a, b, gq, g17);
        if (bary) {       // same scatter as bind_bwd_kernel
            for (int k = 0; k < 3; k++) {
                const int64_t vi = faces[3 * f + k];
                gv[3 * vi] += bary[3 * i + k] * g_xyz[3 * i]; gv[3 * vi + 1] += bary[3 * i + k] * g_xyz[3 * i + 1];
                gv[3 * vi + 2] += bary[3 * i + k] * g_xyz[3 * i + 2];
            }
            g17[10] = g17[11] = g17[12] = 0.f;
        }
        gl_xyz[3 * i] = a.x; gl_xyz[3 * i + 1] = a.y; gl_xyz[3 * i + 2] = a.z;
        gl_scal[3 * i] = b.x; gl_scal[3 * i + 1] = b.y; gl_scal[3 * i + 2] = b.z;
        memcpy(gl_rot + 4 * (size_t)i, gq, 16);
        for (int k = 0; k < 17; k++) gF[(size_t)binding[i] * 17 + k] += g17[k];
    }
    for (int f = 0; f < F; f++) {
        const float* g = gF.data() + (size_t)f * 17;
        V3 g0, g1, g2;
        face_frame_bwd(fr[f], ax[f], g, g[9], v3(g[10], g[11], g[12]), g + 13, g0, g1, g2);
        const V3 gs[3] = {g0, g1, g2};
        for (int k = 0; k < 3; k++) {
            const int64_t vi = faces[3 * f + k];
            gv[3 * vi] += gs[k].x; gv[3 * vi + 1] += gs[k].y; gv[3 * vi + 2] += gs[k].z;
        }
    }
}
