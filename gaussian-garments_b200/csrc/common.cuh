// common.cuh -- shared declarations of the sm_100a rasterizer kernels (internal; the public
// boundary is include/gg_raster.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gg_raster.h"

namespace gg {

constexpr int TILE = GG_TILE;
constexpr int TILE_PIX = TILE * TILE;
constexpr float NEAR_Z = 0.2f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_STOP = 0.0001f;
constexpr float BLUR = 0.3f;

// ---------------------------------------------------------------------------------------------
// workspace layouts (all sub-buffers 256-byte aligned inside one caller-allocated block)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct GeomWS {        // per Gaussian, written by project / sh_color, read by emit / sort_pack
    float2* xy;        // pixel-space mean
    float* depth;      // view-space z
    float4* conic_o;   // conic (A,B,C) + opacity
    float* rgb;        // [N,3]
    uint2* rect;       // x0 | y0<<16 , x1 | y1<<16   (tile rectangle, exclusive max)
    float4* ext;       // alpha >= 1/255 footprint: (half extent x, half extent y, tau = ln(255 o) inflated, kind)
                       // kind 1 = usable, 2 = degenerate conic (no culling information); per Gaussian, so that the
                       // per-INSTANCE record packing does not redo two logf / two sqrt / three divisions
};
inline size_t geom_layout(void* base, int64_t N, GeomWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    GeomWS w;
    w.xy = (float2*)take((size_t)N * 8);
    w.depth = (float*)take((size_t)N * 4);
    w.conic_o = (float4*)take((size_t)N * 16);
    w.rgb = (float*)take((size_t)N * 12);
    w.rect = (uint2*)take((size_t)N * 8);
    w.ext = (float4*)take((size_t)N * 16);
    if (ws) *ws = w;
    return o;
}

struct TileWS {
    uint32_t* count;   // [T]   instances per tile          (zeroed by stage 1)
    uint32_t* fill;    // [T]   emit cursor                 (zeroed by stage 1)
    uint32_t* offset;  // [T+1] exclusive scan of count     (kept for backward)
    uint32_t* misc;    // [8]   misc[0] = K (num_rendered), misc[1] = largest per-tile count
    uint32_t* order;   // [T]   tile indices, heaviest load class first (64 classes of 32 instances): the blend / sort grids
                       //       walk the tiles in this order so that the longest CTAs start first (shorter tail)
};
inline size_t tile_layout(void* base, int64_t T, TileWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    TileWS w;
    w.count = (uint32_t*)take((size_t)T * 4);       // count | fill | misc are adjacent: stage 1 zeroes them with one memset
    w.fill = (uint32_t*)take((size_t)T * 4);
    w.misc = (uint32_t*)take(8 * 4);
    w.offset = (uint32_t*)take((size_t)(T + 1) * 4);
    w.order = (uint32_t*)take((size_t)T * 4);
    if (ws) *ws = w;
    return o;
}

struct ImageWS {
    uint32_t* n_contrib;  // [P] 1-based list position of the last applied Gaussian
    float* final_T;       // [P] transmittance left after blending
};
inline size_t image_layout(void* base, int64_t P, ImageWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    ImageWS w;
    w.n_contrib = (uint32_t*)take((size_t)P * 4);
    w.final_T = (float*)take((size_t)P * 4);
    if (ws) *ws = w;
    return o;
}

// Packed, depth-sorted per-instance records: three float4 planes so that a tile's list is three
// contiguous 16-byte-aligned runs (1-D bulk-TMA friendly) and every store is fully coalesced.
//   p0 = (px, py, A', B')   p1 = (C', opacity, depth, warp-overlap mask bits)   p2 = (r, g, b, gaussian id bits)
// with the conic pre-scaled into the log2 domain (A' = -0.5*log2e*A, B' = -log2e*B, C' = -0.5*log2e*C).
// Warp-overlap mask: bit w set iff the bounding box of the splat's alpha >= 1/255 ellipse touches the
// 8x4 pixel block that warp w of the tile's CTA owns (w&1 -> x half, w>>1 -> y band); a clear bit
// proves no pixel of that warp can pass the alpha test, so the warp skips the entry (exact).
struct RecordWS {
    float4* p0;
    float4* p1;
    float4* p2;
};
inline size_t record_layout(void* base, int64_t K, RecordWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    if (K < 1) K = 1;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    RecordWS w;
    w.p0 = (float4*)take((size_t)K * 16);
    w.p1 = (float4*)take((size_t)K * 16);
    w.p2 = (float4*)take((size_t)K * 16);
    if (ws) *ws = w;
    return o;
}

// Per-Gaussian gradient accumulators of the blend backward (10 floats, vector-atomic friendly):
//   a0 = (dmean2D.x, dmean2D.y, dconic.A, dconic.B)   a1 = (dconic.C, dopacity, dr, dg)
//   a2 = (db, ddepth)
struct AccumWS {
    float4* a0;
    float4* a1;
    float2* a2;
};
inline size_t accum_layout(void* base, int64_t N, AccumWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    if (N < 1) N = 1;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    AccumWS w;
    w.a0 = (float4*)take((size_t)N * 16);
    w.a1 = (float4*)take((size_t)N * 16);
    w.a2 = (float2*)take((size_t)N * 8);
    if (ws) *ws = w;
    return o;
}

// ---------------------------------------------------------------------------------------------
// kernel launchers (defined in the .cu files; each returns the number of kernels it launched)
// ---------------------------------------------------------------------------------------------
int launch_project(const gg_view& v, const gg_inputs& in, const GeomWS& g, const TileWS& t, int32_t* radii,
                   cudaStream_t s);
int launch_tile_scan(int T, const TileWS& t, cudaStream_t s);
int launch_sh_color(const gg_view& v, const gg_inputs& in, const GeomWS& g, const int32_t* radii, cudaStream_t s,
                    int max_blocks = 0);
int launch_emit(const gg_view& v, const GeomWS& g, const TileWS& t, const int32_t* radii, uint64_t* keys,
                uint32_t capacity, cudaStream_t s);
int launch_color_fill(const gg_view& v, const GeomWS& g, const TileWS& t, const RecordWS& r, uint32_t capacity, cudaStream_t s);
int launch_overflow_flag(const TileWS& t, uint32_t capacity, uint32_t* flag, cudaStream_t s);
int launch_sort_pack(const gg_view& v, const GeomWS& g, const TileWS& t, uint64_t* keys, const RecordWS& r,
                     uint32_t capacity, uint32_t max_tile_instances, bool with_color, cudaStream_t s);
int launch_blend_fwd(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                     uint32_t capacity, float* out_color, float* out_depth, float* out_alpha, cudaStream_t s);
int launch_blend_fwd2(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                      uint32_t capacity, float* out_color, float* out_depth, float* out_alpha, cudaStream_t s);
void set_lazy_phase_counters(unsigned long long* p);
int launch_blend_fwd_lazy(const gg_view& v, const gg_inputs& in, const GeomWS& g, const TileWS& t, uint64_t* keys,
                          uint64_t* keys2, const RecordWS& r, const ImageWS& img, uint32_t capacity, float* out_color,
                          float* out_depth, float* out_alpha, cudaStream_t s);
int launch_blend_bwd(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                     const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, const AccumWS& acc,
                     cudaStream_t s);
int launch_blend_bwd2(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                      const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, const AccumWS& acc,
                      int min_blocks, cudaStream_t s);
int launch_preprocess_bwd(const gg_view& v, const gg_inputs& in, const int32_t* radii, const AccumWS& acc,
                          float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors,
                          float* dL_dopacities, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                          cudaStream_t s);
int launch_mark_visible(int N, const float* means3D, const float* viewmatrix, uint8_t* visible, cudaStream_t s);
int launch_photometric_fwd(int W, int H, const float* img, const float* gt, const float* mask, float* m1, float* m2,
                           float* m3, double* sums, cudaStream_t s);
int launch_photometric_l1_u8(int W, int H, const float* img, const uint8_t* gt8, const float* mask, double* sums,
                             float c_l1, const float* g_scalar, float* g_img, cudaStream_t s);
int launch_photometric_finalize(const double* sums, double inv_n, float lambda_dssim, float* out3, cudaStream_t s);
int launch_photometric_bwd(int W, int H, const float* img, const float* gt, const float* mask, const float* m1,
                           const float* m2, const float* m3, float c_l1, float c_ss, const float* g_scalar, float* g_img,
                           cudaStream_t s);
int launch_mesh_bind_forward(int F, int N, const float* verts, const int32_t* faces, const int32_t* binding,
                             const float* lxyz, const float* lscal, const float* lrot, const float* bary,
                             const float* scale_rem, float* frames, float* o_xyz, float* o_scal, float* o_rot,
                             cudaStream_t s);
int launch_mesh_bind_backward(int F, int N, const float* verts, const int32_t* faces, const int32_t* binding,
                              const float* lxyz, const float* lscal, const float* lrot, const float* bary,
                              const float* scale_rem, const float* frames, const float* g_xyz, const float* g_scal,
                              const float* g_rot, float* gF, float* g_verts, float* gl_xyz, float* gl_scal, float* gl_rot,
                              cudaStream_t s);

int launch_gate_wait(uint32_t* gate, cudaStream_t s);
int launch_gate_signal(uint32_t* gate, cudaStream_t s);
int launch_nvls_allreduce(float* mc, uint32_t* const* pads, int rank, int world, int slot0, int64_t n_vec4, float scale,
                          int blocks, cudaStream_t s);
size_t vis_workspace_bytes(int64_t V, int64_t capacity);
int launch_cast_rays(int V, int F, int N, const float* verts, const int32_t* faces, const float* targets,
                     const float* origin, const float* look_at, void* ws_base, int64_t capacity, int force_bruteforce,
                     int32_t* prim, float* t_hit, cudaStream_t s);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// Ampere-style 16-byte async copy (LDGSTS) for padded (bank-conflict-free) staging
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// explicit shared-window loads: keeps the per-iteration address a single 32-bit add (the generic
// path made ptxas rebuild the shared window base -- S2UR SR_CgaCtaId + 4 ULEA -- every iteration)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float ex2_approx(float x) {   // 2^x, flush-to-zero (MUFU.EX2, no denormal fix-up)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- the same primitives on explicit 32-bit shared-window addresses (one pinned base register + compile-time
// offsets: the generic-pointer forms above make ptxas rebuild the window base -- S2R SR_CgaCtaId + LEA -- per use)
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_test_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldsu32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


// Packed records carry the conic pre-scaled into the log2 domain:
//   A' = -0.5*log2(e)*A,  B' = -log2(e)*B,  C' = -0.5*log2(e)*C   =>   G = 2^(A' dx^2 + B' dx dy + C' dy^2)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ---- spherical harmonics (basis of /root/reference/utils/sh_utils.py:25-42,56-111) ----------
#define GG_SH_C0 0.28209479177387814f
#define GG_SH_C1 0.4886025119029199f
#define GG_SH_C2_0 1.0925484305920792f
#define GG_SH_C2_1 (-1.0925484305920792f)
#define GG_SH_C2_2 0.31539156525252005f
#define GG_SH_C2_3 (-1.0925484305920792f)
#define GG_SH_C2_4 0.5462742152960396f
#define GG_SH_C3_0 (-0.5900435899266435f)
#define GG_SH_C3_1 2.890611442640554f
#define GG_SH_C3_2 (-0.4570457994644658f)
#define GG_SH_C3_3 0.3731763325901154f
#define GG_SH_C3_4 (-0.4570457994644658f)
#define GG_SH_C3_5 1.445305721320277f
#define GG_SH_C3_6 (-0.5900435899266435f)

// basis[k], k < (deg+1)^2, for unit direction (x,y,z)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    b[0] = GG_SH_C0;
    if (deg > 0) {
        b[1] = -GG_SH_C1 * y;
        b[2] = GG_SH_C1 * z;
        b[3] = -GG_SH_C1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = GG_SH_C2_0 * xy;
            b[5] = GG_SH_C2_1 * yz;
            b[6] = GG_SH_C2_2 * (2.0f * zz - xx - yy);
            b[7] = GG_SH_C2_3 * xz;
            b[8] = GG_SH_C2_4 * (xx - yy);
            if (deg > 2) {
                b[9] = GG_SH_C3_0 * y * (3.0f * xx - yy);
                b[10] = GG_SH_C3_1 * xy * z;
                b[11] = GG_SH_C3_2 * y * (4.0f * zz - xx - yy);
                b[12] = GG_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = GG_SH_C3_4 * x * (4.0f * zz - xx - yy);
                b[14] = GG_SH_C3_5 * z * (xx - yy);
                b[15] = GG_SH_C3_6 * x * (xx - 3.0f * yy);
            }
        }
    }
}

// 3x3 rotation from an (un-normalised) wxyz quaternion, entries of
// /root/reference/utils/general_utils.py:100-108
__device__ __forceinline__ void quat_to_rot(float r, float x, float y, float z, float* R) {
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z);       R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);       R[7] = 2.f * (y * z + r * x);       R[8] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = (R S)(R S)^T packed xx,xy,xz,yy,yz,zz (/root/reference/scene/gaussian_model.py:27-31)
__device__ __forceinline__ void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* c6) {
    float R[9];
    quat_to_rot(q[0], q[1], q[2], q[3], R);
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[i * 3 + j] = R[i * 3 + j] * (mod * s[j]);
    c6[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    c6[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    c6[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    c6[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    c6[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    c6[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// Half extents (pixels) of the axis-aligned bounding box of the ellipse {alpha >= 1/255}:
// o*exp(-d^T S^-1 d / 2) >= 1/255  <=>  d^T S^-1 d <= 2 tau, tau = ln(255 o)  =>  |dx| <= sqrt(2 tau S_xx).
// Inflated by 1e-3 relative + 0.02 px so that rounding in the blend's exponent can never matter.
// Returns false when the splat cannot reach alpha >= 1/255 anywhere.
__device__ __forceinline__ bool alpha_extent(float opacity, float cov_xx, float cov_yy, float& ex, float& ey) {
    if (!(opacity > ALPHA_MIN)) return false;
    const float two_tau = 2.0f * logf(255.0f * opacity);
    ex = sqrtf(two_tau * cov_xx) * 1.001f + 0.02f;
    ey = sqrtf(two_tau * cov_yy) * 1.001f + 0.02f;
    return true;
}

// Build the three packed planes of one (tile, Gaussian) instance from its sort key and the projected
// per-Gaussian records (shared by sort_pack_kernel and the lazy fused forward).
__device__ __forceinline__ void pack_record(uint64_t key, const float2* __restrict__ xy,
                                            const float4* __restrict__ conic_o, const float4* __restrict__ ext,
                                            const float* __restrict__ rgb, float tile_x, float tile_y, float4& q0,
                                            float4& q1, float4& q2) {
    const uint32_t id = (uint32_t)(key & 0xffffffffu);
    const float depth = __uint_as_float((uint32_t)(key >> 32));
    const float2 m = xy[id];
    const float4 co = conic_o[id];
    const float4 E = ext[id];
    // rgb == nullptr: colours are filled in later (color_fill_kernel) -- the SH -> RGB kernel may still be waiting for
    // the previous step's SH-gradient exchange while the instances are already being sorted (dist.py)
    float r = 0.f, g = 0.f, b = 0.f;
    if (rgb) { r = rgb[3 * (size_t)id]; g = rgb[3 * (size_t)id + 1]; b = rgb[3 * (size_t)id + 2]; }
    // warp-overlap mask: bit w set iff some point of warp w's 8x4 pixel-centre box can reach alpha >= 1/255, i.e.
    // min over the box of f(d) = (A dx^2 + 2 B dx dy + C dy^2)/2 is <= tau = ln(255 o).  First the ellipse's bounding
    // box (cheap reject; extents and tau come precomputed per Gaussian from project_kernel), then the exact box
    // minimum of the convex quadratic (centre inside -> 0, else the best of the four edge minima).  tau is inflated by
    // 0.1 % + 1e-3 so rounding in the blend's exponent cannot matter.
    uint32_t wmask = 0;
    if (E.w == 1.f) {
        const float ex = E.x, ey = E.y, tau = E.z;
        const float cx = m.x - tile_x, cy = m.y - tile_y;                       // centre in tile-local pixels
        const float A = co.x, B = co.y, C = co.z, iA = 1.0f / co.x, iC = 1.0f / co.z;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float X0 = (float)((w & 1) * 8), X1 = X0 + 7.f, Y0 = (float)((w >> 1) * 4), Y1 = Y0 + 3.f;
            if (cx + ex < X0 || cx - ex > X1 || cy + ey < Y0 || cy - ey > Y1) continue;      // bounding-box reject
            bool hit = (cx >= X0 && cx <= X1 && cy >= Y0 && cy <= Y1);
            if (!hit) {
                float fmin_ = 3.0e38f;
#pragma unroll
                for (int e = 0; e < 2; e++) {                                   // vertical edges x = X0 / X1
                    const float u = (e ? X1 : X0) - cx;
                    const float v = fminf(fmaxf(cy - B * u * iC, Y0), Y1) - cy;
                    fmin_ = fminf(fmin_, 0.5f * (A * u * u + C * v * v) + B * u * v);
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {                                   // horizontal edges y = Y0 / Y1
                    const float v = (e ? Y1 : Y0) - cy;
                    const float u = fminf(fmaxf(cx - B * v * iA, X0), X1) - cx;
                    fmin_ = fminf(fmin_, 0.5f * (A * u * u + C * v * v) + B * u * v);
                }
                hit = fmin_ <= tau;
            }
            if (hit) wmask |= 1u << w;
        }
    } else if (E.w == 2.f) {
        wmask = 0xffu;   // degenerate conic: no culling information
    }
    q0 = make_float4(m.x, m.y, (-0.5f * LOG2E) * co.x, -LOG2E * co.y);
    q1 = make_float4((-0.5f * LOG2E) * co.z, co.w, depth, __uint_as_float(wmask));
    q2 = make_float4(r, g, b, __uint_as_float(id));
}

// Shared EWA projection state of one Gaussian (used by project and by preprocess backward)
struct Ewa {
    float tvx, tvy, tvz;       // view-space position
    float tx, ty;              // clamped view-space x,y used in the Jacobian
    float gate_x, gate_y;      // 0 when the 1.3*tanfov clamp is active
    float T00, T01, T02, T10, T11, T12;   // J * W3
    float u0, u1, u2, v0, v1, v2;         // (J W3) Sigma rows
    float a, b, c, det;                   // Sigma2D (+blur) and its determinant
};
__device__ __forceinline__ void ewa_project(const float* V, float x, float y, float z, const float* c6, float focal_x,
                                            float focal_y, float tanfovx, float tanfovy, Ewa& e) {
    e.tvx = V[0] * x + V[4] * y + V[8] * z + V[12];
    e.tvy = V[1] * x + V[5] * y + V[9] * z + V[13];
    // view depth is the sort key: fixed fused evaluation order so that it is bit-identical to the
    // CPU oracle (oracle/gg_oracle.c uses the same fmaf chain) and near-equal depths order the same
    e.tvz = __fmaf_rn(V[10], z, __fmaf_rn(V[6], y, __fmaf_rn(V[2], x, V[14])));
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = e.tvx / e.tvz, tytz = e.tvy / e.tvz;
    e.tx = fminf(limx, fmaxf(-limx, txtz)) * e.tvz;
    e.ty = fminf(limy, fmaxf(-limy, tytz)) * e.tvz;
    e.gate_x = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    e.gate_y = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J00 = focal_x / e.tvz, J02 = -(focal_x * e.tx) / (e.tvz * e.tvz);
    const float J11 = focal_y / e.tvz, J12 = -(focal_y * e.ty) / (e.tvz * e.tvz);
    e.T00 = J00 * V[0] + J02 * V[2]; e.T01 = J00 * V[4] + J02 * V[6]; e.T02 = J00 * V[8] + J02 * V[10];
    e.T10 = J11 * V[1] + J12 * V[2]; e.T11 = J11 * V[5] + J12 * V[6]; e.T12 = J11 * V[9] + J12 * V[10];
    const float S00 = c6[0], S01 = c6[1], S02 = c6[2], S11 = c6[3], S12 = c6[4], S22 = c6[5];
    e.u0 = e.T00 * S00 + e.T01 * S01 + e.T02 * S02;
    e.u1 = e.T00 * S01 + e.T01 * S11 + e.T02 * S12;
    e.u2 = e.T00 * S02 + e.T01 * S12 + e.T02 * S22;
    e.v0 = e.T10 * S00 + e.T11 * S01 + e.T12 * S02;
    e.v1 = e.T10 * S01 + e.T11 * S11 + e.T12 * S12;
    e.v2 = e.T10 * S02 + e.T11 * S12 + e.T12 * S22;
    e.a = e.u0 * e.T00 + e.u1 * e.T01 + e.u2 * e.T02 + BLUR;
    e.b = e.u0 * e.T10 + e.u1 * e.T11 + e.u2 * e.T12;
    e.c = e.v0 * e.T10 + e.v1 * e.T11 + e.v2 * e.T12 + BLUR;
    e.det = e.a * e.c - e.b * e.b;
}
#endif  // __CUDACC__

}  // namespace gg
