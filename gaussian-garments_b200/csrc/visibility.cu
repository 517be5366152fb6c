// visibility.cu -- on-device first-hit ray casting from ONE origin ("next" row N3 of SURVEY.md 8f).
//
// The reference builds its per-frame visibility mask by copying the garment mesh and every Gaussian anchor to the
// host, casting one ray per Gaussian from the camera centre through open3d's RaycastingScene (CPU Embree) and copying
// the mask back (/root/reference/scene/avatar_gaussian_model.py:227-263, /root/reference/inference.py:285-316):
//     ray_o = camera;  ray_d = (x_i - camera) / |x_i - camera|;  ans = scene.cast_rays(rays)
//     vis_i = (ans.primitive_ids[i] == binding[i])                      (avatar model)
//     vis_i = (ans.geometry_ids[i] == garment_of(i)) | (no hit)         (inference)
// Here the cast never leaves the GPU.  All rays share their origin, so "which triangle does ray i hit first" is a
// rasterisation question: project the mesh once through a pinhole at the origin looking at `look_at`, bin the
// triangles' projected bounding boxes into a G x G grid (count -> scan -> fill, the same pattern as the tile binner),
// and let every ray test only the triangles of the cell its direction falls into.  The per-pair test is the exact
// two-sided Moller-Trumbore intersection in world space (closest t, ties to the lower triangle index), so the grid is
// only an exact accelerator: bounding boxes are inflated far beyond fp32 rounding, and whenever the projection is
// not valid (a vertex at or behind the pinhole plane) or the cell lists overflow the workspace, a brute-force
// kernel that tests all F triangles per ray produces the same answer.
// Compiled with -fmad=false: the arithmetic is then bit-identical to the CPU oracle (oracle/raycast_oracle.c).
#include "common.cuh"

namespace gg {

constexpr int VIS_G = 128;                     // grid cells per axis
constexpr int VIS_CELLS = VIS_G * VIS_G;
constexpr float VIS_ZMIN = 1e-4f;              // vertices closer than this to the pinhole plane disable the grid
// barycentric slack: a ray through a shared edge must hit at least one of the two triangles (Embree is watertight;
// plain Moller-Trumbore can reject both by one ulp).  Same constant in oracle/raycast_oracle.c.
#define RAY_EDGE_EPS 1e-6f

struct VisWS {
    float* proj;        // [V,3] (u, w, z) per vertex
    uint32_t* count;    // [CELLS]
    uint32_t* fill;     // [CELLS]
    uint32_t* offset;   // [CELLS+1]
    uint32_t* misc;     // [8]: 0 = total list length, 1 = largest cell, 2 = fallback flag, 4..7 = bounds (ordered ints)
    uint32_t* list;     // [capacity] triangle ids, bucketed by cell
};
inline size_t vis_layout(void* base, int64_t V, int64_t capacity, VisWS* ws) {
    char* p = (char*)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char* q = p ? p + o : nullptr; o += align_up(bytes); return q; };
    VisWS w;
    w.count = (uint32_t*)take((size_t)VIS_CELLS * 4);        // count | fill | misc are adjacent: one memset
    w.fill = (uint32_t*)take((size_t)VIS_CELLS * 4);
    w.misc = (uint32_t*)take(8 * 4);
    w.offset = (uint32_t*)take((size_t)(VIS_CELLS + 1) * 4);
    w.proj = (float*)take((size_t)(V > 0 ? V : 1) * 12);
    w.list = (uint32_t*)take((size_t)(capacity > 0 ? capacity : 1) * 4);
    if (ws) *ws = w;
    return o;
}

#ifdef __CUDACC__
// order-preserving float <-> uint (for atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct Basis { float ox, oy, oz, rx, ry, rz, ux, uy, uz, fx, fy, fz; };
// pinhole at `origin` looking at `look_at` (any deterministic orthonormal completion will do)
__device__ __forceinline__ Basis make_basis(const float* __restrict__ origin, const float* __restrict__ look_at) {
    Basis b;
    b.ox = origin[0]; b.oy = origin[1]; b.oz = origin[2];
    float fx = look_at[0] - b.ox, fy = look_at[1] - b.oy, fz = look_at[2] - b.oz;
    float n = sqrtf(fx * fx + fy * fy + fz * fz);
    if (!(n > 0.f)) { fx = 0.f; fy = 0.f; fz = 1.f; n = 1.f; }
    fx /= n; fy /= n; fz /= n;
    // helper axis: the world axis least aligned with f
    float ax = 1.f, ay = 0.f, az = 0.f;
    if (fabsf(fx) > fabsf(fy) || fabsf(fx) > fabsf(fz)) {
        if (fabsf(fy) <= fabsf(fz)) { ax = 0.f; ay = 1.f; } else { ax = 0.f; az = 1.f; }
    }
    float rx = ay * fz - az * fy, ry = az * fx - ax * fz, rz = ax * fy - ay * fx;
    const float rn = sqrtf(rx * rx + ry * ry + rz * rz);
    rx /= rn; ry /= rn; rz /= rn;
    b.rx = rx; b.ry = ry; b.rz = rz;
    b.ux = fy * rz - fz * ry; b.uy = fz * rx - fx * rz; b.uz = fx * ry - fy * rx;
    b.fx = fx; b.fy = fy; b.fz = fz;
    return b;
}
__device__ __forceinline__ void project(const Basis& b, float x, float y, float z, float& u, float& w, float& d) {
    const float px = x - b.ox, py = y - b.oy, pz = z - b.oz;
    d = px * b.fx + py * b.fy + pz * b.fz;
    const float inv = 1.0f / d;
    u = (px * b.rx + py * b.ry + pz * b.rz) * inv;
    w = (px * b.ux + py * b.uy + pz * b.uz) * inv;
}

// Two-sided Moller-Trumbore.  Same expression order as oracle/raycast_oracle.c (no FMA contraction on either side).
__device__ __forceinline__ bool ray_tri(float ox, float oy, float oz, float dx, float dy, float dz, const float* __restrict__ v,
                                        int i0, int i1, int i2, float& t_out) {
    const float ax = v[3 * (size_t)i0], ay = v[3 * (size_t)i0 + 1], az = v[3 * (size_t)i0 + 2];
    const float e1x = v[3 * (size_t)i1] - ax, e1y = v[3 * (size_t)i1 + 1] - ay, e1z = v[3 * (size_t)i1 + 2] - az;
    const float e2x = v[3 * (size_t)i2] - ax, e2y = v[3 * (size_t)i2 + 1] - ay, e2z = v[3 * (size_t)i2 + 2] - az;
    const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.f) return false;
    const float inv = 1.0f / det;
    const float tx = ox - ax, ty = oy - ay, tz = oz - az;
    const float bu = (tx * px + ty * py + tz * pz) * inv;
    if (bu < -RAY_EDGE_EPS || bu > 1.f + RAY_EDGE_EPS) return false;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const float bv = (dx * qx + dy * qy + dz * qz) * inv;
    if (bv < -RAY_EDGE_EPS || bu + bv > 1.f + RAY_EDGE_EPS) return false;
    const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    if (!(t > 0.f)) return false;
    t_out = t;
    return true;
}

__global__ void __launch_bounds__(256)
vis_vertex_kernel(int V, const float* __restrict__ verts, const float* __restrict__ origin,
                  const float* __restrict__ look_at, float* __restrict__ proj, uint32_t* __restrict__ misc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Basis b = make_basis(origin, look_at);
    float u = 0.f, w = 0.f, d = 1.f;
    bool bad = false;
    if (i < V) {
        project(b, verts[3 * (size_t)i], verts[3 * (size_t)i + 1], verts[3 * (size_t)i + 2], u, w, d);
        bad = !(d > VIS_ZMIN) || !isfinite(u) || !isfinite(w);
        proj[3 * (size_t)i] = u; proj[3 * (size_t)i + 1] = w; proj[3 * (size_t)i + 2] = d;
    }
    const unsigned FULL = 0xffffffffu;
    if (__any_sync(FULL, bad)) { if ((threadIdx.x & 31) == 0) atomicOr(&misc[2], 1u); }
    const bool ok = i < V && !bad;
    const uint32_t lo_u = __reduce_min_sync(FULL, ok ? f2ord(u) : 0xffffffffu), hi_u = __reduce_max_sync(FULL, ok ? f2ord(u) : 0u);
    const uint32_t lo_w = __reduce_min_sync(FULL, ok ? f2ord(w) : 0xffffffffu), hi_w = __reduce_max_sync(FULL, ok ? f2ord(w) : 0u);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&misc[4], lo_u); atomicMax(&misc[5], hi_u);
        atomicMin(&misc[6], lo_w); atomicMax(&misc[7], hi_w);
    }
}

struct GridMap { float u0, w0, su, sw; };     // cell = (coord - c0) * s
__device__ __forceinline__ GridMap grid_map(const uint32_t* __restrict__ misc) {
    GridMap g;
    const float u0 = ord2f(misc[4]), u1 = ord2f(misc[5]), w0 = ord2f(misc[6]), w1 = ord2f(misc[7]);
    const float eu = fmaxf(u1 - u0, 1e-12f), ew = fmaxf(w1 - w0, 1e-12f);
    g.u0 = u0 - 0.01f * eu; g.w0 = w0 - 0.01f * ew;          // 1 % margin around the mesh
    g.su = (float)VIS_G / (1.02f * eu); g.sw = (float)VIS_G / (1.02f * ew);
    return g;
}
// cell range of a triangle's projected bounding box, inflated by a quarter cell (>> fp32 rounding of the projection)
__device__ __forceinline__ bool tri_cells(const GridMap& g, const float* __restrict__ proj, int i0, int i1, int i2,
                                          int& x0, int& x1, int& y0, int& y1) {
    const float ua = proj[3 * (size_t)i0], ub = proj[3 * (size_t)i1], uc = proj[3 * (size_t)i2];
    const float wa = proj[3 * (size_t)i0 + 1], wb = proj[3 * (size_t)i1 + 1], wc = proj[3 * (size_t)i2 + 1];
    const float fx0 = (fminf(ua, fminf(ub, uc)) - g.u0) * g.su - 0.25f, fx1 = (fmaxf(ua, fmaxf(ub, uc)) - g.u0) * g.su + 0.25f;
    const float fy0 = (fminf(wa, fminf(wb, wc)) - g.w0) * g.sw - 0.25f, fy1 = (fmaxf(wa, fmaxf(wb, wc)) - g.w0) * g.sw + 0.25f;
    x0 = max(0, (int)floorf(fx0)); x1 = min(VIS_G - 1, (int)floorf(fx1));
    y0 = max(0, (int)floorf(fy0)); y1 = min(VIS_G - 1, (int)floorf(fy1));
    return x0 <= x1 && y0 <= y1;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
vis_bin_kernel(int F, const int32_t* __restrict__ faces, const float* __restrict__ proj, uint32_t* __restrict__ misc,
               uint32_t* __restrict__ count, const uint32_t* __restrict__ offset, uint32_t* __restrict__ fill,
               uint32_t* __restrict__ list, uint32_t capacity) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F || misc[2]) return;
    if (FILL && misc[0] > capacity) return;            // lists do not fit: the brute-force kernel takes over
    const GridMap g = grid_map(misc);
    int x0, x1, y0, y1;
    if (!tri_cells(g, proj, faces[3 * f], faces[3 * f + 1], faces[3 * f + 2], x0, x1, y0, y1)) return;
    for (int y = y0; y <= y1; y++)
        for (int x = x0; x <= x1; x++) {
            const int c = y * VIS_G + x;
            if (FILL) list[offset[c] + atomicAdd(&fill[c], 1u)] = (uint32_t)f;
            else atomicAdd(&count[c], 1u);
        }
}

__global__ void __launch_bounds__(256)
vis_ray_kernel(int N, const float* __restrict__ targets, const float* __restrict__ origin, const float* __restrict__ look_at,
               const float* __restrict__ verts, const int32_t* __restrict__ faces, const uint32_t* __restrict__ misc,
               const uint32_t* __restrict__ offset, const uint32_t* __restrict__ list, uint32_t capacity,
               int32_t* __restrict__ prim, float* __restrict__ t_hit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || misc[2] || misc[0] > capacity) return;
    const Basis b = make_basis(origin, look_at);
    const float tx = targets[3 * (size_t)i], ty = targets[3 * (size_t)i + 1], tz = targets[3 * (size_t)i + 2];
    float dx = tx - b.ox, dy = ty - b.oy, dz = tz - b.oz;
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= nrm; dy /= nrm; dz /= nrm;                      // ray_d = _d / norm_d  (avatar_gaussian_model.py:247-249)
    int best = -1;
    float best_t = 3.0e38f;
    float u, w, d;
    project(b, tx, ty, tz, u, w, d);
    if (d > 0.f && isfinite(u) && isfinite(w)) {
        const GridMap g = grid_map(misc);
        const float cx = (u - g.u0) * g.su, cy = (w - g.w0) * g.sw;
        if (cx >= 0.f && cy >= 0.f && cx < (float)VIS_G && cy < (float)VIS_G) {
            const int c = (int)cy * VIS_G + (int)cx;
            const uint32_t beg = offset[c], end = offset[c + 1];
            for (uint32_t k = beg; k < end; k++) {
                const int f = (int)list[k];
                float t;
                if (ray_tri(b.ox, b.oy, b.oz, dx, dy, dz, verts, faces[3 * f], faces[3 * f + 1], faces[3 * f + 2], t) &&
                    (t < best_t || (t == best_t && f < best))) {
                    best_t = t;
                    best = f;
                }
            }
        }
    }
    prim[i] = best;
    if (t_hit) t_hit[i] = best >= 0 ? best_t : __int_as_float(0x7f800000);
}

// fallback: every ray against every triangle, triangles staged through shared memory (9 floats each)
constexpr int VIS_BF_TILE = 256;
__global__ void __launch_bounds__(256)
vis_bruteforce_kernel(int N, int F, const float* __restrict__ targets, const float* __restrict__ origin,
                      const float* __restrict__ verts, const int32_t* __restrict__ faces, const uint32_t* __restrict__ misc,
                      uint32_t capacity, int32_t* __restrict__ prim, float* __restrict__ t_hit) {
    if (!(misc[2] || misc[0] > capacity)) return;          // the grid path already answered
    __shared__ float tri[VIS_BF_TILE][9];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float ox = origin[0], oy = origin[1], oz = origin[2];
    float dx = 0.f, dy = 0.f, dz = 1.f;
    if (i < N) {
        dx = targets[3 * (size_t)i] - ox; dy = targets[3 * (size_t)i + 1] - oy; dz = targets[3 * (size_t)i + 2] - oz;
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        dx /= nrm; dy /= nrm; dz /= nrm;
    }
    int best = -1;
    float best_t = 3.0e38f;
    for (int base = 0; base < F; base += VIS_BF_TILE) {
        const int f = base + threadIdx.x;
        if (f < F) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int vi = faces[3 * f + k];
                tri[threadIdx.x][3 * k] = verts[3 * (size_t)vi];
                tri[threadIdx.x][3 * k + 1] = verts[3 * (size_t)vi + 1];
                tri[threadIdx.x][3 * k + 2] = verts[3 * (size_t)vi + 2];
            }
        }
        __syncthreads();
        const int cnt = min(VIS_BF_TILE, F - base);
        if (i < N) {
            for (int j = 0; j < cnt; j++) {
                float t;
                if (ray_tri(ox, oy, oz, dx, dy, dz, &tri[j][0], 0, 1, 2, t) && (t < best_t || (t == best_t && base + j < best))) {
                    best_t = t;
                    best = base + j;
                }
            }
        }
        __syncthreads();
    }
    if (i < N) {
        prim[i] = best;
        if (t_hit) t_hit[i] = best >= 0 ? best_t : __int_as_float(0x7f800000);
    }
}
#endif  // __CUDACC__

size_t vis_workspace_bytes(int64_t V, int64_t capacity) { return vis_layout(nullptr, V, capacity, nullptr); }

int launch_cast_rays(int V, int F, int N, const float* verts, const int32_t* faces, const float* targets,
                     const float* origin, const float* look_at, void* ws_base, int64_t capacity, int force_bruteforce,
                     int32_t* prim, float* t_hit, cudaStream_t s) {
    if (N == 0) return 0;
    VisWS w;
    vis_layout(ws_base, V, capacity, &w);
    int n = 0;
    // count | fill | misc in one memset, then the bounds / flag words
    cudaMemsetAsync(w.count, 0, (size_t)((char*)w.offset - (char*)w.count), s);
    const uint32_t init[4] = {0xffffffffu, 0u, 0xffffffffu, 0u};
    cudaMemcpyAsync(w.misc + 4, init, sizeof(init), cudaMemcpyHostToDevice, s);
    if (force_bruteforce || F == 0 || V == 0) {
        const uint32_t one = 1u;
        cudaMemcpyAsync(w.misc + 2, &one, 4, cudaMemcpyHostToDevice, s);
    } else {
        vis_vertex_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, verts, origin, look_at, w.proj, w.misc); n++;
        vis_bin_kernel<false><<<(F + 255) / 256, 256, 0, s>>>(F, faces, w.proj, w.misc, w.count, w.offset, w.fill, w.list, (uint32_t)capacity); n++;
        TileWS t;
        t.count = w.count; t.fill = w.fill; t.offset = w.offset; t.misc = w.misc; t.order = nullptr;     // misc[0] = total, misc[1] = largest cell
        n += launch_tile_scan(VIS_CELLS, t, s);
        vis_bin_kernel<true><<<(F + 255) / 256, 256, 0, s>>>(F, faces, w.proj, w.misc, w.count, w.offset, w.fill, w.list, (uint32_t)capacity); n++;
        vis_ray_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, targets, origin, look_at, verts, faces, w.misc, w.offset, w.list,
                                                       (uint32_t)capacity, prim, t_hit); n++;
    }
    vis_bruteforce_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, F, targets, origin, verts, faces, w.misc, (uint32_t)capacity, prim, t_hit); n++;
    return n;
}

}  // namespace gg
