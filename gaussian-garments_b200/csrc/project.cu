// project.cu -- forward stage 1:
//   project_kernel   per-Gaussian cull, EWA 3D->2D covariance, conic, radius, tile rectangle, and
//                    per-tile instance counts (replaces upstream preprocessCUDA + tiles_touched)
//   tile_scan_kernel exclusive scan of the per-tile counts -> tile ranges + K
//                    (replaces InclusiveSum over Gaussians AND identifyTileRanges: ranges come
//                    straight out of the scan because instances are bucketed by tile)
//   sh_color_kernel  SH(deg<=3) -> RGB (+0.5, clamp), staged through padded shared memory with
//                    16-byte async copies so the 192 B/Gaussian read is fully coalesced
//   mark_visible_kernel
#include "common.cuh"

namespace gg {

// ---------------------------------------------------------------------------------------------
// Per-tile instance counts.  Must be called by full warps.
constexpr int COOP_TILES = 16;   // splats touching more tiles than this are walked by the whole warp

__device__ __forceinline__ void tile_count_aggregated(bool live, int x0, int y0, int x1, int y1, int gx,
                                                      uint32_t* __restrict__ tile_count) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nt_all = live ? (x1 - x0) * (y1 - y0) : 0;
    const bool big = nt_all > COOP_TILES;
    // small footprints: every lane walks its own rectangle, same-tile lanes share one red.add
    const int nt = big ? 0 : nt_all;
    const int max_nt = __reduce_max_sync(FULL, nt);
    int tx = x0, ty = y0;
    for (int k = 0; k < max_nt; k++) {
        const bool valid = k < nt;
        const int t = valid ? ty * gx + tx : -1 - lane;          // invalid lanes never match anyone
        const unsigned grp = __match_any_sync(FULL, t);
        if (valid && lane == __ffs(grp) - 1) atomicAdd(&tile_count[t], (uint32_t)__popc(grp));
        if (++tx == x1) { tx = x0; ty++; }
    }
    // large footprints (dense / close-up scenes): the warp walks one splat's tiles together, 32 tiles per step
    unsigned bigmask = __ballot_sync(FULL, big);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const int bx0 = __shfl_sync(FULL, x0, src), by0 = __shfl_sync(FULL, y0, src);
        const int bx1 = __shfl_sync(FULL, x1, src), by1 = __shfl_sync(FULL, y1, src);
        const int w = bx1 - bx0, total = w * (by1 - by0);
        for (int k = lane; k < total; k += 32) atomicAdd(&tile_count[(by0 + k / w) * gx + bx0 + k % w], 1u);
    }
}

__global__ void __launch_bounds__(256)
project_kernel(int N, const float* __restrict__ means3D, const float* __restrict__ scales,
               const float* __restrict__ rotations, const float* __restrict__ cov3D_precomp,
               const float* __restrict__ opacities, const float* __restrict__ viewmatrix,
               const float* __restrict__ projmatrix, int W, int H, int gx, int gy, float tanfovx, float tanfovy,
               float mod, GeomWS g, uint32_t* __restrict__ tile_count, int32_t* __restrict__ radii) {
    __shared__ float cam[32];
    if (threadIdx.x < 16) cam[threadIdx.x] = viewmatrix[threadIdx.x];
    else if (threadIdx.x < 32) cam[threadIdx.x] = projmatrix[threadIdx.x - 16];
    __syncthreads();
    const int i_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i_raw < N;
    const int i = in_range ? i_raw : N - 1;     // tail lanes recompute the last Gaussian but never store/count
    const float* V = cam;
    const float* Pm = cam + 16;
    const float x = means3D[3 * (size_t)i], y = means3D[3 * (size_t)i + 1], z = means3D[3 * (size_t)i + 2];

    int rad_out = 0;
    uint2 rect = make_uint2(0u, 0u);
    float2 xy = make_float2(0.f, 0.f);
    float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ext = make_float4(0.f, 0.f, 0.f, 0.f);
    float depth = 0.f;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;

    float c6[6];
    if (cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * (size_t)i + k];
    } else {
        const float s[3] = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
        const float4 q4 = reinterpret_cast<const float4*>(rotations)[i];
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        cov3d_from_scale_rot(s, mod, q, c6);
    }
    const float focal_x = W / (2.0f * tanfovx), focal_y = H / (2.0f * tanfovy);
    Ewa e;
    ewa_project(V, x, y, z, c6, focal_x, focal_y, tanfovx, tanfovy, e);
    if (e.tvz > NEAR_Z && e.det != 0.0f) {
        const float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
        const float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
        const float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
        const float pw = 1.0f / (hw + 0.0000001f);
        const float det_inv = 1.f / e.det;
        const float mid = 0.5f * (e.a + e.c);
        const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - e.det));
        const float rad = ceilf(3.f * sqrtf(lam));
        const float px = ((hx * pw + 1.0f) * W - 1.0f) * 0.5f;
        const float py = ((hy * pw + 1.0f) * H - 1.0f) * 0.5f;
        if (isfinite(rad) && isfinite(px) && isfinite(py)) {
            // (int)(v/16) with clamp to [0, grid]: clamp-then-truncate == truncate-then-clamp
            x0 = (int)fminf(fmaxf((px - rad) / (float)TILE, 0.f), (float)gx);
            y0 = (int)fminf(fmaxf((py - rad) / (float)TILE, 0.f), (float)gy);
            x1 = (int)fminf(fmaxf((px + rad + (float)(TILE - 1)) / (float)TILE, 0.f), (float)gx);
            y1 = (int)fminf(fmaxf((py + rad + (float)(TILE - 1)) / (float)TILE, 0.f), (float)gy);
            if ((x1 - x0) * (y1 - y0) > 0) {
                rad_out = (int)rad;                  // radii / visibility keep the upstream 3-sigma semantics
                const float op = opacities[i];
                xy = make_float2(px, py);
                con = make_float4(e.c * det_inv, -e.b * det_inv, e.a * det_inv, op);
                depth = e.tvz;
                // Exact tile culling: inside the upstream rectangle keep only tiles that the bounding box of
                // the alpha >= 1/255 ellipse reaches.  A dropped tile has no pixel that could pass the alpha
                // test, so images and gradients are unchanged while K shrinks (~16 % on the cfg2 scene).
                float ex, ey;
                if (alpha_extent(op, e.a, e.c, ex, ey)) {
                    // footprint record for the per-instance warp-overlap mask (common.cuh: pack_record)
                    const bool regular = (con.x * con.z - con.y * con.y > 0.f) && con.x > 0.f && con.z > 0.f;
                    ext = make_float4(ex, ey, logf(255.0f * op) * 1.001f + 1e-3f, regular ? 1.f : 2.f);
                    const int sx0 = (int)fmaxf(ceilf((px - ex - (float)(TILE - 1)) / (float)TILE), 0.f);
                    const int sy0 = (int)fmaxf(ceilf((py - ey - (float)(TILE - 1)) / (float)TILE), 0.f);
                    const int sx1 = (int)fminf(floorf((px + ex) / (float)TILE) + 1.f, (float)gx);
                    const int sy1 = (int)fminf(floorf((py + ey) / (float)TILE) + 1.f, (float)gy);
                    x0 = max(x0, sx0); y0 = max(y0, sy0); x1 = min(x1, sx1); y1 = min(y1, sy1);
                    if (x1 <= x0 || y1 <= y0) { x1 = x0; y1 = y0; }
                } else {
                    x1 = x0; y1 = y0;
                }
                rect = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
            }
        }
    }
    if (in_range) {
        radii[i] = rad_out;
        g.xy[i] = xy;
        g.depth[i] = depth;
        g.conic_o[i] = con;
        g.rect[i] = rect;
        g.ext[i] = ext;
    } else {
        rad_out = 0;
    }
    // per-tile instance counts, warp-aggregated: neighbouring Gaussians (6 per mesh face) mostly hit the
    // same tiles, so lanes that target the same counter elect one leader per iteration (match.any)
    tile_count_aggregated(rad_out > 0, x0, y0, x1, y1, gx, tile_count);
    // (Scanning the counters in the last block to retire was tried: the __threadfence every block then needs before
    //  taking its ticket waits for the block's outstanding counter atomics -- 27 % of this kernel's stall samples,
    //  +9.5 us -- more than the separate, latency-optimised scan kernel below costs.)
}

// ---------------------------------------------------------------------------------------------
// One CTA of 1024 threads scans all T tile counts, 8192 per round: every thread owns 8 consecutive counters (two
// 16-byte loads issued back to back, two 16-byte stores), warp-shuffle scan, one shared-memory hop across the 32
// warps.  A 1080p image (8160 tiles) is ONE round: the kernel is a single L2 round trip plus ~100 instructions
// (round 1's version walked its chunk with two scalar, non-unrolled loops: 16 serialized L2 latencies, 9.7 us).
// With `order` != NULL the same launch also emits a launch order for the per-tile kernels: tile indices sorted by
// load class, heaviest first (longest-processing-time-first scheduling).  On the cfg2 scene only ~1 900 of 8 160 tiles
// hold instances, ~700 each and up to ~2 000: started in image order, a 2 000-instance tile that is scheduled late IS
// the tail of the kernel; started first, it overlaps with the rest (measured: blend_bwd 0.272 -> 0.229 ms, blend_fwd
// 0.167 -> 0.140 ms, sort_pack 0.113 -> 0.095 ms).

__global__ void __launch_bounds__(1024) tile_scan_kernel(int T, const uint32_t* __restrict__ count,
                                                         uint32_t* __restrict__ offset, uint32_t* __restrict__ misc,
                                                         uint32_t* __restrict__ order) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    uint32_t vmax = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(count) | reinterpret_cast<uintptr_t>(offset)) & 15u) == 0;
    __syncthreads();
    for (int base = 0; base < T; base += 8192) {
        const int i0 = base + tid * 8;
        uint32_t c[8];
        const bool full = vec && i0 + 8 <= T;
        if (full) {
            const uint4 a = *reinterpret_cast<const uint4*>(count + i0);
            const uint4 b = *reinterpret_cast<const uint4*>(count + i0 + 4);
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = (i0 + k < T) ? count[i0 + k] : 0u;
        }
        uint32_t local = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { local += c[k]; vmax = max(vmax, c[k]); }
        uint32_t v = local;                                   // inclusive warp scan of the per-thread sums
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += n;
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        if (wid == 0) {                                       // scan of the 32 warp totals
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        uint32_t run = s_carry + (wid > 0 ? s_warp[wid - 1] : 0u) + v - local;     // exclusive prefix of this thread's 8
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { o[k] = run; run += c[k]; }
        if (full) {
            *reinterpret_cast<uint4*>(offset + i0) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(offset + i0 + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (i0 + k < T) offset[i0 + k] = o[k];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;                       // running total
        __syncthreads();
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if (lane == 0) s_warp[wid] = vmax;
    __syncthreads();
    if (tid == 0) {
        uint32_t mx = 0;
        for (int w = 0; w < 32; w++) mx = max(mx, s_warp[w]);
        offset[T] = s_carry;
        misc[0] = s_carry;                                    // K = num_rendered
        misc[1] = mx;                                         // largest per-tile instance count (picks the sort variant)
    }
    if (order == nullptr) return;
    // ---- launch order: counting sort of the tiles by load class (64 classes of 32 instances, heaviest first).  The
    // order inside a class is whatever the shared-memory atomics yield: it only decides which CTA starts first.
    __shared__ uint32_t s_hist[64];
    __syncthreads();
    if (tid < 64) s_hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < T; i += 1024) atomicAdd(&s_hist[63 - min(63u, count[i] >> 5)], 1u);
    __syncthreads();
    if (wid == 0) {                                           // exclusive scan of the 64 class sizes (2 per lane)
        const uint32_t a = s_hist[2 * lane], b = s_hist[2 * lane + 1];
        uint32_t v = a + b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += n;
        }
        s_hist[2 * lane] = v - a - b;
        s_hist[2 * lane + 1] = v - b;
    }
    __syncthreads();
    for (int i = tid; i < T; i += 1024) order[atomicAdd(&s_hist[63 - min(63u, count[i] >> 5)], 1u)] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------
// SH -> RGB.  M == 16 fast path: a CTA of 128 threads stages its 128 x 192 B contiguous slab with
// coalesced 16-byte cp.async into rows padded to 13 x 16 B (LDS.128 conflict-free), then every
// thread evaluates its own Gaussian from registers.
constexpr int SH_BLOCK = 128;
constexpr int SH_ROW_U = 13;  // padded row stride in 16-byte units (12 used)

__device__ __forceinline__ void sh_to_rgb(int deg, const float* sh /*[M*3]*/, float dx, float dy, float dz, float* rgb) {
    float b[16];
    sh_basis(deg, dx, dy, dz, b);
    const int nb = (deg + 1) * (deg + 1);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < nb) {
            r0 += b[k] * sh[3 * k];
            r1 += b[k] * sh[3 * k + 1];
            r2 += b[k] * sh[3 * k + 2];
        }
    }
    rgb[0] = fmaxf(r0 + 0.5f, 0.f);
    rgb[1] = fmaxf(r1 + 0.5f, 0.f);
    rgb[2] = fmaxf(r2 + 0.5f, 0.f);
}

// Slab loop: with a full grid (one slab per CTA) the loop runs once; the forked launch (c_api.cu: ForkSet) uses a SMALL
// resident grid striding over the slabs, so that the kernel never queues more CTAs than fit -- the block scheduler
// hands out CTAs in launch order, and a full grid of 2 344 pending CTAs would keep the concurrently running
// emit_kernel's CTAs waiting until the colour kernel has been issued completely.
__global__ void __launch_bounds__(SH_BLOCK)
sh_color16_kernel(int N, int deg, const float* __restrict__ means3D, const float* __restrict__ shs,
                  const float* __restrict__ campos, const int32_t* __restrict__ radii, float* __restrict__ rgb_out) {
    __shared__ __align__(16) float4 rows[SH_BLOCK * SH_ROW_U];
    const int nslabs = (N + SH_BLOCK - 1) / SH_BLOCK;
    const float cx = campos[0], cy = campos[1], cz = campos[2];
    for (int slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
        const int base = slab * SH_BLOCK;
        const int i = base + threadIdx.x;
        const int nG = min(SH_BLOCK, N - base);
        const bool live = (i < N) && (radii[i] > 0);
        if (!__syncthreads_or(live)) continue;   // whole slab culled: skip its 24 KB read
        const float4* src = reinterpret_cast<const float4*>(shs) + (size_t)base * 12;
        for (int u = threadIdx.x; u < nG * 12; u += SH_BLOCK) {
            const int gI = u / 12, j = u - gI * 12;
            cp_async16(&rows[gI * SH_ROW_U + j], src + u);
        }
        cp_async_wait_all();
        __syncthreads();
        if (live) {
            float dx = means3D[3 * (size_t)i] - cx, dy = means3D[3 * (size_t)i + 1] - cy, dz = means3D[3 * (size_t)i + 2] - cz;
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float b[16];
#pragma unroll
            for (int k = 0; k < 16; k++) b[k] = 0.f;          // coefficients above the active degree contribute nothing
            sh_basis(deg, dx * inv, dy * inv, dz * inv, b);
            const int nb = (deg + 1) * (deg + 1);
            // accumulate straight out of the padded row (the 48 coefficients never sit in registers together)
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const float4 q = rows[threadIdx.x * SH_ROW_U + j];
                const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int f = 4 * j + e;                          // flat index = 3 k + channel (compile-time)
                    if (f / 3 < nb) acc[f % 3] += b[f / 3] * v[e];    // coefficients above the active degree are ignored
                }
            }
            rgb_out[3 * (size_t)i] = fmaxf(acc[0] + 0.5f, 0.f);
            rgb_out[3 * (size_t)i + 1] = fmaxf(acc[1] + 0.5f, 0.f);
            rgb_out[3 * (size_t)i + 2] = fmaxf(acc[2] + 0.5f, 0.f);
        }
        __syncthreads();                         // the rows are overwritten by the next slab
    }
}

// generic M (1, 4, 9, ...) or precomputed colours: direct loads
__global__ void __launch_bounds__(256)
sh_color_generic_kernel(int N, int M, int deg, const float* __restrict__ means3D, const float* __restrict__ shs,
                        const float* __restrict__ colors_precomp, const float* __restrict__ campos,
                        const int32_t* __restrict__ radii, float* __restrict__ rgb_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || radii[i] <= 0) return;
    float rgb[3];
    if (colors_precomp) {
        rgb[0] = colors_precomp[3 * (size_t)i]; rgb[1] = colors_precomp[3 * (size_t)i + 1]; rgb[2] = colors_precomp[3 * (size_t)i + 2];
    } else {
        float sh[48];
        const int nb = (deg + 1) * (deg + 1);
#pragma unroll
        for (int k = 0; k < 48; k++) sh[k] = (k < 3 * nb) ? shs[(size_t)i * M * 3 + k] : 0.f;
        float dx = means3D[3 * (size_t)i] - campos[0], dy = means3D[3 * (size_t)i + 1] - campos[1],
              dz = means3D[3 * (size_t)i + 2] - campos[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        sh_to_rgb(deg, sh, dx * inv, dy * inv, dz * inv, rgb);
    }
    rgb_out[3 * (size_t)i] = rgb[0];
    rgb_out[3 * (size_t)i + 1] = rgb[1];
    rgb_out[3 * (size_t)i + 2] = rgb[2];
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(int N, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                    uint8_t* __restrict__ visible) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x = means3D[3 * (size_t)i], y = means3D[3 * (size_t)i + 1], z = means3D[3 * (size_t)i + 2];
    const float tz = viewmatrix[2] * x + viewmatrix[6] * y + viewmatrix[10] * z + viewmatrix[14];
    visible[i] = tz > NEAR_Z ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
int launch_project(const gg_view& v, const gg_inputs& in, const GeomWS& g, const TileWS& t, int32_t* radii,
                   cudaStream_t s) {
    const int N = v.num_gaussians;
    if (N == 0) return 0;
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    project_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, in.means3D, in.scales, in.rotations, in.cov3D_precomp,
                                                   in.opacities, in.viewmatrix, in.projmatrix, v.image_width,
                                                   v.image_height, gx, gy, v.tanfovx, v.tanfovy, v.scale_modifier, g,
                                                   t.count, radii);
    return 1;
}

int launch_tile_scan(int T, const TileWS& t, cudaStream_t s) {
    tile_scan_kernel<<<1, 1024, 0, s>>>(T, t.count, t.offset, t.misc, t.order);
    return 1;
}

int launch_sh_color(const gg_view& v, const gg_inputs& in, const GeomWS& g, const int32_t* radii, cudaStream_t s,
                    int max_blocks) {
    const int N = v.num_gaussians;
    if (N == 0) return 0;
    if (!in.colors_precomp && v.sh_coeffs == 16) {
        int blocks = (N + SH_BLOCK - 1) / SH_BLOCK;
        if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
        sh_color16_kernel<<<blocks, SH_BLOCK, 0, s>>>(N, v.sh_degree, in.means3D, in.shs, in.campos, radii, g.rgb);
    } else {
        sh_color_generic_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, v.sh_coeffs, v.sh_degree, in.means3D, in.shs,
                                                                in.colors_precomp, in.campos, radii, g.rgb);
    }
    return 1;
}

int launch_mark_visible(int N, const float* means3D, const float* viewmatrix, uint8_t* visible, cudaStream_t s) {
    if (N == 0) return 0;
    mark_visible_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, means3D, viewmatrix, visible);
    return 1;
}

}  // namespace gg
