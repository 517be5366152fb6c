// binning.cu -- tile binning and depth ordering, redesigned for B200.
//
// Upstream orders the (tile, depth) instance list with one global 64-bit LSD radix sort
// (>= 6 passes over 12 B/instance).  Here the list is bucketed by tile first -- the per-tile
// ranges are already known from project_kernel's counts + tile_scan_kernel -- so that ordering
// reduces to one independent small sort per tile, done entirely inside a CTA's shared memory
// (B200: 227 KB/SM) and fused with packing the sorted records:
//   emit_kernel       instance -> its tile's range at an atomic cursor; key = depth_bits<<32 | id
//   sort_pack_kernel  per tile: bitonic sort of the 64-bit keys (ties by Gaussian index, i.e. the
//                     order a stable sort of (tile<<32|depth) keys yields), then gather the
//                     Gaussian's projected record and write the three packed planes
// HBM traffic: 8 B write + 8 B read per instance for the keys (vs ~150 B for the radix passes)
// plus the 48 B packed record that both blend passes stream with bulk-TMA.
#include "common.cuh"

namespace gg {

__global__ void __launch_bounds__(256)
emit_kernel(int N, int gx, const int32_t* __restrict__ radii, const uint2* __restrict__ rect,
            const float* __restrict__ depth, const uint32_t* __restrict__ tile_offset,
            uint32_t* __restrict__ tile_fill, uint64_t* __restrict__ keys, uint32_t capacity) {
    // warp-aggregated cursor claims: lanes that target the same tile in the same iteration share one
    // atomicAdd (leader claims popc slots, members take base + their rank)
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = (i < N) && (radii[i] > 0);
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint64_t key = 0;
    if (live) {
        const uint2 r = rect[i];
        x0 = r.x & 0xffff; y0 = r.x >> 16; x1 = r.y & 0xffff; y1 = r.y >> 16;
        key = ((uint64_t)__float_as_uint(depth[i]) << 32) | (uint32_t)i;
    }
    const int nt = (x1 - x0) * (y1 - y0);
    const int max_nt = __reduce_max_sync(FULL, nt);
    int tx = x0, ty = y0;
    for (int k = 0; k < max_nt; k++) {
        const bool valid = k < nt;
        const int t = valid ? ty * gx + tx : -1 - lane;
        const unsigned grp = __match_any_sync(FULL, t);
        const int leader = __ffs(grp) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) base = tile_offset[t] + atomicAdd(&tile_fill[t], (uint32_t)__popc(grp));
        base = __shfl_sync(FULL, base, leader);
        if (valid) {
            const uint32_t slot = base + __popc(grp & ((1u << lane) - 1u));
            if (slot < capacity) keys[slot] = key;
        }
        if (++tx == x1) { tx = x0; ty++; }
    }
}

// ---------------------------------------------------------------------------------------------
// Bitonic network with all comparators ascending ("flip" first step), so elements beyond n act
// as +inf padding that never moves: works for any n without materialising the padding.
template <typename Ptr>
__device__ __forceinline__ void cmpswap(Ptr a, uint32_t lo, uint32_t hi) {
    const uint64_t x = a[lo], y = a[hi];
    if (x > y) {
        a[lo] = y;
        a[hi] = x;
    }
}

template <typename Ptr>
__device__ void bitonic_sort_block(Ptr a, uint32_t n) {
    uint32_t P = 1, lp = 0;
    while (P < n) { P <<= 1; lp++; }
    const uint32_t half = P >> 1;
    for (uint32_t lk = 1; lk <= lp; lk++) {
        const uint32_t k = 1u << lk, hk = k >> 1;
        for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
            const uint32_t blk = i >> (lk - 1), w = i & (hk - 1);
            const uint32_t lo = (blk << lk) + w, hi = (blk << lk) + k - 1 - w;
            if (hi < n) cmpswap(a, lo, hi);
        }
        __syncthreads();
        for (uint32_t d = hk >> 1; d > 0; d >>= 1) {
            for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
                const uint32_t lo = 2 * d * (i / d) + (i % d), hi = lo + d;
                if (hi < n) cmpswap(a, lo, hi);
            }
            __syncthreads();
        }
    }
}

constexpr int SORT_THREADS = 256;
constexpr uint32_t SORT_SMEM_KEYS = 4096;   // 32 KB of keys per CTA; larger tiles sort in L2/global

// Shared-memory variant with ~4x fewer CTA barriers: every step whose comparators stay inside an
// aligned 64-key block (all of k <= 64, and the stride <= 32 tail of every later merge) is done by
// ONE warp per block with __syncwarp only; __syncthreads is needed just around the long strides.
__device__ __forceinline__ void cs_smem(uint64_t* a, uint32_t lo, uint32_t hi, uint32_t n) {
    if (hi < n) {
        const uint64_t x = a[lo], y = a[hi];
        if (x > y) {
            a[lo] = y;
            a[hi] = x;
        }
    }
}
__device__ __forceinline__ void warp_half_cleaners(uint64_t* a, uint32_t base, uint32_t n, uint32_t lane, int ld_from) {
    for (int ld = ld_from; ld >= 0; ld--) {          // strides 2^ld ... 1 inside one 64-key block
        const uint32_t lo = base + (((lane >> ld) << (ld + 1)) | (lane & ((1u << ld) - 1u)));
        cs_smem(a, lo, lo + (1u << ld), n);
        __syncwarp();
    }
}
__device__ void bitonic_sort_smem(uint64_t* a, uint32_t n) {
    uint32_t P = 1, lp = 0;
    while (P < n) { P <<= 1; lp++; }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = SORT_THREADS / 32;
    // phase 1: merges of size 2..64, warp-local
    for (uint32_t base = warp * 64; base < n; base += nwarps * 64) {
        const uint32_t lk_max = lp < 6 ? lp : 6;
        for (uint32_t lk = 1; lk <= lk_max; lk++) {
            const uint32_t k = 1u << lk, hk = k >> 1;
            const uint32_t blk = lane >> (lk - 1), w = lane & (hk - 1);
            cs_smem(a, base + (blk << lk) + w, base + (blk << lk) + k - 1 - w, n);
            __syncwarp();
            if (lk >= 2) warp_half_cleaners(a, base, n, lane, (int)lk - 2);
        }
    }
    __syncthreads();
    // phase 2: merges of size 128..P: long strides CTA-wide, stride <= 32 tail warp-local
    const uint32_t half = P >> 1;
    for (uint32_t lk = 7; lk <= lp; lk++) {
        const uint32_t k = 1u << lk, hk = k >> 1;
        for (uint32_t i = threadIdx.x; i < half; i += SORT_THREADS) {
            const uint32_t blk = i >> (lk - 1), w = i & (hk - 1);
            cs_smem(a, (blk << lk) + w, (blk << lk) + k - 1 - w, n);
        }
        __syncthreads();
        for (uint32_t d = hk >> 1; d >= 64; d >>= 1) {
            for (uint32_t i = threadIdx.x; i < half; i += SORT_THREADS) {
                const uint32_t lo = 2 * d * (i / d) + (i % d);
                cs_smem(a, lo, lo + d, n);
            }
            __syncthreads();
        }
        for (uint32_t base = warp * 64; base < n; base += nwarps * 64) warp_half_cleaners(a, base, n, lane, 5);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_pack_kernel(const uint32_t* __restrict__ tile_offset, uint64_t* __restrict__ keys,
                 const float2* __restrict__ xy, const float4* __restrict__ conic_o, const float* __restrict__ rgb,
                 float4* __restrict__ p0, float4* __restrict__ p1, float4* __restrict__ p2, uint32_t capacity, int gx) {
    __shared__ uint64_t skeys[SORT_SMEM_KEYS];
    const uint32_t t = blockIdx.x;
    const uint32_t off = tile_offset[t];
    uint32_t n = tile_offset[t + 1] - off;
    if (n == 0) return;
    if (off + n > capacity) return;   // overflow is reported by the host wrapper (K > capacity)
    const uint64_t* sorted;
    if (n <= SORT_SMEM_KEYS) {
        for (uint32_t j = threadIdx.x; j < n; j += SORT_THREADS) skeys[j] = keys[off + j];
        __syncthreads();
        if (n > 1) bitonic_sort_smem(skeys, n);
        sorted = skeys;
    } else {
        bitonic_sort_block(keys + off, n);
        sorted = keys + off;
    }
    const float tile_x = (float)((t % (uint32_t)gx) * TILE), tile_y = (float)((t / (uint32_t)gx) * TILE);
    for (uint32_t j = threadIdx.x; j < n; j += SORT_THREADS) {
        const uint64_t key = sorted[j];
        const uint32_t id = (uint32_t)(key & 0xffffffffu);
        const float depth = __uint_as_float((uint32_t)(key >> 32));
        const float2 m = xy[id];
        const float4 co = conic_o[id];
        const float r = rgb[3 * (size_t)id], g = rgb[3 * (size_t)id + 1], b = rgb[3 * (size_t)id + 2];
        // warp-overlap mask from the alpha >= 1/255 ellipse's bounding box (covariance = conic^-1)
        uint32_t wmask = 0;
        {
            const float dc = co.x * co.z - co.y * co.y;
            float ex, ey;
            if (dc > 0.f && alpha_extent(co.w, co.z / dc, co.x / dc, ex, ey)) {
                const float lx0 = m.x - ex - tile_x, lx1 = m.x + ex - tile_x;     // bbox in tile-local pixels
                const float ly0 = m.y - ey - tile_y, ly1 = m.y + ey - tile_y;
                const uint32_t xb = ((lx0 <= 7.f && lx1 >= 0.f) ? 1u : 0u) | ((lx0 <= 15.f && lx1 >= 8.f) ? 2u : 0u);
                uint32_t yb = 0;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (ly0 <= (float)(4 * q + 3) && ly1 >= (float)(4 * q)) yb |= 1u << q;
#pragma unroll
                for (int w = 0; w < 8; w++)
                    if (((xb >> (w & 1)) & 1u) && ((yb >> (w >> 1)) & 1u)) wmask |= 1u << w;
            } else if (!(dc > 0.f)) {
                wmask = 0xffu;   // degenerate conic: no culling information
            }
        }
        p0[off + j] = make_float4(m.x, m.y, (-0.5f * LOG2E) * co.x, -LOG2E * co.y);
        p1[off + j] = make_float4((-0.5f * LOG2E) * co.z, co.w, depth, __uint_as_float(wmask));
        p2[off + j] = make_float4(r, g, b, __uint_as_float(id));
    }
}

int launch_emit(const gg_view& v, const GeomWS& g, const TileWS& t, const int32_t* radii, uint64_t* keys,
                uint32_t capacity, cudaStream_t s) {
    const int N = v.num_gaussians;
    if (N == 0) return 0;
    const int gx = (v.image_width + TILE - 1) / TILE;
    emit_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, gx, radii, g.rect, g.depth, t.offset, t.fill, keys, capacity);
    return 1;
}

int launch_sort_pack(const gg_view& v, const GeomWS& g, const TileWS& t, uint64_t* keys, const RecordWS& r,
                     uint32_t capacity, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0 || v.num_gaussians == 0) return 0;
    sort_pack_kernel<<<T, SORT_THREADS, 0, s>>>(t.offset, keys, g.xy, g.conic_o, g.rgb, r.p0, r.p1, r.p2, capacity, gx);
    return 1;
}

}  // namespace gg
