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
    constexpr int COOP_TILES = 16;
    const int nt_all = (x1 - x0) * (y1 - y0);
    const bool big = nt_all > COOP_TILES;
    const int nt = big ? 0 : nt_all;
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
    // large footprints: the whole warp emits one splat's instances, 32 tiles per step
    unsigned bigmask = __ballot_sync(FULL, big);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const int bx0 = __shfl_sync(FULL, x0, src), by0 = __shfl_sync(FULL, y0, src);
        const int bx1 = __shfl_sync(FULL, x1, src), by1 = __shfl_sync(FULL, y1, src);
        const uint32_t klo = __shfl_sync(FULL, (uint32_t)key, src), khi = __shfl_sync(FULL, (uint32_t)(key >> 32), src);
        const uint64_t bkey = ((uint64_t)khi << 32) | klo;
        const int w = bx1 - bx0, total = w * (by1 - by0);
        for (int k = lane; k < total; k += 32) {
            const int t = (by0 + k / w) * gx + bx0 + k % w;
            const uint32_t slot = tile_offset[t] + atomicAdd(&tile_fill[t], 1u);
            if (slot < capacity) keys[slot] = bkey;
        }
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
constexpr uint32_t SORT_SMEM_KEYS = 4096;   // default: 32 KB of keys per CTA (7 CTAs/SM)
constexpr uint32_t SORT_SMEM_KEYS_MAX = 24576;   // 192 KB dynamic variant for dense scenes; beyond: L2/global

// Shared-memory variant with ~4x fewer CTA barriers: every step whose comparators stay inside an aligned 64-key block
// (all of k <= 64, and the stride <= 32 tail of every later merge) is done by ONE warp per block IN REGISTERS: lane l
// holds keys l and l + 32 of the block, stride 32 is an in-thread exchange, shorter strides are shfl_xor exchanges.
// (The first version ran those steps through shared memory with __syncwarp: a warp's 64-bit accesses at stride < 16 keys
//  span 512 B for 256 B of data -- ncu: 4.5 M bank conflicts on 13.5 M shared wavefronts, LSU data pipe at 67 %.)
// __syncthreads is needed just around the long strides, which are conflict-free (consecutive lanes, consecutive keys).
__device__ __forceinline__ void cs_smem(uint64_t* a, uint32_t lo, uint32_t hi, uint32_t n) {
    if (hi < n) {
        const uint64_t x = a[lo], y = a[hi];
        if (x > y) {
            a[lo] = y;
            a[hi] = x;
        }
    }
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t keep(uint64_t mine, uint64_t other, bool want_min) {
    return (want_min == (other < mine)) ? other : mine;      // ties: either copy is the same key
}
// strides 2^ld_from ... 1 inside one 64-key block held as (r0 = key[lane], r1 = key[lane + 32])
__device__ __forceinline__ void reg_half_cleaners(uint64_t& r0, uint64_t& r1, uint32_t lane, int ld_from) {
    if (ld_from >= 5) {
        const uint64_t lo = r0 < r1 ? r0 : r1, hi = r0 < r1 ? r1 : r0;
        r0 = lo;
        r1 = hi;
        ld_from = 4;
    }
#pragma unroll
    for (int ld = 4; ld >= 0; ld--) {
        if (ld <= ld_from) {
            const bool lower = (lane & (1u << ld)) == 0u;
            r0 = keep(r0, shfl_xor_u64(r0, 1 << ld), lower);
            r1 = keep(r1, shfl_xor_u64(r1, 1 << ld), lower);
        }
    }
}
constexpr uint64_t KEY_INF = ~0ull;
// `src` (optional): the unsorted keys are still in global memory -- the first six merge levels read them straight into
// registers and only the 64-key sorted runs are written to shared memory.
__device__ void bitonic_sort_smem(uint64_t* a, uint32_t n, const uint64_t* __restrict__ src = nullptr) {
    uint32_t P = 1, lp = 0;
    while (P < n) { P <<= 1; lp++; }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = SORT_THREADS / 32;
    // phase 1: merges of size 2..64, warp-local, in registers (keys beyond n are +inf and never move down)
    for (uint32_t base = warp * 64; base < n; base += nwarps * 64) {
        const uint32_t i0 = base + lane, i1 = base + lane + 32;
        uint64_t r0 = i0 < n ? (src ? src[i0] : a[i0]) : KEY_INF;
        uint64_t r1 = i1 < n ? (src ? src[i1] : a[i1]) : KEY_INF;
#pragma unroll
        for (int lk = 1; lk <= 5; lk++) {
            const bool lower = (lane & (1u << (lk - 1))) == 0u;          // flip step: partner = e ^ (2^lk - 1)
            r0 = keep(r0, shfl_xor_u64(r0, (1 << lk) - 1), lower);
            r1 = keep(r1, shfl_xor_u64(r1, (1 << lk) - 1), lower);
            if (lk >= 2) reg_half_cleaners(r0, r1, lane, lk - 2);
        }
        {   // lk = 6: partner of (lane, r) is (lane ^ 31, r ^ 1); the r = 0 copy is the lower one
            const uint64_t q0 = shfl_xor_u64(r1, 31), q1 = shfl_xor_u64(r0, 31);
            r0 = keep(r0, q0, true);
            r1 = keep(r1, q1, false);
            reg_half_cleaners(r0, r1, lane, 4);
        }
        if (i0 < n) a[i0] = r0;
        if (i1 < n) a[i1] = r1;
    }
    __syncthreads();
    // phase 2: merges of size 128..P: long strides CTA-wide in shared memory, stride <= 32 tail warp-local in registers
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
        for (uint32_t base = warp * 64; base < n; base += nwarps * 64) {
            const uint32_t i0 = base + lane, i1 = base + lane + 32;
            uint64_t r0 = i0 < n ? a[i0] : KEY_INF, r1 = i1 < n ? a[i1] : KEY_INF;
            reg_half_cleaners(r0, r1, lane, 5);
            if (i0 < n) a[i0] = r0;
            if (i1 < n) a[i1] = r1;
        }
        __syncthreads();
    }
}

// (A register-shuffle + rank-merge sort -- 32-key runs sorted with warp shuffles, then log2(n/32) levels in which every
//  key binary-searches its rank in the partner run -- was measured against this network on cfg2, where the 1 900
//  non-empty tiles hold ~700 keys each: 0.159 ms vs 0.113 ms.  The searches are dependent shared-memory probes, and the
//  ping-pong buffer halves the CTAs per SM; the network's independent compare-exchanges hide latency better.)
__global__ void __launch_bounds__(SORT_THREADS)
sort_pack_kernel(const uint32_t* __restrict__ tile_offset, uint64_t* __restrict__ keys,
                 const float2* __restrict__ xy, const float4* __restrict__ conic_o, const float4* __restrict__ ext,
                 const float* __restrict__ rgb, float4* __restrict__ p0, float4* __restrict__ p1, float4* __restrict__ p2,
                 uint32_t capacity, int gx, uint32_t smem_keys, const uint32_t* __restrict__ order) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t t = order[blockIdx.x];                      // heaviest tiles first
    const uint32_t off = tile_offset[t];
    uint32_t n = tile_offset[t + 1] - off;
    if (n == 0) return;
    if (off + n > capacity) return;   // overflow is reported by the host wrapper (K > capacity)
    const uint64_t* sorted;
    if (n <= smem_keys) {
        bitonic_sort_smem(skeys, n, keys + off);      // n == 1 included: phase 1 moves the key into shared memory
        sorted = skeys;
    } else {
        bitonic_sort_block(keys + off, n);
        sorted = keys + off;
    }
    const float tile_x = (float)((t % (uint32_t)gx) * TILE), tile_y = (float)((t / (uint32_t)gx) * TILE);
    for (uint32_t j = threadIdx.x; j < n; j += SORT_THREADS) {
        float4 q0, q1, q2;
        pack_record(sorted[j], xy, conic_o, ext, rgb, tile_x, tile_y, q0, q1, q2);
        p0[off + j] = q0;
        p1[off + j] = q1;
        p2[off + j] = q2;
    }
}

// sticky overflow flag for sync-free (CUDA-graph) operation: *flag |= 1 when this view's instance count exceeds the
// capacity its workspaces were laid out for (the instance stages then skipped the overflowing tiles)
__global__ void overflow_flag_kernel(const uint32_t* __restrict__ misc, uint32_t capacity, uint32_t* __restrict__ flag) {
    if (misc[0] > capacity) atomicOr(flag, 1u);
    atomicMax(flag + 1, misc[0]);            // largest K seen (lets the host re-size before re-capturing)
}
int launch_overflow_flag(const TileWS& t, uint32_t capacity, uint32_t* flag, cudaStream_t s) {
    overflow_flag_kernel<<<1, 1, 0, s>>>(t.misc, capacity, flag);
    return 1;
}

int launch_emit(const gg_view& v, const GeomWS& g, const TileWS& t, const int32_t* radii, uint64_t* keys,
                uint32_t capacity, cudaStream_t s) {
    const int N = v.num_gaussians;
    if (N == 0) return 0;
    const int gx = (v.image_width + TILE - 1) / TILE;
    emit_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, gx, radii, g.rect, g.depth, t.offset, t.fill, keys, capacity);
    return 1;
}

// late colours: p2[k].rgb = rgb[id(p2[k])] for every packed instance (ids were parked in p2.w by sort_pack)
__global__ void __launch_bounds__(256)
color_fill_kernel(const uint32_t* __restrict__ total, uint32_t capacity, uint32_t N, const float* __restrict__ rgb,
                  float4* __restrict__ p2) {
    const uint32_t n = min(*total, capacity);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t id = __float_as_uint(p2[k].w);
        if (id >= N) continue;     // a tile that straddles an undersized capacity was never packed (overflow retry follows)
        p2[k] = make_float4(rgb[3 * (size_t)id], rgb[3 * (size_t)id + 1], rgb[3 * (size_t)id + 2], __uint_as_float(id));
    }
}
int launch_color_fill(const gg_view& v, const GeomWS& g, const TileWS& t, const RecordWS& r, uint32_t capacity, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    if (gx * gy == 0 || v.num_gaussians == 0 || capacity == 0) return 0;
    const uint32_t blocks = min((capacity + 255u) / 256u, 148u * 16u);
    color_fill_kernel<<<blocks, 256, 0, s>>>(t.offset + gx * gy, capacity, (uint32_t)v.num_gaussians, g.rgb, r.p2);
    return 1;
}

int launch_sort_pack(const gg_view& v, const GeomWS& g, const TileWS& t, uint64_t* keys, const RecordWS& r,
                     uint32_t capacity, uint32_t max_tile_instances, bool with_color, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0 || v.num_gaussians == 0) return 0;
    // shared-memory budget from the largest tile (0 = unknown -> largest variant): 32 KB keeps 7 CTAs/SM for
    // ordinary scenes; dense scenes trade occupancy for an in-smem sort of up to 24576 keys per tile
    uint32_t smem_keys = SORT_SMEM_KEYS;
    if (max_tile_instances == 0 || max_tile_instances > SORT_SMEM_KEYS) {
        smem_keys = 8192;
        while (smem_keys < SORT_SMEM_KEYS_MAX && smem_keys < max_tile_instances) smem_keys += 8192;
        if (max_tile_instances == 0) smem_keys = SORT_SMEM_KEYS_MAX;
    }
    const size_t bytes = (size_t)smem_keys * sizeof(uint64_t);
    if (bytes > 48 * 1024) {   // opt in to large dynamic shared memory; a failure surfaces as a launch error right below
        static size_t opted[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || opted[dev] < bytes) {
            if (cudaFuncSetAttribute(sort_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM_KEYS_MAX * 8) == cudaSuccess &&
                dev >= 0 && dev < 64)
                opted[dev] = (size_t)SORT_SMEM_KEYS_MAX * 8;
        }
    }
    sort_pack_kernel<<<T, SORT_THREADS, bytes, s>>>(t.offset, keys, g.xy, g.conic_o, g.ext, with_color ? g.rgb : nullptr, r.p0, r.p1,
                                                     r.p2, capacity, gx, smem_keys, t.order);
    return 1;
}

// =============================================================================================
// Lazy fused forward (dense scenes): per tile, bucket the keys by depth in O(n), then sort + pack +
// blend ONE BUCKET AT A TIME and stop as soon as all 256 pixels are saturated.  In the 2M-Gaussian
// 2160p stress config a tile's list holds ~7000 entries of which the blend consumes a few hundred:
// sorting everything (sort_pack_kernel) cost 152 ms/view there, 95 % of the forward+backward time.
// The order produced is identical to the full sort: buckets are a monotone function of depth and each
// bucket is ordered by (depth bits, Gaussian index).  Only the consumed prefix of the packed records
// is written; backward never reads past the tile's last contributor.
// =============================================================================================
constexpr uint32_t LZ_CAP = 2048;     // keys sorted in shared memory per run (16 KB)
constexpr int LZ_CHUNK = 256;         // records staged per blend round (one gather per thread)
constexpr uint32_t LZ_MAXB = 1024;    // depth buckets per tile at most

__global__ void __launch_bounds__(TILE_PIX)
blend_fwd_lazy_kernel(const uint32_t* __restrict__ tile_offset, const uint64_t* __restrict__ keys,
                      uint64_t* __restrict__ keys2, const float2* __restrict__ xy,
                      const float4* __restrict__ conic_o, const float4* __restrict__ ext, const float* __restrict__ rgb,
                      float4* __restrict__ p0,
                      float4* __restrict__ p1, float4* __restrict__ p2, uint32_t capacity, int W, int H, int gx,
                      const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                      float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                      unsigned long long* __restrict__ phase /* diagnostics or NULL: [0] ordering cycles, [1] pack+blend cycles */) {
    __shared__ __align__(16) uint64_t skeys[LZ_CAP];
    __shared__ __align__(16) float4 c0[LZ_CHUNK];
    __shared__ __align__(16) float4 c1[LZ_CHUNK];
    __shared__ __align__(16) float4 c2[LZ_CHUNK];
    __shared__ uint32_t boff[LZ_MAXB + 1];
    __shared__ uint32_t cursor[LZ_MAXB];
    __shared__ uint32_t red_min[8], red_max[8];

    const uint32_t tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const float tile_x = (float)(tx * TILE), tile_y = (float)(ty * TILE);

    const uint32_t off = tile_offset[tile];
    uint32_t n = tile_offset[tile + 1] - off;
    if (off + n > capacity) n = 0;

    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, Ac = 0.f;
    uint32_t last = 0, processed = 0;
    bool done = !inside;
    const uint32_t a0 = smem_u32(&c0[0]), a1 = smem_u32(&c1[0]), a2 = smem_u32(&c2[0]);

    // consume a run of keys that is already in final order; returns true when the whole tile is saturated
    long long t_blend = 0;                                  // thread 0: cycles inside process_run (pack + blend)
    const long long t_start = phase ? clock64() : 0;
    auto process_run = [&](const uint64_t* run, uint32_t cnt) -> bool {
        const long long t_in = phase ? clock64() : 0;
        struct Tick { long long& acc; long long t0; bool on; __device__ ~Tick() { if (on) acc += clock64() - t0; } } tick{t_blend, t_in, phase != nullptr};
        for (uint32_t c = 0; c < cnt; c += LZ_CHUNK) {
            const uint32_t m = min((uint32_t)LZ_CHUNK, cnt - c);
            if (threadIdx.x < m) {
                float4 q0, q1, q2;
                pack_record(run[c + threadIdx.x], xy, conic_o, ext, rgb, tile_x, tile_y, q0, q1, q2);
                c0[threadIdx.x] = q0; c1[threadIdx.x] = q1; c2[threadIdx.x] = q2;
                const size_t dst = (size_t)off + processed + threadIdx.x;
                p0[dst] = q0; p1[dst] = q1; p2[dst] = q2;
            }
            __syncthreads();
            uint32_t mw[LZ_CHUNK / 32];                      // this warp's work list for the chunk (see blend_fwd.cu)
#pragma unroll
            for (int k = 0; k < LZ_CHUNK / 32; k++) {
                const uint32_t e = (uint32_t)(k * 32 + lane);
                const uint32_t wbits = (e < m) ? __float_as_uint(lds32(a1 + 16u * e + 12u)) : 0u;
                mw[k] = __ballot_sync(0xffffffffu, (wbits >> warp) & 1u);
            }
            bool warp_done = false;
#pragma unroll
            for (int k = 0; k < LZ_CHUNK / 32; k++) {
                uint32_t bits = mw[k];
                while (bits && !warp_done) {
                    if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
                    const uint32_t j = (uint32_t)(k * 32 + __ffs(bits) - 1);
                    bits &= bits - 1;
                    const float4 cc = lds128(a1 + 16u * j);
                    const float4 a = lds128(a0 + 16u * j);
                    const float dx = a.x - fx, dy = a.y - fy;
                    const float e2 = dx * (a.z * dx + a.w * dy) + (cc.x * dy) * dy;
                    const float alpha = fminf(ALPHA_MAX, cc.y * ex2_approx(e2));
                    bool ok = !done && e2 <= 0.f && alpha >= ALPHA_MIN;
                    const float test_T = T * (1.f - alpha);
                    if (ok && test_T < T_STOP) {
                        done = true;
                        ok = false;
                    }
                    if (ok) {
                        const float w = alpha * T;
                        const float4 col = lds128(a2 + 16u * j);
                        C0 += col.x * w; C1 += col.y * w; C2 += col.z * w;
                        Dp += cc.z * w;
                        Ac += w;
                        T = test_T;
                        last = processed + j + 1;
                    }
                }
            }
            processed += m;
            if (__syncthreads_count(done) == TILE_PIX) return true;
        }
        return false;
    };

    if (n > 0 && n <= LZ_CAP) {
        for (uint32_t j = threadIdx.x; j < n; j += TILE_PIX) skeys[j] = keys[off + j];
        __syncthreads();
        if (n > 1) bitonic_sort_smem(skeys, n);
        process_run(skeys, n);
    } else if (n > LZ_CAP) {
        // ---- pass A: depth range of the tile
        uint32_t dmin = 0xffffffffu, dmax = 0u;
        for (uint32_t j = threadIdx.x; j < n; j += TILE_PIX) {
            const uint32_t d = (uint32_t)(keys[off + j] >> 32);
            dmin = min(dmin, d);
            dmax = max(dmax, d);
        }
        dmin = __reduce_min_sync(0xffffffffu, dmin);
        dmax = __reduce_max_sync(0xffffffffu, dmax);
        if (lane == 0) { red_min[warp] = dmin; red_max[warp] = dmax; }
        const uint32_t NB = min(LZ_MAXB, (n + LZ_CAP / 2 - 1) / (LZ_CAP / 2));
        for (uint32_t k = threadIdx.x; k <= NB; k += TILE_PIX) boff[k] = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; w++) { dmin = min(dmin, red_min[w]); dmax = max(dmax, red_max[w]); }
        const float fmin_ = __uint_as_float(dmin), fmax_ = __uint_as_float(dmax);   // depths are positive floats
        const float scale = (fmax_ > fmin_) ? (float)NB / (fmax_ - fmin_) : 0.f;
        auto bucket_of = [&](uint64_t key) -> uint32_t {
            const float d = __uint_as_float((uint32_t)(key >> 32));
            return min(NB - 1u, (uint32_t)((d - fmin_) * scale));               // monotone in depth
        };
        // ---- pass B: histogram (shifted by one so the scan yields start offsets)
        for (uint32_t j = threadIdx.x; j < n; j += TILE_PIX) atomicAdd(&boff[bucket_of(keys[off + j]) + 1], 1u);
        __syncthreads();
        if (warp == 0) {                                                        // inclusive scan of <= 1025 counters
            uint32_t carry = 0;
            for (uint32_t base = 0; base <= NB; base += 32) {
                const uint32_t k = base + lane;
                uint32_t v = (k <= NB) ? boff[k] : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
                    if (lane >= d) v += t;
                }
                v += carry;
                if (k <= NB) boff[k] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < NB; k += TILE_PIX) cursor[k] = boff[k];
        __syncthreads();
        // ---- pass C: scatter into bucket order
        for (uint32_t j = threadIdx.x; j < n; j += TILE_PIX) {
            const uint64_t key = keys[off + j];
            keys2[off + atomicAdd(&cursor[bucket_of(key)], 1u)] = key;
        }
        __syncthreads();
        // ---- buckets front to back
        for (uint32_t bk = 0; bk < NB; bk++) {
            const uint32_t bo = boff[bk], cnt = boff[bk + 1] - bo;
            if (cnt == 0) continue;
            bool all_done;
            if (cnt <= LZ_CAP) {
                for (uint32_t j = threadIdx.x; j < cnt; j += TILE_PIX) skeys[j] = keys2[off + bo + j];
                __syncthreads();
                if (cnt > 1) bitonic_sort_smem(skeys, cnt);
                all_done = process_run(skeys, cnt);
            } else {                                     // degenerate depth distribution: sort this bucket in L2
                bitonic_sort_block(keys2 + off + bo, cnt);
                all_done = process_run(keys2 + off + bo, cnt);
            }
            if (all_done) break;
        }
    }

    if (phase && threadIdx.x == 0 && n > 0) {               // everything that was not pack + blend is ordering work
        const long long total = clock64() - t_start;
        atomicAdd(&phase[0], (unsigned long long)(total - t_blend));
        atomicAdd(&phase[1], (unsigned long long)t_blend);
    }
    if (inside) {
        const size_t P = (size_t)W * H, pid = (size_t)py * W + px;
        out_color[pid] = C0 + T * bg[0];
        out_color[P + pid] = C1 + T * bg[1];
        out_color[2 * P + pid] = C2 + T * bg[2];
        out_depth[pid] = Dp;
        out_alpha[pid] = Ac;
        n_contrib[pid] = last;
        final_T[pid] = T;
    }
}

// diagnostics: device pointer to two 64-bit counters the lazy fused forward adds its per-tile phase cycles to
static unsigned long long* g_lazy_phase = nullptr;
void set_lazy_phase_counters(unsigned long long* p) { g_lazy_phase = p; }

int launch_blend_fwd_lazy(const gg_view& v, const gg_inputs& in, const GeomWS& g, const TileWS& t, uint64_t* keys,
                          uint64_t* keys2, const RecordWS& r, const ImageWS& img, uint32_t capacity, float* out_color,
                          float* out_depth, float* out_alpha, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0) return 0;
    blend_fwd_lazy_kernel<<<T, TILE_PIX, 0, s>>>(t.offset, keys, keys2, g.xy, g.conic_o, g.ext, g.rgb, r.p0, r.p1, r.p2,
                                                 capacity, v.image_width, v.image_height, gx, in.bg, out_color,
                                                 out_depth, out_alpha, img.n_contrib, img.final_T, g_lazy_phase);
    return 1;
}

}  // namespace gg
