// blend_fwd2.cu -- per-tile front-to-back alpha blending, decoupled warps.
//
// Same mathematics and the same inputs / outputs as blend_fwd.cu (kept as GG_FWD_KERNEL=v1); what changes is how the
// CTA is organised around the record stream:
//   * the tile's depth-sorted packed records still arrive through a bulk-TMA (cp.async.bulk) ring, but a stage is
//     recycled through a per-stage "empty" mbarrier (8 warp arrivals) instead of __syncthreads_count: a warp whose
//     8x4 pixel block sees few splats (or saturates early) never waits for the busiest warp of the tile;
//   * the per-entry body is branch-free (predicated updates) -- no vote per entry; whether the whole warp has
//     saturated is tested once per 8 entries;
//   * every shared access is an explicit 32-bit address off one pinned base register.
// A warp that is done keeps retiring stages (one wait + one arrive per batch) until all eight are done; the producer
// then stops refilling and drains what is still in flight.
#include "common.cuh"

namespace gg {

constexpr int F2_BATCH = 64;
constexpr int F2_STAGES = 4;
constexpr int F2_WARPS = TILE_PIX / 32;

struct F2Smem {
    float4 s0[F2_STAGES][F2_BATCH];
    float4 s1[F2_STAGES][F2_BATCH];
    float4 s2[F2_STAGES][F2_BATCH];
    uint64_t full[F2_STAGES];
    uint64_t empty[F2_STAGES];
    uint32_t done_warps;
    uint32_t issued;            // batches handed to the TMA so far (producer -> everyone, for the final drain)
};
constexpr uint32_t F_S0 = offsetof(F2Smem, s0), F_S1 = offsetof(F2Smem, s1), F_S2 = offsetof(F2Smem, s2);
constexpr uint32_t F_FULL = offsetof(F2Smem, full), F_EMPTY = offsetof(F2Smem, empty);
constexpr uint32_t F_DONE = offsetof(F2Smem, done_warps), F_ISSUED = offsetof(F2Smem, issued);
constexpr uint32_t F_STAGE_BYTES = F2_BATCH * 16;

__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(TILE_PIX, 5)
blend_fwd2_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ p0,
                  const float4* __restrict__ p1, const float4* __restrict__ p2, uint32_t capacity, int W, int H, int gx,
                  const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                  float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                  const uint32_t* __restrict__ order) {
    __shared__ __align__(128) F2Smem S;
    uint32_t sb = smem_u32(&S);
    asm volatile("" : "+r"(sb));          // pin: one register, never rematerialised

    const uint32_t tile = order[blockIdx.x];              // launch order: heaviest tiles first (tile_scan_kernel)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = (tile % gx) * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = (tile / gx) * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;

    const uint32_t off = tile_offset[tile];
    uint32_t n = tile_offset[tile + 1] - off;
    if (off + n > capacity) n = 0;
    const int nb = (n + F2_BATCH - 1) / F2_BATCH;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < F2_STAGES; s++) {
            mbar_init_a(sb + F_FULL + 8u * s, 1);
            mbar_init_a(sb + F_EMPTY + 8u * s, F2_WARPS);
        }
        sts32(sb + F_DONE, 0u);
        sts32(sb + F_ISSUED, 0u);
        fence_barrier_init();
    }
    __syncthreads();

    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, Ac = 0.f;
    // Tm = the transmittance while the pixel is live, 0 once it has stopped (or lies outside the image): a stopped pixel
    // then fails the T_STOP test by itself (0 * (1 - alpha) < T_STOP) -- no separate flag to carry through the loop
    float Tm = inside ? 1.f : 0.f;
    uint32_t last = 0;
    bool warp_done = false;               // all 32 pixels of this warp are saturated (or outside the image)
    int issued = 0;                       // thread 0 only
    int q = 0;
#pragma unroll 1
    for (; q < nb; q++) {
        const int st = q % F2_STAGES;
        if (threadIdx.x == 0) {
            // producer: refill the stages all 8 warps have released, as long as some warp still blends; block only
            // when the batch this warp needs right now has not been issued yet
            while (issued < nb && issued < q + F2_STAGES && lds_volatile_u32(sb + F_DONE) < F2_WARPS) {
                if (issued >= F2_STAGES) {
                    const int prev = issued - F2_STAGES;
                    const uint32_t eb = sb + F_EMPTY + 8u * (uint32_t)(prev % F2_STAGES);
                    const uint32_t par = (uint32_t)(prev / F2_STAGES) & 1u;
                    if (issued == q) {
                        bool all_done = false;
                        while (!mbar_test_a(eb, par)) {
                            if (lds_volatile_u32(sb + F_DONE) >= F2_WARPS) { all_done = true; break; }
                        }
                        if (all_done) break;
                    } else if (!mbar_test_a(eb, par)) {
                        break;
                    }
                }
                const int ist = issued % F2_STAGES;
                const uint32_t icnt = min((uint32_t)F2_BATCH, n - (uint32_t)issued * F2_BATCH);
                const uint32_t bytes = icnt * 16u, fb = sb + F_FULL + 8u * ist;
                const size_t src = (size_t)off + (size_t)issued * F2_BATCH;
                mbar_expect_tx_a(fb, 3u * bytes);
                bulk_g2s_a(sb + F_S0 + ist * F_STAGE_BYTES, p0 + src, bytes, fb);
                bulk_g2s_a(sb + F_S1 + ist * F_STAGE_BYTES, p1 + src, bytes, fb);
                bulk_g2s_a(sb + F_S2 + ist * F_STAGE_BYTES, p2 + src, bytes, fb);
                issued++;
                sts32(sb + F_ISSUED, (uint32_t)issued);
            }
        }
        // wait for batch q -- or for the news that every warp has saturated (then nobody will issue it)
        {
            const uint32_t fb = sb + F_FULL + 8u * st, par = (uint32_t)(q / F2_STAGES) & 1u;
            bool all_done = false;
            while (!mbar_test_a(fb, par)) {
                if (lds_volatile_u32(sb + F_DONE) >= F2_WARPS) { all_done = true; break; }
            }
            if (__any_sync(0xffffffffu, all_done)) break;
        }
        if (!warp_done) {
            const int cnt = min(F2_BATCH, (int)n - q * F2_BATCH);
            const uint32_t r0 = sb + F_S0 + st * F_STAGE_BYTES;
            // work list of this warp for the batch: bit `warp` of every record's warp-overlap mask (two ballots);
            // records whose alpha >= 1/255 footprint misses this warp's 8x4 block cost nothing
            const uint32_t w0 = (lane < cnt) ? ldsu32(r0 + (F_S1 - F_S0) + 16u * lane + 12u) : 0u;
            const uint32_t w1 = (lane + 32 < cnt) ? ldsu32(r0 + (F_S1 - F_S0) + 16u * (lane + 32) + 12u) : 0u;
            const uint32_t m_lo = __ballot_sync(0xffffffffu, (w0 >> warp) & 1u);
            const uint32_t m_hi = __ballot_sync(0xffffffffu, (w1 >> warp) & 1u);
            const uint32_t idx0 = (uint32_t)(q * F2_BATCH) + 1u;
#pragma unroll 1
            for (int half = 0; half < 2 && !warp_done; half++) {
                uint32_t m = half ? m_hi : m_lo;
                const uint32_t rh = r0 + 512u * (uint32_t)half;
                const uint32_t idxh = idx0 + 32u * (uint32_t)half;
                int since_check = 0;
#pragma unroll 1
                while (m) {
                    const uint32_t bit = (uint32_t)__ffs(m) - 1u;                      // lowest set bit: front to back
                    m &= m - 1u;
                    const uint32_t ra = rh + 16u * bit;
                    const float4 a = lds128(ra);
                    const float4 c = lds128(ra + (F_S1 - F_S0));
                    const float4 col = lds128(ra + (F_S2 - F_S0));
                    const float dx = a.x - fx, dy = a.y - fy;
                    const float e2 = dx * (a.z * dx + a.w * dy) + (c.x * dy) * dy;     // log2-domain exponent
                    const float alpha = fminf(ALPHA_MAX, c.y * ex2_approx(e2));
                    const bool valid = e2 <= 0.f && alpha >= ALPHA_MIN;
                    const float test_T = Tm * (1.f - alpha);
                    const bool ok = valid && test_T >= T_STOP;      // !ok on a valid splat = the stop (NOT applied)
                    const float w = ok ? alpha * Tm : 0.f;
                    C0 = fmaf(col.x, w, C0);
                    C1 = fmaf(col.y, w, C1);
                    C2 = fmaf(col.z, w, C2);
                    Dp = fmaf(c.z, w, Dp);
                    Ac += w;
                    T = ok ? test_T : T;
                    Tm = valid ? (ok ? test_T : 0.f) : Tm;
                    last = ok ? idxh + bit : last;
                    if (++since_check == 8) {                                          // warp-wide saturation test
                        since_check = 0;
                        if (__all_sync(0xffffffffu, Tm == 0.f)) { warp_done = true; break; }
                    }
                }
            }
            if (!warp_done && __all_sync(0xffffffffu, Tm == 0.f)) warp_done = true;
            if (warp_done && lane == 0) atomicAdd(&S.done_warps, 1u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sb + F_EMPTY + 8u * st);          // this warp is done reading stage st
    }
    // a warp that ran out of records without saturating still has to count as finished for the others' exit test
    if (!warp_done && lane == 0) atomicAdd(&S.done_warps, 1u);

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
    // drain: bulk copies that were issued but never consumed must land before the CTA (and its shared memory) retires
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tot = (int)lds_volatile_u32(sb + F_ISSUED);
        // stage st's barrier completed phase k when batch k*STAGES+st landed; wait for the LAST batch issued per stage
        for (int bq = max(0, tot - F2_STAGES); bq < tot; bq++)
            mbar_wait_a(sb + F_FULL + 8u * (uint32_t)(bq % F2_STAGES), (uint32_t)(bq / F2_STAGES) & 1u);
    }
}

int launch_blend_fwd2(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                      uint32_t capacity, float* out_color, float* out_depth, float* out_alpha, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0) return 0;
    blend_fwd2_kernel<<<T, TILE_PIX, 0, s>>>(t.offset, r.p0, r.p1, r.p2, capacity, v.image_width, v.image_height, gx,
                                             in.bg, out_color, out_depth, out_alpha, img.n_contrib, img.final_T, t.order);
    return 1;
}

}  // namespace gg
