// blend_fwd.cu -- per-tile front-to-back alpha blending.
//
// One CTA (256 threads = 16x16 pixels) per tile.  The tile's depth-sorted packed records are
// three contiguous float4 runs, streamed into shared memory by 1-D bulk TMA (cp.async.bulk,
// SASS UBLKCP) through a 3-stage mbarrier ring: one elected thread arms a stage with
// expect_tx and issues three bulk copies; all threads wait on the stage's phase parity, read the
// records as shared-memory broadcasts, and a __syncthreads_count both recycles the stage and
// detects "every pixel saturated" for the early exit.
// A warp covers an 8x4 pixel block (not 16x2) so a small splat touches fewer warps.
#include "common.cuh"

namespace gg {

constexpr int FWD_BATCH = 128;
constexpr int FWD_STAGES = 3;

#ifndef GG_FWD_MINB
#define GG_FWD_MINB 5
#endif
__global__ void __launch_bounds__(TILE_PIX, GG_FWD_MINB)
blend_fwd_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ p0,
                 const float4* __restrict__ p1, const float4* __restrict__ p2, uint32_t capacity, int W, int H, int gx,
                 const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                 float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
    __shared__ __align__(128) float4 s0[FWD_STAGES][FWD_BATCH];
    __shared__ __align__(128) float4 s1[FWD_STAGES][FWD_BATCH];
    __shared__ __align__(128) float4 s2[FWD_STAGES][FWD_BATCH];
    __shared__ __align__(8) uint64_t full[FWD_STAGES];

    const uint32_t tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;

    const uint32_t off = tile_offset[tile];
    uint32_t n = tile_offset[tile + 1] - off;
    if (off + n > capacity) n = 0;
    const int nb = (n + FWD_BATCH - 1) / FWD_BATCH;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < FWD_STAGES; s++) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int b) {   // thread 0 only
        const int st = b % FWD_STAGES;
        const uint32_t cnt = min((uint32_t)FWD_BATCH, n - (uint32_t)b * FWD_BATCH);
        const uint32_t bytes = cnt * 16u;
        const size_t src = (size_t)off + (size_t)b * FWD_BATCH;
        mbar_arrive_expect_tx(&full[st], 3u * bytes);
        bulk_g2s(&s0[st][0], p0 + src, bytes, &full[st]);
        bulk_g2s(&s1[st][0], p1 + src, bytes, &full[st]);
        bulk_g2s(&s2[st][0], p2 + src, bytes, &full[st]);
    };
    if (threadIdx.x == 0)
        for (int b = 0; b < nb && b < FWD_STAGES; b++) issue(b);

    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f, Ac = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    int b = 0;
    for (; b < nb; b++) {
        const int st = b % FWD_STAGES;
        mbar_wait(&full[st], (uint32_t)(b / FWD_STAGES) & 1u);
        // Convergent, predicated inner loop.  (A first version used divergent continue/break here;
        // ncu showed 2.5 active threads per warp: independent thread scheduling let lanes run ahead
        // into later iterations and never reconverge.  The warp vote below re-converges the warp
        // every iteration and doubles as the per-warp early-out.)
        {
            const int cnt = min(FWD_BATCH, (int)n - b * FWD_BATCH);
            const uint32_t a0 = smem_u32(&s0[st][0]), a1 = smem_u32(&s1[st][0]), a2 = smem_u32(&s2[st][0]);
            // Work list of this warp for the batch: bit `warp` of every record's warp-overlap mask, gathered with
            // four ballots; records whose alpha >= 1/255 bounding box misses this warp's 8x4 block cost nothing.
            uint32_t mw[FWD_BATCH / 32];
#pragma unroll
            for (int k = 0; k < FWD_BATCH / 32; k++) {
                const int e = k * 32 + lane;
                const uint32_t wbits = (e < cnt) ? __float_as_uint(lds32(a1 + 16u * e + 12u)) : 0u;
                mw[k] = __ballot_sync(0xffffffffu, (wbits >> warp) & 1u);
            }
            bool warp_done = false;
#pragma unroll
            for (int k = 0; k < FWD_BATCH / 32; k++) {
                uint32_t m = mw[k];
                while (m && !warp_done) {
                    if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }   // also re-converges the warp
                    const int j = k * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const float4 c = lds128(a1 + 16u * j);
                    const float4 a = lds128(a0 + 16u * j);
                    const float dx = a.x - fx, dy = a.y - fy;
                    // log2-domain exponent: A' dx^2 + B' dx dy + C' dy^2  (= power * log2 e)
                    const float p2 = dx * (a.z * dx + a.w * dy) + (c.x * dy) * dy;
                    const float alpha = fminf(ALPHA_MAX, c.y * ex2_approx(p2));
                    bool ok = !done && p2 <= 0.f && alpha >= ALPHA_MIN;
                    const float test_T = T * (1.f - alpha);
                    if (ok && test_T < T_STOP) {
                        done = true;
                        ok = false;
                    }
                    if (ok) {
                        const float w = alpha * T;
                        const float4 col = lds128(a2 + 16u * j);
                        C0 += col.x * w;
                        C1 += col.y * w;
                        C2 += col.z * w;
                        Dp += c.z * w;
                        Ac += w;
                        T = test_T;
                        last = (uint32_t)(b * FWD_BATCH + j + 1);
                    }
                }
            }
        }
        const int n_done = __syncthreads_count(done);   // also: everyone has finished reading stage st
        if (n_done == TILE_PIX) break;
        if (threadIdx.x == 0 && b + FWD_STAGES < nb) issue(b + FWD_STAGES);
    }
    // drain bulk copies that were issued but never consumed (early exit) before the CTA retires
    if (threadIdx.x == 0 && b < nb) {
        for (int bb = b + 1; bb < nb && bb < b + FWD_STAGES; bb++)
            mbar_wait(&full[bb % FWD_STAGES], (uint32_t)(bb / FWD_STAGES) & 1u);
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

int launch_blend_fwd(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                     uint32_t capacity, float* out_color, float* out_depth, float* out_alpha, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0) return 0;
    blend_fwd_kernel<<<T, TILE_PIX, 0, s>>>(t.offset, r.p0, r.p1, r.p2, capacity, v.image_width, v.image_height, gx,
                                            in.bg, out_color, out_depth, out_alpha, img.n_contrib, img.final_T);
    return 1;
}

}  // namespace gg
