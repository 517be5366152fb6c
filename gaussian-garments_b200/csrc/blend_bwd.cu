// blend_bwd.cu -- per-tile back-to-front gradient pass.
//
// One CTA per tile, one pixel per thread.  The tile's packed records are streamed BACKWARDS
// (from the last contributor of any pixel in the tile) through a 2-stage bulk-TMA/mbarrier ring.
// Per Gaussian every thread forms its 10 partial derivatives; a warp that has no contributing
// pixel skips the Gaussian (one ballot), otherwise the warp reduces the 10 values with shuffles
// and lane 0 parks them in a per-warp shared-memory slot.  After a batch the CTA folds the 8 warp
// slots and issues at most three vector reductions (float4, float4, float2 red.global) per
// (tile, Gaussian) instance -- instead of upstream's ~10 scalar atomics per (pixel, Gaussian).
#include "common.cuh"

namespace gg {

constexpr int BWD_BATCH = 64;
constexpr int BWD_STAGES = 2;
constexpr int BWD_WARPS = TILE_PIX / 32;
constexpr int ACC_STRIDE = 12;   // 10 used; keeps float4 alignment

__global__ void __launch_bounds__(TILE_PIX)
blend_bwd_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ p0,
                 const float4* __restrict__ p1, const float4* __restrict__ p2, int W, int H, int gx,
                 const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                 const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha, float4* __restrict__ a0,
                 float4* __restrict__ a1, float2* __restrict__ a2) {
    __shared__ __align__(128) float4 s0[BWD_STAGES][BWD_BATCH];
    __shared__ __align__(128) float4 s1[BWD_STAGES][BWD_BATCH];
    __shared__ __align__(128) float4 s2[BWD_STAGES][BWD_BATCH];
    __shared__ __align__(16) float acc[BWD_WARPS][BWD_BATCH][ACC_STRIDE];
    __shared__ __align__(8) uint64_t full[BWD_STAGES];
    __shared__ uint32_t s_max[BWD_WARPS];

    const uint32_t tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const size_t P = (size_t)W * H, pid = (size_t)py * W + px;

    const uint32_t off = tile_offset[tile];
    const uint32_t my_n = inside ? n_contrib[pid] : 0u;

    // tile-wide last contributor: nothing behind it can receive gradient
    uint32_t m = my_n;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) s_max[warp] = m;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < BWD_STAGES; s++) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k < BWD_WARPS * BWD_BATCH * ACC_STRIDE; k += TILE_PIX) (&acc[0][0][0])[k] = 0.f;
    __syncthreads();
    uint32_t n = 0;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; w++) n = max(n, s_max[w]);
    if (n == 0) return;
    const int nb = (n + BWD_BATCH - 1) / BWD_BATCH;

    // batches are consumed in the order q = 0..nb-1  <->  list batch b = nb-1-q
    auto issue = [&](int q) {   // thread 0 only
        const int st = q % BWD_STAGES;
        const int b = nb - 1 - q;
        const uint32_t cnt = min((uint32_t)BWD_BATCH, n - (uint32_t)b * BWD_BATCH);
        const uint32_t bytes = cnt * 16u;
        const size_t src = (size_t)off + (size_t)b * BWD_BATCH;
        mbar_arrive_expect_tx(&full[st], 3u * bytes);
        bulk_g2s(&s0[st][0], p0 + src, bytes, &full[st]);
        bulk_g2s(&s1[st][0], p1 + src, bytes, &full[st]);
        bulk_g2s(&s2[st][0], p2 + src, bytes, &full[st]);
    };
    if (threadIdx.x == 0)
        for (int q = 0; q < nb && q < BWD_STAGES; q++) issue(q);

    const float T_final = inside ? final_T[pid] : 0.f;
    float T = T_final;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        if (dL_dcolor) { gC0 = dL_dcolor[pid]; gC1 = dL_dcolor[P + pid]; gC2 = dL_dcolor[2 * P + pid]; }
        if (dL_ddepth) gD = dL_ddepth[pid];
        if (dL_dalpha) gA = dL_dalpha[pid];
    }
    const float bg_dot = bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2;
    float ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, acd = 0.f, aca = 0.f;     // values "behind"
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
    const float half_W = 0.5f * W, half_H = 0.5f * H;

    for (int q = 0; q < nb; q++) {
        const int st = q % BWD_STAGES;
        const int b = nb - 1 - q;
        const int cnt = min(BWD_BATCH, (int)n - b * BWD_BATCH);
        mbar_wait(&full[st], (uint32_t)(q / BWD_STAGES) & 1u);
        for (int j = cnt - 1; j >= 0; j--) {
            const uint32_t idx = (uint32_t)(b * BWD_BATCH + j);
            bool contrib = idx < my_n;
            float4 a, c;
            float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
            if (contrib) {
                a = s0[st][j];
                c = s1[st][j];
                dx = a.x - fx;
                dy = a.y - fy;
                const float power = -0.5f * (a.z * dx * dx + c.x * dy * dy) - a.w * dx * dy;
                G = __expf(power);
                alpha = fminf(ALPHA_MAX, c.y * G);
                contrib = (power <= 0.f) && (alpha >= ALPHA_MIN);
            }
            if (!__any_sync(0xffffffffu, contrib)) continue;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f, v9 = 0.f;
            if (contrib) {
                T = T / (1.f - alpha);
                const float w = alpha * T;
                const float4 col = s2[st][j];
                float dL_da = 0.f;
                ac0 = last_alpha * lc0 + (1.f - last_alpha) * ac0; lc0 = col.x; dL_da += (col.x - ac0) * gC0;
                ac1 = last_alpha * lc1 + (1.f - last_alpha) * ac1; lc1 = col.y; dL_da += (col.y - ac1) * gC1;
                ac2 = last_alpha * lc2 + (1.f - last_alpha) * ac2; lc2 = col.z; dL_da += (col.z - ac2) * gC2;
                acd = last_alpha * ld + (1.f - last_alpha) * acd;  ld = c.z;   dL_da += (c.z - acd) * gD;
                aca = last_alpha + (1.f - last_alpha) * aca;                   dL_da += (1.f - aca) * gA;
                dL_da *= T;
                last_alpha = alpha;
                dL_da += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = c.y * dL_da;
                const float gdx = G * dx, gdy = G * dy;
                v0 = dL_dG * (-gdx * a.z - gdy * a.w) * half_W;
                v1 = dL_dG * (-gdy * c.x - gdx * a.w) * half_H;
                v2 = -0.5f * gdx * dx * dL_dG;
                v3 = -gdx * dy * dL_dG;
                v4 = -0.5f * gdy * dy * dL_dG;
                v5 = G * dL_da;
                v6 = w * gC0;
                v7 = w * gC1;
                v8 = w * gC2;
                v9 = w * gD;
            }
            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
            v5 = warp_sum(v5); v6 = warp_sum(v6); v7 = warp_sum(v7); v8 = warp_sum(v8); v9 = warp_sum(v9);
            if (lane == 0) {
                float4* slot = reinterpret_cast<float4*>(&acc[warp][j][0]);
                slot[0] = make_float4(v0, v1, v2, v3);
                slot[1] = make_float4(v4, v5, v6, v7);
                *reinterpret_cast<float2*>(&acc[warp][j][8]) = make_float2(v8, v9);
            }
        }
        __syncthreads();   // warp slots complete
        // fold the 8 warp slots; 3 work items (float4, float4, float2) per record
        for (int item = threadIdx.x; item < cnt * 3; item += TILE_PIX) {
            const int j = item / 3, part = item - j * 3;
            const uint32_t id = __float_as_uint(s1[st][j].w);
            if (part < 2) {
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < BWD_WARPS; w++) {
                    float4* slot = reinterpret_cast<float4*>(&acc[w][j][4 * part]);
                    const float4 q4 = *slot;
                    sum.x += q4.x; sum.y += q4.y; sum.z += q4.z; sum.w += q4.w;
                    *slot = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f)
                    atomicAdd(part == 0 ? &a0[id] : &a1[id], sum);
            } else {
                float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                for (int w = 0; w < BWD_WARPS; w++) {
                    float2* slot = reinterpret_cast<float2*>(&acc[w][j][8]);
                    const float2 q2 = *slot;
                    sum.x += q2.x; sum.y += q2.y;
                    *slot = make_float2(0.f, 0.f);
                }
                if (sum.x != 0.f || sum.y != 0.f) atomicAdd(&a2[id], sum);
            }
        }
        __syncthreads();   // stage st and the slots are free again
        if (threadIdx.x == 0 && q + BWD_STAGES < nb) issue(q + BWD_STAGES);
    }
}

int launch_blend_bwd(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                     const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, const AccumWS& acc,
                     cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0 || v.num_gaussians == 0) return 0;
    blend_bwd_kernel<<<T, TILE_PIX, 0, s>>>(t.offset, r.p0, r.p1, r.p2, v.image_width, v.image_height, gx, in.bg,
                                            img.n_contrib, img.final_T, dL_dcolor, dL_ddepth, dL_dalpha, acc.a0,
                                            acc.a1, acc.a2);
    return 1;
}

}  // namespace gg
