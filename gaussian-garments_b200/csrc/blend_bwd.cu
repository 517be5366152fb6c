// blend_bwd.cu -- per-tile back-to-front gradient pass.
//
// One CTA per tile, one pixel per thread.  The tile's packed records are streamed BACKWARDS
// (from the last contributor of any pixel in the tile) through a 2-stage bulk-TMA/mbarrier ring.
// Per Gaussian every thread forms its 10 partial derivatives; a warp that has no contributing
// pixel skips the Gaussian (one ballot).  Otherwise the warp reduces the 10 values with a
// TRANSPOSED butterfly (reduce-scatter: 5+3+2+1+1 = 12 shuffles instead of 10 x 5 = 50), after which
// ten lanes each hold one finished sum and park it in the warp's shared-memory slot with a single
// predicated store.  After a batch the CTA folds the 8 warp slots and issues at most three vector
// reductions (red.global.add.v4.f32 x2, .v2.f32 x1) per (tile, Gaussian) instance -- instead of
// upstream's ~10 scalar atomics per (pixel, Gaussian).
#include "common.cuh"

// Blackwell packed fp32 (FFMA2 / FADD2 / FMUL2) for the per-pixel gradient terms and the reduce-scatter adds:
// measured -3 % on blend_bwd (issue-bound kernel; fewer issue slots for the same FMA-pipe work).
#ifndef GG_NO_F32X2
#define GG_F32X2 1
#endif

namespace gg {

constexpr int BWD_BATCH = 64;   // = 2 ballot words of the per-warp entry mask
constexpr int BWD_STAGES = 2;
constexpr int BWD_WARPS = TILE_PIX / 32;
constexpr int ACC_STRIDE = 12;   // 10 used; keeps float4 alignment

// Reduce-scatter of 10 per-lane values over the 32 lanes of a warp.  On return lane L holds the
// warp-wide sum of value `slot` (or nothing when slot < 0); lanes 2k and 2k+1 hold the same one.
__device__ __forceinline__ float warp_reduce_scatter10(const float* v, int lane, int& slot) {
    const unsigned FULL = 0xffffffffu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float r0, r1, r2, r3, r4;
    {   // xor 16: lower half keeps values 0..4, upper half keeps 5..9
#ifdef GG_F32X2
        const float2 a01 = __fadd2_rn(make_float2(b4 ? v[5] : v[0], b4 ? v[6] : v[1]),
                                      make_float2(__shfl_xor_sync(FULL, b4 ? v[0] : v[5], 16), __shfl_xor_sync(FULL, b4 ? v[1] : v[6], 16)));
        const float2 a23 = __fadd2_rn(make_float2(b4 ? v[7] : v[2], b4 ? v[8] : v[3]),
                                      make_float2(__shfl_xor_sync(FULL, b4 ? v[2] : v[7], 16), __shfl_xor_sync(FULL, b4 ? v[3] : v[8], 16)));
        r0 = a01.x; r1 = a01.y; r2 = a23.x; r3 = a23.y;
#else
        r0 = (b4 ? v[5] : v[0]) + __shfl_xor_sync(FULL, b4 ? v[0] : v[5], 16);
        r1 = (b4 ? v[6] : v[1]) + __shfl_xor_sync(FULL, b4 ? v[1] : v[6], 16);
        r2 = (b4 ? v[7] : v[2]) + __shfl_xor_sync(FULL, b4 ? v[2] : v[7], 16);
        r3 = (b4 ? v[8] : v[3]) + __shfl_xor_sync(FULL, b4 ? v[3] : v[8], 16);
#endif
        r4 = (b4 ? v[9] : v[4]) + __shfl_xor_sync(FULL, b4 ? v[4] : v[9], 16);
    }
    float s0, s1, s2;
    {   // xor 8: b3 = 0 keeps {r0,r1,r2}, b3 = 1 keeps {r3,r4,-}
#ifdef GG_F32X2
        const float2 a = __fadd2_rn(make_float2(b3 ? r3 : r0, b3 ? r4 : r1),
                                    make_float2(__shfl_xor_sync(FULL, b3 ? r0 : r3, 8), __shfl_xor_sync(FULL, b3 ? r1 : r4, 8)));
        s0 = a.x; s1 = a.y;
#else
        s0 = (b3 ? r3 : r0) + __shfl_xor_sync(FULL, b3 ? r0 : r3, 8);
        s1 = (b3 ? r4 : r1) + __shfl_xor_sync(FULL, b3 ? r1 : r4, 8);
#endif
        s2 = (b3 ? 0.f : r2) + __shfl_xor_sync(FULL, b3 ? r2 : 0.f, 8);
    }
    float t0, t1;
    {   // xor 4: b2 = 0 keeps {s0,s1}, b2 = 1 keeps {s2,-}
#ifdef GG_F32X2
        const float2 a = __fadd2_rn(make_float2(b2 ? s2 : s0, b2 ? 0.f : s1),
                                    make_float2(__shfl_xor_sync(FULL, b2 ? s0 : s2, 4), __shfl_xor_sync(FULL, b2 ? s1 : 0.f, 4)));
        t0 = a.x; t1 = a.y;
#else
        t0 = (b2 ? s2 : s0) + __shfl_xor_sync(FULL, b2 ? s0 : s2, 4);
        t1 = (b2 ? 0.f : s1) + __shfl_xor_sync(FULL, b2 ? s1 : 0.f, 4);
#endif
    }
    float u = (b1 ? t1 : t0) + __shfl_xor_sync(FULL, b1 ? t0 : t1, 2);   // xor 2
    u += __shfl_xor_sync(FULL, u, 1);                                    // xor 1
    // which value did this lane end up with?
    int within;
    if (!b3) within = !b2 ? (b1 ? 1 : 0) : (b1 ? -1 : 2);
    else     within = !b2 ? (b1 ? 4 : 3) : -1;
    slot = within < 0 ? -1 : (b4 ? 5 : 0) + within;
    return u;
}

#ifndef GG_BWD_MINB
#define GG_BWD_MINB 4
#endif
__global__ void __launch_bounds__(TILE_PIX, GG_BWD_MINB)
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
    const float bgT = -T_final * (bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2);
    float aca = 0.f, last_alpha = 0.f;                               // alpha channel "behind", previous alpha
#ifdef GG_F32X2
    float2 ac01 = make_float2(0.f, 0.f), ac2d = make_float2(0.f, 0.f);   // (r,g) and (b,depth) accumulated behind
    float2 lc01 = make_float2(0.f, 0.f), lc2d = make_float2(0.f, 0.f);   // previous contributor's (r,g), (b,depth)
    const float2 g01 = make_float2(gC0, gC1), g2d = make_float2(gC2, gD);
#else
    float ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, acd = 0.f;               // values "behind"
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
#endif
    const float kx = LN2 * 0.5f * W, ky = LN2 * 0.5f * H;           // d/dx of 2^p2 carries ln 2
    const uint32_t acc_warp = smem_u32(&acc[warp][0][0]);

    for (int q = 0; q < nb; q++) {
        const int st = q % BWD_STAGES;
        const int b = nb - 1 - q;
        const int cnt = min(BWD_BATCH, (int)n - b * BWD_BATCH);
        mbar_wait(&full[st], (uint32_t)(q / BWD_STAGES) & 1u);
        const uint32_t r0 = smem_u32(&s0[st][0]), r1 = smem_u32(&s1[st][0]), r2 = smem_u32(&s2[st][0]);
        // entries this warp may touch: bit `warp` of the record's warp-overlap mask (two ballots per batch);
        // entries whose alpha >= 1/255 bounding box misses this warp's 8x4 block cost nothing
        uint32_t m_lo, m_hi;
        {
            const uint32_t w0 = (lane < cnt) ? __float_as_uint(lds32(r1 + 16u * lane + 12u)) : 0u;
            const uint32_t w1 = (lane + 32 < cnt) ? __float_as_uint(lds32(r1 + 16u * (lane + 32) + 12u)) : 0u;
            m_lo = __ballot_sync(0xffffffffu, (w0 >> warp) & 1u);
            m_hi = __ballot_sync(0xffffffffu, (w1 >> warp) & 1u);
        }
        while (m_hi | m_lo) {
            int j;
            if (m_hi) {
                const int bit = 31 - __clz(m_hi);
                m_hi &= ~(1u << bit);
                j = 32 + bit;
            } else {
                const int bit = 31 - __clz(m_lo);
                m_lo &= ~(1u << bit);
                j = bit;
            }
            const uint32_t idx = (uint32_t)(b * BWD_BATCH + j);
            const float4 a = lds128(r0 + 16u * j);
            const float4 c = lds128(r1 + 16u * j);
            const float dx = a.x - fx, dy = a.y - fy;
            const float e2 = dx * (a.z * dx + a.w * dy) + (c.x * dy) * dy;     // log2-domain exponent
            const float G = ex2_approx(e2);
            const float alpha = fminf(ALPHA_MAX, c.y * G);
            const bool contrib = (idx < my_n) && (e2 <= 0.f) && (alpha >= ALPHA_MIN);
            if (!__any_sync(0xffffffffu, contrib)) continue;
            float v[10];
#pragma unroll
            for (int k = 0; k < 10; k++) v[k] = 0.f;
#ifdef GG_F32X2
            if (contrib) {      // Blackwell packed fp32 (FFMA2/FADD2/FMUL2): two lanes of the colour/depth state per instruction
                const float rinv = rcp_approx(1.f - alpha);
                T *= rinv;
                const float w = alpha * T;
                const float4 col = lds128(r2 + 16u * j);
                const float oml = 1.f - last_alpha;
                const float2 la2 = make_float2(last_alpha, last_alpha), om2 = make_float2(oml, oml);
                const float2 neg1 = make_float2(-1.f, -1.f);
                ac01 = __ffma2_rn(la2, lc01, __fmul2_rn(om2, ac01));
                ac2d = __ffma2_rn(la2, lc2d, __fmul2_rn(om2, ac2d));
                lc01 = make_float2(col.x, col.y);
                lc2d = make_float2(col.z, c.z);
                const float2 t01 = __fmul2_rn(__ffma2_rn(ac01, neg1, lc01), g01);
                const float2 t2d = __fmul2_rn(__ffma2_rn(ac2d, neg1, lc2d), g2d);
                const float2 tt = __fadd2_rn(t01, t2d);
                aca = last_alpha + oml * aca;
                float dL_da = (tt.x + tt.y) + (1.f - aca) * gA;
                dL_da = dL_da * T + bgT * rinv;
                last_alpha = alpha;
                const float X = c.y * dL_da * G;          // dL/dG * G
                const float2 d2 = make_float2(dx, dy), dyx = make_float2(dy, dx);
                const float2 inner = __ffma2_rn(make_float2(2.f * a.z, 2.f * c.x), d2, __fmul2_rn(make_float2(a.w, a.w), dyx));
                const float2 v01 = __fmul2_rn(inner, make_float2(X * kx, X * ky));
                const float2 sq = __fmul2_rn(__fmul2_rn(d2, d2), make_float2(-0.5f * X, -0.5f * X));
                const float2 v67 = __fmul2_rn(make_float2(w, w), g01), v89 = __fmul2_rn(make_float2(w, w), g2d);
                v[0] = v01.x; v[1] = v01.y;
                v[2] = sq.x;  v[4] = sq.y;
                v[3] = -X * dx * dy;
                v[5] = G * dL_da;
                v[6] = v67.x; v[7] = v67.y; v[8] = v89.x; v[9] = v89.y;
            }
#else
            if (contrib) {
                const float rinv = rcp_approx(1.f - alpha);
                T *= rinv;
                const float w = alpha * T;
                const float4 col = lds128(r2 + 16u * j);
                const float oml = 1.f - last_alpha;
                float dL_da;
                ac0 = last_alpha * lc0 + oml * ac0; lc0 = col.x; dL_da = (col.x - ac0) * gC0;
                ac1 = last_alpha * lc1 + oml * ac1; lc1 = col.y; dL_da += (col.y - ac1) * gC1;
                ac2 = last_alpha * lc2 + oml * ac2; lc2 = col.z; dL_da += (col.z - ac2) * gC2;
                acd = last_alpha * ld + oml * acd;  ld = c.z;    dL_da += (c.z - acd) * gD;
                aca = last_alpha + oml * aca;                    dL_da += (1.f - aca) * gA;
                dL_da = dL_da * T + bgT * rinv;
                last_alpha = alpha;
                const float X = c.y * dL_da * G;          // dL/dG * G
                v[0] = X * (2.f * a.z * dx + a.w * dy) * kx;
                v[1] = X * (2.f * c.x * dy + a.w * dx) * ky;
                const float Xdx = X * dx;
                v[2] = -0.5f * Xdx * dx;
                v[3] = -Xdx * dy;
                v[4] = -0.5f * X * dy * dy;
                v[5] = G * dL_da;
                v[6] = w * gC0;
                v[7] = w * gC1;
                v[8] = w * gC2;
                v[9] = w * gD;
            }
#endif
            int slot;
            const float sum = warp_reduce_scatter10(v, lane, slot);
            if (slot >= 0 && !(lane & 1))
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_warp + 4u * (uint32_t)(j * ACC_STRIDE + slot)), "f"(sum) : "memory");
            // (measured alternative: ten predicated red.global per active (warp, entry) instead of the shared slots +
            //  per-batch fold -> blend_bwd 0.516 -> 0.605 ms: L2 atomic throughput loses to the 3 vector reds per instance)
        }
        __syncthreads();   // warp slots complete
        // fold the 8 warp slots; 3 work items (float4, float4, float2) per record
        for (int item = threadIdx.x; item < cnt * 3; item += TILE_PIX) {
            const int j = item / 3, part = item - j * 3;
            const uint32_t id = __float_as_uint(s2[st][j].w);
            if (part < 2) {
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < BWD_WARPS; w++) {
                    float4* sl = reinterpret_cast<float4*>(&acc[w][j][4 * part]);
                    const float4 q4 = *sl;
                    sum.x += q4.x; sum.y += q4.y; sum.z += q4.z; sum.w += q4.w;
                    *sl = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f)
                    atomicAdd(part == 0 ? &a0[id] : &a1[id], sum);
            } else {
                float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                for (int w = 0; w < BWD_WARPS; w++) {
                    float2* sl = reinterpret_cast<float2*>(&acc[w][j][8]);
                    const float2 q2 = *sl;
                    sum.x += q2.x; sum.y += q2.y;
                    *sl = make_float2(0.f, 0.f);
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
