// blend_bwd2.cu -- per-tile gradient pass, "evaluate pixel-parallel, reduce splat-parallel" (SURVEY.md 8a row a10).
//
// Round 1's kernel (blend_bwd.cu, kept as GG_BWD_PATH=v1) formed 10 partial derivatives per pixel and reduced them
// over the 32 lanes of a warp with a 12-shuffle / 44-instruction reduce-scatter per (warp, list entry): ~157 warp
// instructions per active pair, issue-bound at 0.50 ms.  This kernel splits the work where the data dependence
// actually is:
//
//   phase 1 (pixel-parallel, one thread per pixel, warp = 8x4 pixel block, list walked back to front)
//       Only the per-pixel recurrences live here.  With v_i = (c_i, z_i, 1), upstream g = (gC, gD, gA), w_i = alpha_i T_i
//       and the suffix sum R_i = sum_{j>i} w_j (g.v_j) + T_final (gC.bg)  [checked on the CPU in
//       oracle/gg_oracle.c::ggo_backward_forward_order], dL/dalpha_i = T_i (g.v_i) - R_i / (1 - alpha_i).
//       A lane produces just TWO numbers per entry:  X = dL/dG * G  and  w.  They go to a warp-private shared-memory
//       panel [16 entries][32 pixels] (conflict-free, padded).
//   phase 2 (splat-parallel, every 16 active entries of a warp: lane = (entry, half of the 8x4 block))
//       All ten per-Gaussian sums are MOMENTS of the panel rows over the pixel lattice:
//           sum X, sum X cx, sum X cy, sum X cx^2, sum X cx cy, sum X cy^2   (cx, cy = compile-time lattice offsets)
//           sum w gC0, sum w gC1, sum w gC2, sum w gD
//       so a lane walks its 16 pixels with immediate-operand FMAs and NO cross-lane traffic, shifts the moments to
//       the splat centre once, merges the two halves with one shuffle per value and issues the vector reds
//       (red.global.add.v4.f32 x2 + .v2.f32) for its entry.
//
// Records stream backwards through a 4-stage bulk-TMA ring.  Warps are decoupled: a stage is recycled through a
// per-stage "empty" mbarrier (8 warp arrivals) instead of __syncthreads, so a warp whose 8x4 block sees few splats
// never waits for the busiest warp of the tile at batch granularity.
#include "common.cuh"

namespace gg {

constexpr int B2_BATCH = 64;          // records per ring stage (= 2 ballot words)
constexpr int B2_STAGES = 4;
constexpr int B2_WARPS = TILE_PIX / 32;
constexpr int B2_ROUND = 16;          // entries per transpose round
constexpr int B2_PITCH = 33;          // padded panel row (floats)

struct B2Smem {
    float4 s0[B2_STAGES][B2_BATCH];
    float4 s1[B2_STAGES][B2_BATCH];
    float4 s2[B2_STAGES][B2_BATCH];
    float4 gpix[B2_WARPS][2][17];                   // (gC0, gC1, gC2, gD) of the warp's 2 x 16 pixels; the odd pitch keeps
                                                    // the two half-warp broadcast reads of phase 2 on different banks
    float4 meta[B2_WARPS][B2_ROUND][2];             // (mx, my, A', B') , (C', opacity, depth, id bits)
    float2 xw[B2_WARPS][B2_ROUND * B2_PITCH];       // (X, w) panel: [entry][pixel], row pitch 33 -> conflict-free
    uint64_t full[B2_STAGES];
    uint64_t empty[B2_STAGES];
    uint32_t s_max[B2_WARPS];
};
// Every shared access below is an explicit 32-bit shared-window address = one pinned base register + a
// compile-time offset (the generic path made ptxas rebuild the window base -- S2R SR_CgaCtaId + LEA -- per use).
constexpr uint32_t O_S0 = offsetof(B2Smem, s0), O_S1 = offsetof(B2Smem, s1), O_S2 = offsetof(B2Smem, s2);
constexpr uint32_t O_GPIX = offsetof(B2Smem, gpix), O_META = offsetof(B2Smem, meta), O_XW = offsetof(B2Smem, xw);
constexpr uint32_t O_FULL = offsetof(B2Smem, full), O_EMPTY = offsetof(B2Smem, empty), O_SMAX = offsetof(B2Smem, s_max);
constexpr uint32_t STAGE_BYTES = B2_BATCH * 16;
constexpr uint32_t XW_WARP_BYTES = B2_ROUND * B2_PITCH * 8, META_WARP_BYTES = B2_ROUND * 32, GPIX_WARP_BYTES = 2 * 17 * 16;

// phase 2: reduce the warp's panel (cnt <= 16 entries) and add the result to the per-Gaussian accumulators.
// Everything it needs besides the panel is re-derived here (once per 16 entries) so that nothing stays live in
// registers across phase 1.
template <bool DA>
__device__ __forceinline__ void b2_flush(int cnt, uint32_t sb, int tile, int W, int H, int gx, float4* __restrict__ a0,
                                         float4* __restrict__ a1, float2* __restrict__ a2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = lane & (B2_ROUND - 1), h = lane >> 4;
    const uint32_t row = sb + O_XW + (uint32_t)warp * XW_WARP_BYTES + 8u * (uint32_t)(e * B2_PITCH + h * 16);
    const uint32_t gp = sb + O_GPIX + (uint32_t)warp * GPIX_WARP_BYTES + 272u * (uint32_t)h;
    float A0 = 0.f, B0 = 0.f, C0 = 0.f, A1 = 0.f, B1 = 0.f, C1 = 0.f;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const float2 xw = lds64(row + 8u * p);
        const float X = xw.x, w = xw.y;
        const float4 g = lds128(gp + 16u * p);
        const float cx = (float)(p & 7) - 3.5f;
        if (p < 8) {
            A0 += X; B0 = fmaf(X, cx, B0); C0 = fmaf(X, cx * cx, C0);
        } else {
            A1 += X; B1 = fmaf(X, cx, B1); C1 = fmaf(X, cx * cx, C1);
        }
        q0 = fmaf(w, g.x, q0); q1 = fmaf(w, g.y, q1); q2 = fmaf(w, g.z, q2);
        if (DA) q3 = fmaf(w, g.w, q3);
    }
    // lattice moments about the centre of this half (two rows: cy = -0.5, +0.5)
    const float S0 = A0 + A1, Mx = B0 + B1, Mxx = C0 + C1;
    const float My = 0.5f * (A1 - A0), Mxy = 0.5f * (B1 - B0), Myy = 0.25f * S0;
    const uint32_t mb = sb + O_META + (uint32_t)warp * META_WARP_BYTES + 32u * (uint32_t)e;
    const float4 m0 = lds128(mb), m1 = lds128(mb + 16u);
    // shift to the splat centre: dx = ux - cx, dy = uy - cy
    const float ox = (float)((tile % gx) * TILE + (warp & 1) * 8) + 3.5f;
    const float oy = (float)((tile / gx) * TILE + (warp >> 1) * 4 + 2 * h) + 0.5f;
    const float ux = m0.x - ox, uy = m0.y - oy;
    const float Sdx = ux * S0 - Mx, Sdy = uy * S0 - My;
    const float Sdxx = ux * (ux * S0 - 2.f * Mx) + Mxx;
    const float Sdyy = uy * (uy * S0 - 2.f * My) + Myy;
    const float Sdxy = ux * (uy * S0 - My) - uy * Mx + Mxy;
    const float kx = LN2 * 0.5f * W, ky = LN2 * 0.5f * H;               // d/dx of 2^e2 carries ln 2
    float v0 = kx * (2.f * m0.z * Sdx + m0.w * Sdy);
    float v1 = ky * (2.f * m1.x * Sdy + m0.w * Sdx);
    float v2 = -0.5f * Sdxx, v3 = -Sdxy, v4 = -0.5f * Sdyy;
    float v5 = S0 / m1.y;                                   // sum G dL/dalpha = (sum X) / opacity
    const unsigned FULL = 0xffffffffu;
    v0 += __shfl_xor_sync(FULL, v0, 16); v1 += __shfl_xor_sync(FULL, v1, 16);
    v2 += __shfl_xor_sync(FULL, v2, 16); v3 += __shfl_xor_sync(FULL, v3, 16);
    v4 += __shfl_xor_sync(FULL, v4, 16); v5 += __shfl_xor_sync(FULL, v5, 16);
    q0 += __shfl_xor_sync(FULL, q0, 16); q1 += __shfl_xor_sync(FULL, q1, 16);
    q2 += __shfl_xor_sync(FULL, q2, 16);
    if (DA) q3 += __shfl_xor_sync(FULL, q3, 16);
    if (e < cnt) {
        const uint32_t id = __float_as_uint(m1.w);
        // every parked entry had at least one contributing pixel (phase 1's vote): no zero tests needed
        if (h == 0) {
            atomicAdd(&a0[id], make_float4(v0, v1, v2, v3));
            atomicAdd(&a2[id], make_float2(q2, q3));
        } else {
            atomicAdd(&a1[id], make_float4(v4, v5, q0, q1));
        }
    }
}

// MINB: resident CTAs per SM the register allocation aims for (4 -> 64 regs, 3 -> 80 regs; the kernel is issue-bound,
// so fewer rematerialised addresses can beat the extra warps).  DA: dL/ddepth or dL/dalpha present (the training
// loops of the reference feed only the colour into the loss, s2_registration.py:252-260 / s3_appearance.py:125-133).
template <int MINB, bool DA>
__global__ void __launch_bounds__(TILE_PIX, MINB)
blend_bwd2_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ p0,
                  const float4* __restrict__ p1, const float4* __restrict__ p2, int W, int H, int gx,
                  const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                  const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha, float4* __restrict__ a0,
                  float4* __restrict__ a1, float2* __restrict__ a2, const uint32_t* __restrict__ order) {
    extern __shared__ __align__(128) unsigned char b2_raw[];
    uint32_t sb = smem_u32(b2_raw);
    asm volatile("" : "+r"(sb));          // pin: one register, never rematerialised

    const uint32_t tile = order[blockIdx.x];              // launch order: heaviest tiles first (tile_scan_kernel)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = (tile % gx) * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = (tile / gx) * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t P = (size_t)W * H, pid = (size_t)py * W + px;

    const uint32_t off = tile_offset[tile];
    const uint32_t my_n = inside ? n_contrib[pid] : 0u;
    const uint32_t warp_n = __reduce_max_sync(0xffffffffu, my_n);   // nothing behind it matters to this warp
    if (lane == 0) sts32(sb + O_SMAX + 4u * warp, warp_n);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < B2_STAGES; s++) {
            mbar_init_a(sb + O_FULL + 8u * s, 1);
            mbar_init_a(sb + O_EMPTY + 8u * s, B2_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t n = 0;
#pragma unroll
    for (int w = 0; w < B2_WARPS; w++) n = max(n, ldsu32(sb + O_SMAX + 4u * w));
    if (n == 0) return;
    const int nb = (n + B2_BATCH - 1) / B2_BATCH;

    const float T_final = inside ? final_T[pid] : 0.f;
    float T = T_final;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        if (dL_dcolor) { gC0 = dL_dcolor[pid]; gC1 = dL_dcolor[P + pid]; gC2 = dL_dcolor[2 * P + pid]; }
        if (DA && dL_ddepth) gD = dL_ddepth[pid];
        if (DA && dL_dalpha) gA = dL_dalpha[pid];
    }
    sts128(sb + O_GPIX + (uint32_t)warp * GPIX_WARP_BYTES + 272u * (uint32_t)(lane >> 4) + 16u * (lane & 15),
           make_float4(gC0, gC1, gC2, gD));
    __syncwarp();
    float R = T_final * (bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2);     // suffix sum, starts with the background term
    const float fx = (float)px, fy = (float)py;
    // running panel addresses (warp-uniform `fill` entries are parked)
    const uint32_t xw_lane = sb + O_XW + (uint32_t)warp * XW_WARP_BYTES + 8u * lane;
    const uint32_t meta_w = sb + O_META + (uint32_t)warp * META_WARP_BYTES;
    int fill = 0;
    int issued = 0;                       // thread 0 only: batches handed to the TMA so far

    // batches are consumed in the order q = 0..nb-1  <->  list batch b = nb-1-q (back to front)
#pragma unroll 1
    for (int q = 0; q < nb; q++) {
        const int st = q % B2_STAGES;
        const int b = nb - 1 - q;
        const int cnt = min(B2_BATCH, (int)n - b * B2_BATCH);
        if (threadIdx.x == 0) {
            // producer: refill every stage that all 8 warps have released; block only when the batch this warp
            // needs right now has not been issued yet
            while (issued < nb && issued < q + B2_STAGES) {
                if (issued >= B2_STAGES) {
                    const int prev = issued - B2_STAGES;
                    const uint32_t eb = sb + O_EMPTY + 8u * (uint32_t)(prev % B2_STAGES);
                    const uint32_t par = (uint32_t)(prev / B2_STAGES) & 1u;
                    if (issued == q) mbar_wait_a(eb, par);
                    else if (!mbar_test_a(eb, par)) break;
                }
                const int ist = issued % B2_STAGES, ib = nb - 1 - issued;
                const uint32_t icnt = min((uint32_t)B2_BATCH, n - (uint32_t)ib * B2_BATCH);
                const uint32_t bytes = icnt * 16u, fb = sb + O_FULL + 8u * ist;
                const size_t src = (size_t)off + (size_t)ib * B2_BATCH;
                mbar_expect_tx_a(fb, 3u * bytes);
                bulk_g2s_a(sb + O_S0 + ist * STAGE_BYTES, p0 + src, bytes, fb);
                bulk_g2s_a(sb + O_S1 + ist * STAGE_BYTES, p1 + src, bytes, fb);
                bulk_g2s_a(sb + O_S2 + ist * STAGE_BYTES, p2 + src, bytes, fb);
                issued++;
            }
        }
        mbar_wait_a(sb + O_FULL + 8u * st, (uint32_t)(q / B2_STAGES) & 1u);
        const uint32_t r0 = sb + O_S0 + st * STAGE_BYTES;
        // work list of this warp: bit `warp` of the record's warp-overlap mask, entries in front of the warp's
        // last contributor only
        const int lim = min(cnt, (int)warp_n - b * B2_BATCH);           // entries [0, lim) of this batch can matter
        const int my_lim = (int)my_n - b * B2_BATCH;                    // entry j is in front of this pixel's stop
        uint32_t m_lo, m_hi;
        {
            const uint32_t w0 = (lane < lim) ? ldsu32(r0 + (O_S1 - O_S0) + 16u * lane + 12u) : 0u;
            const uint32_t w1 = (lane + 32 < lim) ? ldsu32(r0 + (O_S1 - O_S0) + 16u * (lane + 32) + 12u) : 0u;
            m_lo = __ballot_sync(0xffffffffu, (w0 >> warp) & 1u);
            m_hi = __ballot_sync(0xffffffffu, (w1 >> warp) & 1u);
        }
#pragma unroll 1
        for (int half = 1; half >= 0; half--) {
            uint32_t m = half ? m_hi : m_lo;
            const uint32_t rh = r0 + 512u * (uint32_t)half;
            const int lim_h = my_lim - 32 * half;
#pragma unroll 1
            while (m) {
                uint32_t bit;
                asm("bfind.u32 %0, %1;" : "=r"(bit) : "r"(m));                     // highest set bit: back to front
                m ^= 1u << bit;
                const uint32_t ra = rh + 16u * (uint32_t)bit;
                const float4 a = lds128(ra);
                const float4 c = lds128(ra + (O_S1 - O_S0));
                const float dx = a.x - fx, dy = a.y - fy;
                const float e2 = dx * (a.z * dx + a.w * dy) + (c.x * dy) * dy;     // log2-domain exponent
                const float G = ex2_approx(e2);
                const float araw = c.y * G;
                const float alpha = fminf(ALPHA_MAX, araw);
                const bool contrib = ((int)bit < lim_h) && (e2 <= 0.f) && (alpha >= ALPHA_MIN);
                if (!__any_sync(0xffffffffu, contrib)) continue;
                const float4 col = lds128(ra + (O_S2 - O_S0));
                // branch-free: non-contributing lanes leave T and R untouched and park zeros
                const float rinv = contrib ? rcp_approx(1.f - alpha) : 1.f;
                T *= rinv;                                                         // T_i = T_{i+1} / (1 - alpha_i)
                const float w = contrib ? alpha * T : 0.f;
                const float gv = DA ? fmaf(col.x, gC0, fmaf(col.y, gC1, fmaf(col.z, gC2, fmaf(c.z, gD, gA))))
                                    : fmaf(col.x, gC0, fmaf(col.y, gC1, col.z * gC2));
                const float s = T * gv - R * rinv;                                 // dL/dalpha_i
                R = fmaf(w, gv, R);
                const float X = contrib ? araw * s : 0.f;                          // opacity G dL/dalpha = G dL/dG
                sts64(xw_lane + (uint32_t)fill * (B2_PITCH * 8u), X, w);
                if (lane == 0) {
                    sts128(meta_w + 32u * (uint32_t)fill, a);
                    sts128(meta_w + 32u * (uint32_t)fill + 16u, make_float4(c.x, c.y, c.z, col.w));
                }
                if (++fill == B2_ROUND) {
                    __syncwarp();
                    b2_flush<DA>(B2_ROUND, sb, (int)tile, W, H, gx, a0, a1, a2);
                    __syncwarp();
                    fill = 0;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sb + O_EMPTY + 8u * st);          // this warp is done reading stage st
    }
    if (fill > 0) {
        __syncwarp();
        b2_flush<DA>(fill, sb, (int)tile, W, H, gx, a0, a1, a2);
    }
}

template <int MINB, bool DA>
static void b2_launch(int T, cudaStream_t s, const TileWS& t, const RecordWS& r, const ImageWS& img, const gg_view& v,
                      int gx, const float* bg, const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                      const AccumWS& acc) {
    static bool attr_set[64] = {false};        // opt-in to > 48 KB dynamic shared memory, once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        // a failure here leaves attr_set false and surfaces as a launch error (checked by the C ABI right after)
        attr_set[dev] = cudaFuncSetAttribute(blend_bwd2_kernel<MINB, DA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(B2Smem)) == cudaSuccess;
    }
    blend_bwd2_kernel<MINB, DA><<<T, TILE_PIX, sizeof(B2Smem), s>>>(t.offset, r.p0, r.p1, r.p2, v.image_width,
                                                                    v.image_height, gx, bg, img.n_contrib, img.final_T,
                                                                    dL_dcolor, dL_ddepth, dL_dalpha, acc.a0, acc.a1, acc.a2, t.order);
}

int launch_blend_bwd2(const gg_view& v, const gg_inputs& in, const TileWS& t, const RecordWS& r, const ImageWS& img,
                      const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, const AccumWS& acc,
                      int min_blocks, cudaStream_t s) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    if (T == 0 || v.num_gaussians == 0) return 0;
    const bool da = dL_ddepth || dL_dalpha;
    if (min_blocks == 3) {
        if (da) b2_launch<3, true>(T, s, t, r, img, v, gx, in.bg, dL_dcolor, dL_ddepth, dL_dalpha, acc);
        else    b2_launch<3, false>(T, s, t, r, img, v, gx, in.bg, dL_dcolor, dL_ddepth, dL_dalpha, acc);
    } else {
        if (da) b2_launch<4, true>(T, s, t, r, img, v, gx, in.bg, dL_dcolor, dL_ddepth, dL_dalpha, acc);
        else    b2_launch<4, false>(T, s, t, r, img, v, gx, in.bg, dL_dcolor, dL_ddepth, dL_dalpha, acc);
    }
    return 1;
}

}  // namespace gg
