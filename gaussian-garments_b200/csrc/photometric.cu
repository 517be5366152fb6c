// photometric.cu -- fused photometric loss ("next" row N2 of SURVEY.md 8f): masked L1 + SSIM(11x11 Gaussian
// window, sigma 1.5) forward and backward in two kernels.
//
// Replaces, per iteration of the reference (s2_registration.py:259-260, s3_appearance.py:132-133),
//   l1_loss(image, gt, mask)                /root/reference/utils/loss_utils.py:17-21
//   ssim(image, gt, mask)                   /root/reference/utils/loss_utils.py:36-69
// i.e. 5 depthwise 11x11 convolutions over 3xHxW forward (+ their autograd backward) and ~10 elementwise passes.
// Here the window is applied separably out of shared memory: the forward kernel produces the two sums (L1, SSIM)
// and the three per-pixel partial-derivative maps of the SSIM index; the backward kernel convolves those maps once
// and emits dL/dimage directly (the L1 subgradient is fused in).  Zero padding, as conv2d(padding=5).
#include "common.cuh"

namespace gg {

// gaussian(11, 1.5) of utils/loss_utils.py:26-28, float32 as the reference builds it
__constant__ float kWin[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                               2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                               3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};
constexpr int PT = 16;          // output tile edge
constexpr int PH = 5;           // halo
constexpr int PI = PT + 2 * PH; // 26
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;
constexpr int PHOTO_SLOTS = 64;   // accumulator slots per sum (host adds them up)

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    if (w == 0) t = warp_sum(t);
    __syncthreads();
    return t;     // valid in warp 0
}

// grid (ceil(W/16), ceil(H/16), 3 channels), block 256
__global__ void __launch_bounds__(256)
photometric_fwd_kernel(int W, int H, const float* __restrict__ img, const float* __restrict__ gt,
                       const float* __restrict__ mask, float* __restrict__ m1, float* __restrict__ m2,
                       float* __restrict__ m3, double* __restrict__ sums /*[2]: l1, ssim*/) {
    __shared__ float sx[PI][PI + 1], sy[PI][PI + 1];
    __shared__ float hz[5][PI][PT + 1];
    __shared__ float red[8];
    const int ch = blockIdx.z;
    const int ox = blockIdx.x * PT, oy = blockIdx.y * PT;
    const size_t plane = (size_t)W * H;
    const float* I = img + ch * plane;
    const float* G = gt + ch * plane;
    for (int k = threadIdx.x; k < PI * PI; k += 256) {
        const int ly = k / PI, lx = k - ly * PI;
        const int gx_ = ox + lx - PH, gy_ = oy + ly - PH;
        float x = 0.f, y = 0.f;
        if (gx_ >= 0 && gx_ < W && gy_ >= 0 && gy_ < H) {
            const size_t p = (size_t)gy_ * W + gx_;
            const float mk = mask ? mask[p] : 1.f;
            x = I[p] * mk;
            y = G[p] * mk;
        }
        sx[ly][lx] = x;
        sy[ly][lx] = y;
    }
    __syncthreads();
    const bool with_ssim = m1 != nullptr;                       // block-uniform: lambda_dssim == 0 skips the SSIM work
    if (with_ssim)
    for (int k = threadIdx.x; k < PI * PT; k += 256) {        // horizontal pass
        const int ly = k / PT, lx = k - ly * PT;
        float a = 0.f, b = 0.f, c = 0.f, d = 0.f, e = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float w = kWin[t], x = sx[ly][lx + t], y = sy[ly][lx + t];
            a += w * x; b += w * y; c += w * x * x; d += w * y * y; e += w * x * y;
        }
        hz[0][ly][lx] = a; hz[1][ly][lx] = b; hz[2][ly][lx] = c; hz[3][ly][lx] = d; hz[4][ly][lx] = e;
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int px = ox + lx, py = oy + ly;
    float l1 = 0.f, ss = 0.f;
    if (px < W && py < H && !with_ssim) l1 = fabsf(sx[ly + PH][lx + PH] - sy[ly + PH][lx + PH]);
    if (px < W && py < H && with_ssim) {
        float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float w = kWin[t];
            mu1 += w * hz[0][ly + t][lx]; mu2 += w * hz[1][ly + t][lx]; exx += w * hz[2][ly + t][lx];
            eyy += w * hz[3][ly + t][lx]; exy += w * hz[4][ly + t][lx];
        }
        const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = exx - mu1s, s2 = eyy - mu2s, s12 = exy - mu12;
        const float A = mu1s + mu2s + SSIM_C1, B = s1 + s2 + SSIM_C2, Cc = 2.f * mu12 + SSIM_C1, D = 2.f * s12 + SSIM_C2;
        const float iAB = 1.f / (A * B);
        ss = Cc * D * iAB;
        const size_t p = (size_t)py * W + px + ch * plane;
        // d ssim / d(mu1, E[x^2], E[xy]) with y fixed
        m1[p] = ((2.f * mu2 * D - 2.f * mu2 * Cc) * A * B - Cc * D * (2.f * mu1 * B - 2.f * mu1 * A)) * iAB * iAB;
        m2[p] = -Cc * D * iAB / B;
        m3[p] = 2.f * Cc * iAB;
        l1 = fabsf(sx[ly + PH][lx + PH] - sy[ly + PH][lx + PH]);     // |(img - gt) * mask|
    }
    const float tl1 = block_sum_256(l1, red);
    const float tss = block_sum_256(ss, red);
    if (threadIdx.x == 0) {     // 64 accumulator slots per sum: 24k CTAs hammering one address serialise in L2
        const int slot = (blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z * 7) & (PHOTO_SLOTS - 1);
        atomicAdd(&sums[slot], (double)tl1);
        atomicAdd(&sums[PHOTO_SLOTS + slot], (double)tss);
    }
}

// lambda_dssim == 0: plain streaming |(image - gt) * mask| sum, no tiles, no halo
__global__ void __launch_bounds__(256)
photometric_l1_fwd_kernel(size_t n, size_t plane, const float* __restrict__ img, const float* __restrict__ gt,
                          const float* __restrict__ mask, double* __restrict__ sums) {
    __shared__ float red[8];
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float mk = mask ? mask[i % plane] : 1.f;
        acc += fabsf((img[i] - gt[i]) * mk);
    }
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(&sums[blockIdx.x & (PHOTO_SLOTS - 1)], (double)t);
}

// Vector (16-byte) forms of the two L1-only kernels for planes that are a multiple of 4 pixels: 4 independent
// float4 loads per array in flight per thread (the scalar forms above/below ran at 0.37 / 0.49 of the measured HBM
// peak: too few bytes in flight, and a 64-bit modulo per element for the mask index).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {        // read-once data: do not pollute L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
constexpr int L1_UNROLL = 4;
// ground truth either float [3,H,W] or 8-bit [3,H,W] (value / 255: frames as they are stored and shipped over PCIe;
// dequantised on the fly -- a quarter of the bytes and no separate conversion pass)
template <bool U8>
__device__ __forceinline__ float4 ldg_gt(const void* base, size_t i4) {
    if (U8) {
        const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(base) + i4);
        const float k = 1.0f / 255.0f;
        return make_float4(q.x * k, q.y * k, q.z * k, q.w * k);
    }
    return ldg_stream(reinterpret_cast<const float4*>(base) + i4);
}
// grid (blocks, 3 channels)
template <bool U8>
__global__ void __launch_bounds__(256)
photometric_l1_fwd_vec_kernel(size_t plane4, const float4* __restrict__ img, const void* __restrict__ gt,
                              const float4* __restrict__ mask, double* __restrict__ sums) {
    __shared__ float red[8];
    const float4* I = img + blockIdx.y * plane4;
    const size_t G0 = blockIdx.y * plane4;
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t base = (size_t)blockIdx.x * 256 + threadIdx.x; base < plane4; base += stride * L1_UNROLL) {
        float4 a[L1_UNROLL], b[L1_UNROLL], m[L1_UNROLL];
#pragma unroll
        for (int u = 0; u < L1_UNROLL; u++) {
            const size_t i = base + u * stride;
            const bool ok = i < plane4;
            a[u] = ok ? ldg_stream(I + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[u] = ok ? ldg_gt<U8>(gt, G0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = (ok && mask) ? mask[i] : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < L1_UNROLL; u++)
            acc += fabsf((a[u].x - b[u].x) * m[u].x) + fabsf((a[u].y - b[u].y) * m[u].y) +
                   fabsf((a[u].z - b[u].z) * m[u].z) + fabsf((a[u].w - b[u].w) * m[u].w);
    }
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(&sums[(blockIdx.x + 7 * blockIdx.y) & (PHOTO_SLOTS - 1)], (double)t);
}

template <bool U8>
__global__ void __launch_bounds__(256)
photometric_l1_bwd_vec_kernel(size_t plane4, const float4* __restrict__ img, const void* __restrict__ gt,
                              const float4* __restrict__ mask, float c_l1, const float* __restrict__ g_scalar,
                              float4* __restrict__ g_img) {
    if (g_scalar) c_l1 *= g_scalar[0];
    const float4* I = img + blockIdx.y * plane4;
    const size_t G0 = blockIdx.y * plane4;
    float4* O = g_img + blockIdx.y * plane4;
    const size_t stride = (size_t)gridDim.x * 256;
    auto sg = [&](float x, float y, float mk) {
        const float d = (x - y) * mk;
        return mk * c_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    };
    for (size_t base = (size_t)blockIdx.x * 256 + threadIdx.x; base < plane4; base += stride * L1_UNROLL) {
        float4 a[L1_UNROLL], b[L1_UNROLL], m[L1_UNROLL];
#pragma unroll
        for (int u = 0; u < L1_UNROLL; u++) {
            const size_t i = base + u * stride;
            const bool ok = i < plane4;
            a[u] = ok ? ldg_stream(I + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[u] = ok ? ldg_gt<U8>(gt, G0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = (ok && mask) ? mask[i] : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < L1_UNROLL; u++) {
            const size_t i = base + u * stride;
            if (i < plane4)
                O[i] = make_float4(sg(a[u].x, b[u].x, m[u].x), sg(a[u].y, b[u].y, m[u].y), sg(a[u].z, b[u].z, m[u].z),
                                   sg(a[u].w, b[u].w, m[u].w));
        }
    }
}

// dL/dimage = mask * [ c_l1 * sign((img-gt)*mask) + c_ss * (conv(m1) + 2 x conv(m2) + y conv(m3)) ]
__global__ void __launch_bounds__(256)
photometric_bwd_kernel(int W, int H, const float* __restrict__ img, const float* __restrict__ gt,
                       const float* __restrict__ mask, const float* __restrict__ m1, const float* __restrict__ m2,
                       const float* __restrict__ m3, float c_l1, float c_ss, const float* __restrict__ g_scalar,
                       float* __restrict__ g_img) {
    __shared__ float s[3][PI][PI + 1];
    if (g_scalar) {           // upstream dL/dloss stays on the device: no host round trip
        const float g = g_scalar[0];
        c_l1 *= g;
        c_ss *= g;
    }
    __shared__ float hz[3][PI][PT + 1];
    const int ch = blockIdx.z;
    const int ox = blockIdx.x * PT, oy = blockIdx.y * PT;
    const size_t plane = (size_t)W * H;
    if (m1 == nullptr) {                                        // L1 only
        const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
        const int px = ox + lx, py = oy + ly;
        if (px >= W || py >= H) return;
        const size_t pp = (size_t)py * W + px, p = pp + ch * plane;
        const float mk = mask ? mask[pp] : 1.f;
        const float d = (img[p] - gt[p]) * mk;
        g_img[p] = mk * c_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        return;
    }
    for (int k = threadIdx.x; k < PI * PI; k += 256) {
        const int ly = k / PI, lx = k - ly * PI;
        const int gx_ = ox + lx - PH, gy_ = oy + ly - PH;
        float a = 0.f, b = 0.f, c = 0.f;
        if (gx_ >= 0 && gx_ < W && gy_ >= 0 && gy_ < H) {
            const size_t p = (size_t)gy_ * W + gx_ + ch * plane;
            a = m1[p]; b = m2[p]; c = m3[p];
        }
        s[0][ly][lx] = a; s[1][ly][lx] = b; s[2][ly][lx] = c;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < PI * PT; k += 256) {
        const int ly = k / PT, lx = k - ly * PT;
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float w = kWin[t];
            a += w * s[0][ly][lx + t]; b += w * s[1][ly][lx + t]; c += w * s[2][ly][lx + t];
        }
        hz[0][ly][lx] = a; hz[1][ly][lx] = b; hz[2][ly][lx] = c;
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int px = ox + lx, py = oy + ly;
    if (px >= W || py >= H) return;
    float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
    for (int t = 0; t < 11; t++) {
        const float w = kWin[t];
        a += w * hz[0][ly + t][lx]; b += w * hz[1][ly + t][lx]; c += w * hz[2][ly + t][lx];
    }
    const size_t pp = (size_t)py * W + px, p = pp + ch * plane;
    const float mk = mask ? mask[pp] : 1.f;
    const float x = img[p] * mk, y = gt[p] * mk;
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    g_img[p] = mk * (c_l1 * sgn + c_ss * (a + 2.f * x * b + y * c));
}

// one warp folds the 2 x 64 accumulator slots into the three scalars the host mirror returns (replaces ~10 tiny torch
// kernels: slice/sum/divide/cast/combine): out = (total, l1, ssim), total = l1 (1 - lambda) + 1 - ssim lambda
__global__ void __launch_bounds__(32)
photometric_finalize_kernel(const double* __restrict__ sums, double inv_n, float lambda_dssim, float* __restrict__ out3) {
    double a = sums[threadIdx.x] + sums[threadIdx.x + 32];
    double b = sums[PHOTO_SLOTS + threadIdx.x] + sums[PHOTO_SLOTS + threadIdx.x + 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    if (threadIdx.x == 0) {
        const float l1 = (float)(a * inv_n), ss = (float)(b * inv_n);
        out3[0] = l1 * (1.0f - lambda_dssim) + (1.0f - ss * lambda_dssim);
        out3[1] = l1;
        out3[2] = ss;
    }
}
int launch_photometric_finalize(const double* sums, double inv_n, float lambda_dssim, float* out3, cudaStream_t s) {
    photometric_finalize_kernel<<<1, 32, 0, s>>>(sums, inv_n, lambda_dssim, out3);
    return 1;
}

// L1-only loss with an 8-bit ground truth (vector kernels only; the caller guarantees plane % 4 == 0 and alignment)
int launch_photometric_l1_u8(int W, int H, const float* img, const uint8_t* gt8, const float* mask, double* sums,
                             float c_l1, const float* g_scalar, float* g_img, cudaStream_t s) {
    const size_t plane = (size_t)W * H;
    if (g_img == nullptr)
        photometric_l1_fwd_vec_kernel<true><<<dim3(148 * 2, 3), 256, 0, s>>>(plane / 4, (const float4*)img, gt8, (const float4*)mask, sums);
    else
        photometric_l1_bwd_vec_kernel<true><<<dim3(148 * 2, 3), 256, 0, s>>>(plane / 4, (const float4*)img, gt8, (const float4*)mask,
                                                                             c_l1, g_scalar, (float4*)g_img);
    return 1;
}

int launch_photometric_fwd(int W, int H, const float* img, const float* gt, const float* mask, float* m1, float* m2,
                           float* m3, double* sums, cudaStream_t s) {
    if (W <= 0 || H <= 0) return 0;
    if (m1 == nullptr) {
        const size_t plane = (size_t)W * H;
        const bool vec = (plane % 4 == 0) && !(((uintptr_t)img | (uintptr_t)gt | (uintptr_t)mask) & 15);
        if (vec) {
            photometric_l1_fwd_vec_kernel<false><<<dim3(148 * 2, 3), 256, 0, s>>>(plane / 4, (const float4*)img, gt,
                                                                                  (const float4*)mask, sums);
        } else {
            photometric_l1_fwd_kernel<<<148 * 8, 256, 0, s>>>(3 * plane, plane, img, gt, mask, sums);
        }
        return 1;
    }
    dim3 grid((W + PT - 1) / PT, (H + PT - 1) / PT, 3);
    photometric_fwd_kernel<<<grid, 256, 0, s>>>(W, H, img, gt, mask, m1, m2, m3, sums);
    return 1;
}
int launch_photometric_bwd(int W, int H, const float* img, const float* gt, const float* mask, const float* m1,
                           const float* m2, const float* m3, float c_l1, float c_ss, const float* g_scalar, float* g_img,
                           cudaStream_t s) {
    if (W <= 0 || H <= 0) return 0;
    const size_t plane = (size_t)W * H;
    if (m1 == nullptr && plane % 4 == 0 && !(((uintptr_t)img | (uintptr_t)gt | (uintptr_t)mask | (uintptr_t)g_img) & 15)) {
        photometric_l1_bwd_vec_kernel<false><<<dim3(148 * 2, 3), 256, 0, s>>>(plane / 4, (const float4*)img, gt,
                                                                              (const float4*)mask, c_l1, g_scalar, (float4*)g_img);
        return 1;
    }
    dim3 grid((W + PT - 1) / PT, (H + PT - 1) / PT, 3);
    photometric_bwd_kernel<<<grid, 256, 0, s>>>(W, H, img, gt, mask, m1, m2, m3, c_l1, c_ss, g_scalar, g_img);
    return 1;
}

}  // namespace gg
