// allreduce.cu -- in-switch (NVLS) average of the flat gradient bucket across the ranks of one NVSwitch domain.
//
// The only exchange on the path is the per-step sum of the per-view parameter gradients: one flat fp32
// bucket (236 B x N Gaussians = 70.8 MB at cfg2) that every rank's preprocess-backward kernel has just written in
// place (zero-copy grad sinks).  The bucket lives in symmetric memory that is also mapped behind ONE multicast address
// (torch.distributed._symmetric_memory supplies allocation, rendezvous and the signal pads -- plumbing only), so the
// reduction is a single kernel per rank with no staging buffers and no NCCL ring:
//     rank r owns the r-th 1/world slice;  for every 16 bytes of it:
//         v = multimem.ld_reduce.add.v4.f32 [mc + i]     the switch pulls the 16 B from all ranks and adds them
//         v *= 1/world                                    (the averaging is fused here)
//         multimem.st.v4.f32 [mc + i], v                  the switch writes the result into every rank's bucket
// Per GPU ~ bytes/world in + bytes out per phase instead of 2 (world-1)/world x bytes through a ring, and the adds
// happen in the switch.  Cross-rank ordering uses the symmetric-memory signal pads: every block exchanges one flag
// with the same block of every peer on entry (all ranks' gradients are complete) and on exit (all results landed).
// Two instances may be in flight on different streams (geometry block / SH block) -- they use disjoint pad channels.
// (Measured alternative, 8 x B200: one-warp barrier kernels around a wide -- 64 to 256 block -- data kernel.  The
//  exchange got SLOWER, 0.218-0.237 ms vs 0.2045 ms for 70.8 MB, and the step 0.82 vs 0.72 ms: the in-switch reduction
//  saturates at ~350 GB/s algorithmic with ~45 blocks in flight, and a wider grid only takes SM slots from the compute
//  kernels the deferred block overlaps with.)
#include <cstdlib>
#include "common.cuh"

namespace gg {

constexpr int AR_THREADS = 512;

__device__ __forceinline__ uint32_t cas_sys_relaxed(uint32_t* a, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.relaxed.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_sys_release(uint32_t* a, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_sys_acquire(uint32_t* a, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(cmp), "r"(val) : "memory");
    return old;
}

// One flag per (block, peer): thread t < world raises the flag in peer t's pad and waits for peer t's flag in ours.
// Flags toggle 0 -> 1 (raise) -> 0 (consume), so the pads are ready for the next call without a reset.
template <bool RELEASE_ACQUIRE>
__device__ __forceinline__ void cross_rank_barrier(uint32_t* const* __restrict__ pads, int rank, int world, int slot0) {
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        uint32_t* theirs = pads[peer] + slot0 + (int)blockIdx.x * world + rank;
        uint32_t* mine = pads[rank] + slot0 + (int)blockIdx.x * world + peer;
        if (RELEASE_ACQUIRE) {
            while (cas_sys_release(theirs, 0u, 1u) != 0u) {}
            while (cas_sys_acquire(mine, 1u, 0u) != 1u) {}
        } else {
            while (cas_sys_relaxed(theirs, 0u, 1u) != 0u) {}
            while (cas_sys_relaxed(mine, 1u, 0u) != 1u) {}
        }
    }
}

template <int AR_UNROLL>
__global__ void __launch_bounds__(AR_THREADS)
nvls_allreduce_kernel(float* __restrict__ mc, uint32_t* const* __restrict__ pads, int rank, int world, int slot0,
                      int64_t n_vec4, float scale) {
    cross_rank_barrier<false>(pads, rank, world, slot0);      // every rank's producer kernels have finished
    __syncthreads();
    const int64_t per = (n_vec4 + world - 1) / world;
    const int64_t beg = min((int64_t)rank * per, n_vec4), end = min(beg + per, n_vec4);
    // AR_UNROLL independent 16-byte reductions in flight per thread: one round trip through the switch is ~2-3 us, so
    // bytes in flight (blocks x 512 x 16 B x AR_UNROLL) set the bandwidth, not the instruction rate
    const int64_t stride = (int64_t)gridDim.x * AR_THREADS;
    for (int64_t i0 = beg + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x; i0 < end; i0 += stride * AR_UNROLL) {
        float4 v[AR_UNROLL];
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
            const int64_t i = i0 + u * stride;
            if (i < end)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(mc + 4 * i)
                             : "memory");
        }
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
            const int64_t i = i0 + u * stride;
            if (i < end)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i),
                             "f"(v[u].x * scale), "f"(v[u].y * scale), "f"(v[u].z * scale), "f"(v[u].w * scale)
                             : "memory");
        }
    }
    __syncthreads();
    cross_rank_barrier<true>(pads, rank, world, slot0);       // every rank's slice has landed everywhere
}

// ---- device-side gate (CUDA-graph friendly cross-stream dependency) -------------------------------------------------
// gate[0] = G: how many gated colour stages have started; gate[1] = X: how many deferred exchanges have completed;
// gate[2] = timeout flag.  The k-th gated forward may read the SH coefficients once exchanges 1..k-1 are done (X >= k-1).
// A stream-event wait cannot cross a graph boundary; one spinning WARP can (a whole grid parked on the flag could starve
// the exchange kernel of SM slots, one warp cannot).  The spin gives up after ~1 s and raises the timeout flag.
__global__ void __launch_bounds__(32) gate_wait_kernel(uint32_t* __restrict__ gate) {
    if (threadIdx.x == 0) {
        const uint32_t k = atomicAdd(&gate[0], 1u) + 1u;
        uint32_t spins = 0;
        while (*reinterpret_cast<volatile uint32_t*>(&gate[1]) + 1u < k) {
            __nanosleep(200);
            if (++spins > (1u << 22)) { atomicOr(&gate[2], 1u); break; }
        }
        __threadfence();
    }
}
__global__ void gate_signal_kernel(uint32_t* __restrict__ gate) {
    __threadfence();
    atomicAdd(&gate[1], 1u);
}
int launch_gate_wait(uint32_t* gate, cudaStream_t s) {
    gate_wait_kernel<<<1, 32, 0, s>>>(gate);
    return 1;
}
int launch_gate_signal(uint32_t* gate, cudaStream_t s) {
    gate_signal_kernel<<<1, 1, 0, s>>>(gate);
    return 1;
}

int launch_nvls_allreduce(float* mc, uint32_t* const* pads, int rank, int world, int slot0, int64_t n_vec4, float scale,
                          int blocks, cudaStream_t s) {
    if (n_vec4 <= 0) return 0;
    // independent 16-byte reductions in flight per thread: 16 by default (8 x B200, cfg2 step: 0.683 ms vs 0.696 ms with
    // 8; exchange alone 0.275 vs 0.286 ms); GG_AR_UNROLL=8 selects the smaller variant
    static const int unroll = []() { const char* e = getenv("GG_AR_UNROLL"); return (e && atoi(e) == 8) ? 8 : 16; }();
    if (unroll == 16) nvls_allreduce_kernel<16><<<blocks, AR_THREADS, 0, s>>>(mc, pads, rank, world, slot0, n_vec4, scale);
    else nvls_allreduce_kernel<8><<<blocks, AR_THREADS, 0, s>>>(mc, pads, rank, world, slot0, n_vec4, scale);
    return 1;
}

}  // namespace gg
