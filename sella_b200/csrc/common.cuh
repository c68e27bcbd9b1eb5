// Shared device helpers for the sella_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define SB_WARP 32

// status bits written per system by the kernels (same meaning as the reference's
// error conventions): surfaced as exceptions at the Python boundary.
#define SB_ST_OK 0
#define SB_ST_MGS_MAXITER 1        // math.pyx:132-133  (-2)
#define SB_ST_TR_NOCONV 2          // restricted_step.py:116-117
#define SB_ST_EIGH_NOCONV 4        // QL sweep limit hit
#define SB_ST_DAVIDSON_CAP 8       // subspace reached the compiled capacity
#define SB_ST_SINGULAR 16          // tiny dense solve hit a zero pivot
#define SB_ST_DAVIDSON_STALL 32    // eigensolvers.py:99-109 random-restart branch
#define SB_ST_CAPACITY 64          // a vector block is too small for the requested operation

__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier / TMA
__device__ __forceinline__ void sb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "SB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra SB_DONE;\n"
        "bra SB_WAIT;\n"
        "SB_DONE:\n"
        "}\n" ::"r"(sb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine, UBLKCP in SASS);
// bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void sb_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                               uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            sb_smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(sb_smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double sb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double sb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result broadcast to every thread.  `scratch` >= 33 doubles of
// shared memory.  Deterministic (fixed tree).  Contains __syncthreads().
__device__ __forceinline__ double sb_block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = sb_warp_sum(v);
    __syncthreads();  // protect scratch from the previous use
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nwarp) ? scratch[lane] : 0.0;
        t = sb_warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// two sums at once (halves the barrier count on the BLAS-1 chains)
__device__ __forceinline__ void sb_block_sum2(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    a = sb_warp_sum(a);
    b = sb_warp_sum(b);
    __syncthreads();
    if (lane == 0) { scratch[warp] = a; scratch[34 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nwarp) ? scratch[lane] : 0.0;
        double u = (lane < nwarp) ? scratch[34 + lane] : 0.0;
        t = sb_warp_sum(t);
        u = sb_warp_sum(u);
        if (lane == 0) { scratch[32] = t; scratch[33] = u; }
    }
    __syncthreads();
    a = scratch[32];
    b = scratch[33];
}
#define SB_SCRATCH_DOUBLES 68

// number of kernels this library has launched (reported by bench.py as gpu_launches)
#include <atomic>
extern std::atomic<long long> sb_launch_counter;      // host threads may launch concurrently
#define SB_COUNT(k) (sb_launch_counter.fetch_add((k), std::memory_order_relaxed))

static inline int sb_check(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }
#define SB_LAUNCH_CHECK() sb_check(cudaGetLastError())
