// Shared device helpers for the cgp_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define CGP_TILE 128          // square tile edge of every blocked algorithm (padding granularity)
#ifndef CGP_BK
#define CGP_BK 32             // k-chunk staged per pipeline stage (fp64)
#endif
#ifndef CGP_STAGES
#define CGP_STAGES 3          // cp.async ring depth (3 x 64 KB)
#endif
#define CGP_PANEL 4           // block columns per panel of the hybrid left/right-looking schedules
#define CGP_GEMM_THREADS 256  // 8 warps: 2 (M) x 4 (N), warp tile 64 x 32
#define CGP_MAX_D 16          // max input dimension
#define CGP_MAX_K 8           // max k of k-NN

#define CGP_CHECK(call)                                   \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) return (int)e_;            \
    } while (0)

struct cgp_handle {
    int device;
    int sm_count;
    void* pinned;        // pinned staging for launch metadata
    size_t pinned_bytes;
    size_t pinned_used;  // bump pointer, reset per API call
};

// Per (cluster, theta) pair / per cluster metadata, one entry per matrix in a batched launch.
struct MatMeta {
    long long a_off;  // element offset of the L / K matrix in its buffer
    long long u_off;  // element offset of the U (= L^-T) / W matrix in its buffer
    long long d_off;  // element offset of the [npad,128] strip of inverted diagonal blocks
    int n;            // true sample count
    int npad;         // padded to CGP_TILE
    int xoff;         // first row of this cluster in the cluster-sorted X / y
    int theta_idx;    // row of theta[., P] to use
    int out_idx;      // row of lml / grad / info outputs
    int pad_;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// fp64 tensor-core MMA (SASS: DMMA.8x8x4).  A 8x4 row, B 4x8 col, C 8x8.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
