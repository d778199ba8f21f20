// common.cuh -- shared device/host helpers for the pointops_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <limits.h>

#define POB_API extern "C" __attribute__((visibility("default")))

// Entry points never synchronise and never throw: they return the launch status.
#define POB_RETURN_LAST_ERROR() return (int)cudaPeekAtLastError()
#define POB_CHECK(expr)                        \
    do {                                       \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

// status codes above the cudaError_t range for argument errors
enum { POB_ERR_BAD_ARG = 10001, POB_ERR_WORKSPACE = 10002, POB_ERR_UNSUPPORTED = 10003 };

// process-wide count of kernels launched by this library (bench.py's gpu_launches evidence)
extern "C" void pob_count_launches(int n);

namespace pob {

constexpr unsigned FULL = 0xffffffffu;
constexpr float PLACEHOLDER_D2 = 1e10f;  // knn_query_cuda_kernel.cu:85, functions/sampling.py:19

// The one d2 expression of the path: what nvcc 12.9 makes of
//   (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)
// in the reference (knn_query_cuda_kernel.cu:92, sampling_cuda_kernel.cu:54):
//   sub, sub, mul(dy), fma(dx), sub, fma(dz).  Spelled with intrinsics so that no
// compiler version or flag can re-associate it; the oracle uses the same chain.
__device__ __forceinline__ float d2_ref(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// first i in [0,b) with q < off[i]  (knn_query_cuda_kernel.cu:45-56, by bisection)
__device__ __forceinline__ int segment_of(int64_t q, const int* __restrict__ off, int b) {
    int lo = 0, hi = b - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (q < (int64_t)__ldg(off + mid)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// grid for a grid-stride kernel: enough CTAs for `work` items, capped at waves*SMs*ctas_per_sm
inline unsigned grid_for(int64_t work_items, int threads, int ctas_per_sm, int waves = 4) {
    int64_t need = ceil_div(work_items, threads);
    int64_t cap = (int64_t)sm_count() * ctas_per_sm * waves;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace pob
