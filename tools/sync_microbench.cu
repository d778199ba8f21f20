// micro-benchmarks of the per-iteration synchronisation primitives an FPS kernel can use
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define FULL 0xffffffffu
constexpr int ITERS = 20000;

__global__ void k_cluster_sync(int* out) {
    cg::cluster_group cl = cg::this_cluster();
    int acc = 0;
    for (int j = 0; j < ITERS; j++) { cl.sync(); acc += j; }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = acc;
}
__global__ void k_syncthreads(int* out) {
    int acc = 0;
    for (int j = 0; j < ITERS; j++) { __syncthreads(); acc += j; }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = acc;
}
__global__ void k_redux_chain(int* out) {  // 2 dependent REDUX + smem + syncthreads + 2 REDUX (CTA argmax)
    __shared__ unsigned sb[32]; __shared__ int si[32];
    unsigned v = threadIdx.x * 2654435761u; int id = threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = 0; j < ITERS; j++) {
        unsigned m = __reduce_max_sync(FULL, v);
        int i = __reduce_min_sync(FULL, v == m ? id : 0x7fffffff);
        if (lane == 0) { sb[w] = m; si[w] = i; }
        __syncthreads();
        unsigned b = lane < nw ? sb[lane] : 0u; int bi = lane < nw ? si[lane] : 0x7fffffff;
        unsigned m2 = __reduce_max_sync(FULL, b);
        int i2 = __reduce_min_sync(FULL, b == m2 ? bi : 0x7fffffff);
        v = v * 1664525u + m2 + i2;
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = v;
}

// all-to-all exchange of 32-byte messages through st.async + mbarrier (double buffered)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_mbar_exchange(int* out, int wait_all) {
    cg::cluster_group cl = cg::this_cluster();
    const int C = cl.num_blocks(), rank = cl.block_rank();
    __shared__ __align__(16) unsigned slot[2][16][8];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int p = 0; p < 2; p++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[p])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) for (int p = 0; p < 2; p++)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar[p])), "r"(C * 32) : "memory");
    cl.sync();
    unsigned v = rank * 977 + 1;
    for (int j = 0; j < ITERS; j++) {
        const int par = j & 1;
        if (warp == 0 && lane < C) {
            unsigned dst, dbar;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(&slot[par][rank][0])), "r"(lane));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dbar) : "r"(smem_u32(&mbar[par])), "r"(lane));
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                         :: "r"(dst), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(dbar) : "memory");
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                         :: "r"(dst + 16), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(dbar) : "memory");
        }
        const unsigned phase = (j >> 1) & 1;
        if (wait_all || warp == 0) {
            unsigned done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&mbar[par])), "r"(phase) : "memory");
            }
        }
        if (!wait_all) __syncthreads();
        unsigned got = lane < C ? slot[par][lane][0] : 0u;
        unsigned m = __reduce_max_sync(FULL, got);
        v = v * 1664525u + m;
        if (!wait_all) __syncthreads();
        else __syncthreads();   // everyone must have read the slots before re-arming / next-next write
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar[par])), "r"(C * 32) : "memory");
    }
    cl.sync();
    if (tid == 0 && blockIdx.x == 0) out[0] = v;
}

template <typename K, typename... A>
float run(K kern, int C, int threads, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (C > 8) cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, kern, args...);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(err)); return -1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best * 1e6f / ITERS;  // ns per iteration
}

int main() {
    int* out; cudaMalloc(&out, 64);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock %d kHz\n", clk);
    for (int th : {1024, 256, 64}) {
        printf("syncthreads           threads=%4d : %7.1f ns/iter\n", th, run(k_syncthreads, 1, th, out));
        printf("redux chain (CTA amax) threads=%4d : %7.1f ns/iter\n", th, run(k_redux_chain, 1, th, out));
    }
    for (int C : {1, 2, 4, 8, 16}) for (int th : {1024, 256, 64}) {
        printf("cluster.sync  C=%2d threads=%4d : %7.1f ns/iter\n", C, th, run(k_cluster_sync, C, th, out));
    }
    for (int C : {2, 4, 8, 16}) for (int th : {1024, 256, 64}) for (int wa : {1, 0}) {
        printf("mbar exchange C=%2d threads=%4d wait_all=%d : %7.1f ns/iter\n", C, th, wa, run(k_mbar_exchange, C, th, out, wa));
    }
    return 0;
}
