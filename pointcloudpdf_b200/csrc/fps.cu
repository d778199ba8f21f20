// fps.cu -- farthest point sampling on a thread-block cluster (sm_100a).
//
// Replaces farthest_point_sampling_cuda_kernel<bs> (libs/pointops/src/sampling/
// sampling_cuda_kernel.cu:15-129): ONE block per scene that re-reads xyz and tmp from global
// memory every iteration and reduces through ten __syncthreads.  FPS is m-1 strictly
// dependent iterations, so the only lever is the latency of one iteration:
//   * one CLUSTER of up to 16 CTAs per scene (16 SMs instead of 1); the scene's coordinates
//     and running min-distances stay in REGISTERS for the whole kernel (P points per thread),
//     so an iteration touches no global memory at all;
//   * argmax = two REDUX instructions per level (distances are >= 0, so their bit patterns
//     order like unsigned ints; ties resolved to the lowest index by a second REDUX.min),
//     one __syncthreads per CTA, and one DSMEM all-to-all + cluster barrier per iteration
//     carrying {value, index, x, y, z} of each CTA's winner so nobody reloads coordinates.
// Scenes too large for registers (n > C*1024*8) run the same algorithm with the points
// streamed from global/L2 (fps_stream_kernel), using the caller's tmp buffer.
//
// Semantics: idx[s_m] = s_n; tmp = 1e10; tmp[i] = min(tmp[i], d2(i, last));
// next = argmax tmp, LOWEST index among maxima (north_star tie rule; the reference's winner
// depends on its block size, sampling_cuda_kernel.cu:5-10,49-59).  d2 = pob::d2_ref.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace pob {

constexpr int FPS_THREADS = 1024;
constexpr int FPS_MAX_CLUSTER = 16;

struct __align__(16) FpsMsg {  // one CTA's winner, 32 bytes
    unsigned bits;             // float bits of the winning tmp value
    int idx;                   // global point index
    float x, y, z;
    int pad0, pad1, pad2;
};

__device__ __forceinline__ void warp_argmax(unsigned bits, int gi, unsigned& wbits, int& wi) {
    wbits = __reduce_max_sync(FULL, bits);
    wi = __reduce_min_sync(FULL, bits == wbits ? gi : INT_MAX);
}

// One cluster (C CTAs) per scene; P points per thread held in registers.
template <int P>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                   int* __restrict__ idx) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];

    __shared__ unsigned s_bits[32];
    __shared__ int s_idx[32];
    __shared__ FpsMsg s_msg[2][FPS_MAX_CLUSTER];

    // every CTA of the cluster takes the same branch: no barrier is skipped by part of it
    if (e_m <= s_m || e_n <= s_n) return;

    const int stride = C * FPS_THREADS;
    const int first = s_n + rank * FPS_THREADS + tid;
    float px[P], py[P], pz[P], pt[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        const int i = first + p * stride;
        if (i < e_n) {
            px[p] = __ldg(xyz + (int64_t)i * 3);
            py[p] = __ldg(xyz + (int64_t)i * 3 + 1);
            pz[p] = __ldg(xyz + (int64_t)i * 3 + 2);
            pt[p] = PLACEHOLDER_D2;
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            pt[p] = -1.f;  // never a maximum: fminf keeps it at -1
        }
    }
    float ox = __ldg(xyz + (int64_t)s_n * 3), oy = __ldg(xyz + (int64_t)s_n * 3 + 1), oz = __ldg(xyz + (int64_t)s_n * 3 + 2);
    if (rank == 0 && tid == 0) idx[s_m] = s_n;

    for (int j = s_m + 1; j < e_m; j++) {
        // ---- update the running minimum, thread-local argmax (strict '>' keeps the lowest index) ----
        float best = 0.f;
        int bp = -1;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const float d = d2_ref(px[p], py[p], pz[p], ox, oy, oz);
            const float t = fminf(d, pt[p]);
            pt[p] = t;
            if (t > best || (bp < 0 && t >= 0.f)) { best = t; bp = p; }
        }
        // a thread with no valid point offers (0, INT_MAX): it loses every tie against a real point
        const int gi = bp >= 0 ? first + bp * stride : INT_MAX;
        float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
        for (int p = 0; p < P; p++) if (p == bp) { bx = px[p]; by = py[p]; bz = pz[p]; }

        // ---- CTA argmax: REDUX per warp, one barrier, REDUX over the 32 warp winners ----
        unsigned wbits; int wi;
        warp_argmax(__float_as_uint(best), gi, wbits, wi);
        if (lane == 0) { s_bits[warp] = wbits; s_idx[warp] = wi; }
        __syncthreads();
        unsigned cbits; int ci;
        warp_argmax(s_bits[lane], s_idx[lane], cbits, ci);

        // ---- the warp that owns the CTA winner publishes it (with coordinates) to every CTA ----
        const int par = j & 1;
        const unsigned own = __ballot_sync(FULL, gi == ci && ci != INT_MAX);
        if (own) {
            const int ol = __ffs(own) - 1;
            FpsMsg msg;
            msg.bits = cbits; msg.idx = ci;
            msg.x = __shfl_sync(FULL, bx, ol); msg.y = __shfl_sync(FULL, by, ol); msg.z = __shfl_sync(FULL, bz, ol);
            msg.pad0 = msg.pad1 = msg.pad2 = 0;
            if (lane < C) {
                FpsMsg* dst = cluster.map_shared_rank(&s_msg[par][rank], lane);
                *dst = msg;
            }
        } else if (ci == INT_MAX && warp == 0 && lane < C) {
            // CTA without any valid point (rank beyond the scene): publish a losing entry
            FpsMsg msg = {0u, INT_MAX, 0.f, 0.f, 0.f, 0, 0, 0};
            *cluster.map_shared_rank(&s_msg[par][rank], lane) = msg;
        }
        cluster.sync();  // release/acquire: all C messages of this iteration are visible

        // ---- every thread picks the cluster winner from its own CTA's copy ----
        const FpsMsg mine = s_msg[par][lane < C ? lane : 0];
        unsigned gbits; int gidx;
        warp_argmax(lane < C ? mine.bits : 0u, lane < C ? mine.idx : INT_MAX, gbits, gidx);
        const unsigned who = __ballot_sync(FULL, lane < C && mine.idx == gidx);
        const int wl = __ffs(who) - 1;
        ox = __shfl_sync(FULL, mine.x, wl); oy = __shfl_sync(FULL, mine.y, wl); oz = __shfl_sync(FULL, mine.z, wl);
        if (rank == 0 && tid == 0) idx[j] = gidx;
        // s_bits/s_idx are rewritten only after the next iteration's compute; the cluster barrier
        // above already orders this iteration's reads before those writes.
    }
}

// Same algorithm with the points left in global memory (L2-resident for any realistic scene):
// the fallback for scenes that do not fit the register-resident kernel.
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_stream_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                  float* __restrict__ tmp, int* __restrict__ idx) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];
    __shared__ unsigned s_bits[32];
    __shared__ int s_idx[32];
    __shared__ FpsMsg s_msg[2][FPS_MAX_CLUSTER];
    if (e_m <= s_m || e_n <= s_n) return;

    const int stride = C * FPS_THREADS;
    const int first = s_n + rank * FPS_THREADS + tid;
    for (int i = first; i < e_n; i += stride) tmp[i] = PLACEHOLDER_D2;
    float ox = __ldg(xyz + (int64_t)s_n * 3), oy = __ldg(xyz + (int64_t)s_n * 3 + 1), oz = __ldg(xyz + (int64_t)s_n * 3 + 2);
    if (rank == 0 && tid == 0) idx[s_m] = s_n;

    for (int j = s_m + 1; j < e_m; j++) {
        float best = 0.f;
        int gi = INT_MAX;
        for (int i = first; i < e_n; i += stride) {
            const float d = d2_ref(__ldg(xyz + (int64_t)i * 3), __ldg(xyz + (int64_t)i * 3 + 1),
                                   __ldg(xyz + (int64_t)i * 3 + 2), ox, oy, oz);
            const float t = fminf(d, tmp[i]);
            tmp[i] = t;
            if (t > best || gi == INT_MAX) { best = t; gi = i; }
        }
        unsigned wbits; int wi;
        warp_argmax(__float_as_uint(best), gi, wbits, wi);
        if (lane == 0) { s_bits[warp] = wbits; s_idx[warp] = wi; }
        __syncthreads();
        unsigned cbits; int ci;
        warp_argmax(s_bits[lane], s_idx[lane], cbits, ci);
        const int par = j & 1;
        if (warp == 0 && lane < C) {
            FpsMsg msg = {cbits, ci, 0.f, 0.f, 0.f, 0, 0, 0};
            if (ci != INT_MAX) {
                msg.x = __ldg(xyz + (int64_t)ci * 3); msg.y = __ldg(xyz + (int64_t)ci * 3 + 1); msg.z = __ldg(xyz + (int64_t)ci * 3 + 2);
            }
            *cluster.map_shared_rank(&s_msg[par][rank], lane) = msg;
        }
        cluster.sync();
        const FpsMsg mine = s_msg[par][lane < C ? lane : 0];
        unsigned gbits; int gidx;
        warp_argmax(lane < C ? mine.bits : 0u, lane < C ? mine.idx : INT_MAX, gbits, gidx);
        const unsigned who = __ballot_sync(FULL, lane < C && mine.idx == gidx);
        const int wl = __ffs(who) - 1;
        ox = __shfl_sync(FULL, mine.x, wl); oy = __shfl_sync(FULL, mine.y, wl); oz = __shfl_sync(FULL, mine.z, wl);
        if (rank == 0 && tid == 0) idx[j] = gidx;
    }
}

template <typename K>
static int launch_cluster(K kernel, int b, int C, cudaStream_t stream, const float* xyz, const int* offset,
                          const int* new_offset, float* tmp, int* idx, bool pass_tmp) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * C));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (C > 8) POB_CHECK(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    void* args_tmp[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&tmp, (void*)&idx};
    void* args[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&idx};
    POB_CHECK(cudaLaunchKernelExC(&cfg, (const void*)kernel, pass_tmp ? args_tmp : args));
    pob_count_launches(1);
    return 0;
}

}  // namespace pob

using namespace pob;

// farthest_point_sampling_cuda_launcher(b, n, xyz, offset, new_offset, tmp, idx)
// (sampling_cuda_kernel.h:13) + stream.  n_max = largest scene (the reference's `n`); tmp (n
// floats) is only touched when a scene exceeds the register-resident capacity, and needs no
// initialisation.  cluster_hint: 0 = choose, else force 1/2/4/8/16 CTAs per scene.
POB_API int pob_farthest_point_sampling(int b, int64_t n_max, const float* xyz, const int* offset,
                                        const int* new_offset, float* tmp, int* idx, int cluster_hint,
                                        cudaStream_t stream) {
    if (b < 1 || n_max < 0 || !offset || !new_offset || !idx) return POB_ERR_BAD_ARG;
    if (n_max == 0) return 0;
    if (!xyz) return POB_ERR_BAD_ARG;
    int C = cluster_hint;
    if (C != 1 && C != 2 && C != 4 && C != 8 && C != 16) {
        // per-iteration cost ~ 8 warps/SMSP * 9 instr * P  vs  ~400 cycles of cluster exchange
        C = n_max <= 6 * 1024 ? 1 : (n_max <= 24 * 1024 ? 8 : 16);
    }
    const int64_t per_thread = ceil_div(n_max, (int64_t)C * FPS_THREADS);
    if (per_thread > 8) {
        C = 16;
        if (ceil_div(n_max, (int64_t)C * FPS_THREADS) > 8) {
            if (!tmp) return POB_ERR_BAD_ARG;
            return launch_cluster(fps_stream_kernel, b, C, stream, xyz, offset, new_offset, tmp, idx, true);
        }
    }
    const int64_t P = ceil_div(n_max, (int64_t)C * FPS_THREADS);
    if (P <= 1) return launch_cluster(fps_cluster_kernel<1>, b, C, stream, xyz, offset, new_offset, tmp, idx, false);
    if (P <= 2) return launch_cluster(fps_cluster_kernel<2>, b, C, stream, xyz, offset, new_offset, tmp, idx, false);
    if (P <= 4) return launch_cluster(fps_cluster_kernel<4>, b, C, stream, xyz, offset, new_offset, tmp, idx, false);
    if (P <= 6) return launch_cluster(fps_cluster_kernel<6>, b, C, stream, xyz, offset, new_offset, tmp, idx, false);
    return launch_cluster(fps_cluster_kernel<8>, b, C, stream, xyz, offset, new_offset, tmp, idx, false);
}
