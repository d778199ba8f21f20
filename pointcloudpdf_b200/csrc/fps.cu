// fps.cu -- farthest point sampling on a thread-block cluster (sm_100a).
//
// Replaces farthest_point_sampling_cuda_kernel<bs> (libs/pointops/src/sampling/
// sampling_cuda_kernel.cu:15-129): ONE block per scene that re-reads xyz and tmp from global
// memory every iteration and reduces through ten __syncthreads.  FPS is m-1 strictly
// dependent iterations, so the only lever is the latency of one iteration.  Measured on B200
// (profiles/r01_sync_microbench.txt): __syncthreads 15 ns (256 thr), a CTA-wide argmax by REDUX
// ~60-75 ns, any cluster-wide exchange (cluster.sync or st.async+mbarrier) ~190-260 ns.  Design:
//   * one CLUSTER of up to 16 CTAs per scene; the scene's coordinates and running min-distances
//     stay in REGISTERS for the whole kernel (P points per thread, 2 warps per scheduler), so an
//     iteration touches no global memory at all;
//   * EXACT pruning: when the caller passes the kNN search grid of the same cloud, points are
//     taken in cell order, so each warp owns a spatially compact run of 32*P points with a
//     bounding box; a warp whose box is farther from the new sample than its current maximum
//     (conservatively rounded) cannot change and skips the update -- after a few hundred
//     samples almost every warp skips, and the iteration is pure synchronisation latency;
//   * argmax = two REDUX instructions per level (distances are >= 0, so their bit patterns order
//     like unsigned ints; ties resolved to the lowest ORIGINAL index by a REDUX.min), one
//     __syncthreads per CTA, and one DSMEM all-to-all + cluster barrier per iteration carrying
//     {value, index, x, y, z} of each CTA's winner so nobody reloads coordinates.
// Scenes too large for registers (n > 131072) run the same algorithm with the points streamed
// from global/L2 (fps_stream_kernel), using the caller's tmp buffer.
//
// Semantics: idx[s_m] = s_n; tmp = 1e10; tmp[i] = min(tmp[i], d2(i, last));
// next = argmax tmp, LOWEST index among maxima (north_star tie rule; the reference's winner
// depends on its block size, sampling_cuda_kernel.cu:5-10,49-59).  d2 = pob::d2_ref.
#include "common.cuh"
#include "grid.cuh"
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>
#include <cstring>

namespace cg = cooperative_groups;

namespace pob {

constexpr int FPS_MAX_CLUSTER = 16;
constexpr int FPS_STREAM_THREADS = 1024;

struct __align__(16) FpsMsg {  // one CTA's winner, 32 bytes
    unsigned bits;             // float bits of the winning tmp value
    int idx;                   // global (original) point index
    float x, y, z;
    int pad0, pad1, pad2;
};

__device__ __forceinline__ void warp_argmax(unsigned bits, int gi, unsigned& wbits, int& wi) {
    wbits = __reduce_max_sync(FULL, bits);
    wi = __reduce_min_sync(FULL, bits == wbits ? gi : INT_MAX);
}

// ---- DSMEM all-to-all without fences: st.async carries its own completion (mbarrier tx) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, int cta) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void st_async_v4(unsigned remote_addr, unsigned a, unsigned b, unsigned c, unsigned d,
                                            unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar) : "memory");
}

// One cluster (C CTAs of T threads) per scene; P points per thread held in registers.
// Dynamic shared memory: float[3][T*P] copy of this CTA's coordinates (winner lookup).
// (A variant where every warp publishes straight to every CTA -- no CTA-level stage -- was
// measured slower: 988 vs 727 ns/iteration at 80k points; 16x more st.async per iteration.)
template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                   const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                   const float4* __restrict__ sorted, int* __restrict__ idx) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    int rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));  // volatile: never re-read inside the loop
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;

    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];

    extern __shared__ float s_xyz[];  // [3][T*P]
    __shared__ unsigned s_bits[2][NW];  // per-warp entries, double-buffered by iteration parity
    __shared__ int s_idx[2][NW];
    __shared__ int s_slot[2][NW];
    __shared__ __align__(16) FpsMsg s_msg[2][FPS_MAX_CLUSTER];
    __shared__ __align__(8) unsigned long long s_bar[2];

    // every CTA of the cluster takes the same branch: no barrier is skipped by part of it
    if (e_m <= s_m || e_n <= s_n) return;
    const int n = e_n - s_n;

    // ---- load: cell order (compact warps, prunable) when a grid is given, else strided ----
    const bool in_cells = scenes != nullptr && scenes[scene].use_grid;
    float px[P], py[P], pz[P], pt[P];
    int pi[P];
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool box_ok = true;
    const int sbase = in_cells ? __ldg(cell_start + scenes[scene].cell_base) : 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        // position of this register slot in the scene's point sequence
        const int pos = in_cells ? ((rank * NW + warp) * P + p) * 32 + lane : p * (C * T) + rank * T + tid;
        if (pos < n) {
            if (in_cells) {
                const float4 v = __ldg(sorted + sbase + pos);
                px[p] = v.x; py[p] = v.y; pz[p] = v.z; pi[p] = __float_as_int(v.w);
            } else {
                const int i = s_n + pos;
                px[p] = __ldg(xyz + (int64_t)i * 3); py[p] = __ldg(xyz + (int64_t)i * 3 + 1); pz[p] = __ldg(xyz + (int64_t)i * 3 + 2);
                pi[p] = i;
            }
            pt[p] = PLACEHOLDER_D2;
            box_ok = box_ok && isfinite(px[p]) && isfinite(py[p]) && isfinite(pz[p]);
            blo[0] = fminf(blo[0], px[p]); bhi[0] = fmaxf(bhi[0], px[p]);
            blo[1] = fminf(blo[1], py[p]); bhi[1] = fmaxf(bhi[1], py[p]);
            blo[2] = fminf(blo[2], pz[p]); bhi[2] = fmaxf(bhi[2], pz[p]);
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            pt[p] = -1.f;  // never a maximum: fminf keeps it at -1
            pi[p] = INT_MAX;
        }
        const int slot = (warp * P + p) * 32 + lane;
        s_xyz[slot] = px[p]; s_xyz[T * P + slot] = py[p]; s_xyz[2 * T * P + slot] = pz[p];
    }
    // warp bounding box (only meaningful, and only used, in cell order)
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            blo[a] = fminf(blo[a], __shfl_xor_sync(FULL, blo[a], o));
            bhi[a] = fmaxf(bhi[a], __shfl_xor_sync(FULL, bhi[a], o));
        }
    }
    const bool prune = in_cells && __all_sync(FULL, box_ok);

    float ox = __ldg(xyz + (int64_t)s_n * 3), oy = __ldg(xyz + (int64_t)s_n * 3 + 1), oz = __ldg(xyz + (int64_t)s_n * 3 + 2);
    const bool writer = rank == 0 && tid == 0;
    if (writer) idx[s_m] = s_n;
    int* out = idx + s_m;

    // exchange plumbing: one mbarrier per parity, armed for C messages of 32 bytes
    const unsigned bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    unsigned rslot0 = 0, rslot1 = 0, rbar0 = 0, rbar1 = 0;
    if (C > 1) {
        if (tid == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar0, C * (int)sizeof(FpsMsg));
            mbar_expect_tx(bar1, C * (int)sizeof(FpsMsg));
        }
        if (warp == 0 && lane < C) {  // lane l talks to CTA l
            rslot0 = mapa_u32(smem_u32(&s_msg[0][rank]), lane);
            rslot1 = mapa_u32(smem_u32(&s_msg[1][rank]), lane);
            rbar0 = mapa_u32(bar0, lane);
            rbar1 = mapa_u32(bar1, lane);
        }
        cluster.sync();  // every CTA's barriers exist before anyone writes remotely
    } else {
        __syncthreads();
    }

    unsigned wbits = __float_as_uint(PLACEHOLDER_D2);  // this warp's current maximum (cached while it skips)
    int wi = INT_MAX, wslot = 0;
    bool first_pass = true;  // every warp computes its entry once before it may start skipping
    const int iters = e_m - s_m - 1;

    for (int it = 0; it < iters; it++) {
        const int par = it & 1;
        // ---- can the new sample lower anything this warp owns? (exact, conservative) ----
        bool touch = true;
        if (prune && !first_pass) {
            const float ex = fmaxf(fmaxf(blo[0] - ox, ox - bhi[0]), 0.f);
            const float ey = fmaxf(fmaxf(blo[1] - oy, oy - bhi[1]), 0.f);
            const float ez = fmaxf(fmaxf(blo[2] - oz, oz - bhi[2]), 0.f);
            const float b2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
            // every d2_ref(point, sample) >= b2 * (1 - 1e-6); if even b2 * 0.99999 >= max tmp nothing changes
            touch = !(b2 * 0.99999f >= __uint_as_float(wbits));
        }
        first_pass = false;
        if (touch) {
            float m = 0.f;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const float d = d2_ref(px[p], py[p], pz[p], ox, oy, oz);
                pt[p] = fminf(d, pt[p]);
                m = fmaxf(m, pt[p]);
            }
            wbits = __reduce_max_sync(FULL, __float_as_uint(m));
            // lowest original index among this warp's maxima, and where it sits
            int cand = INT_MAX, cp = 0;
#pragma unroll
            for (int p = 0; p < P; p++)
                if (__float_as_uint(pt[p]) == wbits && pi[p] < cand) { cand = pi[p]; cp = p; }
            wi = __reduce_min_sync(FULL, cand);
            const unsigned own = __ballot_sync(FULL, cand == wi && wi != INT_MAX);
            const int ol = own ? __ffs(own) - 1 : 0;
            wslot = __shfl_sync(FULL, (warp * P + cp) * 32 + lane, ol);
        }
        if (lane == 0) { s_bits[par][warp] = wbits; s_idx[par][warp] = wi; s_slot[par][warp] = wslot; }
        __syncthreads();

        // ---- CTA argmax over the NW warp entries (every warp computes it redundantly) ----
        const unsigned eb = lane < NW ? s_bits[par][lane] : 0u;
        const int ei = lane < NW ? s_idx[par][lane] : INT_MAX;
        unsigned cbits; int ci;
        warp_argmax(eb, ei, cbits, ci);
        const unsigned whow = __ballot_sync(FULL, lane < NW && ei == ci);
        const int cslot = s_slot[par][whow ? __ffs(whow) - 1 : 0];
        float cx = s_xyz[cslot], cy = s_xyz[T * P + cslot], cz = s_xyz[2 * T * P + cslot];
        int gidx = ci;
        if (C > 1) {
            // ---- all-to-all of the CTA winners: 32-byte st.async per peer, completion on the
            //      receiver's mbarrier (no cluster-scope fence, nothing waits on global stores) ----
            if (warp == 0 && lane < C) {
                const unsigned rs = par ? rslot1 : rslot0, rb = par ? rbar1 : rbar0;
                st_async_v4(rs, cbits, (unsigned)ci, __float_as_uint(cx), __float_as_uint(cy), rb);
                st_async_v4(rs + 16, __float_as_uint(cz), 0u, 0u, 0u, rb);
            }
            mbar_wait(par ? bar1 : bar0, (unsigned)(it >> 1) & 1u);
            if (tid == 0) mbar_expect_tx(par ? bar1 : bar0, C * (int)sizeof(FpsMsg));  // re-arm for it + 2
            const FpsMsg mine = s_msg[par][lane < C ? lane : 0];
            unsigned gbits;
            warp_argmax(lane < C ? mine.bits : 0u, lane < C ? mine.idx : INT_MAX, gbits, gidx);
            const unsigned who = __ballot_sync(FULL, lane < C && mine.idx == gidx);
            const int wl = who ? __ffs(who) - 1 : 0;
            cx = __shfl_sync(FULL, mine.x, wl); cy = __shfl_sync(FULL, mine.y, wl); cz = __shfl_sync(FULL, mine.z, wl);
        }
        ox = cx; oy = cy; oz = cz;
        if (writer) out[it + 1] = gidx;
    }
    if (C > 1) cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}

// ---------------------------------------------------------------------------------------------
// fps_chain_kernel: the same register-resident layout, but SEVERAL samples per synchronisation
// round.  The serial chain of FPS is "argmax, update, argmax, ...", one cluster-wide exchange
// (~260 ns) per sample.  Observation: every exchange already tells every CTA the maxima of ALL
// groups (CTAs; warps when the cluster is one CTA).  Sort them by the key (value desc, index asc):
// a_1 > a_2 > ... (distinct groups).  a_1 is the next sample.  a_2 is the sample after it iff
//   (i)  a_1 does not lower a_2's own min-distance:  d2(a_2, a_1) >= tmp[a_2], and
//   (ii) the group of a_1, once a_1 is applied, holds nothing above a_2:  V_1 < tmp[a_2],
// because every other group's maximum was <= a_2 and min-distances only ever decrease.  In general
// a_{j+1} follows a_1..a_j iff it is lowered by none of them and V_i < tmp[a_{j+1}] for all i <= j,
// where V_i is ANY upper bound of group i's maximum after its own candidate a_i has been applied.
// Each warp keeps, next to its maximum e_w, the exact value V_w = max_p min(tmp_p, d2(p, top_w))
// (a dry run against its own top point, redone only when the warp's points changed); a CTA
// publishes E = max_w e_w (warp w*) and V = max(V_w*, second largest e_w).  One exchange then
// yields a whole accepted prefix a_1..a_L (L <= 8), all CTAs derive the same L from the same 16
// messages, apply the L samples (same exact pruning), and go round again.  The emitted index
// sequence is IDENTICAL to one-sample-at-a-time FPS (ties: the strict tests fall back to L = 1).
// Measured on S3DIS-shaped rooms: see profiles/ (mean accepted chain length and ns per sample).
constexpr int FPS_KMAX = 8;

struct __align__(16) FpsEntry {  // one group's candidate, 32 bytes
    unsigned bits;               // float bits of the group's maximum min-distance
    int idx;                     // global (original) index of the point attaining it (lowest among ties)
    float x, y, z;               // its coordinates
    unsigned vbits;              // float bits of V: bound on the group's maximum once this point is a sample
    int pad0, pad1;
};

__device__ __forceinline__ unsigned long long fps_key(unsigned bits, int idx) {
    // orders like (value desc, index asc) under unsigned 64-bit '>'
    return ((unsigned long long)bits << 32) | (unsigned)(~idx);
}

//
// This is the round-1 kernel, kept as variant 2 (A/B against fps_merge_kernel below, which supersedes it).  The
// template switches SP (points in shared memory), GP (candidate groups per CTA) and NG (ranking slots) were
// measured in round 1 (profiles/r01d_experiments.md: none of them moved the room pipeline) and only the
// register-resident, one-group-per-CTA form <P, T, false, 1, 16> is instantiated.
template <int P, int T, bool SP, int GP, int NG = FPS_MAX_CLUSTER>
__device__ __forceinline__ void
fps_chain_body(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                 const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                 const float4* __restrict__ sorted, int* __restrict__ idx, unsigned long long* __restrict__ stats) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    int rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    constexpr int WG = NW / GP;          // warps per group
    // NG group slots are ranked per round (unused ones stay zero)
    static_assert(NW <= NG && NW % GP == 0 && (NG == 16 || NG == 32), "group entries are ranked 16 or 32 at a time");

    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];

    extern __shared__ __align__(16) float s_xyz[];  // SP: float4 {x, y, z, idx}[T*P]; else float [3][T*P]
    float4* const s_pts = reinterpret_cast<float4*>(s_xyz);
    __shared__ __align__(16) FpsEntry s_warp[2][NG];          // warp entries by round parity (slots >= NW stay zero)
    __shared__ __align__(16) FpsEntry s_msg[2][NG];           // CTA entries from the peers, by round parity
    __shared__ __align__(16) FpsEntry s_sorted[NW][NG];       // per warp: the candidates in key order
    __shared__ __align__(8) unsigned long long s_bar[2];

    if (e_m <= s_m || e_n <= s_n) return;  // uniform over the cluster
    const int n = e_n - s_n;
    const int total = e_m - s_m;

    const bool in_cells = scenes != nullptr && scenes[scene].use_grid;
    constexpr int PR = SP ? 1 : P;     // register copies of the points only in the register-resident form
    float px[PR], py[PR], pz[PR], pt[P];
    int pi[PR];
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool box_ok = true;
    const int sbase = in_cells ? __ldg(cell_start + scenes[scene].cell_base) : 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        const int pos = in_cells ? ((rank * NW + warp) * P + p) * 32 + lane : p * (C * T) + rank * T + tid;
        float vx = 0.f, vy = 0.f, vz = 0.f;
        int vi = INT_MAX;
        if (pos < n) {
            if (in_cells) {
                const float4 v = __ldg(sorted + sbase + pos);
                vx = v.x; vy = v.y; vz = v.z; vi = __float_as_int(v.w);
            } else {
                const int i = s_n + pos;
                vx = __ldg(xyz + (int64_t)i * 3); vy = __ldg(xyz + (int64_t)i * 3 + 1); vz = __ldg(xyz + (int64_t)i * 3 + 2);
                vi = i;
            }
            pt[p] = PLACEHOLDER_D2;
            box_ok = box_ok && isfinite(vx) && isfinite(vy) && isfinite(vz);
            blo[0] = fminf(blo[0], vx); bhi[0] = fmaxf(bhi[0], vx);
            blo[1] = fminf(blo[1], vy); bhi[1] = fmaxf(bhi[1], vy);
            blo[2] = fminf(blo[2], vz); bhi[2] = fmaxf(bhi[2], vz);
        } else {
            pt[p] = -1.f;  // never a maximum: fminf keeps it at -1
        }
        const int slot = (warp * P + p) * 32 + lane;
        if constexpr (SP) {
            s_pts[slot] = make_float4(vx, vy, vz, __int_as_float(vi));
        } else {
            px[p] = vx; py[p] = vy; pz[p] = vz; pi[p] = vi;
            s_xyz[slot] = vx; s_xyz[T * P + slot] = vy; s_xyz[2 * T * P + slot] = vz;
        }
    }
    const float4* const my_pts = s_pts + warp * P * 32 + lane;   // SP: point p of this lane is my_pts[p * 32]
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            blo[a] = fminf(blo[a], __shfl_xor_sync(FULL, blo[a], o));
            bhi[a] = fmaxf(bhi[a], __shfl_xor_sync(FULL, bhi[a], o));
        }
    }
    const bool prune = in_cells && __all_sync(FULL, box_ok);

    // empty group slots: value 0 and index -1, the smallest key there is -- below every real candidate even
    // when all remaining min-distances are 0 (duplicate points: the real keys are then (0, ~idx) with
    // idx >= 0), and never accepted into a chain (acceptance needs V < 0 as unsigned)
    for (int i = tid; i < 2 * NG * (int)(sizeof(FpsEntry) / 16); i += T) {
        const float4 z = (i & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(0.f, __int_as_float(-1), 0.f, 0.f);
        reinterpret_cast<float4*>(s_warp)[i] = z;
        reinterpret_cast<float4*>(s_msg)[i] = z;
    }
    const bool writer = rank == 0 && warp == 0;
    const unsigned bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    unsigned rslot0 = 0, rslot1 = 0, rbar0 = 0, rbar1 = 0;
    if (writer && lane == 0) idx[s_m] = s_n;  // sampling_cuda_kernel.cu:39
    if (C > 1) {
        if (tid == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar0, C * GP * (int)sizeof(FpsEntry));
            mbar_expect_tx(bar1, C * GP * (int)sizeof(FpsEntry));
        }
        if (warp % WG == 0 && lane < C) {  // a group's first warp publishes for it: lane l talks to CTA l
            rslot0 = mapa_u32(smem_u32(&s_msg[0][rank * GP + warp / WG]), lane);
            rslot1 = mapa_u32(smem_u32(&s_msg[1][rank * GP + warp / WG]), lane);
            rbar0 = mapa_u32(bar0, lane);
            rbar1 = mapa_u32(bar1, lane);
        }
        cluster.sync();
    } else {
        __syncthreads();
    }

    // lane <-> candidate pair (pa < pb), pb-major: (0,1) (0,2) (1,2) (0,3) ...; 28 pairs for 8 candidates
    int pb = 1, pbase = 0;
    while (pbase + pb <= lane) { pbase += pb; pb++; }
    const int pa = lane - pbase;
    FpsEntry* my_sorted = s_sorted[warp];

    // the accepted, not yet applied samples: lane j < L holds a_j.  First one: the scene's first point.
    FpsEntry E;
    E.bits = __float_as_uint(PLACEHOLDER_D2); E.idx = s_n;
    E.x = __ldg(xyz + (int64_t)s_n * 3); E.y = __ldg(xyz + (int64_t)s_n * 3 + 1); E.z = __ldg(xyz + (int64_t)s_n * 3 + 2);
    E.vbits = 0u;
    int L = 1;
    int emitted = 1;  // indices written so far; the last L of them are not applied yet
    unsigned wbits = __float_as_uint(PLACEHOLDER_D2), vbits = __float_as_uint(PLACEHOLDER_D2);
    int wi = INT_MAX, wslot = 0;
    float tx = 0.f, ty = 0.f, tz = 0.f;  // coordinates of the warp's top point
    bool dirty = true;  // the warp entry has to be (re)computed
    unsigned long long n_rounds = 0;
    int* out = idx + s_m;

    for (int round = 0;; round++) {
        // ---- apply the accepted samples: lane j tests sample j against the warp's box (exact pruning) ----
        {
            bool touch = lane < L;
            if (prune && round > 0) {
                const float ex = fmaxf(fmaxf(blo[0] - E.x, E.x - bhi[0]), 0.f);
                const float ey = fmaxf(fmaxf(blo[1] - E.y, E.y - bhi[1]), 0.f);
                const float ez = fmaxf(fmaxf(blo[2] - E.z, E.z - bhi[2]), 0.f);
                const float b2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
                // every d2_ref(point, sample) >= b2 * (1 - 1e-6); if even b2 * 0.99999 >= max tmp nothing changes
                touch = touch && !(b2 * 0.99999f >= __uint_as_float(wbits));
            }
            unsigned tm = __ballot_sync(FULL, touch);
            dirty = dirty || tm != 0u;
            while (tm) {
                const int j = __ffs(tm) - 1;
                tm &= tm - 1;
                const float sx = __shfl_sync(FULL, E.x, j), sy = __shfl_sync(FULL, E.y, j), sz = __shfl_sync(FULL, E.z, j);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    if constexpr (SP) {
                        const float4 q = my_pts[p * 32];
                        pt[p] = fminf(d2_ref(q.x, q.y, q.z, sx, sy, sz), pt[p]);
                    } else {
                        pt[p] = fminf(d2_ref(px[p], py[p], pz[p], sx, sy, sz), pt[p]);
                    }
                }
            }
        }
        if (emitted >= total) break;  // uniform: `emitted` evolves identically in every thread of the cluster
        n_rounds++;
        if (dirty) {
            dirty = false;
            float m = 0.f;
#pragma unroll
            for (int p = 0; p < P; p++) m = fmaxf(m, pt[p]);
            wbits = __reduce_max_sync(FULL, __float_as_uint(m));
            int cand = INT_MAX, cp = 0;
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (__float_as_uint(pt[p]) == wbits) {
                    int gi;
                    if constexpr (SP) gi = __float_as_int(my_pts[p * 32].w); else gi = pi[p];
                    if (gi < cand) { cand = gi; cp = p; }
                }
            }
            wi = __reduce_min_sync(FULL, cand);
            const unsigned own = __ballot_sync(FULL, cand == wi && wi != INT_MAX);
            const int ol = own ? __ffs(own) - 1 : 0;
            wslot = __shfl_sync(FULL, (warp * P + cp) * 32 + lane, ol);
            // V_w: this warp's maximum if its own top point became a sample (dry run, nothing stored)
            if constexpr (SP) {
                const float4 t4 = s_pts[wslot];
                tx = t4.x; ty = t4.y; tz = t4.z;
            } else {
                tx = s_xyz[wslot]; ty = s_xyz[T * P + wslot]; tz = s_xyz[2 * T * P + wslot];
            }
            float vm = 0.f;
#pragma unroll
            for (int p = 0; p < P; p++) {
                if constexpr (SP) {
                    const float4 q = my_pts[p * 32];
                    vm = fmaxf(vm, fminf(d2_ref(q.x, q.y, q.z, tx, ty, tz), pt[p]));
                } else {
                    vm = fmaxf(vm, fminf(d2_ref(px[p], py[p], pz[p], tx, ty, tz), pt[p]));
                }
            }
            vbits = __reduce_max_sync(FULL, __float_as_uint(vm));
        }
        const int par = round & 1;
        if (lane == 0) {   // parity buffers: a warp already in round r + 1 must not overwrite what round r still reads
            FpsEntry e;
            e.bits = wbits; e.idx = wi; e.x = tx; e.y = ty; e.z = tz; e.vbits = vbits; e.pad0 = e.pad1 = 0;
            s_warp[par][warp] = e;
        }
        __syncthreads();  // warp entries of this round visible

        const FpsEntry* src = s_warp[par];
        if (C > 1) {
            if (warp % WG == 0) {
                // this group's candidate: its best warp entry; V = max(V of that warp, runner-up warp maximum)
                const FpsEntry e = s_warp[par][(warp + lane) & (NG - 1)];   // the group's warps are warp .. warp + WG - 1
                const unsigned eb = lane < WG ? e.bits : 0u;
                const int ei = lane < WG ? e.idx : INT_MAX;
                unsigned cbits; int ci;
                warp_argmax(eb, ei, cbits, ci);
                const unsigned whow = __ballot_sync(FULL, lane < WG && eb == cbits && ei == ci);
                const int wl = whow ? __ffs(whow) - 1 : 0;
                const unsigned second = __reduce_max_sync(FULL, (lane < WG && lane != wl) ? eb : 0u);
                const unsigned vw = __shfl_sync(FULL, e.vbits, wl);
                const unsigned cv = vw > second ? vw : second;
                const float cx = __shfl_sync(FULL, e.x, wl), cy = __shfl_sync(FULL, e.y, wl), cz = __shfl_sync(FULL, e.z, wl);
                if (lane < C) {
                    const unsigned rs = par ? rslot1 : rslot0, rb = par ? rbar1 : rbar0;
                    st_async_v4(rs, cbits, (unsigned)ci, __float_as_uint(cx), __float_as_uint(cy), rb);
                    st_async_v4(rs + 16, __float_as_uint(cz), cv, 0u, 0u, rb);
                }
            }
            mbar_wait(par ? bar1 : bar0, (unsigned)(round >> 1) & 1u);
            src = s_msg[par];
        }
        // ---- every warp, redundantly: rank the 16 group slots by key, lay them out in order ----
        const FpsEntry mine = src[lane & (NG - 1)];
        const unsigned long long mykey = fps_key(mine.bits, mine.idx);
        int rk = 0;
#pragma unroll
        for (int c = 0; c < NG; c++) {
            const uint2 o = *reinterpret_cast<const uint2*>(&src[c]);  // {bits, idx}
            rk += fps_key(o.x, (int)o.y) > mykey ? 1 : 0;
        }
        if (lane < NG) my_sorted[rk] = mine;   // equal keys only among empty slots (same content)
        __syncwarp();
        if (C > 1 && tid == 0) mbar_expect_tx(par ? bar1 : bar0, C * GP * (int)sizeof(FpsEntry));  // re-arm for round + 2
        // ---- longest accepted prefix: pair (a < b) needs "a_a leaves a_b untouched" and V_a < tmp[a_b] ----
        bool ok = true;
        if (pb < FPS_KMAX) {
            const FpsEntry ea = my_sorted[pa], eb = my_sorted[pb];
            const float d = d2_ref(eb.x, eb.y, eb.z, ea.x, ea.y, ea.z);   // exactly what the update computes for point b
            ok = !(d < __uint_as_float(eb.bits)) && ea.vbits < eb.bits;
        }
        E = my_sorted[lane & (FPS_KMAX - 1)];
        const unsigned bad = ~__ballot_sync(FULL, ok);
        L = bad ? __shfl_sync(FULL, pb, __ffs(bad) - 1) : FPS_KMAX;  // pairs are pb-major: the first failing pair names L
        if (L > total - emitted) L = total - emitted;
        if (writer && lane < L) out[emitted + lane] = E.idx;
        emitted += L;
        __syncwarp();  // my_sorted is rewritten next round
    }
    if (stats && tid == 0 && rank == 0) {   // diagnostics: rounds and samples per scene (mean chain = samples / rounds)
        atomicAdd(stats, n_rounds);
        atomicAdd(stats + 1, (unsigned long long)(total - 1));
    }
    if (C > 1) cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}

#define POB_FPS_CHAIN_PARAMS                                                                                         \
    const float *__restrict__ xyz, const int *__restrict__ offset, const int *__restrict__ new_offset,              \
        const SceneGrid *__restrict__ scenes, const int *__restrict__ cell_start, const float4 *__restrict__ sorted, \
        int *__restrict__ idx, unsigned long long *__restrict__ stats
template <int P, int T>
__global__ void __launch_bounds__(T, 1) fps_chain_kernel(POB_FPS_CHAIN_PARAMS) {
    fps_chain_body<P, T, false, 1>(xyz, offset, new_offset, scenes, cell_start, sorted, idx, stats);
}
#undef POB_FPS_CHAIN_PARAMS

// Scenes too large for the register-resident kernels: the same algorithm with the points left in global
// memory (L2-resident for any realistic scene) and the SAME exact pruning, per tile of 256 points.  A warp
// owns every 512th tile of the scene's point sequence (cell order when the kNN grid is given: compact
// tiles); per tile it keeps {bbox, max min-distance, argmax index, argmax coordinates} in shared memory.
// An iteration tests the new sample against the tile boxes and touches only tiles it can change (a few
// per cent), so what is left is the per-iteration synchronisation: warp -> CTA -> cluster argmax.
// tmp (caller's buffer, n floats) holds the running min-distances in the kernel's point order.
constexpr int FPS_TILE_PTS = 8;                        // points per lane per tile
constexpr int FPS_TILE = 32 * FPS_TILE_PTS;            // 256 points
constexpr int FPS_STREAM_WARPS = FPS_STREAM_THREADS / 32;
struct __align__(16) FpsTileRec {
    float lo[3], hi[3];
    unsigned bits;   // float bits of the tile's maximum min-distance
    int idx;         // lowest original index attaining it
    float x, y, z;   // its coordinates
    int pad;
};

__global__ void __launch_bounds__(FPS_STREAM_THREADS, 1)
fps_stream_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                  const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                  const float4* __restrict__ sorted, float* __restrict__ tmp, int* __restrict__ idx, int max_tiles_per_warp) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];
    extern __shared__ __align__(16) unsigned char fps_stream_smem[];
    FpsTileRec* recs = reinterpret_cast<FpsTileRec*>(fps_stream_smem) + (size_t)warp * max_tiles_per_warp;
    __shared__ unsigned s_bits[2][FPS_STREAM_WARPS];   // per-warp entries, double-buffered by iteration parity
    __shared__ int s_idx[2][FPS_STREAM_WARPS];
    __shared__ float s_c[2][FPS_STREAM_WARPS][3];
    __shared__ __align__(16) FpsMsg s_msg[2][FPS_MAX_CLUSTER];
    __shared__ __align__(8) unsigned long long s_bar[2];
    if (e_m <= s_m || e_n <= s_n) return;   // uniform over the cluster

    const int n = e_n - s_n;
    const bool in_cells = scenes != nullptr && scenes[scene].use_grid;
    const int sbase = in_cells ? __ldg(cell_start + scenes[scene].cell_base) : 0;
    const int ntiles = (n + FPS_TILE - 1) / FPS_TILE;
    const int gw = rank * FPS_STREAM_WARPS + warp, nw = C * FPS_STREAM_WARPS;
    const int my_tiles = gw < ntiles ? (ntiles - gw + nw - 1) / nw : 0;
    float* tm = tmp + s_n;

    auto load_point = [&](int pos, float& x, float& y, float& z, int& gi) {
        if (in_cells) {
            const float4 v = __ldg(sorted + sbase + pos);
            x = v.x; y = v.y; z = v.z; gi = __float_as_int(v.w);
        } else {
            const int i = s_n + pos;
            x = __ldg(xyz + (int64_t)i * 3); y = __ldg(xyz + (int64_t)i * 3 + 1); z = __ldg(xyz + (int64_t)i * 3 + 2);
            gi = i;
        }
    };
    // (re)compute one tile against a sample (first = true: initialise instead) and refresh its record
    auto visit = [&](int j, bool first, float ox, float oy, float oz) {
        const int base = (gw + j * nw) * FPS_TILE;
        float bd = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        int bi = INT_MAX;
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        bool finite = true;
#pragma unroll
        for (int p = 0; p < FPS_TILE_PTS; p++) {
            const int pos = base + p * 32 + lane;
            if (pos < n) {
                float x, y, z; int gi;
                load_point(pos, x, y, z, gi);
                float t;
                if (first) {
                    t = PLACEHOLDER_D2;
                    tm[pos] = t;
                    finite = finite && isfinite(x) && isfinite(y) && isfinite(z);
                    lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
                    lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
                    lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
                } else {
                    const float old = tm[pos];
                    t = fminf(d2_ref(x, y, z, ox, oy, oz), old);
                    if (t != old) tm[pos] = t;
                }
                if (t > bd || (t == bd && gi < bi)) { bd = t; bi = gi; bx = x; by = y; bz = z; }
            }
        }
        const unsigned mybits = bd >= 0.f ? __float_as_uint(bd) : 0u;
        const unsigned wbits = __reduce_max_sync(FULL, mybits);
        const int wi = __reduce_min_sync(FULL, (bd >= 0.f && mybits == wbits) ? bi : INT_MAX);
        const unsigned own = __ballot_sync(FULL, bd >= 0.f && mybits == wbits && bi == wi);
        const int ol = own ? __ffs(own) - 1 : 0;
        const float wx = __shfl_sync(FULL, bx, ol), wy = __shfl_sync(FULL, by, ol), wz = __shfl_sync(FULL, bz, ol);
        if (first) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo[a] = fminf(lo[a], __shfl_xor_sync(FULL, lo[a], o));
                    hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL, hi[a], o));
                }
            if (!__all_sync(FULL, finite)) {   // a box that every sample "touches": pruning off for this tile
#pragma unroll
                for (int a = 0; a < 3; a++) { lo[a] = -FLT_MAX; hi[a] = FLT_MAX; }
            }
        }
        if (lane == 0) {
            FpsTileRec& r = recs[j];
            if (first) {
#pragma unroll
                for (int a = 0; a < 3; a++) { r.lo[a] = lo[a]; r.hi[a] = hi[a]; }
            }
            r.bits = wbits; r.idx = wi; r.x = wx; r.y = wy; r.z = wz;
        }
    };

    for (int j = 0; j < my_tiles; j++) visit(j, true, 0.f, 0.f, 0.f);
    __syncwarp();
    float ox = __ldg(xyz + (int64_t)s_n * 3), oy = __ldg(xyz + (int64_t)s_n * 3 + 1), oz = __ldg(xyz + (int64_t)s_n * 3 + 2);
    if (rank == 0 && tid == 0) idx[s_m] = s_n;

    // exchange plumbing as in fps_cluster_kernel: st.async + mbarrier, NO cluster-scope fence on the chain
    // (cluster.sync() would make every iteration wait for the tmp stores above to drain: ~2 us)
    const unsigned bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    unsigned rslot0 = 0, rslot1 = 0, rbar0 = 0, rbar1 = 0;
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar0, C * (int)sizeof(FpsMsg));
        mbar_expect_tx(bar1, C * (int)sizeof(FpsMsg));
    }
    if (warp == 0 && lane < C) {  // lane l talks to CTA l
        rslot0 = mapa_u32(smem_u32(&s_msg[0][rank]), lane);
        rslot1 = mapa_u32(smem_u32(&s_msg[1][rank]), lane);
        rbar0 = mapa_u32(bar0, lane);
        rbar1 = mapa_u32(bar1, lane);
    }
    cluster.sync();  // every CTA's barriers exist before anyone writes remotely

    const int iters = e_m - s_m - 1;
    for (int r = 0; r < iters; r++) {
        const int par = r & 1;
        // ---- tiles the new sample can change (exact, conservative; see fps_cluster_kernel) ----
        for (int j0 = 0; j0 < my_tiles; j0 += 32) {
            const int j = j0 + lane;
            bool touch = false;
            if (j < my_tiles) {
                const FpsTileRec& rec = recs[j];
                const float ex = fmaxf(fmaxf(rec.lo[0] - ox, ox - rec.hi[0]), 0.f);
                const float ey = fmaxf(fmaxf(rec.lo[1] - oy, oy - rec.hi[1]), 0.f);
                const float ez = fmaxf(fmaxf(rec.lo[2] - oz, oz - rec.hi[2]), 0.f);
                const float b2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
                touch = !(b2 * 0.99999f >= __uint_as_float(rec.bits));
            }
            unsigned tmask = __ballot_sync(FULL, touch);
            while (tmask) {
                const int jj = j0 + __ffs(tmask) - 1;
                tmask &= tmask - 1;
                visit(jj, false, ox, oy, oz);
            }
            __syncwarp();
        }
        // ---- warp argmax over its tile records ----
        unsigned wb = 0u; int wi = INT_MAX, wj = 0;
        for (int j0 = 0; j0 < my_tiles; j0 += 32) {
            const int j = j0 + lane;
            const unsigned b = j < my_tiles ? recs[j].bits : 0u;
            const int i2 = j < my_tiles ? recs[j].idx : INT_MAX;
            unsigned cb; int ci;
            warp_argmax(b, i2, cb, ci);
            if (cb > wb || (cb == wb && ci < wi)) {
                const unsigned who = __ballot_sync(FULL, j < my_tiles && b == cb && i2 == ci);
                wb = cb; wi = ci; wj = j0 + (who ? __ffs(who) - 1 : 0);
            }
        }
        if (lane == 0) {
            s_bits[par][warp] = wb; s_idx[par][warp] = wi;
            const bool real = my_tiles > 0 && wi != INT_MAX;
            s_c[par][warp][0] = real ? recs[wj].x : 0.f; s_c[par][warp][1] = real ? recs[wj].y : 0.f;
            s_c[par][warp][2] = real ? recs[wj].z : 0.f;
        }
        __syncthreads();
        // ---- CTA argmax (every warp redundantly), all-to-all of the CTA winners, cluster argmax ----
        const unsigned eb = s_bits[par][lane];
        const int ei = s_idx[par][lane];
        unsigned cbits; int ci;
        warp_argmax(eb, ei, cbits, ci);
        const unsigned whow = __ballot_sync(FULL, eb == cbits && ei == ci);
        const int wl0 = whow ? __ffs(whow) - 1 : 0;
        if (warp == 0 && lane < C) {
            const unsigned rs = par ? rslot1 : rslot0, rb = par ? rbar1 : rbar0;
            st_async_v4(rs, cbits, (unsigned)ci, __float_as_uint(s_c[par][wl0][0]), __float_as_uint(s_c[par][wl0][1]), rb);
            st_async_v4(rs + 16, __float_as_uint(s_c[par][wl0][2]), 0u, 0u, 0u, rb);
        }
        mbar_wait(par ? bar1 : bar0, (unsigned)(r >> 1) & 1u);
        if (tid == 0) mbar_expect_tx(par ? bar1 : bar0, C * (int)sizeof(FpsMsg));  // re-arm for r + 2
        const FpsMsg mine = s_msg[par][lane < C ? lane : 0];
        unsigned gbits; int gidx;
        warp_argmax(lane < C ? mine.bits : 0u, lane < C ? mine.idx : INT_MAX, gbits, gidx);
        const unsigned who = __ballot_sync(FULL, lane < C && mine.idx == gidx);
        const int wl = who ? __ffs(who) - 1 : 0;
        ox = __shfl_sync(FULL, mine.x, wl); oy = __shfl_sync(FULL, mine.y, wl); oz = __shfl_sync(FULL, mine.z, wl);
        if (rank == 0 && tid == 0) idx[s_m + 1 + r] = gidx;
    }
    cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}

#include "fps_merge.cuh"

static int launch_cluster(const void* kernel, int b, int C, int threads, size_t smem, cudaStream_t stream, void** args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * C));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (C > 8) POB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (smem > 32 * 1024) POB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // static smem counts against the 48 KB default too
    POB_CHECK(cudaLaunchKernelExC(&cfg, kernel, args));
    pob_count_launches(1);
    return 0;
}

template <int T>
static int launch_chain(int P, int b, int C, cudaStream_t stream, void** args) {
#define POB_FPS_CASE(PP) \
    if (P <= PP) return launch_cluster((const void*)fps_chain_kernel<PP, T>, b, C, T, sizeof(float) * 3 * T * PP, stream, args)
    POB_FPS_CASE(1); POB_FPS_CASE(2); POB_FPS_CASE(4); POB_FPS_CASE(6); POB_FPS_CASE(8); POB_FPS_CASE(12);
    POB_FPS_CASE(16); POB_FPS_CASE(20); POB_FPS_CASE(24); POB_FPS_CASE(32);
#undef POB_FPS_CASE
    return POB_ERR_UNSUPPORTED;
}

template <int T>
static int launch_resident(int P, int b, int C, cudaStream_t stream, void** args) {
#define POB_FPS_CASE(PP) \
    if (P <= PP) return launch_cluster((const void*)fps_cluster_kernel<PP, T>, b, C, T, sizeof(float) * 3 * T * PP, stream, args)
    POB_FPS_CASE(1); POB_FPS_CASE(2); POB_FPS_CASE(4); POB_FPS_CASE(6); POB_FPS_CASE(8); POB_FPS_CASE(12);
    POB_FPS_CASE(16); POB_FPS_CASE(20); POB_FPS_CASE(24); POB_FPS_CASE(32);
#undef POB_FPS_CASE
    return POB_ERR_UNSUPPORTED;
}

template <int T, int D, int KC>
static int launch_merge(int P, int b, int C, cudaStream_t stream, void** args) {
#define POB_FPS_CASE(PP, KERN) \
    if (P <= PP) return launch_cluster((const void*)KERN<PP, T, D, KC>, b, C, T, sizeof(float4) * T * PP, stream, args)
    POB_FPS_CASE(1, fps_merge_kernel); POB_FPS_CASE(2, fps_merge_kernel); POB_FPS_CASE(4, fps_merge_kernel);
    POB_FPS_CASE(6, fps_merge_kernel); POB_FPS_CASE(8, fps_merge_kernel);
    POB_FPS_CASE(12, fps_merge_sp_kernel); POB_FPS_CASE(16, fps_merge_sp_kernel); POB_FPS_CASE(20, fps_merge_sp_kernel);
    POB_FPS_CASE(24, fps_merge_sp_kernel); POB_FPS_CASE(32, fps_merge_sp1_kernel);
    // beyond 32 points per thread the coordinates only fit in shared memory (16 B x 256 x 48 = 196 KB) and two rows share
    // a pruning box: scenes of 131 073 .. 196 608 points (most ScanNet scenes) still run inside ONE cluster, whose round is
    // a third of the grid-wide form's
    POB_FPS_CASE(40, fps_merge_sp1_kernel); POB_FPS_CASE(48, fps_merge_sp1_kernel);
#undef POB_FPS_CASE
    return POB_ERR_UNSUPPORTED;
}

// grid-wide form for one scene: cooperative launch of G CTAs (all resident at once: they spin on each other)
template <int T, int D, int KC>
static int launch_merge_grid(int P, int G, cudaStream_t stream, void** args) {
    const void* kernel = nullptr;
    int PP = 0;
#define POB_FPS_CASE(Q) \
    if (!kernel && P <= Q) { kernel = (const void*)fps_merge_grid_kernel<Q, T, D, KC>; PP = Q; }
    POB_FPS_CASE(4) POB_FPS_CASE(8) POB_FPS_CASE(12) POB_FPS_CASE(16) POB_FPS_CASE(20) POB_FPS_CASE(24) POB_FPS_CASE(32)
#undef POB_FPS_CASE
    if (!kernel) return POB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float4) * T * PP + (size_t)G * (KC + 1) * (sizeof(FpsEnt) + sizeof(unsigned short)) + 16;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)G);
    cfg.blockDim = dim3((unsigned)T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    POB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POB_CHECK(cudaLaunchKernelExC(&cfg, kernel, args));
    pob_count_launches(1);
    return 0;
}

// ---- tiny scenes (<= 2 048 points): one CTA per scene, one sample per iteration, nothing speculative ----
// For the last stage of a room (1 250 -> 312) a round of the cluster kernels (1-3 us of exchange, ranking and list
// upkeep for 2-4 samples) costs more than the plain chain: a CTA of 8 warps keeps its scene in registers (P points per
// thread), and an iteration is update -> 64-bit key max inside the warp (two REDUX) -> one shared-memory slot per warp ->
// ONE __syncthreads (slots are double buffered) -> the same two REDUX over the warp slots -> winner's coordinates from the
// shared-memory copy: ~250-300 ns (1 250 -> 312: 0.098 ms against 0.124 for the round-1 chain kernel on one CTA).  Beyond
// ~2 000 points one SM's issue rate bounds the update (5 000 points on 1 024 threads: 860 ns per sample) and the
// cluster kernels win again.
// Keys are (d2 bits << 32) | (0x7fffffff - index): the maximum is the farthest point, ties to the lowest index.
template <int T, int P>
__global__ void __launch_bounds__(T)
fps_small_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                 int* __restrict__ idx, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) float4 sm_pts[];   // [T * P]
    __shared__ unsigned long long s_best[2][32];
    constexpr int NW = T / 32;
    const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s_n = scene == 0 ? 0 : offset[scene - 1], n = offset[scene] - s_n;
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], m = new_offset[scene] - s_m;
    if (n <= 0 || m <= 0) return;
    float px[P], py[P], pz[P], tmp[P];
#pragma unroll
    for (int k = 0; k < P; k++) {
        const int i = tid + k * T;
        px[k] = py[k] = pz[k] = 0.f;
        tmp[k] = -1.f;                       // never a maximum
        if (i < n) {
            px[k] = __ldg(xyz + (int64_t)(s_n + i) * 3); py[k] = __ldg(xyz + (int64_t)(s_n + i) * 3 + 1); pz[k] = __ldg(xyz + (int64_t)(s_n + i) * 3 + 2);
            tmp[k] = PLACEHOLDER_D2;
            sm_pts[i] = make_float4(px[k], py[k], pz[k], 0.f);
        }
    }
    if (tid < 64) s_best[tid >> 5][tid & 31] = 0ull;
    if (tid == 0) idx[s_m] = s_n;
    __syncthreads();
    float4 cur = sm_pts[0];
    for (int j = 1; j < m; j++) {
        const int par = j & 1;
        unsigned long long best = 0ull;
#pragma unroll
        for (int k = 0; k < P; k++) {
            const float d = d2_ref(px[k], py[k], pz[k], cur.x, cur.y, cur.z);
            tmp[k] = fminf(d, tmp[k]);
            const unsigned long long key = tmp[k] >= 0.f ? ((unsigned long long)__float_as_uint(tmp[k]) << 32) | (unsigned)(0x7fffffff - (s_n + tid + k * T)) : 0ull;
            best = key > best ? key : best;
        }
        {
            const unsigned hi = (unsigned)(best >> 32), mh = __reduce_max_sync(FULL, hi);
            const unsigned ml = __reduce_max_sync(FULL, hi == mh ? (unsigned)best : 0u);
            if (lane == 0) s_best[par][warp] = ((unsigned long long)mh << 32) | ml;
        }
        __syncthreads();   // the slots of the other parity are rewritten only after the next barrier
        const unsigned long long wk = lane < NW ? s_best[par][lane] : 0ull;
        const unsigned hi = (unsigned)(wk >> 32), mh = __reduce_max_sync(FULL, hi);
        const unsigned ml = __reduce_max_sync(FULL, hi == mh ? (unsigned)wk : 0u);
        const int win = 0x7fffffff - (int)ml;
        cur = sm_pts[win - s_n];
        if (tid == 0) idx[s_m + j] = win;
    }
    if (stats && tid == 0) {   // one sample per round; every point meets every sample
        atomicAdd(stats, (unsigned long long)(m - 1));
        atomicAdd(stats + 1, (unsigned long long)(m - 1));
        atomicAdd(stats + 2, (unsigned long long)(m - 1) * (unsigned long long)n);
    }
}

template <int T>
static int launch_small(int P, int b, cudaStream_t stream, const float* xyz, const int* offset, const int* new_offset, int* idx,
                        unsigned long long* stats) {
#define POB_FPS_CASE(PP)                                                                                                    \
    if (P <= PP) {                                                                                                          \
        constexpr size_t smem = sizeof(float4) * T * PP;                                                                    \
        auto kern = fps_small_kernel<T, PP>;                                                                                \
        if (smem > 32 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<b, T, smem, stream>>>(xyz, offset, new_offset, idx, stats);                                                      \
        pob_count_launches(1);                                                                                              \
        POB_RETURN_LAST_ERROR();                                                                                            \
    }
    POB_FPS_CASE(1) POB_FPS_CASE(2) POB_FPS_CASE(3) POB_FPS_CASE(4) POB_FPS_CASE(5) POB_FPS_CASE(6) POB_FPS_CASE(8)
#undef POB_FPS_CASE
    return POB_ERR_UNSUPPORTED;
}

// ---- space-filling-curve order of the input (merged-list kernels) ----
// The cluster kernel prunes by the bounding boxes of rows (32 consecutive points) and warps (P rows), so it wants
// consecutive points to be compact blobs.  The kNN grid's cell-sorted array runs x-fastest: a warp's 640 points are a
// strip across the room.  Sorting the points of every scene along a 24-bit Hilbert curve (8 bits per axis over the
// scene's own extent) costs ~20 us and takes 17-19 % off the kernel (80 000 -> 20 000: 3.94 -> 3.25 ms; fewer rounds
// AND a third fewer touched rows per sample; tools/fps_order_experiment.py).  The samples are the same whatever the
// order.
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x30000ffu;
    v = (v | (v << 8)) & 0x300f00fu;
    v = (v | (v << 4)) & 0x30c30c3u;
    v = (v | (v << 2)) & 0x9249249u;
    return v;
}

__global__ void fps_curve_key_kernel(int64_t n, int b, const float* __restrict__ xyz, const int* __restrict__ offset,
                                      const SceneGrid* __restrict__ scenes, int bits, unsigned long long* __restrict__ keys,
                                      unsigned* __restrict__ vals) {
    const float cells = (float)(1 << bits);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = segment_of(i, offset, b);
        const SceneGrid g = scenes[s];
        unsigned m;
        if (g.use_grid) {
            const float ext = (float)max(g.dx, max(g.dy, g.dz)) * g.h;
            const float sc = ext > 0.f ? cells / ext : 0.f;
            const float q[3] = {(__ldg(xyz + i * 3) - g.lox) * sc, (__ldg(xyz + i * 3 + 1) - g.loy) * sc, (__ldg(xyz + i * 3 + 2) - g.loz) * sc};
            unsigned c[3];
#pragma unroll
            for (int a = 0; a < 3; a++) c[a] = q[a] >= 0.f ? (unsigned)fminf(q[a], cells - 1.f) : 0u;   // NaN -> 0
            // Hilbert index of the cell (Skilling's axes -> transpose form): unlike the plain Morton order,
            // consecutive keys are always face neighbours, so rows and warps have no far jumps inside them
            // (80 000 -> 20 000: 3.25 ms vs 3.51 Morton vs 3.94 cell order; 443 vs 579 vs 680 distances per sample)
#pragma unroll 1
            for (unsigned Q = 1u << (bits - 1); Q > 1u; Q >>= 1) {
                const unsigned Pm = Q - 1u;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (c[a] & Q) c[0] ^= Pm;   // a == 0: the exchange below is the identity
                    else { const unsigned t = (c[0] ^ c[a]) & Pm; c[0] ^= t; c[a] ^= t; }
                }
            }
            c[1] ^= c[0]; c[2] ^= c[1];
            unsigned t = 0u;
#pragma unroll 1
            for (unsigned Q = 1u << (bits - 1); Q > 1u; Q >>= 1)
                if (c[2] & Q) t ^= Q - 1u;
            c[0] ^= t; c[1] ^= t; c[2] ^= t;
            m = (spread10(c[0]) << 2) | (spread10(c[1]) << 1) | spread10(c[2]);
        } else {   // scenes too small for a grid keep their order
            m = (unsigned)min((int64_t)((1 << (3 * bits)) - 1), i - (s == 0 ? 0 : offset[s - 1]));
        }
        keys[i] = ((unsigned long long)s << (3 * bits)) | m;
        vals[i] = (unsigned)i;
    }
}

__global__ void fps_curve_gather_kernel(int64_t n, const float* __restrict__ xyz, const unsigned* __restrict__ vals,
                                         float4* __restrict__ ordered) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = vals[j];
        ordered[j] = make_float4(__ldg(xyz + i * 3), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2), __int_as_float((int)i));
    }
}

// fills the ordered copy inside the grid workspace; returns it through *ordered_out
static int fps_curve_order(int64_t n, int b, const float* xyz, const int* offset, const SceneGrid* scenes, char* region,
                            size_t region_bytes, const float4** ordered_out, cudaStream_t stream) {
    size_t o = 0;
    float4* ordered = (float4*)(region + o);               o = align_up(o + sizeof(float4) * (size_t)n, 256);
    unsigned long long* k0 = (unsigned long long*)(region + o);  o = align_up(o + 8 * (size_t)n, 256);
    unsigned long long* k1 = (unsigned long long*)(region + o);  o = align_up(o + 8 * (size_t)n, 256);
    unsigned* v0 = (unsigned*)(region + o);                 o = align_up(o + 4 * (size_t)n, 256);
    unsigned* v1 = (unsigned*)(region + o);                 o = align_up(o + 4 * (size_t)n, 256);
    if (o > region_bytes) return POB_ERR_WORKSPACE;
    void* temp = region + o;
    size_t temp_bytes = 0;
    int scene_bits = 0;
    // 8 bits per axis: as good as 10 (3.28 vs 3.30 ms at 80 000 points; 5 bits still is, 4 are not) and one radix pass less
    constexpr int bits = 8;
    while ((1 << scene_bits) < b) scene_bits++;
    // double-buffer form: the sort ping-pongs between the two key / value arrays and needs only its histograms as
    // temporary storage (the plain form allocates another n keys + values there)
    cub::DoubleBuffer<unsigned long long> dk(k0, k1);
    cub::DoubleBuffer<unsigned> dv(v0, v1);
    POB_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, dk, dv, (int64_t)n, 0, 3 * bits + scene_bits, stream));
    if (o + temp_bytes > region_bytes) return POB_ERR_WORKSPACE;
    fps_curve_key_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, b, xyz, offset, scenes, bits, k0, v0);
    POB_CHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, dk, dv, (int64_t)n, 0, 3 * bits + scene_bits, stream));
    fps_curve_gather_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, xyz, dv.Current(), ordered);
    pob_count_launches(2);
    *ordered_out = ordered;
    POB_RETURN_LAST_ERROR();
}

}  // namespace pob

using namespace pob;

// farthest_point_sampling_cuda_launcher(b, n, xyz, offset, new_offset, tmp, idx)
// (sampling_cuda_kernel.h:13) + the optional kNN grid of the same (xyz, offset) + per-call options + stream.
// n_max = largest scene (the reference's `n`); tmp (n floats) is only touched when a scene
// exceeds the register-resident capacity (131072 points) and needs no initialisation.
// grid_workspace: NULL, or the workspace pob_knn_grid_build filled for the same xyz/offset with
// the same n, b, cell_pts -- enables exact spatial pruning; results are identical either way.  The launcher writes
// its Hilbert-ordered copy of the points into the workspace's own FPS region (disjoint from what the kNN queries read).
// cluster_hint: 0 = choose, else force 1/2/4/8/16 CTAs per scene.
// variant: POB_FPS_AUTO (0) / POB_FPS_MERGE (1) / POB_FPS_CHAIN (2) / POB_FPS_SINGLE (3) / POB_FPS_MERGE_CELLS (4): same samples, different
// schedules (A/B and fallback); stats: NULL or 2 x u64 on the device {rounds, samples} accumulated by the launch.
POB_API int pob_farthest_point_sampling(int b, int64_t n_max, const float* xyz, const int* offset,
                                        const int* new_offset, float* tmp, int* idx, int cluster_hint,
                                        void* grid_workspace, int64_t n, float cell_pts, int variant,
                                        void* stats_u64x4, cudaStream_t stream) {
    if (b < 1 || n_max < 0 || !offset || !new_offset || !idx || variant < 0 || variant > 4) return POB_ERR_BAD_ARG;
    const bool reorder = variant != 4;   // POB_FPS_MERGE_CELLS: the merged-list kernel on the grid's cell order (A/B)
    if (variant == 4) variant = 1;
    if (n_max == 0) return 0;
    if (!xyz) return POB_ERR_BAD_ARG;
    const SceneGrid* scenes = nullptr;
    const int* cell_start = nullptr;
    const float4* sorted = nullptr;
    if (grid_workspace) {
        if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
        const GridLayout L = grid_layout(n, b, cell_pts);
        const char* ws = (const char*)grid_workspace;
        scenes = (const SceneGrid*)(ws + L.off_scene);
        cell_start = (const int*)(ws + L.off_start);
        sorted = (const float4*)(ws + L.off_sorted);
    }
    const bool hinted = cluster_hint == 1 || cluster_hint == 2 || cluster_hint == 4 || cluster_hint == 8 || cluster_hint == 16;
    if (variant == 0 && !hinted && n_max <= 2048) {   // tiny scenes: the plain one-sample-per-iteration CTA kernel is the fastest form
        return launch_small<256>((int)ceil_div(n_max, 256), b, stream, xyz, offset, new_offset, idx, (unsigned long long*)stats_u64x4);
    }
    // the merged-list kernels read a Hilbert-ordered copy (built on first use, in the workspace's FPS region)
    const int* m_cell_start = cell_start;
    const float4* m_sorted = sorted;
    bool ordered_done = false;
    auto order_points = [&]() -> int {
        if (ordered_done || !grid_workspace || !reorder) return 0;
        const GridLayout L = grid_layout(n, b, cell_pts);
        const int rc = fps_curve_order(n, b, xyz, offset, scenes, (char*)grid_workspace + L.off_fps, L.fps_bytes, &m_sorted, stream);
        if (rc) return rc;
        m_cell_start = nullptr;
        ordered_done = true;
        return 0;
    };
    constexpr int T = 256;      // 2 warps per scheduler: the per-iteration overhead scales with warps
    // points per thread: 32 with the coordinates in registers (round-1 kernels), 48 for the merged-list kernel with
    // its coordinates in shared memory
    const int PMAX = variant <= 1 ? 48 : 32;
    int C = cluster_hint;
    if (C != 1 && C != 2 && C != 4 && C != 8 && C != 16) {
        if (variant != 3) {
            // a round costs about the same whatever C is, and accepts more samples the more warps compete:
            // 16 CTAs unless the scene is tiny or there are so many scenes that 16-CTA clusters could not all
            // be resident (8 GPCs)
            C = n_max <= 2048 ? 1 : (b <= 8 ? 16 : (b <= 16 ? 8 : 4));
        } else {
            // ~200 ns of cluster exchange per iteration buys a 1/C share of the update work
            C = n_max <= 4096 ? 1 : (n_max <= 16384 ? 4 : (n_max <= 40960 ? 8 : 16));
        }
    }
    while (C < 16 && ceil_div(n_max, (int64_t)C * T) > PMAX) C *= 2;
    const int64_t P = ceil_div(n_max, (int64_t)C * T);
    constexpr int FPS_D = 2, FPS_KC = 4;
    if (P > PMAX && variant <= 1 && n_max <= (int64_t)sm_count() * T * 32) {
        // beyond one cluster's registers: the grid-wide form, one cooperative launch per scene (a scene that fits a
        // cluster returns at once there and is sampled by the cluster launch below, and vice versa).  Workspace in
        // tmp: a counter (zeroed here) and the message buffers.
        if (!tmp) return POB_ERR_BAD_ARG;
        const int G = sm_count();
        const int64_t Pg = ceil_div(n_max, (int64_t)G * T);
        if (n < (int64_t)(256 + 2 * (size_t)G * (FPS_KC + 1) * sizeof(FpsEnt) + 3) / 4) return POB_ERR_WORKSPACE;
        unsigned* counter = (unsigned*)tmp;
        FpsEnt* gmsg = (FpsEnt*)((char*)tmp + 256);
        unsigned long long* stats = (unsigned long long*)stats_u64x4;
        const int cap = b > 1 ? FPS_MAX_CLUSTER * T * PMAX : 0;
        if (const int rc = order_points()) return rc;
        for (int s = 0; s < b; s++) {
            POB_CHECK(cudaMemsetAsync(counter, 0, 256, stream));
            int scene = s, cap_points = cap;
            void* gargs[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&m_cell_start, (void*)&m_sorted,
                             (void*)&idx, (void*)&stats, (void*)&scene, (void*)&cap_points, (void*)&gmsg, (void*)&counter};
            const int rc = launch_merge_grid<T, FPS_D, FPS_KC>((int)Pg, G, stream, gargs);
            if (rc) return rc;
        }
        if (b == 1) return 0;
        void* cargs[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&m_cell_start,
                         (void*)&m_sorted, (void*)&idx, (void*)&stats};
        return launch_merge<T, FPS_D, FPS_KC>(PMAX, b, FPS_MAX_CLUSTER, stream, cargs);
    }
    if (P > PMAX) {
        if (!tmp) return POB_ERR_BAD_ARG;
        const int Cs = 16;
        const int64_t ntiles = ceil_div(n_max, (int64_t)FPS_TILE);
        int tiles_per_warp = (int)ceil_div(ntiles, (int64_t)Cs * FPS_STREAM_WARPS);
        if (tiles_per_warp < 1) tiles_per_warp = 1;
        const size_t smem = sizeof(FpsTileRec) * (size_t)FPS_STREAM_WARPS * tiles_per_warp;
        if (smem > 200 * 1024) return POB_ERR_UNSUPPORTED;   // > 17 M points in one scene
        void* args[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&cell_start,
                        (void*)&sorted, (void*)&tmp, (void*)&idx, (void*)&tiles_per_warp};
        return launch_cluster((const void*)fps_stream_kernel, b, Cs, FPS_STREAM_THREADS, smem, stream, args);
    }
    if (variant != 3) {
        unsigned long long* stats = (unsigned long long*)stats_u64x4;
        void* cargs[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&cell_start,
                         (void*)&sorted, (void*)&idx, (void*)&stats};
        // tiny scenes run on one CTA without a grid: there every warp is touched by every sample and the round-1
        // kernel's shorter round wins (0.13 vs 0.30 ms at 1 250 -> 312 points)
        if (variant == 2 || (variant == 0 && C == 1)) return launch_chain<T>((int)P, b, C, stream, cargs);
        if (const int rc = order_points()) return rc;
        void* margs[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&m_cell_start,
                         (void*)&m_sorted, (void*)&idx, (void*)&stats};
        return launch_merge<T, FPS_D, FPS_KC>((int)P, b, C, stream, margs);
    }
    void* args[] = {(void*)&xyz, (void*)&offset, (void*)&new_offset, (void*)&scenes, (void*)&cell_start,
                    (void*)&sorted, (void*)&idx};
    return launch_resident<T>((int)P, b, C, stream, args);
}


