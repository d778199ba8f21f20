// fps_merge.cuh -- merged-list farthest point sampling (included by fps.cu inside namespace pob).
//
// FPS is "argmax, update, argmax, ..." -- one cluster-wide exchange per sample in the plain form.  The round-1
// chain kernel accepted ~4.5 samples per exchange (every CTA offers ONE candidate plus a bound on what would
// follow it).  This kernel generalises that to whole LISTS and reaches ~20 samples per exchange on S3DIS-shaped
// rooms (tools/fps_merge_sim.py: 80 000 -> 20 000 points in 1 010 rounds instead of 4 444, output identical
// to one-at-a-time FPS in every configuration simulated):
//
//   * every warp owns a spatially compact run of 32*P points (cell order of the kNN grid) and keeps, next to
//     the real min-distances pt[], a SPECULATIVE copy st[] on which it runs FPS locally: a_1 = its argmax,
//     apply a_1 to st, a_2 = next argmax, ...  The list (a_1 .. a_D, then the value of a_{D+1} as a TERMINAL
//     bound) is what the warp would contribute if no other warp's sample ever reached into its box;
//   * one exchange merges all lists by the key (value desc, original index asc).  Walking the merged order,
//     entry e_j is exactly the next global FPS sample as long as (i) it is a real entry, not a terminal, and
//     (ii) no EARLIER entry e_i of ANOTHER warp lowers it: d2(e_j, e_i) >= value(e_j)  (samples of its own
//     warp are already accounted for in its value).  Proof sketch: with e_1..e_{j-1} applied, every warp's
//     true maximum is <= the value of its first not-yet-taken list entry (min-distances only decrease), which
//     sorts after e_j; e_j itself is untouched, so it is the global maximum, lowest index among ties.
//     The walk stops at the first terminal or conflict; at most 32 samples are taken per exchange;
//   * afterwards a warp applies only the accepted samples whose reach intersects its bounding box (exact,
//     conservatively rounded -- as in the round-1 kernels): its own ones to pt[] (st[] has them already), foreign
//     ones to pt[] AND st[].  The list stays valid as long as no foreign sample lowers one of its pending entries
//     (a point test per entry and sample): entries taken by the merge are popped, one local step per popped entry
//     tops the list up again.  Only when a pending entry IS lowered the list is void: st = pt, D local steps
//     (5 of 128 warps per round at 80 000 points; restarting on every touch would hit 60).
//
// Two-level merge: a CTA ranks its 8 x (D+1) warp entries and publishes the top KC (+ the next one as its
// terminal) to all CTAs of the cluster with st.async + mbarrier (32-byte entries); every CTA then ranks the
// 16 x (KC+1) published entries cooperatively (all warps, counting ranks), lays out the top 32 in order and
// tests the 496 pairs in parallel.  Keys are made unique, so ranks are positions:
//   key = value bits << 32 | (0x7fffffff - idx) << 1 | real       (terminal copies sort right after the entry)
// Duplicate points (all remaining min-distances 0) repeat the same index, as plain FPS does: entries of value
// 0 beyond the head of a list are published as terminals, so such clouds advance one sample per warp and round.
#pragma once

struct __align__(16) FpsEnt {   // 32 bytes = two 16-byte halves {key, owner} {x, y, z}
    unsigned klo, khi;          // khi: float bits of the min-distance; klo: ((0x7fffffff - idx) << 1) | real
    int owner;                  // cluster-wide warp id the point lives in (-1: none)
    int pad;
    float x, y, z, w;
};
static_assert(sizeof(FpsEnt) == 32, "FpsEnt layout");

__device__ __forceinline__ unsigned long long ent_key(const FpsEnt* e) {
    const uint2 k = *reinterpret_cast<const uint2*>(e);
    return ((unsigned long long)k.y << 32) | k.x;
}
__device__ __forceinline__ FpsEnt make_ent(unsigned bits, int idx, float x, float y, float z, int owner, bool real) {
    FpsEnt e;
    e.khi = idx < 0 ? 0u : bits;
    e.klo = idx < 0 ? 0u : (((0x7fffffffu - (unsigned)idx) << 1) | (real ? 1u : 0u));
    e.owner = owner; e.pad = 0;
    e.x = x; e.y = y; e.z = z; e.w = 0.f;
    return e;
}
__device__ __forceinline__ void ent_store(FpsEnt* dst, const FpsEnt& e) {
    reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(&e)[0];
    reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(&e)[1];
}
__device__ __forceinline__ FpsEnt ent_load(const FpsEnt* src) {
    FpsEnt e;
    reinterpret_cast<uint4*>(&e)[0] = reinterpret_cast<const uint4*>(src)[0];
    reinterpret_cast<uint4*>(&e)[1] = reinterpret_cast<const uint4*>(src)[1];
    return e;
}

// argmax of a warp's speculative min-distances: value bits, lowest ORIGINAL index among the maxima, its
// coordinates (from the float4 {x, y, z, idx} copy in shared memory).  ni = -1: the warp holds no point.
template <int P>
__device__ __forceinline__ void fps_warp_argmax(const float (&st)[P], const float4* __restrict__ warp_pts, int lane,
                                                unsigned& nb, int& ni, float& nx, float& ny, float& nz) {
    float m4[4] = {0.f, 0.f, 0.f, 0.f};   // padding slots hold -1 and never win; four chains, not one of length P
#pragma unroll
    for (int p = 0; p < P; p++) m4[p & 3] = fmaxf(m4[p & 3], st[p]);
    const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    const unsigned mb = __float_as_uint(m);
    nb = __reduce_max_sync(FULL, mb);
    int cand = INT_MAX, cp = 0;
    if (mb == nb) {   // usually one lane
#pragma unroll
        for (int p = 0; p < P; p++) {
            if (__float_as_uint(st[p]) == nb) {
                const int gi = __float_as_int(warp_pts[p * 32 + lane].w);
                if (gi < cand) { cand = gi; cp = p; }
            }
        }
    }
    ni = __reduce_min_sync(FULL, cand);
    const unsigned own = __ballot_sync(FULL, cand == ni && ni != INT_MAX);
    if (own == 0u) { nb = 0u; ni = -1; nx = ny = nz = 0.f; return; }
    const int slot = __shfl_sync(FULL, cp * 32 + lane, __ffs(own) - 1);
    const float4 q = warp_pts[slot];
    nx = q.x; ny = q.y; nz = q.z;
}

// rows of register slots the sample (sx, sy, sz) can lower, given that every value in the warp is <= bound:
// lane p tests row p's box (same conservative rounding as the warp-level test)
template <int P>
__device__ __forceinline__ unsigned fps_row_mask(const float (&rlo)[3], const float (&rhi)[3], float sx, float sy, float sz,
                                                 float bound, int lane) {
    const float ex = fmaxf(fmaxf(rlo[0] - sx, sx - rhi[0]), 0.f);
    const float ey = fmaxf(fmaxf(rlo[1] - sy, sy - rhi[1]), 0.f);
    const float ez = fmaxf(fmaxf(rlo[2] - sz, sz - rhi[2]), 0.f);
    const float b2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
    return __ballot_sync(FULL, lane < P && !(b2 * 0.99999f >= bound));
}

// d2 of this lane's point of row p against (sx, sy, sz): coordinates from registers, or (SP) from shared memory
#define POB_FPS_D2(p, sx, sy, sz) \
    (SP ? [&] { const float4 q_ = warp_pts[(p) * 32 + lane]; return d2_ref(q_.x, q_.y, q_.z, sx, sy, sz); }() \
        : d2_ref(px[SP ? 0 : (p)], py[SP ? 0 : (p)], pz[SP ? 0 : (p)], sx, sy, sz))

// for every row p covered by a set bit of `mask` (warp-uniform; bit b covers the RG consecutive rows b*RG .. b*RG+RG-1):
// BODY with the compile-time index p; bits are skipped four at a time.  NB = number of bits = P / RG.
#define POB_FPS_FOR_ROWS(mask, BODY)                                                          \
    _Pragma("unroll") for (int g_ = 0; g_ < NB; g_ += 4) {                                    \
        if ((mask) & (0xFu << g_)) {                                                          \
            _Pragma("unroll") for (int b_ = g_; b_ < (g_ + 4 < NB ? g_ + 4 : NB); b_++) {     \
                if ((mask) & (1u << b_)) {                                                    \
                    _Pragma("unroll") for (int r_ = 0; r_ < RG; r_++) {                       \
                        const int p = b_ * RG + r_;                                           \
                        BODY                                                                  \
                    }                                                                         \
                }                                                                             \
            }                                                                                 \
        }                                                                                     \
    }

// One cluster of C CTAs (T = 256 threads) per scene; P points per thread in registers; D list entries per warp;
// KC entries per CTA and exchange.  Dynamic shared memory: float4 {x, y, z, idx}[T * P].
//
// GX = true is the GRID-WIDE form for scenes beyond what one cluster holds in registers (131 072 points): one
// cooperative launch of G <= 148 CTAs per scene (1.2 M points at 32 per thread), same lists, same merge rule.  The
// exchange goes through global memory instead of DSMEM -- every CTA writes its message to gmsg[parity][cta],
// arrives on a monotonic counter (release) and spins until all G have (acquire); then all G * (KC + 1) entries are
// staged in shared memory.  Ranking 740 entries by counting would cost 2 000 compares per thread, so it is
// thresholded first: T* = the largest "first terminal" key of any CTA is where the walk must stop anyway, only
// real entries above it (a few dozen) are compacted and ranked.
#ifndef FPS_ONE_STEP
#define FPS_ONE_STEP 1
#endif
template <int P, int T, int D, int KC, bool GX, bool SP>
__device__ __forceinline__ void
fps_merge_body(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
               const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
               const float4* __restrict__ sorted, int* __restrict__ idx, unsigned long long* __restrict__ stats,
               int scene_arg, int cap_points, FpsEnt* __restrict__ gmsg, unsigned* __restrict__ gcounter) {
    int C, rank, scene;
    if constexpr (GX) {
        C = (int)gridDim.x; rank = (int)blockIdx.x; scene = scene_arg;
    } else {
        C = (int)cg::this_cluster().num_blocks();
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        scene = blockIdx.x / C;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;                 // warps per CTA
    constexpr int WL = D + 1;                  // entries of a warp list (D real + terminal)
    constexpr int CE = NW * WL;                // entries ranked per CTA
    constexpr int ME = KC + 1;                 // entries per CTA message (KC + its terminal)
    constexpr int NSLOT = FPS_MAX_CLUSTER * ME;
    constexpr int LMAX = 32;                   // samples accepted per exchange, at most
    constexpr int RG = P > 32 ? 2 : 1;         // rows per pruning box (one box per lane: P / RG <= 32 boxes)
    constexpr int NB = P / RG;
    static_assert(P % RG == 0 && NB <= 32, "one row box per lane");
    constexpr unsigned ALLROWS = NB < 32 ? (1u << NB) - 1u : FULL;
    static_assert(T == 256, "the pair test maps 256 threads onto 32 x 8 (j, i mod 8)");
    static_assert((NW & (NW - 1)) == 0 && WL * NW <= 32 && CE >= ME, "CTA ranking: NW lanes per warp-list entry");
    static_assert(NSLOT % NW == 0 && NSLOT / NW <= 32, "cluster ranking: each warp ranks NSLOT / NW entries");
    const int gw = rank * NW + warp;           // cluster-wide warp id = owner tag

    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];

    extern __shared__ __align__(16) float4 s_pts[];            // [T * P] {x, y, z, idx}; GX: then FpsEnt[C * ME], u16[C * ME]
    __shared__ __align__(16) FpsEnt s_wl[CE];                  // warp lists, WL entries each
    __shared__ __align__(16) FpsEnt s_cs[ME];                  // this CTA's message
    __shared__ __align__(16) FpsEnt s_msg_store[GX ? 1 : 2 * NSLOT];   // cluster form: messages of all CTAs, by round parity
    FpsEnt (*s_msg)[NSLOT] = reinterpret_cast<FpsEnt (*)[NSLOT]>(s_msg_store);
    FpsEnt* const s_all = reinterpret_cast<FpsEnt*>(s_pts + T * P);                    // GX: all C * ME entries of the round
    unsigned short* const s_cand = reinterpret_cast<unsigned short*>(s_all + C * ME);   // GX: indices of the entries above T*
    __shared__ unsigned long long s_red[NW];
    __shared__ int s_cnt;
    __shared__ __align__(16) FpsEnt s_sorted[LMAX];            // merged order (top LMAX); [0, L) = accepted samples
    __shared__ __align__(16) unsigned s_failw[2][NW];          // per warp: which of its four positions are not accepted
    __shared__ __align__(8) unsigned long long s_bar[2];

    if (e_m <= s_m || e_n <= s_n) return;  // uniform over the cluster / grid
    const int n = e_n - s_n;
    const int total = e_m - s_m;
    // a call with scenes on both sides of the cluster capacity launches both forms; each takes its own scenes
    if (GX ? n <= cap_points : n > C * T * P) return;

    // the points arrive spatially ordered with their original index in .w: either the kNN grid's cell-sorted array
    // (cell_start != NULL: the scene starts at its first cell) or the launcher's Hilbert-ordered copy (cell_start == NULL:
    // every scene sits at its own offset)
    const bool in_cells = sorted != nullptr && (cell_start == nullptr || (scenes != nullptr && scenes[scene].use_grid));
    // SP (shared-memory points): the coordinates are read from the float4 copy in shared memory where they are used
    // instead of living in 3 * P registers -- with row pruning an update touches a few rows only, so the extra LDS.128
    // per row is noise, and at <= 128 registers a second CTA (another room's FPS cluster, or the feature path's
    // kernels) fits on the SM next to this latency-bound one
    constexpr int PR = SP ? 1 : P;
    float px[PR], py[PR], pz[PR], pt[P], st[P];
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    // second pruning level: lane b < NB keeps the bounding box of rows b*RG .. b*RG+RG-1, a ROW being the 32 consecutive
    // points (one per lane) of a register slot -- in cell order a patch of a few cells, far tighter than the warp's box
    float rlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool box_ok = true;
    const int sbase = !in_cells ? 0 : (cell_start ? __ldg(cell_start + scenes[scene].cell_base) : s_n);
    float4* const warp_pts = s_pts + warp * P * 32;
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
        st[p] = pt[p];
        if constexpr (!SP) { px[p] = vx; py[p] = vy; pz[p] = vz; }
        warp_pts[p * 32 + lane] = make_float4(vx, vy, vz, __int_as_float(vi));
        {
            const bool real = pos < n;
            float l3[3] = {real ? vx : FLT_MAX, real ? vy : FLT_MAX, real ? vz : FLT_MAX};
            float h3[3] = {real ? vx : -FLT_MAX, real ? vy : -FLT_MAX, real ? vz : -FLT_MAX};
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    l3[a] = fminf(l3[a], __shfl_xor_sync(FULL, l3[a], o));
                    h3[a] = fmaxf(h3[a], __shfl_xor_sync(FULL, h3[a], o));
                }
            }
            if (lane == p / RG) {
#pragma unroll
                for (int a = 0; a < 3; a++) { rlo[a] = fminf(rlo[a], l3[a]); rhi[a] = fmaxf(rhi[a], h3[a]); }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            blo[a] = fminf(blo[a], __shfl_xor_sync(FULL, blo[a], o));
            bhi[a] = fmaxf(bhi[a], __shfl_xor_sync(FULL, bhi[a], o));
        }
    }
    const bool prune = in_cells && __all_sync(FULL, box_ok);

    // every slot starts EMPTY: key 0 (below every real key), terminal, owner -1
    {
        const uint4 z0 = make_uint4(0u, 0u, 0xffffffffu, 0u), z1 = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (!GX)
            for (int i = tid; i < 2 * NSLOT; i += T) { reinterpret_cast<uint4*>(&s_msg[0][0])[2 * i] = z0; reinterpret_cast<uint4*>(&s_msg[0][0])[2 * i + 1] = z1; }
        if (tid == 0) s_cnt = 0;
        for (int i = tid; i < CE; i += T) { reinterpret_cast<uint4*>(s_wl)[2 * i] = z0; reinterpret_cast<uint4*>(s_wl)[2 * i + 1] = z1; }
        for (int i = tid; i < LMAX; i += T) { reinterpret_cast<uint4*>(s_sorted)[2 * i] = z0; reinterpret_cast<uint4*>(s_sorted)[2 * i + 1] = z1; }
        for (int i = tid; i < ME; i += T) { reinterpret_cast<uint4*>(s_cs)[2 * i] = z0; reinterpret_cast<uint4*>(s_cs)[2 * i + 1] = z1; }
    }
    __syncthreads();
    // the first sample of a scene is its first point (sampling_cuda_kernel.cu:39): "accepted" by nobody's list
    if (tid == 0) {
        const FpsEnt e0 = make_ent(__float_as_uint(PLACEHOLDER_D2), s_n, __ldg(xyz + (int64_t)s_n * 3), __ldg(xyz + (int64_t)s_n * 3 + 1),
                                   __ldg(xyz + (int64_t)s_n * 3 + 2), -1, true);
        ent_store(&s_sorted[0], e0);
    }
    const bool writer = rank == 0 && warp == 0;
    if (writer && lane == 0) idx[s_m] = s_n;
    const unsigned bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    unsigned rslot0 = 0, rslot1 = 0, rbar0 = 0, rbar1 = 0;
    if (!GX && C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (tid == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar0, C * ME * (int)sizeof(FpsEnt));
            mbar_expect_tx(bar1, C * ME * (int)sizeof(FpsEnt));
        }
        if (warp == 0 && (lane & 15) < C) {   // lanes l and l + 16 talk to CTA l (first / second half of every entry)
            rslot0 = mapa_u32(smem_u32(&s_msg[0][rank * ME]), lane & 15) + 16u * (unsigned)(lane >> 4);
            rslot1 = mapa_u32(smem_u32(&s_msg[1][rank * ME]), lane & 15) + 16u * (unsigned)(lane >> 4);
            rbar0 = mapa_u32(bar0, lane & 15);
            rbar1 = mapa_u32(bar1, lane & 15);
        }
        cluster.sync();
    } else {
        __syncthreads();
    }

    int L = 1;            // accepted, not yet applied samples: s_sorted[0 .. L)
    int emitted = 1;
    int len = 0;          // entries of this warp's list in s_wl (the terminal sits at position len)
    unsigned nb = 0u; int ni = -1; float nx = 0.f, ny = 0.f, nz = 0.f;   // argmax of st: the next entry to append
    float wmaxf = PLACEHOLDER_D2;   // this warp's real maximum (value of its list head)
    bool first = true;
    bool stale = true;              // nxt no longer is the argmax of st (a foreign sample lowered that point)
    unsigned long long n_rounds = 0, n_evals = 0;   // diagnostics
    int* out = idx + s_m;
    FpsEnt* const my_wl = s_wl + warp * WL;

    for (int round = 0; emitted < total; round++) {
        const int par = round & 1;
        n_rounds++;
        // ================= phase A: apply the accepted samples, keep the local list current =================
        {
            const int4 h = reinterpret_cast<const int4*>(&s_sorted[lane])[0];
            const float4 c = reinterpret_cast<const float4*>(&s_sorted[lane])[1];
            const bool valid = lane < L;
            bool touch = valid;
            if (prune && !first) {
                const float ex = fmaxf(fmaxf(blo[0] - c.x, c.x - bhi[0]), 0.f);
                const float ey = fmaxf(fmaxf(blo[1] - c.y, c.y - bhi[1]), 0.f);
                const float ez = fmaxf(fmaxf(blo[2] - c.z, c.z - bhi[2]), 0.f);
                const float b2 = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
                // every d2_ref(point, sample) >= b2 * (1 - 1e-6); if even b2 * 0.99999 >= max tmp nothing changes
                touch = valid && !(b2 * 0.99999f >= wmaxf);
            }
            const bool own = valid && h.z == gw;
            unsigned tm = __ballot_sync(FULL, touch);
            const unsigned om = __ballot_sync(FULL, own);
            const int c_own = __popc(om);
            // the merge took this warp's first c_own entries: the list shifts (st has them applied already)
            if (c_own && !first) {
                FpsEnt t;
                const bool mv = lane + c_own <= len;
                if (mv) t = ent_load(my_wl + lane + c_own);
                __syncwarp();
                if (mv) {
                    if (lane == 0 && len - c_own > 0) t.klo |= 1u;   // a value-0 entry that was parked as terminal heads the list now
                    ent_store(my_wl + lane, t);
                } else if (lane < WL) {
                    ent_store(my_wl + lane, make_ent(0u, -1, 0.f, 0.f, 0.f, -1, false));
                }
                len -= c_own;
                __syncwarp();
            }
            // lane i <= len watches entry i of the list (i == len: the terminal = nxt) against the foreign samples
            float wx = 0.f, wy = 0.f, wz = 0.f, wv = -1.f;
            if (lane <= len && !first) {
                const FpsEnt* e = my_wl + lane;
                const float4 wc = reinterpret_cast<const float4*>(e)[1];
                wx = wc.x; wy = wc.y; wz = wc.z;
                wv = (e->klo | e->khi) ? __uint_as_float(e->khi) : -1.f;   // EMPTY slots are never "lowered"
            }
            bool lowered = false;
            const unsigned fm = tm & ~om;
            while (tm) {
                const int j = __ffs(tm) - 1;
                const bool foreign = (fm >> j) & 1u;
                tm &= tm - 1;
                const float4 s = reinterpret_cast<const float4*>(&s_sorted[j])[1];
                const unsigned rmask = (prune && !first) ? fps_row_mask<NB>(rlo, rhi, s.x, s.y, s.z, wmaxf, lane) : ALLROWS;
                if (foreign) {
                    lowered = lowered || d2_ref(wx, wy, wz, s.x, s.y, s.z) < wv;
                    POB_FPS_FOR_ROWS(rmask, {
                        const float d = POB_FPS_D2(p, s.x, s.y, s.z);
                        pt[p] = fminf(d, pt[p]);
                        st[p] = fminf(d, st[p]);
                    })
                } else {
                    POB_FPS_FOR_ROWS(rmask, { pt[p] = fminf(POB_FPS_D2(p, s.x, s.y, s.z), pt[p]); })
                }
                if (stats) n_evals += 32u * RG * (unsigned)__popc(rmask);
            }
            const unsigned lm = __ballot_sync(FULL, lowered);
            const bool restart = first || (lm & ((1u << len) - 1u)) != 0u;   // a pending entry was lowered: the list is void
            stale = stale || ((lm >> len) & 1u);   // nxt itself was lowered: its value still bounds, recompute before use
            int fill = 0;
            if (restart) {
#pragma unroll
                for (int p = 0; p < P; p++) st[p] = pt[p];
                len = 0;
                stale = true;
                fill = D;
            } else {
                fill = D - len;                                               // one local step per popped entry
            }
            // at most ONE local step per warp and round: a round waits for its slowest warp, and that warp is the one
            // refilling a void list -- the second entry follows next round (same samples per exchange in
            // tools/fps_merge_sim.py: 18.51 vs 18.58, the slowest warp's work -21 %)
            if (FPS_ONE_STEP && fill > 1) fill = 1;
            for (int d = 0; d < fill; d++) {
                if (stale) { fps_warp_argmax<P>(st, warp_pts, lane, nb, ni, nx, ny, nz); stale = false; }
                if (lane == 0) ent_store(my_wl + len, make_ent(nb, ni, nx, ny, nz, gw, !(nb == 0u && len > 0)));
                len++;
                {   // nb is the maximum of st: rows farther than that from the new local sample cannot change
                    const unsigned rmask = prune ? fps_row_mask<NB>(rlo, rhi, nx, ny, nz, __uint_as_float(nb), lane) : ALLROWS;
                    POB_FPS_FOR_ROWS(rmask, { st[p] = fminf(POB_FPS_D2(p, nx, ny, nz), st[p]); })
                    if (stats) n_evals += 32u * RG * (unsigned)__popc(rmask);
                }
                fps_warp_argmax<P>(st, warp_pts, lane, nb, ni, nx, ny, nz);
            }
            if (fill) {
                if (lane == 0) ent_store(my_wl + len, make_ent(nb, ni, nx, ny, nz, gw, false));   // terminal: bound on what follows
                else if (lane > len && lane < WL) ent_store(my_wl + lane, make_ent(0u, -1, 0.f, 0.f, 0.f, -1, false));   // a shorter list than before: the slots behind the terminal are EMPTY
                __syncwarp();
            }
            wmaxf = __uint_as_float(my_wl[0].khi);
            first = false;
        }
        __syncthreads();   // (1) warp lists visible

        // ================= phase B: rank the CTA's CE entries, top ME -> s_cs =================
        {
            const int e = lane / NW, q = lane % NW;     // NW lanes per entry of this warp's list, one per warp compared against
            const int myslot = warp * WL + e;
            int cnt = 0;
            if (e < WL) {
                const unsigned long long mk = ent_key(s_wl + myslot);
#pragma unroll
                for (int i = 0; i < WL; i++) {
                    const int os = q * WL + i;
                    const unsigned long long ok = ent_key(s_wl + os);
                    cnt += (ok > mk || (ok == mk && os < myslot)) ? 1 : 0;
                }
            }
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
            if (e < WL && q == 0 && cnt < ME) {
                FpsEnt t = ent_load(s_wl + myslot);
                if (cnt == ME - 1) t.klo &= ~1u;        // the (KC+1)-th is the CTA's terminal: bounds everything not published
                ent_store(s_cs + cnt, t);
            }
        }
        __syncthreads();   // (2) s_cs complete

        // ================= exchange =================
        if constexpr (GX) {
            if (warp == 0) {   // message -> global, then arrive (release): the stores of the whole warp precede lane 0's fence
                if (lane < 2 * ME) __stcg(reinterpret_cast<uint4*>(gmsg + ((size_t)par * C + rank) * ME) + lane, reinterpret_cast<const uint4*>(s_cs)[lane]);
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(gcounter, 1u);
                    const unsigned target = (unsigned)C * (unsigned)(round + 1);
                    unsigned seen;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gcounter) : "memory"); } while (seen < target);
                }
            }
            __syncthreads();   // all G messages of this round are visible
        } else if (C > 1) {
            if (warp == 0 && (lane & 15) < C) {
                const unsigned rs = par ? rslot1 : rslot0, rb = par ? rbar1 : rbar0;
#pragma unroll
                for (int e = 0; e < ME; e++) {
                    const uint4 v = reinterpret_cast<const uint4*>(s_cs + e)[lane >> 4];
                    st_async_v4(rs + 32u * (unsigned)e, v.x, v.y, v.z, v.w, rb);
                }
            }
            mbar_wait(par ? bar1 : bar0, (unsigned)(round >> 1) & 1u);
            if (tid == 0) mbar_expect_tx(par ? bar1 : bar0, C * ME * (int)sizeof(FpsEnt));  // re-arm for round + 2
        } else {
            if (tid < 2 * ME) reinterpret_cast<uint4*>(&s_msg[par][0])[tid] = reinterpret_cast<const uint4*>(s_cs)[tid];
            __syncthreads();
        }

        // ================= phase D1: rank the published entries, top LMAX -> s_sorted =================
        if constexpr (GX) {
            const FpsEnt* gsrc = gmsg + (size_t)par * C * ME;
            const int NS = C * ME;
            for (int i = tid; i < NS; i += T) {
                reinterpret_cast<uint4*>(s_all + i)[0] = __ldcg(reinterpret_cast<const uint4*>(gsrc + i));
                reinterpret_cast<uint4*>(s_all + i)[1] = __ldcg(reinterpret_cast<const uint4*>(gsrc + i) + 1);
            }
            // T* = max over CTAs of the key of the first terminal of their (sorted) message: the walk ends there at the latest
            unsigned long long tk = 0ull;
            for (int c = tid; c < C; c += T) {
                bool found = false;
#pragma unroll
                for (int e = 0; e < ME; e++) {
                    const uint2 k2 = __ldcg(reinterpret_cast<const uint2*>(gsrc + c * ME + e));
                    const unsigned long long k = ((unsigned long long)k2.y << 32) | k2.x;
                    if (!found && !(k2.x & 1u)) { found = true; if (k > tk) tk = k; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(FULL, tk, o); tk = v > tk ? v : tk; }
            if (lane == 0) s_red[warp] = tk;
            __syncthreads();   // (a) s_all and the warp maxima are in
            unsigned long long tstar = 0ull;
#pragma unroll
            for (int w = 0; w < NW; w++) tstar = s_red[w] > tstar ? s_red[w] : tstar;
            for (int i0 = 0; i0 < NS; i0 += T) {   // uniform trip count
                const int i = i0 + tid;
                bool elig = false;
                if (i < NS) { const unsigned long long k = ent_key(s_all + i); elig = (k & 1ull) && k > tstar; }
                const unsigned em = __ballot_sync(FULL, elig);
                int base = 0;
                if (lane == 0 && em) base = atomicAdd(&s_cnt, __popc(em));
                base = __shfl_sync(FULL, base, 0);
                if (elig) s_cand[base + __popc(em & ((1u << lane) - 1u))] = (unsigned short)i;
            }
            __syncthreads();   // (b) candidates compacted
            const int M = s_cnt;
            for (int ci = warp; ci < M; ci += NW) {   // a warp per candidate: count the candidates above it
                const FpsEnt* me = s_all + s_cand[ci];
                const unsigned long long mk = ent_key(me);
                int cnt = 0;
                for (int j = lane; j < M; j += 32) cnt += ent_key(s_all + s_cand[j]) > mk ? 1 : 0;
                const int tot = __reduce_add_sync(FULL, cnt);
                if (tot < LMAX && lane < 2) reinterpret_cast<uint4*>(s_sorted + tot)[lane] = reinterpret_cast<const uint4*>(me)[lane];
            }
            if (M < LMAX && tid < 2)   // the walk stops after the last candidate: an EMPTY terminal
                reinterpret_cast<uint4*>(s_sorted + M)[tid] = tid == 0 ? make_uint4(0u, 0u, 0xffffffffu, 0u) : make_uint4(0u, 0u, 0u, 0u);
        } else {
            constexpr int EPW = NSLOT / NW;             // entries ranked by each warp
            constexpr int LPE = 32 / EPW;               // lanes per entry
            constexpr int CMPN = (NSLOT + LPE - 1) / LPE;
            const int le = lane / LPE, sub = lane % LPE;
            const bool act = le < EPW;
            const FpsEnt* src = s_msg[par];
            const int g = warp * EPW + (act ? le : 0);
            const unsigned long long mk = ent_key(src + g);
            int cnt = 0;
            const int lo = sub * CMPN;
#pragma unroll
            for (int i = 0; i < CMPN; i++) {
                const int o = lo + i;
                if (o < NSLOT) cnt += ent_key(src + o) > mk ? 1 : 0;
            }
            int tot = cnt;
#pragma unroll
            for (int s = 1; s < LPE; s++) tot += __shfl_sync(FULL, cnt, (le * LPE + s) & 31);
            if (act && sub == 0 && tot < LMAX) ent_store(s_sorted + tot, ent_load(src + g));
        }
        __syncthreads();   // (3) merged order laid out

        // ================= phase D2: longest accepted prefix =================
        {
            const int j = tid >> 3, i0 = tid & 7;       // thread (j, i0) tests entry j against entries i0, i0 + 8, i0 + 16, i0 + 24
            const int4 hj = reinterpret_cast<const int4*>(&s_sorted[j])[0];
            const float4 cj = reinterpret_cast<const float4*>(&s_sorted[j])[1];
            const float vj = __int_as_float(hj.y);
            bool fail = i0 == 0 && !(hj.x & 1);         // a terminal ends the walk
#pragma unroll
            for (int t = 0; t < LMAX / 8; t++) {
                const int i = i0 + 8 * t;
                const int4 hi = reinterpret_cast<const int4*>(&s_sorted[i])[0];
                const float4 ci = reinterpret_cast<const float4*>(&s_sorted[i])[1];
                // exactly what the update computes for point j against sample i
                fail = fail || (i < j && hi.z != hj.z && d2_ref(cj.x, cj.y, cj.z, ci.x, ci.y, ci.z) < vj);
            }
            const unsigned fm = __ballot_sync(FULL, fail);   // 8 lanes per j, four j per warp
            if (lane == 0) {
                const unsigned nib = ((fm & 0xffu) ? 1u : 0u) | ((fm & 0xff00u) ? 2u : 0u) | ((fm & 0xff0000u) ? 4u : 0u) |
                                     ((fm & 0xff000000u) ? 8u : 0u);
                s_failw[par][warp] = nib << (4 * warp);
            }
        }
        __syncthreads();   // (4) every warp's verdict on its four positions is in
        int Ln;
        {
            const uint4 f0 = reinterpret_cast<const uint4*>(s_failw[par])[0], f1 = reinterpret_cast<const uint4*>(s_failw[par])[1];
            const unsigned fmask = f0.x | f0.y | f0.z | f0.w | f1.x | f1.y | f1.z | f1.w;   // bit j: position j is not accepted
            Ln = fmask ? __ffs(fmask) - 1 : LMAX;
        }
        if (GX && tid == 0) s_cnt = 0;   // next read is three barriers away
        if (Ln < 1) asm volatile("trap;");               // the top entry is always a real one: cannot happen
        if (Ln > total - emitted) Ln = total - emitted;
        if (writer && lane < Ln) out[emitted + lane] = 0x7fffffff - (int)(s_sorted[lane].klo >> 1);
        emitted += Ln;
        L = Ln;
    }
    if (stats) {   // diagnostics: rounds and samples per scene (mean chain = samples / rounds), point distances evaluated
        if (tid == 0 && rank == 0) {
            atomicAdd(stats, n_rounds);
            atomicAdd(stats + 1, (unsigned long long)(total - 1));
        }
        if (lane == 0) atomicAdd(stats + 2, n_evals);
    }
    if (!GX && C > 1) cg::this_cluster().sync();  // nobody leaves while a peer may still be writing into its shared memory
}

#define POB_FPS_MERGE_PARAMS                                                                                  \
    const float *__restrict__ xyz, const int *__restrict__ offset, const int *__restrict__ new_offset,           \
        const SceneGrid *__restrict__ scenes, const int *__restrict__ cell_start, const float4 *__restrict__ sorted, \
        int *__restrict__ idx, unsigned long long *__restrict__ stats
// register-resident points (small P: few registers anyway)
template <int P, int T, int D, int KC>
__global__ void __launch_bounds__(T, 1) fps_merge_kernel(POB_FPS_MERGE_PARAMS) {
    fps_merge_body<P, T, D, KC, false, false>(xyz, offset, new_offset, scenes, cell_start, sorted, idx, stats, 0, 0, nullptr, nullptr);
}
// shared-memory points, two CTAs per SM (<= 128 registers)
template <int P, int T, int D, int KC>
__global__ void __launch_bounds__(T, 2) fps_merge_sp_kernel(POB_FPS_MERGE_PARAMS) {
    fps_merge_body<P, T, D, KC, false, true>(xyz, offset, new_offset, scenes, cell_start, sorted, idx, stats, 0, 0, nullptr, nullptr);
}
// shared-memory points, one CTA per SM (P = 32: 64 registers of running minima alone)
template <int P, int T, int D, int KC>
__global__ void __launch_bounds__(T, 1) fps_merge_sp1_kernel(POB_FPS_MERGE_PARAMS) {
    fps_merge_body<P, T, D, KC, false, true>(xyz, offset, new_offset, scenes, cell_start, sorted, idx, stats, 0, 0, nullptr, nullptr);
}

// grid-wide form: cooperative launch of G CTAs for ONE scene; workspace = {counter (zeroed by the caller), gmsg[2][G][KC+1]}
template <int P, int T, int D, int KC>
__global__ void __launch_bounds__(T, 1)
fps_merge_grid_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                      const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                      const float4* __restrict__ sorted, int* __restrict__ idx, unsigned long long* __restrict__ stats,
                      int scene, int cap_points, FpsEnt* __restrict__ gmsg, unsigned* __restrict__ gcounter) {
    fps_merge_body<P, T, D, KC, true, true>(xyz, offset, new_offset, scenes, cell_start, sorted, idx, stats, scene, cap_points, gmsg, gcounter);
}
#undef POB_FPS_MERGE_PARAMS
