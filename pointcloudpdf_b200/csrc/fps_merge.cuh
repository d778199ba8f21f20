// fps_merge.cuh -- merged-list farthest point sampling (included by fps.cu inside namespace pob).
//
// FPS is "argmax, update, argmax, ..." -- one cluster-wide exchange per sample in the plain form.  The round-1
// chain kernel accepted ~4.5 samples per exchange (every CTA offers ONE candidate plus a bound on what would
// follow it).  This kernel generalises that to whole LISTS and reaches ~20 samples per exchange on S3DIS-shaped
// rooms (scratch/fps_merge_sim.py: 80 000 -> 20 000 points in 1 010 rounds instead of 4 444, output identical
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
//   * afterwards a warp applies to pt[] only the accepted samples whose reach intersects its bounding box
//     (exact, conservatively rounded -- as in the round-1 kernels).  If one of them came from another warp its
//     list is void: st = pt, D local steps again.  If only its own entries were taken the list just shifts.
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
    float m = 0.f;   // padding slots hold -1 and never win
#pragma unroll
    for (int p = 0; p < P; p++) m = fmaxf(m, st[p]);
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

// One cluster of C CTAs (T = 256 threads) per scene; P points per thread in registers; D list entries per warp;
// KC entries per CTA and exchange.  Dynamic shared memory: float4 {x, y, z, idx}[T * P].
template <int P, int T, int D, int KC>
__global__ void __launch_bounds__(T, 1)
fps_merge_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset,
                 const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                 const float4* __restrict__ sorted, int* __restrict__ idx, unsigned long long* __restrict__ stats) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    int rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int scene = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;                 // warps per CTA
    constexpr int WL = D + 1;                  // entries of a warp list (D real + terminal)
    constexpr int CE = NW * WL;                // entries ranked per CTA
    constexpr int ME = KC + 1;                 // entries per CTA message (KC + its terminal)
    constexpr int NSLOT = FPS_MAX_CLUSTER * ME;
    constexpr int LMAX = 32;                   // samples accepted per exchange, at most
    static_assert(T == 256, "the pair test maps 256 threads onto 32 x 8 (j, i mod 8)");
    static_assert((NW & (NW - 1)) == 0 && WL * NW <= 32 && CE >= ME, "CTA ranking: NW lanes per warp-list entry");
    static_assert(NSLOT % NW == 0 && NSLOT / NW <= 32, "cluster ranking: each warp ranks NSLOT / NW entries");
    const int gw = rank * NW + warp;           // cluster-wide warp id = owner tag

    const int s_n = scene == 0 ? 0 : offset[scene - 1], e_n = offset[scene];
    const int s_m = scene == 0 ? 0 : new_offset[scene - 1], e_m = new_offset[scene];

    extern __shared__ __align__(16) float4 s_pts[];            // [T * P] {x, y, z, idx}
    __shared__ __align__(16) FpsEnt s_wl[CE];                  // warp lists, WL entries each
    __shared__ __align__(16) FpsEnt s_cs[ME];                  // this CTA's message
    __shared__ __align__(16) FpsEnt s_msg[2][NSLOT];           // messages of all CTAs, by round parity
    __shared__ __align__(16) FpsEnt s_sorted[LMAX];            // merged order (top LMAX); [0, L) = accepted samples
    __shared__ int s_fail[2];                                  // first position that is not accepted, by parity
    __shared__ __align__(8) unsigned long long s_bar[2];

    if (e_m <= s_m || e_n <= s_n) return;  // uniform over the cluster
    const int n = e_n - s_n;
    const int total = e_m - s_m;

    const bool in_cells = scenes != nullptr && scenes[scene].use_grid;
    float px[P], py[P], pz[P], pt[P], st[P];
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool box_ok = true;
    const int sbase = in_cells ? __ldg(cell_start + scenes[scene].cell_base) : 0;
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
        px[p] = vx; py[p] = vy; pz[p] = vz;
        warp_pts[p * 32 + lane] = make_float4(vx, vy, vz, __int_as_float(vi));
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
        for (int i = tid; i < 2 * NSLOT; i += T) { reinterpret_cast<uint4*>(&s_msg[0][0])[2 * i] = z0; reinterpret_cast<uint4*>(&s_msg[0][0])[2 * i + 1] = z1; }
        for (int i = tid; i < CE; i += T) { reinterpret_cast<uint4*>(s_wl)[2 * i] = z0; reinterpret_cast<uint4*>(s_wl)[2 * i + 1] = z1; }
        for (int i = tid; i < LMAX; i += T) { reinterpret_cast<uint4*>(s_sorted)[2 * i] = z0; reinterpret_cast<uint4*>(s_sorted)[2 * i + 1] = z1; }
        for (int i = tid; i < ME; i += T) { reinterpret_cast<uint4*>(s_cs)[2 * i] = z0; reinterpret_cast<uint4*>(s_cs)[2 * i + 1] = z1; }
        if (tid < 2) s_fail[tid] = LMAX;
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
    if (C > 1) {
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
    unsigned long long n_rounds = 0;
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
            const bool restart = first || (tm & ~om) != 0u;
            while (tm) {
                const int j = __ffs(tm) - 1;
                tm &= tm - 1;
                const float4 s = reinterpret_cast<const float4*>(&s_sorted[j])[1];
#pragma unroll
                for (int p = 0; p < P; p++) pt[p] = fminf(d2_ref(px[p], py[p], pz[p], s.x, s.y, s.z), pt[p]);
            }
            const int c_own = __popc(om);
            int fill = 0;
            if (restart) {
#pragma unroll
                for (int p = 0; p < P; p++) st[p] = pt[p];
                len = 0;
                fps_warp_argmax<P>(st, warp_pts, lane, nb, ni, nx, ny, nz);
                fill = D;
            } else if (c_own) {
                // its own head entries were taken (and nobody else reached in): the list shifts, st stays valid
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
                if (len == 0) fill = D;
                __syncwarp();
            }
            for (int d = 0; d < fill; d++) {
                if (lane == 0) ent_store(my_wl + len, make_ent(nb, ni, nx, ny, nz, gw, !(nb == 0u && len > 0)));
                len++;
#pragma unroll
                for (int p = 0; p < P; p++) st[p] = fminf(d2_ref(px[p], py[p], pz[p], nx, ny, nz), st[p]);
                fps_warp_argmax<P>(st, warp_pts, lane, nb, ni, nx, ny, nz);
            }
            if (fill) {
                if (lane == 0) ent_store(my_wl + len, make_ent(nb, ni, nx, ny, nz, gw, false));   // terminal: bound on what follows
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
        if (C > 1) {
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

        // ================= phase D1: rank all NSLOT published entries, top LMAX -> s_sorted =================
        {
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
            const int j = tid >> 3, i0 = tid & 7;
            const int4 hj = reinterpret_cast<const int4*>(&s_sorted[j])[0];
            const float4 cj = reinterpret_cast<const float4*>(&s_sorted[j])[1];
            const float vj = __int_as_float(hj.y);
            bool fail = i0 == 0 && !(hj.x & 1);         // a terminal ends the walk
            for (int i = i0; i < j; i += 8) {
                const int4 hi = reinterpret_cast<const int4*>(&s_sorted[i])[0];
                const float4 ci = reinterpret_cast<const float4*>(&s_sorted[i])[1];
                // exactly what the update computes for point j against sample i
                if (hi.z != hj.z && d2_ref(cj.x, cj.y, cj.z, ci.x, ci.y, ci.z) < vj) fail = true;
            }
            const unsigned fm = __ballot_sync(FULL, fail);
            if (fm && lane == 0) atomicMin(&s_fail[par], (warp * 32 + __ffs(fm) - 1) >> 3);
        }
        __syncthreads();   // (4) s_fail final
        int Ln = s_fail[par];
        if (tid == 0) s_fail[par ^ 1] = LMAX;
        if (Ln < 1) asm volatile("trap;");               // the top entry is always a real one: cannot happen
        if (Ln > total - emitted) Ln = total - emitted;
        if (writer && lane < Ln) out[emitted + lane] = 0x7fffffff - (int)(s_sorted[lane].klo >> 1);
        emitted += Ln;
        L = Ln;
    }
    if (stats && tid == 0 && rank == 0) {   // diagnostics: rounds and samples per scene (mean chain = samples / rounds)
        atomicAdd(stats, n_rounds);
        atomicAdd(stats + 1, (unsigned long long)(total - 1));
    }
    if (C > 1) cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}
