// knn.cu -- exact k-nearest-neighbour query on grid-binned candidates (sm_100a).
//
// Replaces knn_query_cuda_kernel (libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104):
// one *thread* per query brute-forcing its whole scene with a local-memory heap, by
//   1. a per-scene uniform grid built on the device (bbox -> cell size -> counting sort of
//      the points into cell order, stored as float4 {x,y,z,idx}), and
//   2. one *warp* per query scanning the cube of cells around it with coalesced float4
//      loads; the k best live in registers as a sorted list striped across the 32 lanes
//      (k <= 32*KPL) and are updated with ballot + shuffle insertions.
// Exactness: candidates are compared on the key (d2, idx) with d2 from the single FMA
// chain pob::d2_ref, so the result is independent of visiting order; the cube radius R
// doubles until the k-th best d2 is provably below the squared distance to any unvisited
// cell, with the cell-assignment rounding accounted for (see bound2()).  FP32 on CUDA
// cores on purpose: tensor cores would change the rounding of d2 and with it the indices.
#include "common.cuh"
#include "grid.cuh"

namespace pob {

// ---- order-preserving float <-> int key, for atomicMin/atomicMax on coordinates ----
__device__ __forceinline__ int fkey(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// the one cell-coordinate function, shared by build and query (monotone in v)
__device__ __forceinline__ int cell_coord(float v, float lo, float inv_h, int dim) {
    const int c = __float2int_rd(__fmul_rn(__fsub_rn(v, lo), inv_h));
    return min(max(c, 0), dim - 1);
}

// ------------------------------------------------------------------- build kernels --

__global__ void grid_init_kernel(int* __restrict__ bbox, int b, int* __restrict__ cnt, int64_t cap1,
                                 int* __restrict__ tiles, int64_t ntiles) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = t; i < cap1; i += stride) cnt[i] = 0;
    for (int64_t i = t; i < ntiles; i += stride) tiles[i] = 0;
    for (int64_t i = t; i < (int64_t)b * 6; i += stride) bbox[i] = (i % 6) < 3 ? INT_MAX : INT_MIN;
}

__global__ void __launch_bounds__(256) grid_bbox_kernel(const float* __restrict__ xyz, const int* __restrict__ offset,
                                                        int* __restrict__ bbox) {
    const int s = blockIdx.y;
    const int start = s == 0 ? 0 : offset[s - 1], end = offset[s];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = start + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float v = __ldg(xyz + (int64_t)i * 3 + a);
            if (isfinite(v)) { mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
        }
    }
    __shared__ float red[6][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = mn[a], hi = mx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(FULL, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o));
        }
        if (lane == 0) { red[a][w] = lo; red[3 + a][w] = hi; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int a = threadIdx.x;
        float v = red[a][0];
        for (int k = 1; k < 8; k++) v = a < 3 ? fminf(v, red[a][k]) : fmaxf(v, red[a][k]);
        if (a < 3) { if (v != FLT_MAX) atomicMin(bbox + s * 6 + a, fkey(v)); }
        else       { if (v != -FLT_MAX) atomicMax(bbox + s * 6 + a, fkey(v)); }
    }
}

// one thread: per-scene cell size and dims, cells of all scenes laid out back to back.
__global__ void grid_params_kernel(const int* __restrict__ offset, int b, const int* __restrict__ bbox,
                                   float cell_pts, int64_t cap, SceneGrid* __restrict__ scenes) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t base = 0;
    for (int s = 0; s < b; s++) {
        SceneGrid g;
        g.start = s == 0 ? 0 : offset[s - 1];
        g.end = offset[s];
        g.pad0 = g.pad1 = g.pad2 = g.pad3 = 0;
        const int nb = g.end - g.start;
        const int* bb = bbox + s * 6;
        const bool has_box = bb[0] != INT_MAX && bb[1] != INT_MAX && bb[2] != INT_MAX;
        g.lox = g.loy = g.loz = 0.f; g.inv_h = 1.f; g.h = 1.f; g.dx = g.dy = g.dz = 1; g.cell_base = (int)base;
        g.use_grid = 0;
        if (nb > BRUTE_MAX_POINTS && has_box) {
            const float lo[3] = {fkey_inv(bb[0]), fkey_inv(bb[1]), fkey_inv(bb[2])};
            const float L[3] = {fkey_inv(bb[3]) - lo[0], fkey_inv(bb[4]) - lo[1], fkey_inv(bb[5]) - lo[2]};
            const float budget = floorf((float)nb / cell_pts) + 1.f;
            const float Lmax = fmaxf(L[0], fmaxf(L[1], L[2]));
            const float pair = fmaxf(L[0] * L[1], fmaxf(L[0] * L[2], L[1] * L[2]));
            float h = fmaxf(cbrtf(L[0] * L[1] * L[2] / budget), fmaxf(sqrtf(pair / budget), Lmax / budget));
            if (!(h > 0.f) || !isfinite(h)) h = 1.f;
            int d[3];
            for (int it = 0; it < 4096; it++) {
                float cells = 1.f;
                for (int a = 0; a < 3; a++) {
                    float da = floorf(L[a] / h) + 1.f;
                    if (!(da >= 1.f)) da = 1.f;
                    if (da > (float)MAX_DIM) da = (float)MAX_DIM + 1.f;  // force another growth step
                    d[a] = (int)da;
                    cells *= da;
                }
                if (cells <= budget && d[0] <= MAX_DIM && d[1] <= MAX_DIM && d[2] <= MAX_DIM) break;
                h *= 1.1f;
                if (!isfinite(h)) { d[0] = d[1] = d[2] = 1; h = 1.f; break; }
            }
            const int64_t cells = (int64_t)d[0] * d[1] * d[2];
            if (base + cells <= cap) {
                g.lox = lo[0]; g.loy = lo[1]; g.loz = lo[2];
                g.h = h; g.inv_h = 1.f / h;
                g.dx = d[0]; g.dy = d[1]; g.dz = d[2];
                g.use_grid = 1;
                base += cells;
            }
        }
        scenes[s] = g;
    }
}

__global__ void __launch_bounds__(256) grid_count_kernel(int64_t n, int b, const float* __restrict__ xyz,
                                                         const int* __restrict__ offset,
                                                         const SceneGrid* __restrict__ scenes,
                                                         int* __restrict__ pcell, int* __restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = segment_of(i, offset, b);
        const SceneGrid g = scenes[s];
        int cell = -1;
        if (g.use_grid && i < g.end) {
            const float x = __ldg(xyz + i * 3), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
            const int cx = cell_coord(x, g.lox, g.inv_h, g.dx);
            const int cy = cell_coord(y, g.loy, g.inv_h, g.dy);
            const int cz = cell_coord(z, g.loz, g.inv_h, g.dz);
            cell = g.cell_base + (cz * g.dy + cy) * g.dx + cx;
            atomicAdd(cnt + cell, 1);
        }
        pcell[i] = cell;
    }
}

// exclusive scan of cnt[0..total) -> start[0..total), in three phases over SCAN_TILE tiles
__device__ __forceinline__ int block_exclusive_scan_512(int v, int* total_out) {
    __shared__ int wsum[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < 16 ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(FULL, s, o);
            if (lane >= o) s += t;
        }
        if (lane < 16) wsum[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const int wprefix = w == 0 ? 0 : wsum[w - 1];
    if (total_out) *total_out = wsum[15];
    __syncthreads();
    return wprefix + inc - v;
}

__global__ void __launch_bounds__(512) scan_tile_sums_kernel(const int* __restrict__ cnt, int64_t total,
                                                             int* __restrict__ tiles) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + k < total) s += cnt[base + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    __shared__ int ws[16];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 16; k++) t += ws[k];
        tiles[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(512) scan_tiles_kernel(int* __restrict__ tiles, int64_t ntiles) {
    int carry = 0;
    for (int64_t base = 0; base < ntiles; base += 512) {
        const int64_t i = base + threadIdx.x;
        const int v = i < ntiles ? tiles[i] : 0;
        int tot;
        const int ex = block_exclusive_scan_512(v, &tot);
        if (i < ntiles) tiles[i] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(512) scan_apply_kernel(const int* __restrict__ cnt, int64_t total,
                                                         const int* __restrict__ tiles, int* __restrict__ start) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base + k < total ? cnt[base + k] : 0; s += v[k]; }
    int ex = block_exclusive_scan_512(s, nullptr) + tiles[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < total) start[base + k] = ex; ex += v[k]; }
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(int64_t n, const float* __restrict__ xyz,
                                                           const int* __restrict__ pcell, int* __restrict__ cnt,
                                                           const int* __restrict__ start, float4* __restrict__ sorted) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int cell = pcell[i];
        if (cell < 0) continue;
        const int r = atomicSub(cnt + cell, 1) - 1;  // order inside a cell is irrelevant to the result
        sorted[start[cell] + r] = make_float4(__ldg(xyz + i * 3), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2),
                                              __int_as_float((int)i));
    }
}

// ------------------------------------------------------------- warp-wide top-k list --

template <int KPL>
struct TopK {
    float d[KPL];
    int i[KPL];
};

// strict lexicographic (d2, idx) order; idx are unique so there are no equal keys
__device__ __forceinline__ bool key_less(float da, int ia, float db, int ib) {
    return da < db || (da == db && ia < ib);
}

template <int KPL>
__device__ __forceinline__ void topk_insert(TopK<KPL>& t, float cd, int ci, int lane) {
    int pos = 0;  // number of residents ordered before the candidate
#pragma unroll
    for (int r = 0; r < KPL; r++) pos += __popc(__ballot_sync(FULL, key_less(t.d[r], t.i[r], cd, ci)));
#pragma unroll
    for (int r = KPL - 1; r >= 0; r--) {
        float pd = __shfl_up_sync(FULL, t.d[r], 1);
        int pi = __shfl_up_sync(FULL, t.i[r], 1);
        if (r > 0) {
            const float wd = __shfl_sync(FULL, t.d[r - 1], 31);
            const int wi = __shfl_sync(FULL, t.i[r - 1], 31);
            if (lane == 0) { pd = wd; pi = wi; }
        }
        const int e = r * 32 + lane;
        if (e == pos) { t.d[r] = cd; t.i[r] = ci; }
        else if (e > pos) { t.d[r] = pd; t.i[r] = pi; }
    }
}

template <int KPL>
__device__ __forceinline__ void topk_kth(const TopK<KPL>& t, int k, float& td, int& ti) {
    const int rk = (k - 1) >> 5, lk = (k - 1) & 31;
    float d = t.d[0];
    int i = t.i[0];
#pragma unroll
    for (int r = 1; r < KPL; r++) if (rk == r) { d = t.d[r]; i = t.i[r]; }
    td = __shfl_sync(FULL, d, lk);
    ti = __shfl_sync(FULL, i, lk);
}

// compare-exchange with the lane `j` away: keep the smaller key of the pair when keep_min, else the larger
// (equal keys only occur between identical padding entries, where either choice is the same value)
__device__ __forceinline__ void key_cmpx(float& d, int& i, int j, bool keep_min) {
    const float od = __shfl_xor_sync(FULL, d, j);
    const int oi = __shfl_xor_sync(FULL, i, j);
    if (key_less(od, oi, d, i) == keep_min) { d = od; i = oi; }
}

// Batch path (k <= 32, list = one key per lane): bitonic-sort the up-to-32 passing candidates, merge with
// the resident list, keep the 32 smallest.  ~200 warp instructions whatever the number of candidates,
// against ~28 per serial insertion: the first cells of a query, where nearly every candidate passes,
// were 80 % of the kernel's instructions (ncu: 2 060 warp instructions per query at k = 8, IPC 3.0).
__device__ __forceinline__ void topk_merge32(TopK<1>& t, bool pass, float cd, int ci, int lane) {
    float d = pass ? cd : PLACEHOLDER_D2;
    int i = pass ? ci : INT_MAX;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1)
            key_cmpx(d, i, j, ((lane & j) == 0) == ((lane & k2) == 0));   // ascending overall (lane & 32 == 0)
    // candidates descending against residents ascending: the elementwise minimum is a bitonic sequence
    // holding the 32 smallest keys of the union
    const float rd = __shfl_sync(FULL, d, 31 - lane);
    const int ri = __shfl_sync(FULL, i, 31 - lane);
    if (key_less(rd, ri, t.d[0], t.i[0])) { t.d[0] = rd; t.i[0] = ri; }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) key_cmpx(t.d[0], t.i[0], j, (lane & j) == 0);
}

// offer one candidate per lane (valid lanes only); warp-uniform control flow
template <int KPL>
__device__ __forceinline__ void topk_offer(TopK<KPL>& t, int k, float& tau_d, int& tau_i, bool valid, float cd,
                                           int ci, int lane) {
    bool pass = valid && cd < PLACEHOLDER_D2 && key_less(cd, ci, tau_d, tau_i);
    unsigned mask = __ballot_sync(FULL, pass);
    if constexpr (KPL == 1) {
        if (__popc(mask) >= 8) {
            topk_merge32(t, pass, cd, ci, lane);
            topk_kth<KPL>(t, k, tau_d, tau_i);
            return;
        }
    }
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float sd = __shfl_sync(FULL, cd, src);
        const int si = __shfl_sync(FULL, ci, src);
        if (key_less(sd, si, tau_d, tau_i)) {  // tau may have tightened since the ballot
            topk_insert<KPL>(t, sd, si, lane);
            topk_kth<KPL>(t, k, tau_d, tau_i);
        }
    }
}

// Squared distance below which no point outside the scanned cube of radius R can lie.
// A point whose cell differs from the query's by more than R along some axis is separated
// from it by more than (R - 2*eta*MAX_DIM) cells along that axis, eta ~ 1.8e-7 being the
// relative error of cell_coord's two roundings plus inv_h (both points go through the same
// monotone function, also when clamped).  0.002 and the 1e-4 shave dominate those terms and
// the <= 3 ulp error of d2_ref itself.
__device__ __forceinline__ float bound2(int R, float h) {
    const float r = ((float)R - 0.002f) * h;
    return r * r * 0.9999f;
}

// Fused "query -> all-gather": when npeers > 0 every result row is stored straight into the result buffers of ALL
// ranks (peer-mapped device memory: NVLink / NVSwitch P2P stores) at row row_base + q, instead of into a local
// shard that an NCCL all-gather would then copy around.  The stores of one warp are a contiguous k * 4-byte row per
// peer; they drain over NVLink while the next queries are being searched (the one place where a
// kernel is directly followed by a collective).
constexpr int KNN_MAX_PEERS = 16;
struct KnnPeers {
    int npeers;
    int64_t row_base;
    int* idx[KNN_MAX_PEERS];
    float* dist[KNN_MAX_PEERS];
};

template <int KPL>
__global__ void __launch_bounds__(256) knn_grid_kernel(int64_t m, int k, int b, const float* __restrict__ xyz,
                                                       const float* __restrict__ new_xyz,
                                                       const int* __restrict__ new_offset,
                                                       const SceneGrid* __restrict__ scenes,
                                                       const int* __restrict__ cell_start,
                                                       const float4* __restrict__ sorted, int* __restrict__ idx_out,
                                                       float* __restrict__ dist_out, float* __restrict__ weight_out,
                                                       int take_sqrt, int force_brute,
                                                       unsigned long long* __restrict__ stats, const KnnPeers peers) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= m) return;
    unsigned evals = 0;   // diagnostics: candidates this query's warp evaluated a distance for
    const int s = segment_of(q, new_offset, b);
    const SceneGrid g = scenes[s];
    const float qx = __ldg(new_xyz + q * 3), qy = __ldg(new_xyz + q * 3 + 1), qz = __ldg(new_xyz + q * 3 + 2);

    TopK<KPL> t;
    float tau_d;
    int tau_i;
    auto reset = [&]() {
#pragma unroll
        for (int r = 0; r < KPL; r++) { t.d[r] = PLACEHOLDER_D2; t.i[r] = INT_MAX; }
        tau_d = PLACEHOLDER_D2; tau_i = INT_MAX;
    };
    reset();

    if (!g.use_grid || force_brute) {
        // direct scan of the scene (tiny scenes, or the exact brute-force entry point)
        for (int base = g.start; base < g.end; base += 32) {
            const int i = base + lane;
            const bool valid = i < g.end;
            float cd = 0.f;
            if (valid) cd = d2_ref(qx, qy, qz, __ldg(xyz + (int64_t)i * 3), __ldg(xyz + (int64_t)i * 3 + 1),
                                   __ldg(xyz + (int64_t)i * 3 + 2));
            topk_offer<KPL>(t, k, tau_d, tau_i, valid, cd, i, lane);
        }
        evals += (unsigned)(g.end - g.start);
    } else {
        const int cx = cell_coord(qx, g.lox, g.inv_h, g.dx);
        const int cy = cell_coord(qy, g.loy, g.inv_h, g.dy);
        const int cz = cell_coord(qz, g.loz, g.inv_h, g.dz);
        const int* cs = cell_start + g.cell_base;
        for (int R = 1;; R <<= 1) {
            const int x0 = max(cx - R, 0), x1 = min(cx + R, g.dx - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dy - 1);
            const int z0 = max(cz - R, 0), z1 = min(cz + R, g.dz - 1);
            const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
            for (int rbase = 0; rbase < nrows; rbase += 32) {
                // lane <-> one x-row of cells: a contiguous range of the sorted array
                const int row = rbase + lane;
                int beg = 0, cnt = 0;
                if (row < nrows) {
                    const int zz = z0 + row / ny, yy = y0 + row % ny;
                    const int c = (zz * g.dy + yy) * g.dx;
                    beg = __ldg(cs + c + x0);
                    cnt = __ldg(cs + c + x1 + 1) - beg;
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                const int shift = beg - (incl - cnt);  // position = shift + ordinal, for ordinals of this row
                evals += (unsigned)total;
                for (int obase = 0; obase < total; obase += 32) {
                    const int o = obase + lane;
                    // smallest j with incl[j] > o
                    int j = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int v = __shfl_sync(FULL, incl, j + step - 1);
                        if (v <= o) j += step;
                    }
                    const int pos = __shfl_sync(FULL, shift, j & 31) + o;
                    const bool valid = o < total;
                    float cd = 0.f;
                    int ci = 0;
                    if (valid) {
                        const float4 p = __ldg(sorted + pos);
                        cd = d2_ref(qx, qy, qz, p.x, p.y, p.z);
                        ci = __float_as_int(p.w);
                    }
                    topk_offer<KPL>(t, k, tau_d, tau_i, valid, cd, ci, lane);
                }
            }
            const bool covered = x0 == 0 && y0 == 0 && z0 == 0 && x1 == g.dx - 1 && y1 == g.dy - 1 && z1 == g.dz - 1;
            if (covered) break;
            if (tau_i != INT_MAX && tau_d < bound2(R, g.h)) break;
            reset();  // rescan the doubled cube from scratch (rare; keeps the list duplicate-free)
        }
    }

    if (stats && lane == 0) atomicAdd(stats, (unsigned long long)evals);
    float recip[KPL], rsum = 0.f;
#pragma unroll
    for (int r = 0; r < KPL; r++) {
        const int e = r * 32 + lane;
        recip[r] = 0.f;
        if (e < k) {
            const bool real = t.i[r] != INT_MAX;
            const float d2 = real ? t.d[r] : PLACEHOLDER_D2;
            if (peers.npeers == 0) {
                idx_out[q * k + e] = real ? t.i[r] : -1;
                if (dist_out) dist_out[q * k + e] = take_sqrt ? __fsqrt_rn(d2) : d2;
            } else {
                const int64_t at = (peers.row_base + q) * k + e;
                const int vi = real ? t.i[r] : -1;
                const float vd = take_sqrt ? __fsqrt_rn(d2) : d2;
                for (int g = 0; g < peers.npeers; g++) {
                    peers.idx[g][at] = vi;
                    if (peers.dist[g]) peers.dist[g][at] = vd;
                }
            }
            // functions/interpolation.py:15: 1 / (sqrt(d2) + 1e-8)
            recip[r] = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d2), 1e-8f));
            rsum += recip[r];
        }
    }
    if (weight_out) {  // fused inverse-distance weights (functions/interpolation.py:16-17)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(FULL, rsum, o);
#pragma unroll
        for (int r = 0; r < KPL; r++) {
            const int e = r * 32 + lane;
            if (e < k) weight_out[q * k + e] = __fdiv_rn(recip[r], rsum);
        }
    }
}

// --------------------------------------------------------------------------------------------
// Radius queries (pointops.ball_query / random_ball_query; off the PTv1 path) on
// the same grid: one warp per query scans the cube of cells that covers the outer radius.

// calls f(valid, point) once per batch of 32 candidates of the cube of radius R cells around (cx, cy, cz)
template <class F>
__device__ __forceinline__ void scan_cube(const SceneGrid& g, const int* __restrict__ cs, const float4* __restrict__ sorted,
                                          int cx, int cy, int cz, int R, int lane, F&& f) {
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.dx - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dy - 1);
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.dz - 1);
    const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
    for (int rbase = 0; rbase < nrows; rbase += 32) {
        const int row = rbase + lane;
        int beg = 0, cnt = 0;
        if (row < nrows) {
            const int zz = z0 + row / ny, yy = y0 + row % ny;
            const int c = (zz * g.dy + yy) * g.dx;
            beg = __ldg(cs + c + x0);
            cnt = __ldg(cs + c + x1 + 1) - beg;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const int shift = beg - (incl - cnt);
        for (int obase = 0; obase < total; obase += 32) {
            const int o = obase + lane;
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(FULL, incl, j + step - 1);
                if (v <= o) j += step;
            }
            const int pos = __shfl_sync(FULL, shift, j & 31) + o;
            const bool valid = o < total;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) p = __ldg(sorted + pos);
            f(valid, p);
        }
    }
}

// cube radius (in cells) outside of which no point can have d2 < r2 (same bound as the kNN stop test)
__device__ __forceinline__ int cover_radius(float r2, float h, int maxdim) {
    int R = 1;
    while (R < maxdim && !(bound2(R, h) > r2)) R++;
    return R;
}

// the reference's acceptance test (ball_query_cuda_kernel.cu:99): the 1e-5 literal is a double
__device__ __forceinline__ bool ball_accepts(float d2, float min_r2, float max_r2) {
    return (double)d2 <= 1e-5 || (d2 >= min_r2 && d2 < max_r2);
}

constexpr int BALL_MAX_CAND = 2048;   // the reference's per-thread candi_dist[2048] / candi_idx[2048]
constexpr int BALL_WARPS = 4;

// ball_query_utils::reheap / heap_sort (ball_query_cuda_kernel.cu:16-42), literally.
__device__ __forceinline__ void ref_heap_sort(float* dist, int* idx, int k) {
    for (int i = k - 1; i > 0; i--) {
        float td = dist[0]; dist[0] = dist[i]; dist[i] = td;
        int ti = idx[0]; idx[0] = idx[i]; idx[i] = ti;
        int root = 0, child = 1;
        while (child < i) {
            if (child + 1 < i && dist[child + 1] > dist[child]) child++;
            if (dist[root] > dist[child]) break;
            td = dist[root]; dist[root] = dist[child]; dist[child] = td;
            ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
            root = child;
            child = 2 * root + 1;
        }
    }
}

// ball_query_cuda_kernel (ball_query_cuda_kernel.cu:58-124).  The reference appends the accepted points in
// scan (= ascending index) order and then runs heap_sort on that list WITHOUT building a heap first, so its
// output is only partially ordered by distance -- but it is a deterministic function of the index-ordered
// list, and a drop-in has to return the same rows.  Hence: gather the candidates of the cube (any order),
// put them in ascending index order (rank by counting), let one lane run the reference's heap_sort on the
// shared-memory copy, then emit: up to nsample -> the list padded with (-1, 1e10); more -> every
// (cnt / nsample)-th entry, with the candidate INDEX written into dist2 exactly as :120 does.
__global__ void __launch_bounds__(BALL_WARPS * 32)
ball_query_kernel(int64_t m, int nsample, float min_radius, float max_radius, int b, const float* __restrict__ xyz,
                  const float* __restrict__ new_xyz, const int* __restrict__ new_offset,
                  const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                  const float4* __restrict__ sorted, int* __restrict__ idx_out, float* __restrict__ dist2_out,
                  int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char ball_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* cd = reinterpret_cast<float*>(ball_smem) + (size_t)warp * 4 * BALL_MAX_CAND;   // per warp: cd | ci | sd | si
    int* ci = reinterpret_cast<int*>(cd + BALL_MAX_CAND);
    float* sd = cd + 2 * BALL_MAX_CAND;
    int* si = reinterpret_cast<int*>(cd + 3 * BALL_MAX_CAND);
    const int64_t q = (int64_t)blockIdx.x * BALL_WARPS + warp;
    if (q >= m) return;
    const int s = segment_of(q, new_offset, b);
    const SceneGrid g = scenes[s];
    const float qx = __ldg(new_xyz + q * 3), qy = __ldg(new_xyz + q * 3 + 1), qz = __ldg(new_xyz + q * 3 + 2);
    const float max_r2 = __fmul_rn(max_radius, max_radius), min_r2 = __fmul_rn(min_radius, min_radius);
    int cnt = 0;
    auto offer = [&](bool valid, float d2, int i) {
        const bool take = valid && ball_accepts(d2, min_r2, max_r2);
        const unsigned mask = __ballot_sync(FULL, take);
        const int at = cnt + __popc(mask & ((1u << lane) - 1u));
        if (take && at < BALL_MAX_CAND) { cd[at] = d2; ci[at] = i; }
        cnt += __popc(mask);
    };
    if (!g.use_grid) {
        for (int base = g.start; base < g.end; base += 32) {
            const int i = base + lane;
            const bool valid = i < g.end;
            float d2 = 0.f;
            if (valid) d2 = d2_ref(qx, qy, qz, __ldg(xyz + (int64_t)i * 3), __ldg(xyz + (int64_t)i * 3 + 1),
                                   __ldg(xyz + (int64_t)i * 3 + 2));
            offer(valid, d2, i);
        }
    } else {
        const int R = cover_radius(fmaxf(max_r2, 1.0001e-5f), g.h, MAX_DIM);
        scan_cube(g, cell_start + g.cell_base, sorted, cell_coord(qx, g.lox, g.inv_h, g.dx),
                  cell_coord(qy, g.loy, g.inv_h, g.dy), cell_coord(qz, g.loz, g.inv_h, g.dz), R, lane,
                  [&](bool valid, const float4& p) { offer(valid, d2_ref(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w)); });
    }
    __syncwarp();
    if (cnt > BALL_MAX_CAND) {   // the reference overruns its stack arrays here; report instead
        if (lane == 0) atomicExch(overflow, 1);
        cnt = BALL_MAX_CAND;
    }
    // ascending index order = the reference's scan order
    for (int e = lane; e < cnt; e += 32) {
        const int i = ci[e];
        int rank = 0;
        for (int f2 = 0; f2 < cnt; f2++) rank += ci[f2] < i ? 1 : 0;
        sd[rank] = cd[e];
        si[rank] = i;
    }
    __syncwarp();
    if (lane == 0) ref_heap_sort(sd, si, cnt);
    __syncwarp();
    int* io = idx_out + q * nsample;
    float* dout = dist2_out + q * nsample;
    if (cnt <= nsample) {
        for (int e = lane; e < nsample; e += 32) {
            io[e] = e < cnt ? si[e] : -1;
            dout[e] = e < cnt ? sd[e] : PLACEHOLDER_D2;
        }
    } else {
        const float sep = (float)cnt / (float)nsample;
        for (int e = lane; e < nsample; e += 32) {
            const int j = (int)(sep * (float)e);
            io[e] = si[j];
            dout[e] = (float)si[j];
        }
    }
}

__global__ void invert_order_kernel(int64_t n, const int* __restrict__ order, int* __restrict__ inv) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = order[p];
        if (i >= 0 && i < n) inv[i] = (int)p;
    }
}

// random_ball_query_cuda_kernel (random_ball_query_cuda_kernel.cu:58-107): the first nsample accepted points
// in the order of the given permutation = the nsample accepted points with the smallest POSITION in
// `order`: a top-k on the key (0, position) over the candidates of the cube.
template <int KPL>
__global__ void __launch_bounds__(256)
random_ball_query_kernel(int64_t m, int nsample, float min_radius, float max_radius, int b, const int* __restrict__ order,
                         const int* __restrict__ inv_order, const float* __restrict__ xyz,
                         const float* __restrict__ new_xyz, const int* __restrict__ new_offset,
                         const SceneGrid* __restrict__ scenes, const int* __restrict__ cell_start,
                         const float4* __restrict__ sorted, int* __restrict__ idx_out, float* __restrict__ dist2_out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= m) return;
    const int s = segment_of(q, new_offset, b);
    const SceneGrid g = scenes[s];
    const float qx = __ldg(new_xyz + q * 3), qy = __ldg(new_xyz + q * 3 + 1), qz = __ldg(new_xyz + q * 3 + 2);
    const float max_r2 = __fmul_rn(max_radius, max_radius), min_r2 = __fmul_rn(min_radius, min_radius);
    TopK<KPL> t;
#pragma unroll
    for (int r = 0; r < KPL; r++) { t.d[r] = PLACEHOLDER_D2; t.i[r] = INT_MAX; }
    float tau_d = PLACEHOLDER_D2;
    int tau_i = INT_MAX;
    auto offer = [&](bool valid, float d2, int i) {
        const bool take = valid && ball_accepts(d2, min_r2, max_r2);
        topk_offer<KPL>(t, nsample, tau_d, tau_i, take, 0.f, take ? __ldg(inv_order + i) : 0, lane);
    };
    if (!g.use_grid) {
        for (int base = g.start; base < g.end; base += 32) {
            const int i = base + lane;
            const bool valid = i < g.end;
            float d2 = 0.f;
            if (valid) d2 = d2_ref(qx, qy, qz, __ldg(xyz + (int64_t)i * 3), __ldg(xyz + (int64_t)i * 3 + 1),
                                   __ldg(xyz + (int64_t)i * 3 + 2));
            offer(valid, d2, i);
        }
    } else {
        const int R = cover_radius(fmaxf(max_r2, 1.0001e-5f), g.h, MAX_DIM);
        scan_cube(g, cell_start + g.cell_base, sorted, cell_coord(qx, g.lox, g.inv_h, g.dx),
                  cell_coord(qy, g.loy, g.inv_h, g.dy), cell_coord(qz, g.loz, g.inv_h, g.dz), R, lane,
                  [&](bool valid, const float4& p) { offer(valid, d2_ref(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w)); });
    }
#pragma unroll
    for (int r = 0; r < KPL; r++) {
        const int e = r * 32 + lane;
        if (e < nsample) {
            const bool real = t.i[r] != INT_MAX;
            int i = -1;
            float d2 = PLACEHOLDER_D2;
            if (real) {
                i = __ldg(order + t.i[r]);
                d2 = d2_ref(qx, qy, qz, __ldg(xyz + (int64_t)i * 3), __ldg(xyz + (int64_t)i * 3 + 1), __ldg(xyz + (int64_t)i * 3 + 2));
            }
            idx_out[q * nsample + e] = i;
            dist2_out[q * nsample + e] = d2;
        }
    }
}

}  // namespace pob

using namespace pob;

// --------------------------------------------------------------------- C ABI --

POB_API size_t pob_knn_grid_workspace_bytes(int64_t n, int b, float cell_pts) {
    if (n < 0 || b < 1) return 0;
    return grid_layout(n, b, cell_pts).total;
}

POB_API int pob_knn_grid_build(int64_t n, int b, const float* xyz, const int* offset, float cell_pts,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n < 0 || b < 1 || !offset || !workspace || (n > 0 && !xyz)) return POB_ERR_BAD_ARG;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    const GridLayout L = grid_layout(n, b, cell_pts);
    if (workspace_bytes < L.total) return POB_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    SceneGrid* scenes = (SceneGrid*)(ws + L.off_scene);
    int* bbox = (int*)(ws + L.off_bbox);
    int* cnt = (int*)(ws + L.off_cnt);
    int* start = (int*)(ws + L.off_start);
    int* tiles = (int*)(ws + L.off_tiles);
    int* pcell = (int*)(ws + L.off_pcell);
    float4* sorted = (float4*)(ws + L.off_sorted);
    const int64_t cap1 = L.cap + 1;
    const int64_t ntiles = ceil_div(cap1, SCAN_TILE);

    grid_init_kernel<<<grid_for(cap1, 256, 8), 256, 0, stream>>>(bbox, b, cnt, cap1, tiles, ntiles + 1);
    if (n > 0) {
        const unsigned bx = (unsigned)min((int64_t)64, max((int64_t)1, ceil_div(ceil_div(n, b), 1024)));
        grid_bbox_kernel<<<dim3(bx, (unsigned)b), 256, 0, stream>>>(xyz, offset, bbox);
    }
    grid_params_kernel<<<1, 32, 0, stream>>>(offset, b, bbox, cell_pts, L.cap, scenes);
    if (n > 0) grid_count_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, b, xyz, offset, scenes, pcell, cnt);
    if (ntiles > 1) {
        scan_tile_sums_kernel<<<(unsigned)ntiles, 512, 0, stream>>>(cnt, cap1, tiles);
        scan_tiles_kernel<<<1, 512, 0, stream>>>(tiles, ntiles);
    }
    scan_apply_kernel<<<(unsigned)ntiles, 512, 0, stream>>>(cnt, cap1, tiles, start);
    if (n > 0) grid_scatter_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, xyz, pcell, cnt, start, sorted);
    pob_count_launches((n > 0 ? 6 : 3) + (ntiles > 1 ? 2 : 0));
    POB_RETURN_LAST_ERROR();
}

static int knn_launch(int64_t m, int k, int b, const float* xyz, const float* new_xyz, const int* new_offset,
                      const SceneGrid* scenes, const int* cell_start, const float4* sorted, int* idx, float* dist,
                      float* weight, int take_sqrt, int force_brute, cudaStream_t stream,
                      unsigned long long* stats = nullptr, const KnnPeers* peers_in = nullptr) {
    KnnPeers peers = {};
    if (peers_in) peers = *peers_in;
    const unsigned blocks = (unsigned)ceil_div(m, 8);
#define POB_KNN_LAUNCH(KPL)                                                                                       \
    knn_grid_kernel<KPL><<<blocks, 256, 0, stream>>>(m, k, b, xyz, new_xyz, new_offset, scenes, cell_start, sorted, \
                                                     idx, dist, weight, take_sqrt, force_brute, stats, peers)
    if (k <= 32) POB_KNN_LAUNCH(1);
    else if (k <= 64) POB_KNN_LAUNCH(2);
    else if (k <= 128) POB_KNN_LAUNCH(4);
    else POB_KNN_LAUNCH(8);
#undef POB_KNN_LAUNCH
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

POB_API int pob_knn_grid_query(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                               const int* new_offset, float cell_pts, const void* workspace, int* idx, float* dist,
                               float* weight, int take_sqrt, void* stats_u64, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || nsample > 256 || b < 1 || !workspace) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!new_xyz || !new_offset || !idx) return POB_ERR_BAD_ARG;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    const GridLayout L = grid_layout(n, b, cell_pts);
    const char* ws = (const char*)workspace;
    return knn_launch(m, nsample, b, xyz, new_xyz, new_offset, (const SceneGrid*)(ws + L.off_scene),
                      (const int*)(ws + L.off_start), (const float4*)(ws + L.off_sorted), idx, dist, weight, take_sqrt,
                      0, stream, (unsigned long long*)stats_u64);
}

// Query a built grid and store every result row into the (n_total, nsample) result buffers of ALL ranks: peer_idx /
// peer_dist are HOST arrays of npeers device pointers (peer-mapped: cudaMalloc + IPC / symmetric memory; the caller's
// own buffer is one of them), row_base is where this rank's first query row goes.  peer_dist may be NULL, or hold
// NULLs.  Synchronisation with the peers (nobody reads before everybody has written) is the caller's.
POB_API int pob_knn_grid_query_scatter(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                                       const int* new_offset, float cell_pts, const void* workspace, int64_t row_base,
                                       int npeers, int* const* peer_idx, float* const* peer_dist, int take_sqrt,
                                       cudaStream_t stream) {
    if (m < 0 || nsample < 1 || nsample > 256 || b < 1 || !workspace || npeers < 1 || npeers > KNN_MAX_PEERS || !peer_idx ||
        row_base < 0)
        return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!new_xyz || !new_offset) return POB_ERR_BAD_ARG;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    const GridLayout L = grid_layout(n, b, cell_pts);
    const char* ws = (const char*)workspace;
    KnnPeers peers = {};
    peers.npeers = npeers;
    peers.row_base = row_base;
    for (int g = 0; g < npeers; g++) {
        if (!peer_idx[g]) return POB_ERR_BAD_ARG;
        peers.idx[g] = peer_idx[g];
        peers.dist[g] = peer_dist ? peer_dist[g] : nullptr;
    }
    return knn_launch(m, nsample, b, xyz, new_xyz, new_offset, (const SceneGrid*)(ws + L.off_scene),
                      (const int*)(ws + L.off_start), (const float4*)(ws + L.off_sorted), nullptr, nullptr, nullptr, take_sqrt,
                      0, stream, nullptr, &peers);
}

// Reference-shaped entry point: knn_query_cuda_launcher (knn_query_cuda_kernel.h:13) plus the
// sizes a grid build needs, a caller-owned workspace and the stream.  dist receives d2 like the
// reference (take_sqrt=0) or sqrt(d2) fused (take_sqrt=1, what functions/query.py:24 returns).
POB_API int pob_knn_query(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                          const int* offset, const int* new_offset, int* idx, float* dist, int take_sqrt,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const float cell_pts = 2.0f;
    int rc = pob_knn_grid_build(n, b, xyz, offset, cell_pts, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return pob_knn_grid_query(m, nsample, n, b, xyz, new_xyz, new_offset, cell_pts, workspace, idx, dist, nullptr,
                              take_sqrt, nullptr, stream);
}

// Exhaustive variant (same key, same d2): every query scans its whole scene.  Needs only the
// SceneGrid table, which it fills itself in the first bytes of the workspace (64*b bytes).
namespace pob {
__global__ void scene_table_kernel(const int* __restrict__ offset, int b, SceneGrid* __restrict__ scenes) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b) return;
    SceneGrid g = {};
    g.start = s == 0 ? 0 : offset[s - 1];
    g.end = offset[s];
    g.use_grid = 0;
    g.dx = g.dy = g.dz = 1;
    scenes[s] = g;
}
}  // namespace pob

POB_API int pob_knn_query_bruteforce(int64_t m, int nsample, int b, const float* xyz, const float* new_xyz,
                                     const int* offset, const int* new_offset, int* idx, float* dist, int take_sqrt,
                                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || nsample > 256 || b < 1 || !workspace) return POB_ERR_BAD_ARG;
    if (workspace_bytes < sizeof(SceneGrid) * (size_t)b) return POB_ERR_WORKSPACE;
    if (m == 0) return 0;
    SceneGrid* scenes = (SceneGrid*)workspace;
    scene_table_kernel<<<(unsigned)ceil_div(b, 128), 128, 0, stream>>>(offset, b, scenes);
    pob_count_launches(1);
    return knn_launch(m, nsample, b, xyz, new_xyz, new_offset, scenes, nullptr, nullptr, idx, dist, nullptr, take_sqrt,
                      1, stream);
}

// ball_query_cuda_launcher(m, nsample, min_radius, max_radius, xyz, new_xyz, offset, new_offset, idx, dist2)
// (src/ball_query/ball_query_cuda_kernel.h) on a grid workspace built by pob_knn_grid_build for (xyz, offset).
// dist2 receives what the reference writes (d2, or the candidate index where it subsamples, see the kernel).
// *overflow_flag (device int, caller-zeroed, may be NULL) is set when a query had more than 2048 candidates,
// where the reference overruns its stack arrays; such rows use the first 2048 found.
POB_API int pob_ball_query(int64_t m, int nsample, float min_radius, float max_radius, int64_t n, int b, const float* xyz,
                           const float* new_xyz, const int* new_offset, float cell_pts, const void* workspace, int* idx,
                           float* dist2, int* overflow_flag, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || b < 1 || !workspace || !(min_radius < max_radius)) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!xyz || !new_xyz || !new_offset || !idx || !dist2 || !overflow_flag) return POB_ERR_BAD_ARG;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    const GridLayout L = grid_layout(n, b, cell_pts);
    const char* ws = (const char*)workspace;
    const size_t smem = 4 * sizeof(float) * BALL_WARPS * BALL_MAX_CAND;   // candidates + index-ordered copy, 32 KB per warp
    POB_CHECK(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_query_kernel<<<(unsigned)ceil_div(m, BALL_WARPS), BALL_WARPS * 32, smem, stream>>>(
        m, nsample, min_radius, max_radius, b, xyz, new_xyz, new_offset, (const SceneGrid*)(ws + L.off_scene),
        (const int*)(ws + L.off_start), (const float4*)(ws + L.off_sorted), idx, dist2, overflow_flag);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// random_ball_query_cuda_launcher(m, nsample, min_radius, max_radius, order, xyz, new_xyz, offset, new_offset,
// idx, dist2) (src/random_ball_query/random_ball_query_cuda_kernel.h).  order (n): a permutation of each
// scene's rows (global indices, scene-major); inv_scratch (n ints, caller-owned) receives its inverse.
POB_API int pob_random_ball_query(int64_t m, int nsample, float min_radius, float max_radius, int64_t n, int b,
                                  const int* order, const float* xyz, const float* new_xyz, const int* new_offset,
                                  float cell_pts, const void* workspace, int* inv_scratch, int* idx, float* dist2,
                                  cudaStream_t stream) {
    if (m < 0 || nsample < 1 || nsample > 256 || b < 1 || !workspace || !(min_radius < max_radius)) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!order || !xyz || !new_xyz || !new_offset || !inv_scratch || !idx || !dist2) return POB_ERR_BAD_ARG;
    if (!(cell_pts >= 0.25f)) cell_pts = 0.25f;
    const GridLayout L = grid_layout(n, b, cell_pts);
    const char* ws = (const char*)workspace;
    invert_order_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, order, inv_scratch);
    const unsigned blocks = (unsigned)ceil_div(m, 8);
#define POB_RBQ(KPL)                                                                                                  \
    random_ball_query_kernel<KPL><<<blocks, 256, 0, stream>>>(m, nsample, min_radius, max_radius, b, order, inv_scratch, \
        xyz, new_xyz, new_offset, (const SceneGrid*)(ws + L.off_scene), (const int*)(ws + L.off_start),                \
        (const float4*)(ws + L.off_sorted), idx, dist2)
    if (nsample <= 32) POB_RBQ(1);
    else if (nsample <= 64) POB_RBQ(2);
    else if (nsample <= 128) POB_RBQ(4);
    else POB_RBQ(8);
#undef POB_RBQ
    pob_count_launches(2);
    POB_RETURN_LAST_ERROR();
}
