// ptlayer.cu -- the inference form of PointTransformerLayer (vector attention) as ONE kernel, plus
// the two small epilogue kernels the eval plan of the PTv1 mirror needs (sm_100a).
//
// Reference caller: pointcept/models/point_transformer/point_transformer_seg.py:48-81.  Between
// the q/k/v linears and the layer output the reference runs, per block, two kNN-group gathers,
// linear_p (Linear(3,3) -> BN -> ReLU -> Linear(3,C)), the relation r = k_j - q_i + p_r, linear_w
// (BN -> ReLU -> Linear(C,C/8) -> BN -> ReLU -> Linear(C/8,C/8)), a softmax over the neighbours and
// the einsum aggregation: ~16 eager kernels and ~10 (n, ns, C) temporaries.  In eval mode every
// BatchNorm is a per-channel affine map, so the whole chain is a function of one point's
// neighbourhood: one warp per point, neighbours processed one after the other,
//   lane <-> channels (c = lane + 32 j for C <= 64; c = 4 lane + i + 128 j above, 128-bit loads),
//   the C -> C/8 projection as per-lane partial sums + a shuffle reduce-scatter,
//   the C/8 -> C/8 projection from a per-warp shared-memory scratch,
//   softmax over neighbours ONLINE (running max / sum per lane-owned weight channel), so k and v
//   rows are gathered once and nothing of size (n, ns, *) is ever written.
// Traffic: q, out rows streamed once; k, v rows gathered ns times (L2-resident tables).
// The dense q/k/v linears stay on cuBLAS (north_star); this is the part that was eager glue.
#include "common.cuh"
#include <cstdlib>

namespace pob {

constexpr int PTL_THREADS = 256;
constexpr int PTL_WARPS = PTL_THREADS / 32;

// packed parameter block (floats): p1[16] | chan[8][C] | w1[WC][C] | b1[WC] | w2t[WC][WC] | b2[WC]
//   p1   = A (3x3 row-major, Linear(3,3) with its BatchNorm folded in), c (3), 4 pad
//   chan = wx, wy, wz, bp (Linear(3,C) columns + bias), aw, bw (first BN of linear_w as affine),
//          oa, ob (affine applied to the layer output before ReLU when out_affine != 0: bn2)
//   w1   = Linear(C, WC) weight with the second BN folded in, b1 its bias
//   w2t  = TRANSPOSE of the Linear(WC, WC) weight (w2t[o][o'] = W[o'][o]), b2 its bias
__host__ __device__ inline int64_t ptl_param_floats(int c, int wc) {
    return 16 + 8 * (int64_t)c + (int64_t)wc * c + wc + (int64_t)wc * wc + wc;
}

template <int VEC> struct Vf;
template <> struct Vf<1> {
    float v[1];
    static __device__ __forceinline__ Vf ld(const float* p) { Vf r; r.v[0] = __ldg(p); return r; }
    static __device__ __forceinline__ Vf lds(const float* p) { Vf r; r.v[0] = *p; return r; }
    __device__ __forceinline__ void st(float* p) const { *p = v[0]; }
};
template <> struct Vf<4> {
    float v[4];
    static __device__ __forceinline__ Vf ld(const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        Vf r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
    }
    static __device__ __forceinline__ Vf lds(const float* p) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        Vf r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
    }
    __device__ __forceinline__ void st(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// Sum W per-lane values over the 32 lanes of a warp, leaving each lane with max(W/32, 1) of the
// totals: halving exchanges while more than one value is left (lane bit `OFF` picks the half it
// keeps), plain butterfly adds after that.  Returns through `base` the index of a[0]'s total.
template <int CNT, int OFF>
struct ReduceScatter {
    template <int W>
    static __device__ __forceinline__ void run(float (&a)[W], int lane, int& base) {
        if constexpr (CNT > 1) {
            constexpr int HALF = CNT / 2;
            const bool up = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < HALF; i++) {
                const float send = up ? a[i] : a[i + HALF];
                const float keep = up ? a[i + HALF] : a[i];
                a[i] = keep + __shfl_xor_sync(FULL, send, OFF);
            }
            base += up ? HALF : 0;
            if constexpr (OFF > 1) ReduceScatter<HALF, OFF / 2>::run(a, lane, base);
        } else {
            a[0] += __shfl_xor_sync(FULL, a[0], OFF);
            if constexpr (OFF > 1) ReduceScatter<1, OFF / 2>::run(a, lane, base);
        }
    }
};

// NSW warps share one point: warp `sub` of the group takes neighbours sub, sub + NSW, ... with its own
// online softmax, and the NSW partial states (max, denominator, weighted sum) are merged through
// shared memory.  The deep stages of PTv1 have few points (1 250 x 256 channels, 312 x 512): with one
// warp per point the 16 neighbours are a serial chain on a handful of warps per SM (~79 / 110 us per
// call, ncu launch list r01b); spreading them over the CTA turns that latency into parallelism.
// NSW = 1 is the plain one-warp-per-point form (no merge, no block barrier in the loop).
template <int VEC, int R, int NS, int WC, int NSW>
__global__ void __launch_bounds__((NSW > PTL_WARPS ? NSW : PTL_WARPS) * 32)
pt_layer_fwd_kernel(int64_t n, const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                    const float* __restrict__ v, int64_t ldv, const float* __restrict__ xyz,
                    const int* __restrict__ idx, const float* __restrict__ params, int out_affine,
                    float* __restrict__ out, int64_t ldo) {
    constexpr int C = 32 * VEC * R;
    constexpr int NV = WC >= 32 ? WC / 32 : 1;
    constexpr int JS = 32 * VEC;  // channel stride between a lane's j-slices
    constexpr int WARPS = NSW > PTL_WARPS ? NSW : PTL_WARPS;
    constexpr int THREADS = WARPS * 32;
    constexpr int PPC = WARPS / NSW;  // points per CTA iteration
    static_assert(NS % NSW == 0 && WARPS % NSW == 0, "neighbours split evenly over the warps of a point");
    extern __shared__ __align__(16) float smem[];
    float* w1s = smem;              // [WC][C]
    float* b1s = w1s + WC * C;      // [WC]
    float* w2ts = b1s + WC;         // [WC][WC]
    float* b2s = w2ts + WC * WC;    // [WC]
    float* us = b2s + WC;           // [WARPS][WC]
    float* mxs = us + WARPS * WC;   // [WARPS][WC]  per-warp running max      (NSW > 1 only)
    float* dens = mxs + WARPS * WC; // [WARPS][WC]  per-warp denominator, rescaled to the common max
    float* part = dens + WARPS * WC;// [WARPS][C]   per-warp weighted sum, rescaled
    {
        const float4* src = reinterpret_cast<const float4*>(params + 16 + 8 * C);
        float4* dst = reinterpret_cast<float4*>(smem);
        constexpr int N4 = (WC * C + WC + WC * WC + WC) / 4;
        for (int i = threadIdx.x; i < N4; i += THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / NSW, sub = warp % NSW;  // point slot within the CTA, share of its neighbours
    const int c0 = VEC == 1 ? lane : 4 * lane;  // channel (j = 0, i = 0); channel(j, i) = c0 + i + JS * j
    const int o0 = c0 % WC;                     // the VEC weight channels this lane's channels use
    float* uw = us + warp * WC;
    const float* chan = params + 16;
    float A[9], cb[3];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = __ldg(params + i);
#pragma unroll
    for (int i = 0; i < 3; i++) cb[i] = __ldg(params + 9 + i);

    for (int64_t pbase = (int64_t)blockIdx.x * PPC; pbase < n; pbase += (int64_t)gridDim.x * PPC) {
        const int64_t p = pbase + grp;
        const bool active = p < n;   // the tail CTA may hold empty point slots; they still meet the barriers
        // lane s < NS: neighbour s -- its row, and the hidden layer of linear_p on its relative position
        int js = -1;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        if (active && lane < NS) {
            js = __ldg(idx + p * NS + lane);
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if (js >= 0) {   // placeholder neighbours group to a zero row (functions/grouping.py:41-57)
                rx = __ldg(xyz + (int64_t)js * 3) - __ldg(xyz + p * 3);
                ry = __ldg(xyz + (int64_t)js * 3 + 1) - __ldg(xyz + p * 3 + 1);
                rz = __ldg(xyz + (int64_t)js * 3 + 2) - __ldg(xyz + p * 3 + 2);
            }
            h0 = fmaxf(fmaf(A[0], rx, fmaf(A[1], ry, fmaf(A[2], rz, cb[0]))), 0.f);
            h1 = fmaxf(fmaf(A[3], rx, fmaf(A[4], ry, fmaf(A[5], rz, cb[1]))), 0.f);
            h2 = fmaxf(fmaf(A[6], rx, fmaf(A[7], ry, fmaf(A[8], rz, cb[2]))), 0.f);
        }
        Vf<VEC> qv[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (active) qv[j] = Vf<VEC>::ld(q + p * ldq + c0 + JS * j);
            else {
#pragma unroll
                for (int i = 0; i < VEC; i++) qv[j].v[i] = 0.f;
            }
        }
        float mx[VEC], den[VEC], o_acc[R][VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) { mx[i] = -INFINITY; den[i] = 0.f; }
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int i = 0; i < VEC; i++) o_acc[j][i] = 0.f;

#pragma unroll 1
        for (int s = sub; s < (active ? NS : 0); s += NSW) {
            const int jn = __shfl_sync(FULL, js, s);
            const float g0 = __shfl_sync(FULL, h0, s), g1 = __shfl_sync(FULL, h1, s), g2 = __shfl_sync(FULL, h2, s);
            float val[R][VEC];   // v_j + p_r, what the attention weights multiply
            float acc[WC];
#pragma unroll
            for (int o = 0; o < WC; o++) acc[o] = 0.f;
#pragma unroll
            for (int j = 0; j < R; j++) {
                const int c = c0 + JS * j;
                Vf<VEC> kk, vv;
                if (jn >= 0) {
                    kk = Vf<VEC>::ld(k + (int64_t)jn * ldk + c);
                    vv = Vf<VEC>::ld(v + (int64_t)jn * ldv + c);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; i++) { kk.v[i] = 0.f; vv.v[i] = 0.f; }
                }
                const Vf<VEC> wx = Vf<VEC>::ld(chan + c), wy = Vf<VEC>::ld(chan + C + c), wz = Vf<VEC>::ld(chan + 2 * C + c),
                              bp = Vf<VEC>::ld(chan + 3 * C + c), aw = Vf<VEC>::ld(chan + 4 * C + c),
                              bw = Vf<VEC>::ld(chan + 5 * C + c);
                float t[VEC];
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const float pr = fmaf(wx.v[i], g0, fmaf(wy.v[i], g1, fmaf(wz.v[i], g2, bp.v[i])));
                    const float r = (kk.v[i] - qv[j].v[i]) + pr;
                    t[i] = fmaxf(fmaf(aw.v[i], r, bw.v[i]), 0.f);
                    val[j][i] = vv.v[i] + pr;
                }
#pragma unroll
                for (int o = 0; o < WC; o++) {
                    const Vf<VEC> w = Vf<VEC>::lds(w1s + o * C + c);
#pragma unroll
                    for (int i = 0; i < VEC; i++) acc[o] = fmaf(t[i], w.v[i], acc[o]);
                }
            }
            int base = 0;
            ReduceScatter<WC, 16>::run(acc, lane, base);
#pragma unroll
            for (int i = 0; i < NV; i++) uw[base + i] = fmaxf(acc[i] + b1s[base + i], 0.f);
            __syncwarp();
            float lg[VEC];
            {
                const Vf<VEC> b2 = Vf<VEC>::lds(b2s + o0);
#pragma unroll
                for (int i = 0; i < VEC; i++) lg[i] = b2.v[i];
            }
#pragma unroll
            for (int o = 0; o < WC; o++) {
                const float uo = uw[o];
                const Vf<VEC> w2 = Vf<VEC>::lds(w2ts + o * WC + o0);
#pragma unroll
                for (int i = 0; i < VEC; i++) lg[i] = fmaf(uo, w2.v[i], lg[i]);
            }
            __syncwarp();
            // online softmax over the neighbours, per weight channel owned by this lane
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const float mn = fmaxf(mx[i], lg[i]);
                const float sc = expf(mx[i] - mn), e = expf(lg[i] - mn);
                den[i] = fmaf(den[i], sc, e);
#pragma unroll
                for (int j = 0; j < R; j++) o_acc[j][i] = fmaf(o_acc[j][i], sc, val[j][i] * e);
                mx[i] = mn;
            }
        }
        if constexpr (NSW == 1) {
#pragma unroll
            for (int j = 0; j < R; j++) {
                const int c = c0 + JS * j;
                Vf<VEC> o;
#pragma unroll
                for (int i = 0; i < VEC; i++) o.v[i] = o_acc[j][i] / den[i];
                if (out_affine) {
                    const Vf<VEC> oa = Vf<VEC>::ld(chan + 6 * C + c), ob = Vf<VEC>::ld(chan + 7 * C + c);
#pragma unroll
                    for (int i = 0; i < VEC; i++) o.v[i] = fmaxf(fmaf(oa.v[i], o.v[i], ob.v[i]), 0.f);
                }
                if (active) o.st(out + p * ldo + c);
            }
        } else {
            // ---- merge the NSW partial softmax states of the point ----
            const bool owner = c0 < WC;  // lanes whose first-slice channels ARE the weight channels o0 .. o0 + VEC - 1
            if (active && owner) {
#pragma unroll
                for (int i = 0; i < VEC; i++) mxs[warp * WC + o0 + i] = mx[i];
            }
            __syncthreads();
            if (active) {
                float f[VEC];
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    float m = mx[i];
#pragma unroll
                    for (int w = 0; w < NSW; w++) m = fmaxf(m, mxs[(grp * NSW + w) * WC + o0 + i]);
                    f[i] = expf(mx[i] - m);
                }
                if (owner) {
#pragma unroll
                    for (int i = 0; i < VEC; i++) dens[warp * WC + o0 + i] = den[i] * f[i];
                }
#pragma unroll
                for (int j = 0; j < R; j++) {
                    Vf<VEC> o;
#pragma unroll
                    for (int i = 0; i < VEC; i++) o.v[i] = o_acc[j][i] * f[i];
                    o.st(part + warp * C + c0 + JS * j);
                }
            }
            __syncthreads();
            if (active) {
                // the NSW * 32 threads of the point sum the partial rows, 4 channels at a time
                for (int item = sub * 32 + lane; item < C / 4; item += NSW * 32) {
                    const int c = item * 4, o = c % WC;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
#pragma unroll
                    for (int w = 0; w < NSW; w++) {
                        const float4 t = *reinterpret_cast<const float4*>(part + (grp * NSW + w) * C + c);
                        const float4 e = *reinterpret_cast<const float4*>(dens + (grp * NSW + w) * WC + o);
                        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                        d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
                    }
                    a.x /= d.x; a.y /= d.y; a.z /= d.z; a.w /= d.w;
                    if (out_affine) {
                        const float4 oa = __ldg(reinterpret_cast<const float4*>(chan + 6 * C + c));
                        const float4 ob = __ldg(reinterpret_cast<const float4*>(chan + 7 * C + c));
                        a.x = fmaxf(fmaf(oa.x, a.x, ob.x), 0.f); a.y = fmaxf(fmaf(oa.y, a.y, ob.y), 0.f);
                        a.z = fmaxf(fmaf(oa.z, a.z, ob.z), 0.f); a.w = fmaxf(fmaf(oa.w, a.w, ob.w), 0.f);
                    }
                    *reinterpret_cast<float4*>(out + p * ldo + c) = a;
                }
            }
        }
    }
}

template <int VEC, int R, int NS, int WC, int NSW>
static int launch_pt_layer(int64_t n, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                           const float* xyz, const int* idx, const float* params, int out_affine, float* out, int64_t ldo,
                           cudaStream_t stream) {
    constexpr int C = 32 * VEC * R;
    constexpr int WARPS = NSW > PTL_WARPS ? NSW : PTL_WARPS;
    constexpr int PPC = WARPS / NSW;
    constexpr size_t smem = sizeof(float) * (WC * C + WC + WC * WC + WC + WARPS * WC +
                                             (NSW > 1 ? 2 * WARPS * WC + WARPS * C : 0));
    auto kern = pt_layer_fwd_kernel<VEC, R, NS, WC, NSW>;
    // the opt-in is per device, so it is (cheaply) repeated on every launch; occupancy is a property
    // of the instantiation on sm_100a and is looked up once
    if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int occ = 0;
        POB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    int64_t grid = ceil_div(n, PPC);
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// pt_layer_tile_kernel: the same layer as a CTA-tiled computation.  The warp-per-point kernel above
// pays a shuffle reduce-scatter, a redundant C/8 x C/8 projection and two exp per lane for EVERY
// (point, neighbour) pair, on one dependent instruction stream per point; ncu (profiles/r01b) shows it
// latency-bound at every stage (75-110 us per call whether n is 80 000 or 312).  Here a CTA takes a
// tile of P points = ROWS = P * NS (point, neighbour) rows and runs the layer as five block phases:
//   0  indices, relative positions and the hidden layer of linear_p for the ROWS rows        -> smem
//   1  gather k, build t = relu(aw * (k_j - q_i + p_r) + bw) for a 128-channel chunk         -> smem
//   2  register-tiled FP32 GEMM (ROWS x CK) . (CK x C/8), 2 rows x 4 outputs per thread,
//      accumulated over the chunks; u = relu(. + b1)                                          -> smem
//   3  l = W2 u + b2 (each thread: one output column, 8 rows), softmax over the NS rows of a point
//   4  out[i, c] = sum_s (v_j + p_r)[c] * w[s, c % (C/8)], v rows gathered 128 bits per lane, bn2 + ReLU
// so k and v rows are gathered once each, every weight is read from shared memory once per tile, and
// the per-pair work is the C * C/8 FMAs of the projection plus ~20 instructions.
// FP32 on CUDA cores throughout (north_star: f32 means f32; the projection is 0.16 GFLOP per call).

template <int C, int NS>
struct PtTile {
    static constexpr int WC = C / 8;
    // 2048 projection outputs per 256 threads (2 rows x 4 outputs each); the two deepest stages have so few
    // points (1 250, 312) that half-size CTAs are used to get more independent tiles per SM
    static constexpr int PTT_THREADS = C >= 256 ? 128 : 256;
    static constexpr int ROWS = C <= 64 ? 256 : (C == 128 ? 128 : (C == 256 ? 32 : 16));
    static constexpr int P = ROWS / NS;                 // points per tile
    static constexpr int CK = C <= 64 ? 32 : (C == 128 ? 64 : 128);   // channels per chunk (sizes the t tile: ~35 KB)
    static constexpr int NCH = C / CK, CK4 = CK / 4, C4 = C / 4;
    static constexpr int TS = CK + 4;                   // row stride of the t / W1 chunks (floats): 16-byte rows, 4-bank skew
    static constexpr int OGS = WC / 4;                  // output groups (4 outputs per thread)
    static constexpr int RGS = PTT_THREADS / OGS;       // row groups
    static constexpr int RT = ROWS / RGS;               // rows per thread in the GEMM
    static constexpr int RSTEP = PTT_THREADS / WC;      // phase 3: row stride between a thread's items
    static constexpr int IPT = ROWS / RSTEP;            // phase 3: rows per thread
    static constexpr int ITEMS = ROWS * CK4 / PTT_THREADS;   // phase 1: 128-bit items per thread per chunk
    static constexpr int BATCH = ITEMS < 8 ? ITEMS : 8;      // gathers in flight per thread
    static_assert(RT >= 1 && RGS * RT == ROWS && IPT >= 1 && RSTEP * IPT == ROWS, "tile shape");
    static_assert(ITEMS >= 1 && ITEMS % BATCH == 0 && PTT_THREADS % CK4 == 0, "phase 1 shape");
    static_assert(2 * ROWS * WC <= ROWS * TS, "u and the logits alias the t tile");
    static constexpr int F_IDX = 0;                               // int[ROWS]
    static constexpr int F_H = F_IDX + ROWS;                      // float4[ROWS]
    static constexpr int F_Q = F_H + 4 * ROWS;                    // [P][C]      q rows of the tile's points
    static constexpr int F_T = F_Q + P * C;                       // [ROWS][TS]  t chunk; after the GEMM: u, logits
    static constexpr int F_U = F_T;                               // [ROWS][WC]
    static constexpr int F_L = F_T + ROWS * WC;                   // [ROWS][WC]  logits, then softmax weights
    static constexpr int F_W1 = F_T + ROWS * TS;                  // [WC][TS]
    static constexpr int F_W2 = F_W1 + WC * TS;                   // [WC][WC] (transposed: [o][o'])
    static constexpr int F_B = F_W2 + WC * WC;                    // b1[WC], b2[WC]
    static constexpr int FLOATS = F_B + 2 * WC;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

template <int C, int NS>
__global__ void __launch_bounds__((PtTile<C, NS>::PTT_THREADS), 3)
pt_layer_tile_kernel(int64_t n, const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                     const float* __restrict__ v, int64_t ldv, const float* __restrict__ xyz,
                     const int* __restrict__ idx, const float* __restrict__ params, int out_affine,
                     float* __restrict__ out, int64_t ldo) {
    using T = PtTile<C, NS>;
    constexpr int PTT_THREADS = T::PTT_THREADS;
    constexpr int WC = T::WC, ROWS = T::ROWS, P = T::P, CK = T::CK, CK4 = T::CK4, C4 = T::C4, TS = T::TS;
    constexpr int OGS = T::OGS, RGS = T::RGS, RT = T::RT, RSTEP = T::RSTEP, IPT = T::IPT;
    extern __shared__ __align__(16) float smem[];
    int* sIdx = reinterpret_cast<int*>(smem + T::F_IDX);
    float* sH = smem + T::F_H;
    float* sQ = smem + T::F_Q;
    float* sT = smem + T::F_T;
    float* sW1 = smem + T::F_W1;
    float* sU = smem + T::F_U;
    float* sL = smem + T::F_L;
    float* sW2 = smem + T::F_W2;
    float* sB = smem + T::F_B;
    const int tid = threadIdx.x;
    const float* chan = params + 16;
    const float* W1 = params + 16 + 8 * C;
    const int64_t pbase = (int64_t)blockIdx.x * P;

    // ---- phase 0: rows, and the small weights ----
    {
        float A[9], cb[3];
#pragma unroll
        for (int i = 0; i < 9; i++) A[i] = __ldg(params + i);
#pragma unroll
        for (int i = 0; i < 3; i++) cb[i] = __ldg(params + 9 + i);
        for (int row = tid; row < ROWS; row += PTT_THREADS) {
            const int64_t p = pbase + row / NS;
            int j = -1;
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if (p < n) {
                j = __ldg(idx + p * NS + row % NS);
                if (j >= 0) {   // placeholder neighbours group to a zero row (functions/grouping.py:41-57)
                    rx = __ldg(xyz + (int64_t)j * 3) - __ldg(xyz + p * 3);
                    ry = __ldg(xyz + (int64_t)j * 3 + 1) - __ldg(xyz + p * 3 + 1);
                    rz = __ldg(xyz + (int64_t)j * 3 + 2) - __ldg(xyz + p * 3 + 2);
                }
            }
            float4 h;
            h.x = fmaxf(fmaf(A[0], rx, fmaf(A[1], ry, fmaf(A[2], rz, cb[0]))), 0.f);
            h.y = fmaxf(fmaf(A[3], rx, fmaf(A[4], ry, fmaf(A[5], rz, cb[1]))), 0.f);
            h.z = fmaxf(fmaf(A[6], rx, fmaf(A[7], ry, fmaf(A[8], rz, cb[2]))), 0.f);
            h.w = 0.f;
            sIdx[row] = j;
            *reinterpret_cast<float4*>(sH + 4 * row) = h;
        }
        for (int it = tid; it < P * C4; it += PTT_THREADS) {
            const int64_t p = pbase + it / C4;
            *reinterpret_cast<float4*>(sQ + 4 * it) = p < n ? ld4(q + p * ldq + 4 * (it % C4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* tail = W1 + WC * C;   // b1[WC] | W2t[WC][WC] | b2[WC]
        for (int i = tid; i < WC; i += PTT_THREADS) { sB[i] = __ldg(tail + i); sB[WC + i] = __ldg(tail + WC + WC * WC + i); }
        for (int i = tid; i < WC * WC / 4; i += PTT_THREADS) *reinterpret_cast<float4*>(sW2 + 4 * i) = ld4(tail + WC + 4 * i);
    }
    __syncthreads();

    // ---- phases 1 + 2, one 128-channel chunk at a time ----
    const int og = tid % OGS, rg = tid / OGS;   // GEMM: outputs og + i * OGS, rows rg + i * RGS
    float acc[RT][4];
#pragma unroll
    for (int i = 0; i < RT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

#pragma unroll 1
    for (int ch = 0; ch < T::NCH; ch++) {
        {
            const int c4 = tid % CK4;                 // constant per thread: PTT_THREADS % CK4 == 0
            const int c = ch * CK + 4 * c4;
            const float4 wx = ld4(chan + c), wy = ld4(chan + C + c), wz = ld4(chan + 2 * C + c), bp = ld4(chan + 3 * C + c),
                         aw = ld4(chan + 4 * C + c), bw = ld4(chan + 5 * C + c);
#pragma unroll 1
            for (int i0 = 0; i0 < T::ITEMS; i0 += T::BATCH) {
                float4 kk[T::BATCH];
#pragma unroll
                for (int b = 0; b < T::BATCH; b++) {   // all gathers of the batch in flight before any is used
                    const int row = (tid + (i0 + b) * PTT_THREADS) / CK4;
                    const int j = sIdx[row];
                    kk[b] = j >= 0 ? ld4(k + (int64_t)j * ldk + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int b = 0; b < T::BATCH; b++) {
                    const int row = (tid + (i0 + b) * PTT_THREADS) / CK4;
                    const float4 h = lds4(sH + 4 * row);
                    const float4 qq = lds4(sQ + (row / NS) * C + c);
                    float4 t;
                    t.x = fmaxf(fmaf(aw.x, (kk[b].x - qq.x) + fmaf(wx.x, h.x, fmaf(wy.x, h.y, fmaf(wz.x, h.z, bp.x))), bw.x), 0.f);
                    t.y = fmaxf(fmaf(aw.y, (kk[b].y - qq.y) + fmaf(wx.y, h.x, fmaf(wy.y, h.y, fmaf(wz.y, h.z, bp.y))), bw.y), 0.f);
                    t.z = fmaxf(fmaf(aw.z, (kk[b].z - qq.z) + fmaf(wx.z, h.x, fmaf(wy.z, h.y, fmaf(wz.z, h.z, bp.z))), bw.z), 0.f);
                    t.w = fmaxf(fmaf(aw.w, (kk[b].w - qq.w) + fmaf(wx.w, h.x, fmaf(wy.w, h.y, fmaf(wz.w, h.z, bp.w))), bw.w), 0.f);
                    *reinterpret_cast<float4*>(sT + row * TS + 4 * c4) = t;
                }
            }
            for (int it = tid; it < WC * CK4; it += PTT_THREADS) {
                const int o = it / CK4, cc = it % CK4;
                *reinterpret_cast<float4*>(sW1 + o * TS + 4 * cc) = ld4(W1 + o * C + ch * CK + 4 * cc);
            }
        }
        __syncthreads();
        {
            const float* tb = sT + rg * TS;
            const float* wb = sW1 + og * TS;
#pragma unroll 4
            for (int kk = 0; kk < CK4; kk++) {
                float4 t[RT], w[4];
#pragma unroll
                for (int i = 0; i < RT; i++) t[i] = lds4(tb + i * RGS * TS + 4 * kk);
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = lds4(wb + j * OGS * TS + 4 * kk);
#pragma unroll
                for (int i = 0; i < RT; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = dot4(t[i], w[j], acc[i][j]);
            }
        }
        __syncthreads();   // sT / sW1 are rewritten by the next chunk
    }
#pragma unroll
    for (int i = 0; i < RT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            sU[(rg + i * RGS) * WC + og + j * OGS] = fmaxf(acc[i][j] + sB[og + j * OGS], 0.f);
    __syncthreads();

    // ---- phase 3a: l = W2 u + b2; a thread owns output column o' for IPT rows ----
    {
        const int oc = tid % WC, r0 = tid / WC;
        float lg[IPT];
        const float b2 = sB[WC + oc];
#pragma unroll
        for (int i = 0; i < IPT; i++) lg[i] = b2;
#pragma unroll 2
        for (int o = 0; o < WC; o += 4) {
            const float4 w = make_float4(sW2[o * WC + oc], sW2[(o + 1) * WC + oc], sW2[(o + 2) * WC + oc], sW2[(o + 3) * WC + oc]);
#pragma unroll
            for (int i = 0; i < IPT; i++) lg[i] = dot4(lds4(sU + (r0 + i * RSTEP) * WC + o), w, lg[i]);
        }
#pragma unroll
        for (int i = 0; i < IPT; i++) sL[(r0 + i * RSTEP) * WC + oc] = lg[i];
    }
    __syncthreads();
    // ---- phase 3b: softmax over the NS neighbours of a point, per weight channel ----
    for (int it = tid; it < P * WC; it += PTT_THREADS) {
        const int pl = it / WC, oc = it % WC;
        float* col = sL + pl * NS * WC + oc;
        float e[NS];
        float m = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < NS; s2++) { e[s2] = col[s2 * WC]; m = fmaxf(m, e[s2]); }
        float sum = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < NS; s2++) { e[s2] = expf(e[s2] - m); sum += e[s2]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int s2 = 0; s2 < NS; s2++) col[s2 * WC] = e[s2] * inv;
    }
    __syncthreads();

    // ---- phase 4: aggregation; a thread owns 4 channels of one point ----
    {
        for (int it = tid; it < P * C4; it += PTT_THREADS) {
            const int pl = it / C4, cc = it % C4, c = 4 * cc;
            const int64_t p = pbase + pl;
            if (p >= n) continue;
            const float4 wx = ld4(chan + c), wy = ld4(chan + C + c), wz = ld4(chan + 2 * C + c), bp = ld4(chan + 3 * C + c);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            const int wo = c % WC;
#pragma unroll
            for (int s0 = 0; s0 < NS; s0 += 8) {
                float4 vv[8];
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const int j = sIdx[pl * NS + s0 + b];
                    vv[b] = j >= 0 ? ld4(v + (int64_t)j * ldv + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const int row = pl * NS + s0 + b;
                    const float4 h = lds4(sH + 4 * row);
                    const float4 w = lds4(sL + row * WC + wo);
                    a.x = fmaf(vv[b].x + fmaf(wx.x, h.x, fmaf(wy.x, h.y, fmaf(wz.x, h.z, bp.x))), w.x, a.x);
                    a.y = fmaf(vv[b].y + fmaf(wx.y, h.x, fmaf(wy.y, h.y, fmaf(wz.y, h.z, bp.y))), w.y, a.y);
                    a.z = fmaf(vv[b].z + fmaf(wx.z, h.x, fmaf(wy.z, h.y, fmaf(wz.z, h.z, bp.z))), w.z, a.z);
                    a.w = fmaf(vv[b].w + fmaf(wx.w, h.x, fmaf(wy.w, h.y, fmaf(wz.w, h.z, bp.w))), w.w, a.w);
                }
            }
            if (out_affine) {
                const float4 oa = ld4(chan + 6 * C + c), ob = ld4(chan + 7 * C + c);
                a.x = fmaxf(fmaf(oa.x, a.x, ob.x), 0.f); a.y = fmaxf(fmaf(oa.y, a.y, ob.y), 0.f);
                a.z = fmaxf(fmaf(oa.z, a.z, ob.z), 0.f); a.w = fmaxf(fmaf(oa.w, a.w, ob.w), 0.f);
            }
            *reinterpret_cast<float4*>(out + p * ldo + c) = a;
        }
    }
}

template <int C, int NS>
static int launch_pt_layer_tile(int64_t n, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                                int64_t ldv, const float* xyz, const int* idx, const float* params, int out_affine,
                                float* out, int64_t ldo, cudaStream_t stream) {
    using T = PtTile<C, NS>;
    constexpr size_t smem = sizeof(float) * T::FLOATS;
    auto kern = pt_layer_tile_kernel<C, NS>;
    if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = ceil_div(n, T::P);
    kern<<<(unsigned)grid, T::PTT_THREADS, smem, stream>>>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// How many warps share a point: enough that the launch fills the machine (~16k warps), at most the
// neighbour count, and at most 8 for C = 512 (16 warps = 512 threads would cap the kernel at 128 registers).
static inline int pt_layer_split(int64_t n, int nsample, int c, int hint) {
    const int top = c >= 512 ? 8 : nsample;
    int nsw = hint;
    if (nsw != 1 && nsw != 4 && nsw != 8 && nsw != 16) nsw = n >= 16384 ? 1 : (n >= 4096 ? 4 : top);
    if (nsw > top) nsw = top;
    if (nsw == 8 && top == 16) nsw = 4;   // instantiated: 1, 4 and `top`
    return nsw;
}

// out[r, c] = relu?( x[r, c] * scale[c] + shift[c] + res[r, c] ), 4 channels per thread
__global__ void __launch_bounds__(256)
affine_act_kernel(int64_t total4, int c4, const float4* __restrict__ x, const float4* __restrict__ scale,
                  const float4* __restrict__ shift, const float4* __restrict__ res, int relu, float4* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total4; t += (int64_t)gridDim.x * 256) {
        const int cc = (int)(t % c4);
        float4 a = x[t];
        if (scale) { const float4 s = __ldg(scale + cc); a.x *= s.x; a.y *= s.y; a.z *= s.z; a.w *= s.w; }
        if (shift) { const float4 s = __ldg(shift + cc); a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w; }
        if (res) { const float4 s = __ldg(res + t); a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w; }
        if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        out[t] = a;
    }
}

// out[p, :] = base[p, :] + sum_{i<k} in[idx[p,i], :] * w[p,i]; one thread per 4 channels
__global__ void __launch_bounds__(256)
interpolation_add_kernel(int64_t total4, int c4, int k, const float4* __restrict__ in, const int* __restrict__ idx,
                         const float* __restrict__ w, const float4* __restrict__ base, float4* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total4; t += (int64_t)gridDim.x * 256) {
        const int64_t p = t / c4;
        const int cc = (int)(t - p * c4);
        float4 a = base ? __ldg(base + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < k; i++) {
            const int src = __ldg(idx + p * k + i);
            const float wi = __ldg(w + p * k + i);
            const float4 f = __ldg(in + (int64_t)src * c4 + cc);
            a.x = fmaf(f.x, wi, a.x); a.y = fmaf(f.y, wi, a.y); a.z = fmaf(f.z, wi, a.z); a.w = fmaf(f.w, wi, a.w);
        }
        out[t] = a;
    }
}

// out[m, c] = max_s relu( scale[c] * (z[idx[m,s], c] + wxyz[c] . (xyz[idx[m,s]] - new_xyz[m])) + shift[c] )
// one thread per 4 channels of an output row; the ns gathered rows come from the L2-resident z
__global__ void __launch_bounds__(256)
transition_down_pool_kernel(int64_t total4, int c4, int ns, const float4* __restrict__ z, const float* __restrict__ xyz,
                            const float* __restrict__ new_xyz, const int* __restrict__ idx, const float* __restrict__ wxyz,
                            const float4* __restrict__ scale, const float4* __restrict__ shift, float4* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total4; t += (int64_t)gridDim.x * 256) {
        const int64_t m = t / c4;
        const int cc = (int)(t - m * c4);
        const float4 sc = __ldg(scale + cc), sh = __ldg(shift + cc);
        float w[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int d = 0; d < 3; d++) w[i][d] = __ldg(wxyz + (cc * 4 + i) * 3 + d);
        const float qx = __ldg(new_xyz + m * 3), qy = __ldg(new_xyz + m * 3 + 1), qz = __ldg(new_xyz + m * 3 + 2);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int s = 0; s < ns; s++) {
            const int j = __ldg(idx + m * ns + s);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j >= 0) {
                const float rx = __ldg(xyz + (int64_t)j * 3) - qx, ry = __ldg(xyz + (int64_t)j * 3 + 1) - qy,
                            rz = __ldg(xyz + (int64_t)j * 3 + 2) - qz;
                a = __ldg(z + (int64_t)j * c4 + cc);
                a.x += fmaf(w[0][0], rx, fmaf(w[0][1], ry, w[0][2] * rz));
                a.y += fmaf(w[1][0], rx, fmaf(w[1][1], ry, w[1][2] * rz));
                a.z += fmaf(w[2][0], rx, fmaf(w[2][1], ry, w[2][2] * rz));
                a.w += fmaf(w[3][0], rx, fmaf(w[3][1], ry, w[3][2] * rz));
            }
            best.x = fmaxf(best.x, fmaf(a.x, sc.x, sh.x)); best.y = fmaxf(best.y, fmaf(a.y, sc.y, sh.y));
            best.z = fmaxf(best.z, fmaf(a.z, sc.z, sh.z)); best.w = fmaxf(best.w, fmaf(a.w, sc.w, sh.w));
        }
        best.x = fmaxf(best.x, 0.f); best.y = fmaxf(best.y, 0.f); best.z = fmaxf(best.z, 0.f); best.w = fmaxf(best.w, 0.f);
        out[t] = best;
    }
}

static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace pob

using namespace pob;

// `split` (per call; tests / tuning): 0 (default) = the CTA-tiled kernel (16-byte aligned rows; otherwise
// warp-per-point); 1, 4, 8, 16 = force the warp-per-point kernel with that many warps sharing a point (clamped to
// what the shape instantiates); -1 = warp-per-point with the split chosen from n.  Only f32 summation order differs.

POB_API int64_t pob_pt_layer_param_floats(int c, int w_c) {
    if (c < 1 || w_c < 1) return 0;
    return ptl_param_floats(c, w_c);
}

// Eval-mode PointTransformerLayer forward (point_transformer_seg.py:48-81 with every BatchNorm in
// inference mode).  q, k, v: (n, c) rows with row strides ldq / ldk / ldv floats (so the three can
// be column blocks of one (n, 3c) GEMM output); xyz (n, 3); idx (n, nsample) self-kNN rows, -1 =
// placeholder (zero k / v row and zero relative position, as pointops.grouping masks them);
// params: the packed block described at ptl_param_floats; out (n, c) with row stride ldo.
// Supported: c in {32, 64, 128, 256, 512} with w_c = c / 8 (share_planes = 8), nsample in {8, 16}.
POB_API int pob_pt_layer_forward(int64_t n, int nsample, int c, int w_c, const float* q, int64_t ldq, const float* k,
                                 int64_t ldk, const float* v, int64_t ldv, const float* xyz, const int* idx,
                                 const float* params, int out_affine, float* out, int64_t ldo, int split,
                                 cudaStream_t stream) {
    const int g_ptl_split = split;
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!q || !k || !v || !xyz || !idx || !params || !out) return POB_ERR_BAD_ARG;
    if (w_c * 8 != c) return POB_ERR_UNSUPPORTED;
    if (c >= 128 && (!al16(q) || !al16(k) || !al16(v) || !al16(out) || (ldq | ldk | ldv | ldo) % 4)) return POB_ERR_BAD_ARG;
    if (!al16(params)) return POB_ERR_BAD_ARG;
    const bool vec_ok = al16(q) && al16(k) && al16(v) && al16(out) && (ldq | ldk | ldv | ldo) % 4 == 0;
    if (vec_ok && g_ptl_split == 0) {   // default: the CTA-tiled kernel
#define POB_PTT(CC, NSS) \
    if (c == CC && nsample == NSS) return launch_pt_layer_tile<CC, NSS>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo, stream)
        POB_PTT(32, 8); POB_PTT(64, 8); POB_PTT(128, 8); POB_PTT(256, 8); POB_PTT(512, 8);
        POB_PTT(32, 16); POB_PTT(64, 16); POB_PTT(128, 16); POB_PTT(256, 16); POB_PTT(512, 16);
#undef POB_PTT
    }
    const int nsw = vec_ok ? pt_layer_split(n, nsample, c, g_ptl_split) : 1;   // the merge writes 128-bit rows
#define POB_PTL(VEC, R, NS, WC, TOP)                                                                                        \
    do {                                                                                                                    \
        if (nsw == 1) return launch_pt_layer<VEC, R, NS, WC, 1>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine,   \
                                                                out, ldo, stream);                                          \
        if (nsw == 4) return launch_pt_layer<VEC, R, NS, WC, 4>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine,   \
                                                                out, ldo, stream);                                          \
        return launch_pt_layer<VEC, R, NS, WC, TOP>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo,     \
                                                    stream);                                                                \
    } while (0)
    if (nsample == 8) {
        switch (c) {
            case 32: POB_PTL(1, 1, 8, 4, 8);
            case 64: POB_PTL(1, 2, 8, 8, 8);
            case 128: POB_PTL(4, 1, 8, 16, 8);
            case 256: POB_PTL(4, 2, 8, 32, 8);
            case 512: POB_PTL(4, 4, 8, 64, 8);
        }
    } else if (nsample == 16) {
        switch (c) {
            case 32: POB_PTL(1, 1, 16, 4, 16);
            case 64: POB_PTL(1, 2, 16, 8, 16);
            case 128: POB_PTL(4, 1, 16, 16, 16);
            case 256: POB_PTL(4, 2, 16, 32, 16);
            case 512: POB_PTL(4, 4, 16, 64, 8);
        }
    }
#undef POB_PTL
    return POB_ERR_UNSUPPORTED;
}

// out = relu?(x * scale + shift + residual); scale / shift (c) and residual (rows, c) optional (NULL).
// c % 4 == 0, 16-byte aligned pointers.  In-place (out == x) is allowed.
POB_API int pob_affine_act(int64_t rows, int c, const float* x, const float* scale, const float* shift,
                           const float* residual, int relu, float* out, cudaStream_t stream) {
    if (rows < 0 || c < 1) return POB_ERR_BAD_ARG;
    if (rows == 0) return 0;
    if (!x || !out) return POB_ERR_BAD_ARG;
    if (c % 4 || !al16(x) || !al16(out) || !al16(scale) || !al16(shift) || !al16(residual)) return POB_ERR_UNSUPPORTED;
    const int64_t total4 = rows * (c / 4);
    affine_act_kernel<<<grid_for(total4, 256, 8), 256, 0, stream>>>(total4, c / 4, (const float4*)x, (const float4*)scale,
                                                                   (const float4*)shift, (const float4*)residual, relu,
                                                                   (float4*)out);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// TransitionDown (point_transformer_seg.py:106-119) after its Linear(3 + C, C') has been split by
// linearity: z (n, c) = feat @ W[:, 3:]^T on the UNGATHERED points; wxyz (c, 3) = W[:, :3];
// out (m, c) = max_s relu(scale * (z[idx[m,s]] + wxyz (xyz[idx[m,s]] - new_xyz[m])) + shift); idx < 0
// contributes a zero grouped row.  c % 4 == 0.
POB_API int pob_transition_down_pool(int64_t m, int nsample, int c, const float* z, const float* xyz, const float* new_xyz,
                                     const int* idx, const float* wxyz, const float* scale, const float* shift,
                                     float* out, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!z || !xyz || !new_xyz || !idx || !wxyz || !scale || !shift || !out) return POB_ERR_BAD_ARG;
    if (c % 4 || !al16(z) || !al16(out) || !al16(scale) || !al16(shift)) return POB_ERR_UNSUPPORTED;
    const int64_t total4 = m * (c / 4);
    transition_down_pool_kernel<<<grid_for(total4, 256, 8), 256, 0, stream>>>(total4, c / 4, nsample, (const float4*)z, xyz,
                                                                             new_xyz, idx, wxyz, (const float4*)scale,
                                                                             (const float4*)shift, (float4*)out);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// output = base + three-NN interpolation of input (interpolation_forward with the skip connection of
// TransitionUp, point_transformer_seg.py:168-170, folded in); base may be NULL.  c % 4 == 0.
POB_API int pob_interpolation_add_forward(int64_t n, int c, int k, const float* input, const int* idx, const float* weight,
                                          const float* base, float* output, cudaStream_t stream) {
    if (n < 0 || c < 1 || k < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!input || !idx || !weight || !output) return POB_ERR_BAD_ARG;
    if (c % 4 || !al16(input) || !al16(output) || !al16(base)) return POB_ERR_UNSUPPORTED;
    const int64_t total4 = n * (c / 4);
    interpolation_add_kernel<<<grid_for(total4, 256, 8), 256, 0, stream>>>(total4, c / 4, k, (const float4*)input, idx,
                                                                          weight, (const float4*)base, (float4*)output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}
