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

template <int VEC, int R, int NS, int WC>
__global__ void __launch_bounds__(PTL_THREADS)
pt_layer_fwd_kernel(int64_t n, const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                    const float* __restrict__ v, int64_t ldv, const float* __restrict__ xyz,
                    const int* __restrict__ idx, const float* __restrict__ params, int out_affine,
                    float* __restrict__ out, int64_t ldo) {
    constexpr int C = 32 * VEC * R;
    constexpr int NV = WC >= 32 ? WC / 32 : 1;
    constexpr int JS = 32 * VEC;  // channel stride between a lane's j-slices
    extern __shared__ __align__(16) float smem[];
    float* w1s = smem;              // [WC][C]
    float* b1s = w1s + WC * C;      // [WC]
    float* w2ts = b1s + WC;         // [WC][WC]
    float* b2s = w2ts + WC * WC;    // [WC]
    float* us = b2s + WC;           // [PTL_WARPS][WC]
    {
        const float4* src = reinterpret_cast<const float4*>(params + 16 + 8 * C);
        float4* dst = reinterpret_cast<float4*>(smem);
        constexpr int N4 = (WC * C + WC + WC * WC + WC) / 4;
        for (int i = threadIdx.x; i < N4; i += PTL_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = VEC == 1 ? lane : 4 * lane;  // channel (j = 0, i = 0); channel(j, i) = c0 + i + JS * j
    const int o0 = c0 % WC;                     // the VEC weight channels this lane's channels use
    float* uw = us + warp * WC;
    const float* chan = params + 16;
    float A[9], cb[3];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = __ldg(params + i);
#pragma unroll
    for (int i = 0; i < 3; i++) cb[i] = __ldg(params + 9 + i);

    for (int64_t p = (int64_t)blockIdx.x * PTL_WARPS + warp; p < n; p += (int64_t)gridDim.x * PTL_WARPS) {
        // lane s < NS: neighbour s -- its row, and the hidden layer of linear_p on its relative position
        int js = -1;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        if (lane < NS) {
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
        for (int j = 0; j < R; j++) qv[j] = Vf<VEC>::ld(q + p * ldq + c0 + JS * j);
        float mx[VEC], den[VEC], o_acc[R][VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) { mx[i] = -INFINITY; den[i] = 0.f; }
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int i = 0; i < VEC; i++) o_acc[j][i] = 0.f;

#pragma unroll 1
        for (int s = 0; s < NS; s++) {
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
            o.st(out + p * ldo + c);
        }
    }
}

template <int VEC, int R, int NS, int WC>
static int launch_pt_layer(int64_t n, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                           const float* xyz, const int* idx, const float* params, int out_affine, float* out, int64_t ldo,
                           cudaStream_t stream) {
    constexpr int C = 32 * VEC * R;
    constexpr size_t smem = sizeof(float) * (WC * C + WC + WC * WC + WC + PTL_WARPS * WC);
    auto kern = pt_layer_fwd_kernel<VEC, R, NS, WC>;
    // the opt-in is per device, so it is (cheaply) repeated on every launch; occupancy is a property
    // of the instantiation on sm_100a and is looked up once
    if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int occ = 0;
        POB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PTL_THREADS, smem));
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    int64_t grid = ceil_div(n, PTL_WARPS);
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, PTL_THREADS, smem, stream>>>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
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
                                 const float* params, int out_affine, float* out, int64_t ldo, cudaStream_t stream) {
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!q || !k || !v || !xyz || !idx || !params || !out) return POB_ERR_BAD_ARG;
    if (w_c * 8 != c) return POB_ERR_UNSUPPORTED;
    if (c >= 128 && (!al16(q) || !al16(k) || !al16(v) || !al16(out) || (ldq | ldk | ldv | ldo) % 4)) return POB_ERR_BAD_ARG;
    if (!al16(params)) return POB_ERR_BAD_ARG;
#define POB_PTL(VEC, R, NS, WC) \
    return launch_pt_layer<VEC, R, NS, WC>(n, q, ldq, k, ldk, v, ldv, xyz, idx, params, out_affine, out, ldo, stream)
    if (nsample == 8) {
        switch (c) {
            case 32: POB_PTL(1, 1, 8, 4);
            case 64: POB_PTL(1, 2, 8, 8);
            case 128: POB_PTL(4, 1, 8, 16);
            case 256: POB_PTL(4, 2, 8, 32);
            case 512: POB_PTL(4, 4, 8, 64);
        }
    } else if (nsample == 16) {
        switch (c) {
            case 32: POB_PTL(1, 1, 16, 4);
            case 64: POB_PTL(1, 2, 16, 8);
            case 128: POB_PTL(4, 1, 16, 16);
            case 256: POB_PTL(4, 2, 16, 32);
            case 512: POB_PTL(4, 4, 16, 64);
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
