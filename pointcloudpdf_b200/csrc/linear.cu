// linear.cu -- FP32 linear layer with a fused epilogue for the frozen (inference) form of PTv1.
//
//   out[m, n] = act( sum_k A[m, k] * Wt[k, n] + bias[n] + residual[m, n] ),   act = ReLU or identity
//
// replaces, per Bottleneck of pointcept/models/point_transformer/point_transformer_seg.py:174-195 in eval
// mode, "linear1 + bn1 + ReLU" (cuBLAS SIMT GEMM + a cuBLASLt bias/ReLU pass), the q/k/v linears
// (point_transformer_seg.py:29-31, one GEMM on concatenated weights) and "linear3 + bn3 + skip + ReLU"
// (GEMM + pob_affine_act).  The shapes of this path are skinny (80 000 x 32 x 32 ... 312 x 512 x 1536: 0.16 to
// 0.5 GFLOP each) and cuBLAS' generic SIMT tiles run them at 10-20 % of the FP32 rate plus a second pass
// for the epilogue; here the tile is picked from the row count and the epilogue is applied on the
// accumulators.  FP32 FFMA on CUDA cores on purpose: "f32 means f32" (no TF32 rounding of the operands).
//
// Kernel: CTA tile BM x BN, thread tile 4 x 4, k-tiles of BK staged in shared memory k-major
// (As[k][m], Bs[k][n]; every fragment read is one conflict-free LDS.128, same-row lanes broadcast), next
// k-tile prefetched into registers while the current one is multiplied (one __syncthreads per k-tile).
// For short matrices the k-tile is split over KS thread groups inside the CTA (intra-CTA split-K, partial
// tiles summed through shared memory in a fixed order: deterministic) so that 312 rows still put 8 warps on
// every SM.  Wt is the weight stored K x N (N contiguous); weights are constants, the caller transposes once.
#include "common.cuh"

namespace pob {

template <bool VEC>
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int valid) {
    // valid = number of in-range elements starting at p (<= 0: none).  VEC: 16-byte aligned, valid is 0 or >= 4.
    if (VEC) {
        return valid > 0 ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        float4 v;
        v.x = valid > 0 ? __ldg(p) : 0.f;
        v.y = valid > 1 ? __ldg(p + 1) : 0.f;
        v.z = valid > 2 ? __ldg(p + 2) : 0.f;
        v.w = valid > 3 ? __ldg(p + 3) : 0.f;
        return v;
    }
}

template <int BM, int BN, int BK, int KS, bool VEC_K, bool VEC_N>
__global__ void __launch_bounds__((BM / 4) * (BN / 4) * KS)
linear_tile_kernel(int64_t M, int K, int N, int col_tiles, const float* __restrict__ A, int64_t lda,
                   const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ residual,
                   int64_t ldr, int relu, float* __restrict__ out, int64_t ldo) {
    constexpr int TX = BN / 4, TY = BM / 4, G = TX * TY, NT = G * KS;
    constexpr int KPG = BK / KS;                 // k-steps of one thread group per k-tile
    constexpr int AS = BM + 4;                   // row stride of As[k][.] (multiple of 4: float4 reads stay aligned)
    constexpr int FA = BM * BK / 4, FB = BK * BN / 4;           // float4s per tile
    constexpr int LA = (FA + NT - 1) / NT, LB = (FB + NT - 1) / NT;
    constexpr int TILE_FLOATS = 2 * BK * (AS + BN);
    constexpr int RED_FLOATS = KS > 1 ? KS * BM * BN : 0;
    constexpr int SMEM_FLOATS = TILE_FLOATS > RED_FLOATS ? TILE_FLOATS : RED_FLOATS;
    static_assert(BK % KS == 0 && BM % 4 == 0 && BN % 4 == 0, "tile shape");
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    float* As = smem;                            // [2][BK][AS]
    float* Bs = smem + 2 * BK * AS;              // [2][BK][BN]

    const int tid = threadIdx.x;
    const int g = tid / G, r = tid % G, tx = r % TX, ty = r / TX;
    const int64_t m0 = (int64_t)(blockIdx.x / col_tiles) * BM;
    const int n0 = (int)(blockIdx.x % col_tiles) * BN;

    float4 ra[LA], rb[LB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int f = tid + i * NT;
            if (FA % NT == 0 || f < FA) {
                const int row = f % BM, kq = f / BM;
                const int64_t gm = m0 + row;
                const int gk = k0 + kq * 4;
                ra[i] = load4<VEC_K>(A + (gm < M ? gm : 0) * lda + gk, gm < M ? K - gk : 0);
            }
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int f = tid + i * NT;
            if (FB % NT == 0 || f < FB) {
                const int kk = f / (BN / 4), nq = f % (BN / 4);
                const int gk = k0 + kk, gn = n0 + nq * 4;
                rb[i] = load4<VEC_N>(Wt + (int64_t)(gk < K ? gk : 0) * N + gn, gk < K ? N - gn : 0);
            }
        }
    };
    auto stash = [&](int buf) {
        float* a = As + buf * BK * AS;
        float* b = Bs + buf * BK * BN;
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int f = tid + i * NT;
            if (FA % NT == 0 || f < FA) {
                const int row = f % BM, kq = f / BM;
                a[(kq * 4 + 0) * AS + row] = ra[i].x;
                a[(kq * 4 + 1) * AS + row] = ra[i].y;
                a[(kq * 4 + 2) * AS + row] = ra[i].z;
                a[(kq * 4 + 3) * AS + row] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int f = tid + i * NT;
            if (FB % NT == 0 || f < FB) {
                const int kk = f / (BN / 4), nq = f % (BN / 4);
                *reinterpret_cast<float4*>(b + kk * BN + nq * 4) = rb[i];
            }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int tiles = (K + BK - 1) / BK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int t = 0; t < tiles; ++t) {
        const bool more = t + 1 < tiles;
        if (more) fetch((t + 1) * BK);           // global loads in flight under the FFMAs below
        const float* a = As + (t & 1) * BK * AS + (g * KPG) * AS + ty * 4;
        const float* b = Bs + (t & 1) * BK * BN + (g * KPG) * BN + tx * 4;
#pragma unroll
        for (int kk = 0; kk < KPG; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(a + kk * AS);
            const float4 bv = *reinterpret_cast<const float4*>(b + kk * BN);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (more) stash((t + 1) & 1);            // the other buffer: last read before the previous barrier
        __syncthreads();
    }

    // epilogue on a float4 of 4 consecutive columns of one row
    auto finish = [&](int64_t gm, int gn, float4 v) {
        if (gm >= M || gn >= N) return;
        const int valid = N - gn;
        if (bias) {
            const float4 bb = load4<VEC_N>(bias + gn, valid);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (residual) {
            const float4 rr = load4<VEC_N>(residual + gm * ldr + gn, valid);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        float* o = out + gm * ldo + gn;
        if (VEC_N) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            o[0] = v.x;
            if (valid > 1) o[1] = v.y;
            if (valid > 2) o[2] = v.z;
            if (valid > 3) o[3] = v.w;
        }
    };

    if (KS == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            finish(m0 + ty * 4 + i, n0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    } else {
        // the loop ended on a barrier: the tile buffers are free; partial tiles -> red[g][row][col]
        float* red = smem;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(red + ((g * BM) + ty * 4 + i) * BN + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __syncthreads();
        for (int f = tid; f < BM * BN / 4; f += NT) {
            const int row = f / (BN / 4), nq = f % (BN / 4);
            float4 s = *reinterpret_cast<const float4*>(red + row * BN + nq * 4);
#pragma unroll
            for (int h = 1; h < KS; ++h) {
                const float4 p = *reinterpret_cast<const float4*>(red + ((h * BM) + row) * BN + nq * 4);
                s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
            }
            finish(m0 + row, n0 + nq * 4, s);
        }
    }
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int BM, int BN, int BK, int KS>
static int launch_linear(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                         const float* residual, int64_t ldr, int relu, float* out, int64_t ldo, cudaStream_t stream) {
    const bool vec_k = K % 4 == 0 && lda % 4 == 0 && al16p(A);
    const bool vec_n = N % 4 == 0 && ldo % 4 == 0 && al16p(Wt) && al16p(out) && al16p(bias) &&
                       (!residual || (ldr % 4 == 0 && al16p(residual)));
    const int col_tiles = (N + BN - 1) / BN;
    const int64_t row_tiles = (M + BM - 1) / BM;
    const int64_t ctas = row_tiles * col_tiles;
    if (ctas > 0x7fffffffLL) return POB_ERR_UNSUPPORTED;
    constexpr int NT = (BM / 4) * (BN / 4) * KS;
    const dim3 grid((unsigned)ctas), block(NT);
#define POB_LINEAR_LAUNCH(VK, VN)                                                                                  \
    linear_tile_kernel<BM, BN, BK, KS, VK, VN><<<grid, block, 0, stream>>>(M, K, N, col_tiles, A, lda, Wt, bias,    \
                                                                         residual, ldr, relu, out, ldo)
    if (vec_k && vec_n) POB_LINEAR_LAUNCH(true, true);
    else if (vec_k) POB_LINEAR_LAUNCH(true, false);
    else if (vec_n) POB_LINEAR_LAUNCH(false, true);
    else POB_LINEAR_LAUNCH(false, false);
#undef POB_LINEAR_LAUNCH
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

}  // namespace pob

using namespace pob;

static int g_linear_force = 0;   // 0 = pick the tile from the shape; 1..7 = force a configuration (tests, tuning)

POB_API int pob_linear_set_config(int config) {
    if (config < 0 || config > 7) return POB_ERR_BAD_ARG;
    g_linear_force = config;
    return 0;
}

// out (M, N) = act(A (M, K) @ Wt (K, N) + bias (N) + residual (M, N)); bias / residual may be NULL; relu != 0
// applies max(., 0).  Row strides lda / ldr / ldo in floats (unit column stride); Wt is dense.  Any K, N >= 1:
// 128-bit paths when K, N, the strides and the pointers are 16-byte friendly, scalar loads / stores otherwise.
// out must not alias A or residual (both are read through the read-only data path).
POB_API int pob_linear_forward(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                               const float* residual, int64_t ldr, int relu, float* out, int64_t ldo,
                               cudaStream_t stream) {
    if (M < 0 || K < 1 || N < 1 || lda < K || ldo < N || (residual && ldr < N)) return POB_ERR_BAD_ARG;
    if (M == 0) return 0;
    if (!A || !Wt || !out) return POB_ERR_BAD_ARG;
    int cfg = g_linear_force;
    if (cfg == 0) cfg = M >= 16384 ? 1 : M >= 4096 ? 2 : M >= 1024 ? 3 : 4;
    switch (cfg) {
        case 1: return launch_linear<128, 32, 16, 1>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 2: return launch_linear<64, 32, 16, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 3: return launch_linear<32, 32, 32, 4>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 4: return launch_linear<16, 32, 32, 8>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 5: return launch_linear<64, 64, 16, 1>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 6: return launch_linear<32, 64, 32, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        default: return launch_linear<16, 64, 32, 4>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
    }
}
