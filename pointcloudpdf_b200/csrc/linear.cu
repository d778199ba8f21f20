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

template <int BM, int BN, int BK, int KS, int TM, int TN, bool VEC_K, bool VEC_N>
__global__ void __launch_bounds__((BM / TM) * (BN / TN) * KS, TM * TN > 16 ? 512 / ((BM / TM) * (BN / TN) * KS) : 0)   // big register tiles: <= 128 registers
linear_tile_kernel(int64_t M, int K, int N, int col_tiles, const float* __restrict__ A, int64_t lda,
                   const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ residual,
                   int64_t ldr, int relu, float* __restrict__ out, int64_t ldo) {
    constexpr int TX = BN / TN, TY = BM / TM, G = TX * TY, NT = G * KS;
    constexpr int RH = TM / 4, CH = TN / 4;      // a thread's rows / columns come in RH / CH runs of 4, BM/RH (BN/CH) apart
    constexpr int KPG = BK / KS;                 // k-steps of one thread group per k-tile
    constexpr int AS = BM + 4;                   // row stride of As[k][.] (multiple of 4: float4 reads stay aligned)
    constexpr int FA = BM * BK / 4, FB = BK * BN / 4;           // float4s per tile
    constexpr int LA = (FA + NT - 1) / NT, LB = (FB + NT - 1) / NT;
    constexpr int TILE_FLOATS = 2 * BK * (AS + BN);
    constexpr int RED_FLOATS = KS > 1 ? (KS / 2) * BM * BN : 0;
    constexpr int SMEM_FLOATS = TILE_FLOATS > RED_FLOATS ? TILE_FLOATS : RED_FLOATS;
    static_assert(BK % KS == 0 && (TM == 4 || TM == 8) && (TN == 4 || TN == 8) && BM % TM == 0 && BN % TN == 0, "tile shape");
    static_assert((KS & (KS - 1)) == 0 && SMEM_FLOATS * 4 <= 48 * 1024, "split-K is a power of two; static shared memory");
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

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

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
            float ar[TM], br[TN];
#pragma unroll
            for (int h = 0; h < RH; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(a + kk * AS + h * (BM / RH));
                ar[h * 4 + 0] = v.x; ar[h * 4 + 1] = v.y; ar[h * 4 + 2] = v.z; ar[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < CH; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(b + kk * BN + h * (BN / CH));
                br[h * 4 + 0] = v.x; br[h * 4 + 1] = v.y; br[h * 4 + 2] = v.z; br[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
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

    if (KS > 1) {
        // split-K: pairwise tree over the thread groups, accumulators exchanged as float4 [quad][thread] (conflict
        // free); the loop above ended on a barrier, so the tile buffers are free.  Fixed order: deterministic.
        float4* red = reinterpret_cast<float4*>(smem);
        constexpr int Q = TM * TN / 4;
#pragma unroll
        for (int half = KS / 2; half >= 1; half >>= 1) {
            if (g >= half && g < 2 * half) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int h = 0; h < CH; ++h)
                        red[((g - half) * Q + i * CH + h) * G + r] =
                            make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            }
            __syncthreads();
            if (g < half) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int h = 0; h < CH; ++h) {
                        const float4 p = red[(g * Q + i * CH + h) * G + r];
                        acc[i][h * 4] += p.x; acc[i][h * 4 + 1] += p.y; acc[i][h * 4 + 2] += p.z; acc[i][h * 4 + 3] += p.w;
                    }
            }
            if (half > 1) __syncthreads();       // the next round overwrites what this one read
        }
        if (g != 0) return;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int h = 0; h < CH; ++h)
            finish(m0 + (i / 4) * (BM / RH) + ty * 4 + (i % 4), n0 + h * (BN / CH) + tx * 4,
                   make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]));
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int BM, int BN, int BK, int KS, int TM, int TN>
static int launch_linear(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                         const float* residual, int64_t ldr, int relu, float* out, int64_t ldo, cudaStream_t stream) {
    const int col_tiles = (N + BN - 1) / BN;
    const int64_t row_tiles = (M + BM - 1) / BM;
    const int64_t ctas = row_tiles * col_tiles;
    if (ctas > 0x7fffffffLL) return POB_ERR_UNSUPPORTED;
    constexpr int NT = (BM / TM) * (BN / TN) * KS;
    const dim3 grid((unsigned)ctas), block(NT);
    linear_tile_kernel<BM, BN, BK, KS, TM, TN, true, true><<<grid, block, 0, stream>>>(M, K, N, col_tiles, A, lda, Wt, bias,
                                                                                      residual, ldr, relu, out, ldo);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// shapes that are not 16-byte friendly (the 6-channel input layer, the 13-class head): one tile shape, scalar
// loads on the unaligned side
static int launch_linear_unaligned(int64_t M, int K, int N, bool vec_k, bool vec_n, const float* A, int64_t lda,
                                   const float* Wt, const float* bias, const float* residual, int64_t ldr, int relu,
                                   float* out, int64_t ldo, cudaStream_t stream) {
    constexpr int BM = 128, BN = 32;
    const int col_tiles = (N + BN - 1) / BN;
    const int64_t ctas = ((M + BM - 1) / BM) * col_tiles;
    if (ctas > 0x7fffffffLL) return POB_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)ctas), block(256);
#define POB_LINEAR_LAUNCH(VK, VN)                                                                                   \
    linear_tile_kernel<BM, BN, 16, 1, 4, 4, VK, VN><<<grid, block, 0, stream>>>(M, K, N, col_tiles, A, lda, Wt, bias, \
                                                                              residual, ldr, relu, out, ldo)
    if (vec_k) POB_LINEAR_LAUNCH(true, false);
    else if (vec_n) POB_LINEAR_LAUNCH(false, true);
    else POB_LINEAR_LAUNCH(false, false);
#undef POB_LINEAR_LAUNCH
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

}  // namespace pob

using namespace pob;

#define POB_LINEAR_CONFIGS 16   // `config` (per call): 0 = pick the tile from the shape; 1..16 = force one (tests, tuning)

// Tile choice, from the per-shape timings on B200 (profiles/r01d_linear_time.txt): the shapes are skinny
// and small (0.16-0.5 GFLOP), so a launch is latency bound -- what matters is >= ~2 CTAs per SM with as many
// warps as possible in flight, split-K where the rows alone do not provide them.  8 x 8 register tiles only
// pay for the widest outputs.
static int pick_linear_config(int64_t M, int K, int N) {
    (void)K;
    if (M >= 40000) return N <= 32 ? 1 : (N <= 64 ? 5 : 8);
    if (M >= 10000) return N <= 128 ? 5 : 10;
    if (M >= 2500) return N <= 256 ? 6 : 10;
    if (M >= 600) return N <= 512 ? 7 : 11;
    return N <= 256 ? 7 : (N <= 512 ? 4 : 16);
}

// out (M, N) = act(A (M, K) @ Wt (K, N) + bias (N) + residual (M, N)); bias / residual may be NULL; relu != 0
// applies max(., 0).  Row strides lda / ldr / ldo in floats (unit column stride); Wt is dense.  Any K, N >= 1:
// 128-bit paths when K, N, the strides and the pointers are 16-byte friendly, scalar loads / stores otherwise.
// out must not alias A or residual (both are read through the read-only data path).
POB_API int pob_linear_forward(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                               const float* residual, int64_t ldr, int relu, float* out, int64_t ldo, int config,
                               cudaStream_t stream) {
    if (M < 0 || K < 1 || N < 1 || lda < K || ldo < N || (residual && ldr < N) || config < 0 || config > POB_LINEAR_CONFIGS)
        return POB_ERR_BAD_ARG;
    const int g_linear_force = config;
    if (M == 0) return 0;
    if (!A || !Wt || !out) return POB_ERR_BAD_ARG;
    const bool vec_k = K % 4 == 0 && lda % 4 == 0 && al16p(A);
    const bool vec_n = N % 4 == 0 && ldo % 4 == 0 && al16p(Wt) && al16p(out) && al16p(bias) &&
                       (!residual || (ldr % 4 == 0 && al16p(residual)));
    if (!(vec_k && vec_n))
        return launch_linear_unaligned(M, K, N, vec_k, vec_n, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
    const int cfg = g_linear_force ? g_linear_force : pick_linear_config(M, K, N);
#define POB_LINEAR_CASE(ID, BM, BN, BK, KS, TM, TN) \
    case ID: return launch_linear<BM, BN, BK, KS, TM, TN>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream)
    switch (cfg) {
        POB_LINEAR_CASE(1, 128, 32, 16, 1, 4, 4);     // 256 threads
        POB_LINEAR_CASE(2, 64, 32, 16, 2, 4, 4);
        POB_LINEAR_CASE(3, 32, 32, 32, 4, 4, 4);
        POB_LINEAR_CASE(4, 16, 32, 32, 8, 4, 4);
        POB_LINEAR_CASE(5, 64, 64, 16, 1, 4, 4);
        POB_LINEAR_CASE(6, 32, 64, 32, 2, 4, 4);
        POB_LINEAR_CASE(7, 16, 64, 32, 4, 4, 4);
        POB_LINEAR_CASE(8, 128, 32, 16, 1, 8, 4);     // 128 threads, 8 x 4 register tiles
        POB_LINEAR_CASE(9, 256, 32, 16, 1, 8, 4);     // 256 threads
        POB_LINEAR_CASE(10, 128, 64, 16, 1, 8, 8);    // 128 threads, 8 x 8 register tiles
        POB_LINEAR_CASE(11, 64, 64, 16, 2, 8, 8);     // 128 threads
        POB_LINEAR_CASE(12, 64, 64, 32, 4, 8, 8);     // 256 threads
        POB_LINEAR_CASE(13, 32, 64, 32, 8, 8, 8);     // 256 threads
        POB_LINEAR_CASE(14, 128, 64, 16, 2, 8, 8);    // 256 threads
        POB_LINEAR_CASE(15, 64, 128, 16, 2, 8, 8);    // 256 threads
        default: POB_LINEAR_CASE(16, 32, 128, 32, 4, 8, 8);   // 256 threads
    }
#undef POB_LINEAR_CASE
}
