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

// ---------------------------------------------------------------------------------------------
// Tensor-core form (round 2): the same contract on mma.sync m16n8k8 TF32 with the "3xTF32" split, which keeps the
// result f32-accurate: every f32 operand is split a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi), and
//      a * b  ~=  a_lo * b_hi + a_hi * b_lo + a_hi * b_hi          (the dropped a_lo * b_lo is ~2^-22 relative)
// is accumulated in f32 -- measured error vs float64 ~5e-7 of the largest output, against ~1e-7 for FFMA and ~1e-3 for
// plain TF32 (which the parity bars of this path exclude).  Why: these linears are short, skinny GEMMs (0.16 - 0.5 GFLOP,
// K up to 512) whose FFMA k-loop is 1 000 dependent issue slots per 32-wide k-tile and thread; three MMAs per 16 x 8 x 8
// block replace 128 FFMAs per lane, so the k-loop is bound by the shared-memory fragment loads instead, and the 80 000-row
// layers become memory-bound.  (tcgen05 / TMEM would be the Blackwell form of a LARGE contraction; at a quarter GFLOP
// per launch the launch and the k-loop latency decide, and warp-level MMA needs no TMEM allocation, descriptors or TMA
// maps per launch.)
// CTA tile BM x BN, WM x WN warps, k-tiles of 32 staged by cp.async (16-byte, zero-filled tails) into padded shared
// memory: As[m][32 + 4] and Bs[k][BN + 8] make every fragment load conflict-free.  Epilogue on the accumulators.
__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32)
linear_mma_kernel(int64_t M, int K, int N, int col_tiles, const float* __restrict__ A, int64_t lda,
                  const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ residual,
                  int64_t ldr, int relu, float* __restrict__ out, int64_t ldo) {
    constexpr int BK = 32, NT = WM * WN * 32;
    constexpr int TM = BM / WM, TN = BN / WN;          // warp tile
    constexpr int MT = TM / 16, NTL = TN / 8;          // MMA tiles per warp
    constexpr int AS = BK + 4, BS = BN + 8;
    constexpr int STAGE_FLOATS = BM * AS + BK * BS;
    static_assert(TM % 16 == 0 && TN % 8 == 0 && STAGES >= 2, "warp tile is a multiple of 16 x 8");
    extern __shared__ __align__(16) float lin_smem[];   // [STAGES][BM * AS + BK * BS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    const int64_t m0 = (int64_t)(blockIdx.x / col_tiles) * BM;
    const int n0 = (int)(blockIdx.x % col_tiles) * BN;
    const int tiles = (K + BK - 1) / BK;

    // STAGES - 1 k-tiles are in flight while one is multiplied.  Two stages is what is instantiated: deeper pipelines
    // were measured no faster (the deep shapes wait on the dependent MMA chain, not on memory) and cost CTAs per SM.
    auto load_tile = [&](int kt) {
        if (kt < tiles) {
            float* as = lin_smem + (kt % STAGES) * STAGE_FLOATS;
            float* bs = as + BM * AS;
            const int k0 = kt * BK;
            for (int c = tid; c < BM * (BK / 4); c += NT) {
                const int row = c / (BK / 4), kq = c % (BK / 4);
                const int64_t gm = m0 + row;
                const int gk = k0 + kq * 4;
                const int bytes = (gm < M && gk < K) ? min(16, (K - gk) * 4) : 0;
                cp_async16(as + row * AS + kq * 4, A + (gm < M ? gm : 0) * lda + (gk < K ? gk : 0), bytes);
            }
            for (int c = tid; c < BK * (BN / 4); c += NT) {
                const int kk = c / (BN / 4), nq = c % (BN / 4);
                const int gk = k0 + kk, gn = n0 + nq * 4;
                const int bytes = (gk < K && gn < N) ? min(16, (N - gn) * 4) : 0;
                cp_async16(bs + kk * BS + nq * 4, Wt + (int64_t)(gk < K ? gk : 0) * N + (gn < N ? gn : 0), bytes);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // an empty group keeps the wait arithmetic uniform
    };

    float acc[MT][NTL][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) load_tile(s);
    for (int kt = 0; kt < tiles; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");   // tile kt has landed
        __syncthreads();                      // ... for everyone, and tile kt - 1 is no longer being read
        load_tile(kt + STAGES - 1);           // into the buffer of tile kt - 1
        const float* a = lin_smem + (kt % STAGES) * STAGE_FLOATS + (wm * TM) * AS;
        const float* b = lin_smem + (kt % STAGES) * STAGE_FLOATS + BM * AS + wn * TN;
#pragma unroll
        for (int k8 = 0; k8 < BK; k8 += 8) {
            unsigned ah[MT][4], al[MT][4], bh[NTL][2], bl[NTL][2];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const float* ap = a + (i * 16 + g) * AS + k8 + t;
                const float v[4] = {ap[0], ap[8 * AS], ap[4], ap[8 * AS + 4]};   // (g, t) (g+8, t) (g, t+4) (g+8, t+4)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    ah[i][r] = f2tf32(v[r]);
                    al[i][r] = f2tf32(v[r] - __uint_as_float(ah[i][r]));
                }
            }
#pragma unroll
            for (int j = 0; j < NTL; ++j) {
                const float* bp = b + (k8 + t) * BS + j * 8 + g;
                const float v[2] = {bp[0], bp[4 * BS]};                             // (k = t, n = g) (k = t + 4, n = g)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    bh[j][r] = f2tf32(v[r]);
                    bl[j][r] = f2tf32(v[r] - __uint_as_float(bh[j][r]));
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NTL; ++j) {
                    mma_tf32(acc[i][j], al[i], bh[j]);    // small terms first
                    mma_tf32(acc[i][j], ah[i], bl[j]);
                    mma_tf32(acc[i][j], ah[i], bh[j]);
                }
        }
    }

    // epilogue: accumulator r of MMA tile (i, j) is row (g + 8 * (r / 2)), column (2 * t + r % 2)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t gm = m0 + wm * TM + i * 16 + g + 8 * h;
                const int gn = n0 + wn * TN + j * 8 + 2 * t;
                if (gm >= M || gn >= N) continue;   // N % 4 == 0 on this path: a column pair is inside or outside together
                float2 v = make_float2(acc[i][j][2 * h], acc[i][j][2 * h + 1]);
                if (bias) { const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + gn)); v.x += bb.x; v.y += bb.y; }
                if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + gm * ldr + gn)); v.x += rr.x; v.y += rr.y; }
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                *reinterpret_cast<float2*>(out + gm * ldo + gn) = v;
            }
}

template <int BM, int BN, int WM, int WN, int STAGES>
static int launch_linear_mma(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                             const float* residual, int64_t ldr, int relu, float* out, int64_t ldo, cudaStream_t stream) {
    const int col_tiles = (N + BN - 1) / BN;
    const int64_t ctas = ((M + BM - 1) / BM) * col_tiles;
    if (ctas > 0x7fffffffLL) return POB_ERR_UNSUPPORTED;
    constexpr size_t smem = sizeof(float) * STAGES * (BM * (32 + 4) + 32 * (BN + 8));
    auto kern = linear_mma_kernel<BM, BN, WM, WN, STAGES>;
    if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: every launch
    kern<<<(unsigned)ctas, WM * WN * 32, smem, stream>>>(M, K, N, col_tiles, A, lda, Wt, bias, residual, ldr, relu, out, ldo);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
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

#define POB_LINEAR_CONFIGS 21   // `config` (per call): 0 = pick from the shape; 1..16 = force an FFMA tile, 17..20 = a tensor-core
                                // (3xTF32) tile, 21 = the FFMA tile the shape would pick (tests, tuning, A/B)

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
    // default (gpurun_out/r02_linear_time.json): the tensor-core (3xTF32) form for the long, residual-free shapes --
    // there the legacy HMMA path (about the A100 TF32 rate per SM, so 3 MMAs per product buy ~1.3x over FFMA at best)
    // is 10-30 % ahead; the short, deep shapes are bound by the launch and the dependent chain over K, where the
    // split-K FFMA tiles below are as fast or faster.
    int cfg = g_linear_force;
    if (cfg == 0) cfg = (M >= 10000 && !residual && N % 32 == 0) ? (N == 32 ? 17 : (N == 96 ? 18 : 19)) : 21;
    if (cfg == 21) cfg = pick_linear_config(M, K, N);
    switch (cfg) {
        case 17: return launch_linear_mma<128, 32, 4, 1, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 18: return launch_linear_mma<64, 96, 2, 3, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);   // q/k/v: N = 3C
        case 19: return launch_linear_mma<64, 64, 2, 2, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        case 20: return launch_linear_mma<32, 32, 2, 2, 2>(M, K, N, A, lda, Wt, bias, residual, ldr, relu, out, ldo, stream);
        default: break;
    }
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
