// score.cu -- per-point open-set uncertainty scoring in ONE pass over the logits (sm_100a).
//
// Fuses what the reference computes with 3 + 3 + ~8 eager torch kernels and several (N,K)
// temporaries:
//   a9  MaxProbability  msp: -max log_softmax(logits); ml: -max logits
//       (pointcept/recognizers/max_probability/max_probability_v1m1_base.py:17-29)
//   a10 PointPdfV1 score: softmax(cat[logits, conf])[:, K]
//       (pointcept/recognizers/ours/pointpdf_v1m1_base.py:106-113)
//   a11 pseudo-label scoring prefix: max softmax prob, max logit, and per scene their sum,
//       sum of squares, min, max -> mean, unbiased std, stop = mean - beta*std, min-max
//       normalised max logit ((ml-min)/(max-min+1e-6))  (pointpdf_v1m1_base.py:199-222)
// The logits tile of a CTA (256 rows) is staged through shared memory with coalesced 128-bit
// loads (rows are K*4 bytes, K = 13 / 20: not vector-aligned on their own); each thread then
// owns one row.  Arithmetic is f32 (round 1 ran exp/log in f64: 14 us for 4.8 MB): expf / logf
// are the accurate (<= 2 ulp) library forms, the sum of K <= 64 terms in [0, 1] with the maximum
// term exactly 1 carries ~1e-7 relative error, log of a value in [1, K] turns that into ~1e-7
// absolute -- an order of magnitude inside the 1e-6 contract (the tests compare with torch's own
// f32 log_softmax, which has the same kind of error).  Per-scene statistics are accumulated in
// f64 by block reduction + atomics into a workspace the launcher zeroes with one memset node
// (keys are transformed so that 0 is the identity of both min and max), and the LAST block to
// finish turns them into the per-scene table -- no init / scene kernels.
// HBM-bound: 4*N*(K[+1]) bytes in, 4 bytes per requested output out.
#include "common.cuh"

namespace pob {

constexpr int SCORE_ROWS = 256;
constexpr int SCORE_MAX_K_SMEM = 44;  // (K|1) * 256 * 4 bytes <= 46 KB of static-limit shared memory

struct __align__(16) SceneStats {  // 64 bytes per scene; all-zero is the identity of every field
    double sum_msp, sumsq_msp, sum_ml, sumsq_ml;
    unsigned max_ml_key;       // atomicMax of ukey(max_logit)
    unsigned min_ml_keyc;      // atomicMax of ~ukey(max_logit)  (= min of ukey)
    unsigned done;             // workspace[b].done: blocks finished (last-block-done pattern); unused in the scene slots
    int pad[5];
};
static_assert(sizeof(SceneStats) == 64, "SceneStats layout");

__device__ __forceinline__ int okey(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float okey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
// unsigned, order-preserving, never 0 for a finite float's... 0 only for the most negative key: fine as identity
__device__ __forceinline__ unsigned ukey(float f) { return (unsigned)okey(f) ^ 0x80000000u; }
__device__ __forceinline__ float ukey_inv(unsigned u) { return okey_inv((int)(u ^ 0x80000000u)); }

struct RowScore {
    float msp_score, ml_score, pdf_score, msp_prob, max_logit;
    int pred;
};

template <typename Get>
__device__ __forceinline__ RowScore score_row(Get get, int K, bool has_conf, float conf) {
    float mx = get(0);
    int arg = 0;
    for (int c = 1; c < K; c++) {
        const float v = get(c);
        if (v > mx) { mx = v; arg = c; }  // first maximum, like torch.max
    }
    float s0 = 0.f, s1 = 0.f;             // two chains: halves the dependent-add latency
    int c = 0;
    for (; c + 1 < K; c += 2) { s0 += expf(get(c) - mx); s1 += expf(get(c + 1) - mx); }
    if (c < K) s0 += expf(get(c) - mx);
    const float sum = s0 + s1;            // in [1, K]: the maximum contributes exactly 1
    RowScore r;
    r.max_logit = mx;
    r.pred = arg;
    r.ml_score = -mx;
    r.msp_score = logf(sum);              // -(max - logsumexp)
    r.msp_prob = __fdiv_rn(1.0f, sum);    // max softmax probability
    r.pdf_score = 0.f;
    if (has_conf) {
        // softmax over K+1 entries, last column: exp(conf - M) / (sum*exp(mx - M) + exp(conf - M))
        const float M = fmaxf(mx, conf);
        const float e = expf(conf - M);
        r.pdf_score = __fdiv_rn(e, fmaf(sum, expf(mx - M), e));
    }
    return r;
}

__global__ void __launch_bounds__(SCORE_ROWS)
score_fused_kernel(int64_t n, int K, int b, int use_smem, const float* __restrict__ logits,
                   const float* __restrict__ conf, const int* __restrict__ offset, float* __restrict__ msp_score,
                   float* __restrict__ ml_score, float* __restrict__ pdf_score, float* __restrict__ msp_prob,
                   float* __restrict__ max_logit, int* __restrict__ pred, SceneStats* __restrict__ stats,
                   float beta, float* __restrict__ scene_out) {
    __shared__ float tile[SCORE_ROWS * (SCORE_MAX_K_SMEM | 1)];
    __shared__ double red[4][SCORE_ROWS / 32];
    __shared__ unsigned redi[2][SCORE_ROWS / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * SCORE_ROWS;
    const int rows = (int)min((int64_t)SCORE_ROWS, n - row0);
    const int Kp = K | 1;  // odd row pitch: conflict-free column walks
    if (use_smem) {
        const float* src = logits + row0 * K;  // 1024*K bytes per CTA: 16-byte aligned when logits is
        const int total = rows * K;
        const int nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? total / 4 : 0;
        for (int v = tid; v < nvec; v += SCORE_ROWS) {
            const float4 x = __ldcs(reinterpret_cast<const float4*>(src) + v);
            const float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int f = v * 4 + k;
                tile[(f / K) * Kp + f % K] = e[k];
            }
        }
        for (int f = nvec * 4 + tid; f < total; f += SCORE_ROWS) tile[(f / K) * Kp + f % K] = __ldcs(src + f);
        __syncthreads();
    }
    const int64_t row = row0 + tid;
    const bool live = tid < rows;
    RowScore r = {};
    if (live) {
        const bool has_conf = conf != nullptr;
        const float cf = has_conf ? __ldg(conf + row) : 0.f;
        if (use_smem) {
            const float* t = tile + tid * Kp;
            r = score_row([&](int c) { return t[c]; }, K, has_conf, cf);
        } else {
            const float* g = logits + row * K;
            r = score_row([&](int c) { return __ldg(g + c); }, K, has_conf, cf);
        }
        if (msp_score) msp_score[row] = r.msp_score;
        if (ml_score) ml_score[row] = r.ml_score;
        if (pdf_score) pdf_score[row] = r.pdf_score;
        if (msp_prob) msp_prob[row] = r.msp_prob;
        if (max_logit) max_logit[row] = r.max_logit;
        if (pred) pred[row] = r.pred;
    }
    if (!stats) return;
    // ---- per-scene statistics of msp_prob and max_logit ----
    const int s_first = segment_of(row0, offset, b);
    const int s_end = segment_of(row0 + rows - 1, offset, b);
    if (s_first == s_end) {
        double v[4] = {0, 0, 0, 0};
        unsigned kmin = 0u, kmax = 0u;   // kmin holds ~ukey: both reduce with max, identity 0
        if (live) {
            v[0] = r.msp_prob; v[1] = (double)r.msp_prob * r.msp_prob;
            v[2] = r.max_logit; v[3] = (double)r.max_logit * r.max_logit;
            kmax = ukey(r.max_logit); kmin = ~kmax;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] += __shfl_xor_sync(FULL, v[k], o);
        }
        kmin = __reduce_max_sync(FULL, kmin);
        kmax = __reduce_max_sync(FULL, kmax);
        const int lane = tid & 31, w = tid >> 5;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) red[k][w] = v[k];
            redi[0][w] = kmin; redi[1][w] = kmax;
        }
        __syncthreads();
        if (tid < 4) {
            double acc = 0;
            for (int k = 0; k < SCORE_ROWS / 32; k++) acc += red[tid][k];
            atomicAdd(&stats[s_first].sum_msp + tid, acc);
        } else if (tid == 4) {
            unsigned k0 = 0u;
            for (int k = 0; k < SCORE_ROWS / 32; k++) k0 = max(k0, redi[0][k]);
            atomicMax(&stats[s_first].min_ml_keyc, k0);
        } else if (tid == 5) {
            unsigned k1 = 0u;
            for (int k = 0; k < SCORE_ROWS / 32; k++) k1 = max(k1, redi[1][k]);
            atomicMax(&stats[s_first].max_ml_key, k1);
        }
    } else if (live) {  // tile straddles a scene boundary: rare, per-thread atomics
        SceneStats* st = stats + segment_of(row, offset, b);
        atomicAdd(&st->sum_msp, (double)r.msp_prob);
        atomicAdd(&st->sumsq_msp, (double)r.msp_prob * r.msp_prob);
        atomicAdd(&st->sum_ml, (double)r.max_logit);
        atomicAdd(&st->sumsq_ml, (double)r.max_logit * r.max_logit);
        atomicMax(&st->min_ml_keyc, ~ukey(r.max_logit));
        atomicMax(&st->max_ml_key, ukey(r.max_logit));
    }
    // ---- the last block to get here turns the accumulators into the per-scene table ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&stats[b].done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int s = tid; s < b; s += SCORE_ROWS) {
        const volatile SceneStats* vs = stats + s;
        const double sum_msp = vs->sum_msp, sumsq_msp = vs->sumsq_msp, sum_ml = vs->sum_ml, sumsq_ml = vs->sumsq_ml;
        const double cnt = (double)(offset[s] - (s == 0 ? 0 : offset[s - 1]));
        const float mn = ukey_inv(~vs->min_ml_keyc), mx = ukey_inv(vs->max_ml_key);
        const double den = (double)__fadd_rn(__fsub_rn(mx, mn), 1e-6f);
        const double mean_msp = sum_msp / cnt;
        const double var_msp = (sumsq_msp - sum_msp * sum_msp / cnt) / (cnt - 1.0);
        const double mean_raw = sum_ml / cnt;
        const double var_raw = (sumsq_ml - sum_ml * sum_ml / cnt) / (cnt - 1.0);
        const double std_msp = sqrt(fmax(var_msp, 0.0));
        const double mean_ml = (mean_raw - (double)mn) / den;
        const double std_ml = sqrt(fmax(var_raw, 0.0)) / den;
        float* o = scene_out + (int64_t)s * 8;
        o[0] = (float)mean_msp; o[1] = (float)std_msp; o[2] = (float)(mean_msp - (double)beta * std_msp);
        o[3] = (float)mean_ml;  o[4] = (float)std_ml;  o[5] = (float)(mean_ml - (double)beta * std_ml);
        o[6] = mn; o[7] = mx;
    }
}

// per scene (written by the last block of score_fused_kernel): [msp_mean, msp_std, msp_stop, ml_mean, ml_std,
// ml_stop, ml_min, ml_max]  (ml_* of the min-max normalised max logit; std unbiased like torch.std;
// stop = mean - beta*std)

// ml_norm[i] = (max_logit[i] - min_s) / (max_s - min_s + 1e-6), the exact f32 ops of
// pointpdf_v1m1_base.py:213-215
__global__ void score_normalise_kernel(int64_t n, int b, const int* __restrict__ offset,
                                       const float* __restrict__ max_logit, const float* __restrict__ scene_out,
                                       float* __restrict__ ml_norm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = segment_of(i, offset, b);
        const float mn = scene_out[s * 8 + 6], mx = scene_out[s * 8 + 7];
        ml_norm[i] = __fdiv_rn(__fsub_rn(max_logit[i], mn), __fadd_rn(__fsub_rn(mx, mn), 1e-6f));
    }
}

}  // namespace pob

using namespace pob;

// b scene accumulators + one slot holding the finished-blocks counter; zeroed by pob_score_fused itself
POB_API size_t pob_score_workspace_bytes(int b) { return b < 1 ? 0 : sizeof(SceneStats) * (size_t)(b + 1); }

// One pass over logits (n, K) [+ conf (n)].  Every output pointer may be NULL (skipped).
// Scene statistics need offset (b), workspace (pob_score_workspace_bytes(b)), scene_out (b*8 f32);
// ml_norm additionally needs max_logit.  beta: PointPdfV1's stop coefficient.
POB_API int pob_score_fused(int64_t n, int K, int b, const float* logits, const float* conf, const int* offset,
                            float beta, float* msp_score, float* ml_score, float* pdf_score, float* msp_prob,
                            float* max_logit, int* pred, float* ml_norm, float* scene_out, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream) {
    if (n < 0 || K < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!logits) return POB_ERR_BAD_ARG;
    const bool want_stats = scene_out != nullptr;
    if (want_stats && (!offset || b < 1 || !workspace)) return POB_ERR_BAD_ARG;
    if (want_stats && workspace_bytes < pob_score_workspace_bytes(b)) return POB_ERR_WORKSPACE;
    if (ml_norm && (!want_stats || !max_logit)) return POB_ERR_BAD_ARG;
    if (pdf_score && !conf) return POB_ERR_BAD_ARG;
    SceneStats* stats = want_stats ? (SceneStats*)workspace : nullptr;
    if (want_stats) POB_CHECK(cudaMemsetAsync(workspace, 0, pob_score_workspace_bytes(b), stream));
    score_fused_kernel<<<(unsigned)ceil_div(n, SCORE_ROWS), SCORE_ROWS, 0, stream>>>(
        n, K, b, K <= SCORE_MAX_K_SMEM ? 1 : 0, logits, conf, offset, msp_score, ml_score, pdf_score, msp_prob,
        max_logit, pred, stats, beta, scene_out);
    if (want_stats && ml_norm)
        score_normalise_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, b, offset, max_logit, scene_out, ml_norm);
    pob_count_launches(1 + (want_stats && ml_norm ? 1 : 0));
    POB_RETURN_LAST_ERROR();
}
