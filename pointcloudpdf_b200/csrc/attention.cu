// attention.cu -- pointops.attention_relation_step / attention_fusion_step (sm_100a).
//
// Replaces libs/pointops/src/attention/attention_cuda_kernel.cu (one thread and one atomicAdd per
// (pair, group, channel) element).  These two operators serve Point Transformer v2's grouped vector
// attention, not the PTv1 path; they are complete and tested but not tuned further:
//   relation   out[r, g]          = sum_c query[it[r], g, c] * key[ir[r], g, c] * weight[c]
//   fusion     out[it[r], g, c]  += weight[r, g] * value[ir[r], g, c]
// A warp owns one (pair, group): lanes stride the channels (coalesced rows), reductions over c are
// shuffles, so the forward of `relation` and grad_weight of `fusion` need no atomics at all; true
// scatters (rows addressed through it / ir) stay atomic.
#include "common.cuh"

namespace pob {

constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(ATT_THREADS)
att_relation_fwd_kernel(int64_t m, int g, int c, const float* __restrict__ query, const float* __restrict__ key,
                        const float* __restrict__ weight, const int* __restrict__ it, const int* __restrict__ ir,
                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t total = m * g, nwarps = (int64_t)gridDim.x * (ATT_THREADS / 32);
    for (int64_t w = (int64_t)blockIdx.x * (ATT_THREADS / 32) + (threadIdx.x >> 5); w < total; w += nwarps) {
        const int64_t r = w / g;
        const int gi = (int)(w - r * g);
        const float* qrow = query + ((int64_t)__ldg(it + r) * g + gi) * c;
        const float* krow = key + ((int64_t)__ldg(ir + r) * g + gi) * c;
        float acc = 0.f;
        for (int cc = lane; cc < c; cc += 32) acc = fmaf(__ldg(qrow + cc) * __ldg(krow + cc), __ldg(weight + cc), acc);
        acc = warp_sum(acc);
        if (lane == 0) out[w] = acc;
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
att_relation_bwd_kernel(int64_t m, int g, int c, const float* __restrict__ query, float* __restrict__ grad_query,
                        const float* __restrict__ key, float* __restrict__ grad_key, const float* __restrict__ weight,
                        float* __restrict__ grad_weight, const int* __restrict__ it, const int* __restrict__ ir,
                        const float* __restrict__ grad_out) {
    extern __shared__ float gw_s[];   // [c] per-CTA partial of grad_weight
    for (int cc = threadIdx.x; cc < c; cc += ATT_THREADS) gw_s[cc] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t total = m * g, nwarps = (int64_t)gridDim.x * (ATT_THREADS / 32);
    for (int64_t w = (int64_t)blockIdx.x * (ATT_THREADS / 32) + (threadIdx.x >> 5); w < total; w += nwarps) {
        const int64_t r = w / g;
        const int gi = (int)(w - r * g);
        const int64_t qo = ((int64_t)__ldg(it + r) * g + gi) * c, ko = ((int64_t)__ldg(ir + r) * g + gi) * c;
        const float go = __ldg(grad_out + w);
        for (int cc = lane; cc < c; cc += 32) {
            const float qv = __ldg(query + qo + cc), kv = __ldg(key + ko + cc), wv = __ldg(weight + cc);
            atomicAdd(grad_query + qo + cc, go * kv * wv);
            atomicAdd(grad_key + ko + cc, go * qv * wv);
            atomicAdd(gw_s + cc, go * kv * qv);
        }
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < c; cc += ATT_THREADS) atomicAdd(grad_weight + cc, gw_s[cc]);
}

__global__ void __launch_bounds__(ATT_THREADS)
att_fusion_fwd_kernel(int64_t m, int g, int c, const float* __restrict__ weight, const float* __restrict__ value,
                      const int* __restrict__ it, const int* __restrict__ ir, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t total = m * g, nwarps = (int64_t)gridDim.x * (ATT_THREADS / 32);
    for (int64_t w = (int64_t)blockIdx.x * (ATT_THREADS / 32) + (threadIdx.x >> 5); w < total; w += nwarps) {
        const int64_t r = w / g;
        const int gi = (int)(w - r * g);
        const int64_t oo = ((int64_t)__ldg(it + r) * g + gi) * c, vo = ((int64_t)__ldg(ir + r) * g + gi) * c;
        const float wv = __ldg(weight + w);
        for (int cc = lane; cc < c; cc += 32) atomicAdd(out + oo + cc, wv * __ldg(value + vo + cc));
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
att_fusion_bwd_kernel(int64_t m, int g, int c, const float* __restrict__ weight, float* __restrict__ grad_weight,
                      const float* __restrict__ value, float* __restrict__ grad_value, const int* __restrict__ it,
                      const int* __restrict__ ir, const float* __restrict__ grad_out) {
    const int lane = threadIdx.x & 31;
    const int64_t total = m * g, nwarps = (int64_t)gridDim.x * (ATT_THREADS / 32);
    for (int64_t w = (int64_t)blockIdx.x * (ATT_THREADS / 32) + (threadIdx.x >> 5); w < total; w += nwarps) {
        const int64_t r = w / g;
        const int gi = (int)(w - r * g);
        const int64_t oo = ((int64_t)__ldg(it + r) * g + gi) * c, vo = ((int64_t)__ldg(ir + r) * g + gi) * c;
        const float wv = __ldg(weight + w);
        float acc = 0.f;
        for (int cc = lane; cc < c; cc += 32) {
            const float go = __ldg(grad_out + oo + cc);
            acc = fmaf(go, __ldg(value + vo + cc), acc);
            atomicAdd(grad_value + vo + cc, go * wv);
        }
        acc = warp_sum(acc);
        if (lane == 0) grad_weight[w] = acc;   // one (pair, group) per warp: plain store, no atomics
    }
}

static inline unsigned att_grid(int64_t m, int g) { return grid_for(m * g * 32, ATT_THREADS, 8); }

}  // namespace pob

using namespace pob;

// attention_relation_step_forward_cuda_launcher(m, g, c, query, key, weight, index_target, index_refer, output)
// (src/attention/attention_cuda_kernel.h).  output (m, g) is fully written (no zero-init needed).
POB_API int pob_attention_relation_step_forward(int64_t m, int g, int c, const float* query, const float* key,
                                                const float* weight, const int* index_target, const int* index_refer,
                                                float* output, cudaStream_t stream) {
    if (m < 0 || g < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!query || !key || !weight || !index_target || !index_refer || !output) return POB_ERR_BAD_ARG;
    att_relation_fwd_kernel<<<att_grid(m, g), ATT_THREADS, 0, stream>>>(m, g, c, query, key, weight, index_target,
                                                                        index_refer, output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// ..._backward_cuda_launcher: grad_query, grad_key (n, g, c) and grad_weight (c) are ACCUMULATED into (caller zeroes).
POB_API int pob_attention_relation_step_backward(int64_t m, int g, int c, const float* query, float* grad_query,
                                                 const float* key, float* grad_key, const float* weight,
                                                 float* grad_weight, const int* index_target, const int* index_refer,
                                                 const float* grad_output, cudaStream_t stream) {
    if (m < 0 || g < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!query || !grad_query || !key || !grad_key || !weight || !grad_weight || !index_target || !index_refer || !grad_output)
        return POB_ERR_BAD_ARG;
    if ((size_t)c * sizeof(float) > 48 * 1024) return POB_ERR_UNSUPPORTED;
    att_relation_bwd_kernel<<<att_grid(m, g), ATT_THREADS, sizeof(float) * c, stream>>>(
        m, g, c, query, grad_query, key, grad_key, weight, grad_weight, index_target, index_refer, grad_output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// attention_fusion_step_forward_cuda_launcher: output (n, g, c) is ACCUMULATED into (caller zeroes).
POB_API int pob_attention_fusion_step_forward(int64_t m, int g, int c, const float* weight, const float* value,
                                              const int* index_target, const int* index_refer, float* output,
                                              cudaStream_t stream) {
    if (m < 0 || g < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!weight || !value || !index_target || !index_refer || !output) return POB_ERR_BAD_ARG;
    att_fusion_fwd_kernel<<<att_grid(m, g), ATT_THREADS, 0, stream>>>(m, g, c, weight, value, index_target, index_refer,
                                                                      output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// ..._backward_cuda_launcher: grad_weight (m, g) is written; grad_value (n, g, c) is ACCUMULATED into.
POB_API int pob_attention_fusion_step_backward(int64_t m, int g, int c, const float* weight, float* grad_weight,
                                               const float* value, float* grad_value, const int* index_target,
                                               const int* index_refer, const float* grad_output, cudaStream_t stream) {
    if (m < 0 || g < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!weight || !grad_weight || !value || !grad_value || !index_target || !index_refer || !grad_output)
        return POB_ERR_BAD_ARG;
    att_fusion_bwd_kernel<<<att_grid(m, g), ATT_THREADS, 0, stream>>>(m, g, c, weight, grad_weight, value, grad_value,
                                                                      index_target, index_refer, grad_output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}
