/*
 * pointops_b200.h -- C ABI of libpointops_b200.so, the B200 (sm_100a) replacement for the
 * native half of the reference package `pointops` (JinfengX/PointCloudPDF, libs/pointops).
 *
 * Boundary.  The reference binds 16 pybind functions `*_cuda(sizes..., at::Tensor...)`
 * (libs/pointops/src/pointops_api.cpp:15-32); each one only unwraps data pointers and calls an
 * `extern "C"` launcher taking raw device pointers and int sizes.  This header declares the
 * drop-in for those launchers: the same argument meaning and order, with
 *   - sizes widened to int64_t (the reference's int products overflow above 2^31 elements),
 *   - a trailing cudaStream_t (the reference launches on the legacy default stream),
 *   - an int return: 0 on success, a cudaError_t from the launch, or a POB_ERR_* code;
 *   - where a kernel needs scratch, a caller-owned workspace with a *_workspace_bytes query.
 * No entry point allocates, frees, synchronises or reads device data on the host, and the library keeps
 * NO tuning state between calls (every schedule / tile / variant choice is an argument of the call): all of
 * them are re-entrant, stream-ordered and CUDA-graph capturable.  The only process-wide data are an atomic
 * launch counter (pob_kernel_launch_count) and one environment switch read once at load time
 * (POINTOPS_B200_NO_PIPE: aggregation forward without the bulk-async ring).  All pointers are device
 * pointers unless stated otherwise.  Tensors are dense row-major, f32 / i32 as in the reference.
 *
 * Batched-by-offset convention (libs/pointops/functions/query.py:9-24): a batch is the
 * concatenation of scenes; `offset[b]` holds cumulative end rows.  Every neighbour search is
 * restricted to the query's own scene.
 */
#ifndef POINTOPS_B200_H
#define POINTOPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define POB_ERR_BAD_ARG 10001      /* null pointer / negative size / unsupported value */
#define POB_ERR_WORKSPACE 10002    /* workspace smaller than *_workspace_bytes() */
#define POB_ERR_UNSUPPORTED 10003

/* ABI version; bumped on any signature change. */
int pob_version(void);
/* Kernels launched by this library in this process so far (monotonic; for bench accounting). */
long long pob_kernel_launch_count(void);
/* Human-readable text for a return code (cudaGetErrorString for CUDA codes). Host pointer. */
const char* pob_error_string(int code);

/* ------------------------------------------------------------------ kNN query ----------
 * Replaces knn_query_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)
 * (src/knn_query/knn_query_cuda_kernel.h:13; kernel knn_query_cuda_kernel.cu:60-104).
 * idx (m, nsample) i32: the nsample nearest rows of the query's scene by the key (d2, idx),
 * ascending; tail filled with -1 / 1e10 when the scene holds fewer points.  d2 is the f32 chain
 * fma(dz,dz, fma(dx,dx, dy*dy)), what the reference binary executes.  dist receives d2
 * (take_sqrt = 0, like the reference kernel) or sqrt(d2) (take_sqrt = 1, what
 * functions/query.py:24 hands to callers); may be NULL.  nsample <= 256.
 * Extra arguments vs the reference: n (rows of xyz) and b (scenes) size the search grid.
 * The workspace (about 70 n bytes + 1 MB) also holds the scratch region pob_farthest_point_sampling uses when it
 * is handed this grid (its curve-ordered copy of the points and the radix-sort buffers).      */
size_t pob_knn_grid_workspace_bytes(int64_t n, int b, float cell_pts);
int pob_knn_grid_build(int64_t n, int b, const float* xyz, const int* offset, float cell_pts,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Query a built grid (reusable for any new_xyz / nsample against the same xyz, offset).
 * weight (m, nsample), optional: fused inverse-distance weights of
 * functions/interpolation.py:15-17, w = r / sum(r), r = 1 / (sqrt(d2) + 1e-8).
 * stats_u64, optional (else NULL): device uint64 the launch adds the number of candidate distances it
 * actually evaluated to (the grid visits a few hundred per query; brute force would visit n_scene).  */
int pob_knn_grid_query(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                       const int* new_offset, float cell_pts, const void* workspace, int* idx, float* dist,
                       float* weight, int take_sqrt, void* stats_u64, cudaStream_t stream);
/* Query a built grid and store every result row into the (n_total, nsample) result buffers of ALL ranks of a
 * multi-GPU job (the large-scene form of SURVEY.md 8e: queries sharded, reference set replicated): peer_idx /
 * peer_dist are HOST arrays of npeers <= 16 device pointers into peer-mapped memory (the caller's own buffer among
 * them; peer_dist may be NULL), row_base = row of this rank's first query.  Replaces "local kernel + NCCL
 * all-gather" by P2P stores issued from the query kernel itself.  The caller synchronises the ranks afterwards.  */
int pob_knn_grid_query_scatter(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                               const int* new_offset, float cell_pts, const void* workspace, int64_t row_base,
                               int npeers, int* const* peer_idx, float* const* peer_dist, int take_sqrt,
                               cudaStream_t stream);
/* Build + query in one call, cell_pts = 2; workspace >= pob_knn_grid_workspace_bytes(n, b, 2). */
int pob_knn_query(int64_t m, int nsample, int64_t n, int b, const float* xyz, const float* new_xyz,
                  const int* offset, const int* new_offset, int* idx, float* dist, int take_sqrt,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Exhaustive scan with the same key and arithmetic (O(m * n_scene)); workspace >= 64 * b bytes. */
int pob_knn_query_bruteforce(int64_t m, int nsample, int b, const float* xyz, const float* new_xyz,
                             const int* offset, const int* new_offset, int* idx, float* dist, int take_sqrt,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---------------------------------------------------------------- radius queries --------
 * (off the PTv1 path; pointops.ball_query / random_ball_query are used by other Pointcept backbones and by
 * PDF's pseudo-label neighbour graph)  Both run on a grid workspace built by
 * pob_knn_grid_build for the same (xyz, offset, n, b, cell_pts); a point is accepted when
 * d2 <= 1e-5 or min_radius^2 <= d2 < max_radius^2 (ball_query_cuda_kernel.cu:99).
 *
 * pob_ball_query replaces ball_query_cuda_launcher(m, nsample, min_radius, max_radius, xyz, new_xyz, offset,
 * new_offset, idx, dist2) (src/ball_query/ball_query_cuda_kernel.h): the accepted points in ascending index
 * order, passed through the reference's heap_sort exactly as the kernel does (no heapify first, so the list
 * is only partially ordered by distance -- reproduced, not fixed); at most nsample of them -> that list,
 * padded with idx -1 / dist2 1e10; more -> every (cnt / nsample)-th entry, and dist2 then holds the
 * candidate INDEX as a float, as ball_query_cuda_kernel.cu:120 writes it.
 * *overflow_flag (device int, caller-zeroed) is set when a query saw more than 2048 candidates -- the size of
 * the reference's per-thread stack arrays, which it overruns; such rows use the first 2048 found.
 *
 * pob_random_ball_query replaces random_ball_query_cuda_launcher(m, nsample, min_radius, max_radius, order,
 * xyz, new_xyz, offset, new_offset, idx, dist2) (src/random_ball_query/random_ball_query_cuda_kernel.h):
 * the first nsample accepted points in the order of the permutation `order` (n global row indices, each
 * scene's rows permuted within the scene); inv_scratch (n ints) is caller-owned scratch; nsample <= 256. */
int pob_ball_query(int64_t m, int nsample, float min_radius, float max_radius, int64_t n, int b, const float* xyz,
                   const float* new_xyz, const int* new_offset, float cell_pts, const void* workspace, int* idx,
                   float* dist2, int* overflow_flag, cudaStream_t stream);
int pob_random_ball_query(int64_t m, int nsample, float min_radius, float max_radius, int64_t n, int b,
                          const int* order, const float* xyz, const float* new_xyz, const int* new_offset,
                          float cell_pts, const void* workspace, int* inv_scratch, int* idx, float* dist2,
                          cudaStream_t stream);

/* ------------------------------------------------------- farthest point sampling -------
 * Replaces farthest_point_sampling_cuda_launcher(b, n, xyz, offset, new_offset, tmp, idx)
 * (src/sampling/sampling_cuda_kernel.h:13; kernel sampling_cuda_kernel.cu:15-129).
 * n_max = size of the largest scene (the reference's `n`); an over-estimate is allowed, an
 * under-estimate is not.  idx (new_offset[b-1]) i32 receives GLOBAL row indices, scene-major;
 * first sample of a scene is its first row; ties go to the lowest index.  tmp (n f32) needs no
 * initialisation and is only used (as scratch of the grid-wide / streamed forms) when a scene exceeds what one
 * 16-CTA cluster holds (196608 points with the merged-list kernel, 131072 with variants 2 / 3); it may be NULL
 * below 131072.  cluster_hint: 0 = auto, or 1/2/4/8/16 CTAs per
 * scene.  A scene requesting 0 samples writes nothing (reference quirk C5 not reproduced).
 * grid_workspace (optional, else NULL): the workspace pob_knn_grid_build filled for the same
 * xyz / offset, with its n (rows of xyz) and cell_pts; the kernel then walks the points in cell
 * order and skips, exactly, every warp whose bounding box is out of the new sample's reach.
 * The result is bit-identical with or without it.  The MERGE variant re-orders the points along a
 * Hilbert curve first (~20 us; compact rows and warps prune better: -17..19 % kernel time) and keeps that
 * copy in the workspace's own FPS region -- the workspace is written, the kNN arrays in it are not.   */
/* variant: which schedule of the same algorithm runs (the sampled indices never depend on it):
 *   POB_FPS_AUTO   (0) the library chooses (= MERGE)
 *   POB_FPS_MERGE  (1) merged-list kernel: every warp offers a list of its next local candidates, one cluster
 *                      exchange accepts the exact global prefix (~20 samples per exchange on room-shaped clouds)
 *   POB_FPS_CHAIN  (2) round-1 kernel: one candidate + bound per CTA and exchange (~4.5 samples per exchange)
 *   POB_FPS_SINGLE (3) one sample per exchange
 *   POB_FPS_MERGE_CELLS (4) MERGE without the Hilbert re-ordering (points in the grid's cell order; A/B)
 * stats_u64x4: NULL, or a device pointer to 4 x uint64 {rounds, samples, point distances evaluated, reserved}
 * the launch accumulates into (mean samples accepted per cluster-wide exchange = samples / rounds; the third
 * counter is filled by the MERGE variant only).  Per-call arguments: the library keeps no tuning state
 * between calls.                                                                                          */
#define POB_FPS_AUTO 0
#define POB_FPS_MERGE 1
#define POB_FPS_CHAIN 2
#define POB_FPS_SINGLE 3
#define POB_FPS_MERGE_CELLS 4
int pob_farthest_point_sampling(int b, int64_t n_max, const float* xyz, const int* offset,
                                const int* new_offset, float* tmp, int* idx, int cluster_hint,
                                void* grid_workspace, int64_t n, float cell_pts, int variant,
                                void* stats_u64x4, cudaStream_t stream);

/* --------------------------------------------------------- grouping (pointops.grouping2) --
 * grouping_{forward,backward}_cuda_launcher (src/grouping/grouping_cuda_kernel.h:14-15).
 * forward: output[m,s,:] = input[idx[m,s],:].  backward: grad_input[idx[m,s],:] += grad_output
 * (grad_input must be zeroed by the caller, as functions/grouping.py:31 does).  idx >= 0.       */
int pob_grouping_forward(int64_t m, int nsample, int c, const float* input, const int* idx, float* output,
                         cudaStream_t stream);
int pob_grouping_backward(int64_t m, int nsample, int c, const float* grad_output, const int* idx,
                          float* grad_input, cudaStream_t stream);

/* ------------------------------------------------------------------------ subtraction ----
 * subtraction_{forward,backward}_cuda_launcher (src/subtraction/subtraction_cuda_kernel.h:14-15).
 * forward: output[n,s,:] = input1[n,:] - input2[idx[n,s],:].
 * backward: grad_input1[n,:] = sum_s grad_output[n,s,:] (overwritten, no zeroing needed);
 *           grad_input2[idx[n,s],:] -= grad_output[n,s,:] (accumulated; caller zeroes).         */
int pob_subtraction_forward(int64_t n, int nsample, int c, const float* input1, const float* input2,
                            const int* idx, float* output, cudaStream_t stream);
int pob_subtraction_backward(int64_t n, int nsample, int c, const int* idx, const float* grad_output,
                             float* grad_input1, float* grad_input2, cudaStream_t stream);

/* ------------------------------------------------- vector-attention aggregation ----------
 * aggregation_{forward,backward}_cuda_launcher (src/aggregation/aggregation_cuda_kernel.h:14-15).
 * forward: output[n,c] = sum_s (input[idx[n,s],c] + position[n,s,c]) * weight[n,s,c % w_c]
 *          (output overwritten; the reference accumulates into a zeroed buffer).  idx < 0 (a kNN
 *          placeholder) contributes a zero input row, matching pointops.grouping; the reference
 *          kernel reads out of bounds there.
 * backward: grad_input[idx[n,s],c] += g*w (accumulated; caller zeroes); grad_position = g*w and
 *           grad_weight[n,s,j] = sum_{c % w_c == j} g * (input + position) (both overwritten).  */
int pob_aggregation_forward(int64_t n, int nsample, int c, int w_c, const float* input, const float* position,
                            const float* weight, const int* idx, float* output, cudaStream_t stream);
int pob_aggregation_backward(int64_t n, int nsample, int c, int w_c, const float* input, const float* position,
                             const float* weight, const int* idx, const float* grad_output, float* grad_input,
                             float* grad_position, float* grad_weight, cudaStream_t stream);

/* ------------------------------------------------------- three-NN interpolation ----------
 * interpolation_{forward,backward}_cuda_launcher (src/interpolation/interpolation_cuda_kernel.h:14-15).
 * forward: output[n,c] = sum_{i<k} input[idx[n,i],c] * weight[n,i] (overwritten).
 * backward: grad_input[idx[n,i],c] += grad_output[n,c] * weight[n,i] (accumulated).             */
int pob_interpolation_forward(int64_t n, int c, int k, const float* input, const int* idx, const float* weight,
                              float* output, cudaStream_t stream);
int pob_interpolation_backward(int64_t n, int c, int k, const float* grad_output, const int* idx,
                               const float* weight, float* grad_input, cudaStream_t stream);

/* ------------------------------------- fused grouping-with-xyz (additive entry point) ------
 * One kernel for pointops.grouping (functions/grouping.py:36-60, pure torch there: 2 cats, 2
 * gathers, mask, einsum, cat).  output (m, nsample, 3 + c) f32 when with_xyz else (m, nsample, c):
 *   [ (xyz[idx] - new_xyz[m]) , feat[idx] ]  with all-zero rows where idx < 0.
 * feat_dtype: 0 = f32, 1 = f16, 2 = bf16 (autocast); output is always f32 (torch promotion).
 * backward: grad_feat[idx[m,s],:] += grad_output[m,s,(3 if with_xyz):] for idx >= 0 (f32,
 * accumulated; caller zeroes).                                                                  */
int pob_group_xyz_forward(int64_t m, int nsample, int c, int with_xyz, const void* feat, int feat_dtype,
                          const float* xyz, const float* new_xyz, const int* idx, float* output,
                          cudaStream_t stream);
int pob_group_xyz_backward(int64_t m, int nsample, int c, int with_xyz, const float* grad_output, const int* idx,
                           float* grad_feat, cudaStream_t stream);
/* The coordinate half alone: output (m, nsample, 3) = xyz[idx] - new_xyz[m], zeros for idx < 0
 * (functions/grouping.py:49-57); no gradient (the reference gives xyz none).                    */
int pob_group_relxyz_forward(int64_t m, int nsample, const float* xyz, const float* new_xyz, const int* idx,
                             float* output, cudaStream_t stream);

/* --------------------- inference form of PointTransformerLayer (additive entry points) ------
 * The caller of the operators above, pointcept/models/point_transformer/point_transformer_seg.py:48-81,
 * spends its time between the q/k/v linears and the layer output in ~16 eager kernels per block
 * (two kNN-group gathers, linear_p, the relation k_j - q_i + p_r, linear_w, softmax over the
 * neighbours, the einsum aggregation).  With every BatchNorm in eval mode that chain is one kernel:
 *   h   = relu(A (xyz[j] - xyz[i]) + c)            Linear(3,3) + BN folded   (j = idx[i,s]; 0 if j < 0)
 *   p_r = Wp h + bp                                 Linear(3, C)
 *   t   = relu(aw * (k[j] - q[i] + p_r) + bw)       first BN of linear_w as an affine map
 *   u   = relu(W1 t + b1)                           Linear(C, C/8) + BN folded
 *   l   = W2 u + b2 ;  w = softmax over s of l      Linear(C/8, C/8)
 *   out[i,c] = sum_s (v[j,c] + p_r[c]) * w[s, c % (C/8)]   ( then relu(oa * out + ob) if out_affine: bn2 )
 * params is ONE packed f32 block, 16-byte aligned, pob_pt_layer_param_floats(c, w_c) floats:
 *   A[9] c[3] pad[4] | wx[C] wy[C] wz[C] bp[C] aw[C] bw[C] oa[C] ob[C] | W1[w_c][C] | b1[w_c] |
 *   W2^T[w_c][w_c] (W2T[o][o'] = W2[o'][o]) | b2[w_c]
 * q, k, v (n, c) with row strides ldq / ldk / ldv floats (column blocks of one (n, 3c) GEMM output are
 * fine); out (n, c) with row stride ldo.  Supported: c in {32,64,128,256,512}, w_c = c/8, nsample in {8,16};
 * anything else returns POB_ERR_UNSUPPORTED (callers then run the unfused operator sequence).      */
int64_t pob_pt_layer_param_floats(int c, int w_c);
/* Tuning / test hook.  0 (default): the CTA-tiled kernel (needs 16-byte aligned rows, else falls back);
 * 1, 4, 8, 16: the warp-per-point kernel with that many warps sharing one point (clamped to what the
 * shape instantiates); -1: warp-per-point with the split chosen from n.  Only the f32 summation order
 * depends on it.                                                                                   */
int pob_pt_layer_forward(int64_t n, int nsample, int c, int w_c, const float* q, int64_t ldq, const float* k,
                         int64_t ldk, const float* v, int64_t ldv, const float* xyz, const int* idx,
                         const float* params, int out_affine, float* out, int64_t ldo, int split, cudaStream_t stream);
/* out = relu?(x * scale + shift + residual) over (rows, c); scale, shift (c) and residual (rows, c) may be
 * NULL; in place allowed.  The BN / skip / ReLU tail of Bottleneck (point_transformer_seg.py:188-195).  */
int pob_affine_act(int64_t rows, int c, const float* x, const float* scale, const float* shift,
                   const float* residual, int relu, float* out, cudaStream_t stream);
/* TransitionDown (point_transformer_seg.py:106-119) with its Linear(3 + C, C') split by linearity:
 * z (n, c) = feat @ W[:, 3:]^T computed by the caller on the UNGATHERED points (cuBLAS), wxyz (c, 3) =
 * W[:, :3]; out (m, c) = max_s relu(scale * (z[idx[m,s]] + wxyz (xyz[idx[m,s]] - new_xyz[m])) + shift)
 * = gather + coordinate columns + BatchNorm(eval) + ReLU + MaxPool1d(nsample) in one pass; idx < 0
 * contributes a zero grouped row (pointops.grouping's mask).  c % 4 == 0.                           */
int pob_transition_down_pool(int64_t m, int nsample, int c, const float* z, const float* xyz, const float* new_xyz,
                             const int* idx, const float* wxyz, const float* scale, const float* shift,
                             float* out, cudaStream_t stream);
/* output = base + interpolation_forward(input, idx, weight): TransitionUp's skip connection folded in
 * (point_transformer_seg.py:168-170); base may be NULL.                                             */
int pob_interpolation_add_forward(int64_t n, int c, int k, const float* input, const int* idx, const float* weight,
                                  const float* base, float* output, cudaStream_t stream);

/* FP32 linear layer with fused epilogue (csrc/linear.cu): the eval-mode "Linear + BatchNorm + ReLU",
 * q/k/v and "Linear + BatchNorm + skip + ReLU" GEMMs of Bottleneck / TransitionDown / TransitionUp
 * (point_transformer_seg.py:87-95,128-147,178-195), which the reference runs as cuBLAS GEMM + eager
 * BN / ReLU / add kernels.      out (M, N) = act(A (M, K) @ Wt (K, N) + bias (N) + residual (M, N))
 * bias / residual may be NULL; relu != 0 applies max(., 0).  lda / ldr / ldo are row strides in floats;
 * Wt is the weight stored dense K x N (the caller transposes the constant once).  FP32 FFMA, operands are
 * not rounded to TF32.  Any K, N >= 1 (scalar loads / stores when K, N, strides or pointers are not
 * 16-byte friendly).  out must not alias A or residual (both are read through the read-only path).
 * config (per call; the library keeps no tuning state): test / tuning hook, 0 (default) picks the CTA tile from the shape, 1..16 force one of
 * the instantiated tiles (BM x BN from 16x32 to 256x32 / 128x64, register tiles 4x4, 8x4 or 8x8, intra-CTA
 * split-K 1..8; csrc/linear.cu lists them).  Only the f32 summation order depends on it.                   */
int pob_linear_forward(int64_t M, int K, int N, const float* A, int64_t lda, const float* Wt, const float* bias,
                       const float* residual, int64_t ldr, int relu, float* out, int64_t ldo, int config,
                       cudaStream_t stream);

/* ------------------------------------------------------- grouped vector attention steps --
 * (off the PTv1 path: Point Transformer v2's operators)  Replace the four launchers of
 * src/attention/attention_cuda_kernel.h with the same argument order:
 *   relation   output[r, g]          = sum_c query[it[r], g, c] * key[ir[r], g, c] * weight[c]     (written)
 *   fusion     output[it[r], g, c]  += weight[r, g] * value[ir[r], g, c]                           (accumulated)
 * query / key / value (n, g, c), weight (c) resp. (m, g), index_target / index_refer (m) int32.
 * Backward: grad_query, grad_key, grad_value and relation's grad_weight (c) are accumulated into (the caller
 * zeroes them, as the reference's wrappers do); fusion's grad_weight (m, g) is written.                 */
int pob_attention_relation_step_forward(int64_t m, int g, int c, const float* query, const float* key,
                                        const float* weight, const int* index_target, const int* index_refer,
                                        float* output, cudaStream_t stream);
int pob_attention_relation_step_backward(int64_t m, int g, int c, const float* query, float* grad_query,
                                         const float* key, float* grad_key, const float* weight, float* grad_weight,
                                         const int* index_target, const int* index_refer, const float* grad_output,
                                         cudaStream_t stream);
int pob_attention_fusion_step_forward(int64_t m, int g, int c, const float* weight, const float* value,
                                      const int* index_target, const int* index_refer, float* output,
                                      cudaStream_t stream);
int pob_attention_fusion_step_backward(int64_t m, int g, int c, const float* weight, float* grad_weight,
                                       const float* value, float* grad_value, const int* index_target,
                                       const int* index_refer, const float* grad_output, cudaStream_t stream);

/* ------------------------------------------ fused open-set scoring (additive entry point) --
 * One pass over logits (n, K) [+ conf (n)] replacing
 *   MaxProbability msp / ml  (pointcept/recognizers/max_probability/max_probability_v1m1_base.py:17-29),
 *   PointPdfV1 score         (pointcept/recognizers/ours/pointpdf_v1m1_base.py:106-113),
 *   the scoring prefix of PointPdfV1.pseudo_labeling (pointpdf_v1m1_base.py:199-222).
 * Per point (each pointer optional): msp_score = -max log_softmax, ml_score = -max logit,
 * pdf_score = softmax(cat[logits, conf])[K] (needs conf), msp_prob = max softmax, max_logit,
 * pred = argmax (first maximum), ml_norm = (max_logit - min_s) / (max_s - min_s + 1e-6).
 * Per scene, scene_out (b, 8) f32: msp mean, msp std (unbiased), msp stop = mean - beta*std,
 * then the same three for ml_norm, then min_s, max_s of the max logit.  Scene outputs need
 * offset, b and a workspace of pob_score_workspace_bytes(b).                                    */
size_t pob_score_workspace_bytes(int b);
int pob_score_fused(int64_t n, int K, int b, const float* logits, const float* conf, const int* offset, float beta,
                    float* msp_score, float* ml_score, float* pdf_score, float* msp_prob, float* max_logit,
                    int* pred, float* ml_norm, float* scene_out, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream);

/* ----------------------------------------------------------- data path ----
 * pob_grid_hash: GridSample's voxel coordinates and FNV64 key (pointcept/datasets/transform.py:813-823, 911-925):
 * grid_coord (n, 3) i32 = floor(coord / grid_size) - min_cell (float64 division, as numpy does it), may be NULL;
 * key (n) i64 = fnv_hash_vec(grid_coord) with the sign bit flipped, so that a signed sort orders it like numpy's
 * unsigned argsort.  min_cell: device int64[3].
 * pob_scatter_mean: torch_scatter.scatter_mean(src, index, dim=0, dim_size) as the tester averages fragment scores
 * (pointcept/engines/test.py:243-248): out (dim_size, c) and count (dim_size) zeroed by the caller; rows whose index
 * is outside [0, dim_size) are ignored.                                                                      */
int pob_grid_hash(int64_t n, const float* coord, double gx, double gy, double gz, const long long* min_cell,
                  int* grid_coord, long long* key, cudaStream_t stream);
int pob_scatter_mean(int64_t rows, int c, const float* src, const long long* index, int64_t dim_size, float* out,
                     float* count, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTOPS_B200_H */
