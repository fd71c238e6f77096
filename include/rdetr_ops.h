/*
 * rdetr_ops.h -- C ABI of librdetr_ops.so, the sm_100a implementation of the Relation-DETR hot path.
 *
 * These entry points are what the reference's native binding for this path would bind instead of
 * its pybind11 module `MultiScaleDeformableAttention`
 *   (upstream models/bricks/ops/cuda/ms_deform_attn_cuda.cu:148-151:
 *    ms_deform_attn_forward / ms_deform_attn_backward),
 * plus a fused replacement for the eager PositionRelationEmbedding.forward
 *   (upstream models/bricks/relation_transformer.py:520-532), which has no native code upstream.
 *
 * Conventions (all functions):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the device that owns
 *     `value` / `src_boxes` unless the name ends in `_host`;
 *   - the caller owns every buffer; the library never allocates device memory, never synchronises
 *     the device and only enqueues work on the stream it is given (cudaStream_t passed as void*);
 *   - return value 0 = success, otherwise one of RDETR_ERR_*; rdetr_last_error() returns a
 *     thread-local, human-readable description of the last failure on the calling thread;
 *   - re-entrant: no global mutable state besides that thread-local message, so forward may run on
 *     the Python thread while backward runs on the autograd thread;
 *   - no fallback: an unsupported shape or dtype returns RDETR_ERR_UNSUPPORTED.
 */
#ifndef RDETR_OPS_H_
#define RDETR_OPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDETR_ABI_VERSION 1

#define RDETR_OK 0
#define RDETR_ERR_INVALID_ARGUMENT 1 /* null pointer, non-positive size, misaligned buffer */
#define RDETR_ERR_UNSUPPORTED 2      /* shape / dtype outside what the kernels are built for */
#define RDETR_ERR_CUDA 3             /* a CUDA runtime call or the launch itself failed */
#define RDETR_ERR_WORKSPACE 4        /* workspace missing or smaller than *_workspace_bytes() */

/* value_dtype: element type of value / out / grad_out / grad_value.  sampling locations,
 * attention weights and their gradients are always fp32. */
#define RDETR_DTYPE_F32 0
#define RDETR_DTYPE_BF16 1

/* flags for the relation kernels */
#define RDETR_REL_EXACT 0 /* evaluation order of the reference: (e*scale)/dim_t, sinf/cosf, true division */
#define RDETR_REL_FAST 1  /* restructured arithmetic, looser documented bound */

typedef void *rdetr_stream_t; /* a cudaStream_t */

int rdetr_abi_version(void);
const char *rdetr_last_error(void);

/*
 * Multi-scale deformable attention, forward.
 * Replaces ms_deform_attn_cuda_forward (ms_deform_attn_cuda.cu:12-72) and the kernel it launches
 * (ms_deform_im2col_cuda.cuh:226-288).  `im2col_step` of the reference has no equivalent: the whole
 * batch is one launch.
 *
 *   value               [B, S, M, D]        value_dtype, contiguous, 16-byte aligned
 *   spatial_shapes      [L, 2] int64 (h, w) device memory, read inside the kernel
 *   level_start_index   [L]    int64        device memory, read inside the kernel
 *   sampling_locations  [B, Nq, M, L, P, 2] fp32, last dim (x, y) normalised to [0, 1]
 *   attention_weights   [B, Nq, M, L, P]    fp32
 *   out                 [B, Nq, M*D]        value_dtype; every element is written
 * Supported: D == 32, 1 <= L <= 8, 1 <= P <= 8, S*M*D < 2^31.
 */
int rdetr_msda_forward(const void *value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const float *sampling_locations,
                       const float *attention_weights, void *out, int B, int S, int M, int D, int L,
                       int Nq, int P, int value_dtype, rdetr_stream_t stream);

/*
 * Multi-scale deformable attention, backward.
 * Replaces ms_deform_attn_cuda_backward (ms_deform_attn_cuda.cu:75-145) and
 * ms_deformable_col2im_gpu_kernel_* (ms_deform_im2col_cuda.cuh:290-909).
 *
 *   grad_out   [B, Nq, M*D]         value_dtype
 *   grad_value [B, S, M, D]         value_dtype; ZEROED INSIDE this call, then accumulated
 *   grad_loc   [B, Nq, M, L, P, 2]  fp32; every element is written (0 for samples outside the map)
 *   grad_attn  [B, Nq, M, L, P]     fp32; every element is written
 *   workspace  rdetr_msda_backward_workspace_bytes() bytes of device scratch (may be NULL when that
 *              returns 0); contents are undefined on return.
 */
size_t rdetr_msda_backward_workspace_bytes(int B, int S, int M, int D, int L, int Nq, int P,
                                           int value_dtype);
int rdetr_msda_backward(const void *value, const int64_t *spatial_shapes,
                        const int64_t *level_start_index, const float *sampling_locations,
                        const float *attention_weights, const void *grad_out, void *grad_value,
                        float *grad_loc, float *grad_attn, int B, int S, int M, int D, int L, int Nq,
                        int P, int value_dtype, void *workspace, size_t workspace_bytes,
                        rdetr_stream_t stream);

/*
 * Process-wide tuning knobs of the MSDA kernels (not per-call state; safe to call from any thread).
 * When Nq == S (encoder self-attention: every query is a pixel of the pyramid) and P == 4, L in {4, 5}, the
 * library uses tiled kernels (one CTA per (image, head, 8x8 query tile), per-level windows in shared memory;
 * DESIGN.md section 4).  mode 0 = automatic (default; also RDETR_MSDA_TILE in the environment), 1 = flat kernels
 * only, 2 = tiled kernels whenever the shape allows.  rows = shared-memory rows (128 bytes each) a CTA may give
 * to level windows; 0 restores the default.  Results do not depend on either knob beyond fp32 summation order.
 */
int rdetr_msda_set_tile_mode(int mode);
int rdetr_msda_set_tile_rows(int rows);
/*
 * Backward only, experimental.  mode 2 (also RDETR_MSDA_COARSE=2 in the environment): pyramid levels of at most
 * 1 056 pixels (25x42 and 13x21 of the 800x1333 pyramid, 19x32 of the 1200x2000 one) have their grad_value
 * accumulated in shared memory by a second kernel on a library-owned high-priority side stream, next to the scatter
 * kernel, instead of one L2 vector reduction per bilinear corner (csrc/msda_bwd_coarse.cu).  Which levels qualify is
 * decided on the device from spatial_shapes.  The side stream forks from and joins the caller's stream with events
 * inside the call (also under stream capture), so the call keeps its stream semantics.  Parity-tested
 * (tests/test_msda_coarse_gpu.py) but SLOWER on B200, hence off by default: mode 0
 * (default) and 1 = never, 2 = whenever P divides 8.  Results do not depend on the knob beyond fp32 summation order.
 */
int rdetr_msda_set_coarse_mode(int mode);
/*
 * Backward with bf16 value only.  A pyramid level whose grad_value rows receive at most `max_updates_per_row` updates
 * on average (Nq * P * 4 bilinear corners over H * W rows: 21 for the finest level of both shipped pyramids when
 * Nq = S, 85 for the next; a quarter of the limit applies when Nq != S, where queries cluster on objects and the
 * average under-states the busiest rows; levels of fewer than 1 024 pixels never qualify) is scattered straight into the bf16 grad_value with packed bf16x2 vector reductions
 * (64 bytes per row: 81 G rows/s at the L2 against 48 G rows/s for 128-byte fp32 rows, profiles/r02al_microbench_red.txt)
 * instead of into the fp32 workspace; the other levels keep fp32 accumulation (a coarse row receives ~1e3 updates, which
 * bf16 accumulation would swallow).  Every addition then rounds to bf16: the directly scattered rows carry about
 * sqrt(updates) * 2^-9 relative error instead of the final rounding's 2^-9 (tests/test_msda_bf16_scatter_gpu.py bounds it
 * level by level).  Default 100: the two finest levels of an encoder call, bf16 backward 1.75 -> 1.47 ms at configs[1]
 * (profiles/r02an_exp_bf16_scatter.txt); 0 = never (also RDETR_MSDA_BF16_SCATTER in the environment).  The criterion is
 * an average: an input that sends most of its samples to a few rows of a fine level (nothing a detector produces, but
 * possible) loses accuracy on those rows -- use 0 for such inputs.  Process-wide, safe to call from any thread.  fp32
 * calls are unaffected.
 */
int rdetr_msda_set_bf16_scatter(int max_updates_per_row);

/*
 * Multi-scale deformable attention with the module prologue folded in ("N2"):
 * softmax over the L*P logits of every (b, q, head), sampling location = reference point + offset, and
 * the key-padding mask, which the reference computes with separate elementwise kernels before the op
 * (models/bricks/ms_deform_attn.py:318-349).  Evaluation order follows torch:
 *   attn = exp(z - max z) / sum;  ref_dim 2: loc = ref + off / (W_l, H_l);
 *   ref_dim 4: loc = ref_xy + ((off / P) * ref_wh) * 0.5
 *
 *   reference_points  [B, Nq, L, ref_dim] fp32 (ref_dim 2 or 4; no gradient is produced for it: every
 *                     call site of the reference passes detached / constant reference points)
 *   sampling_offsets  [B, Nq, M, L, P, 2]   dtype     attention_logits [B, Nq, M, L*P] dtype (pre-softmax)
 *   key_padding_mask  NULL or [B, S] bytes (torch.bool), non-zero = padded pixel, read as value 0
 *   dtype             element type of value, out, offsets, logits and of all gradients
 *   grad_offsets / grad_logits: every element is written; grad_value as in rdetr_msda_backward
 *   workspace         rdetr_msda_backward_workspace_bytes(..., dtype) bytes
 */
int rdetr_msda_fused_forward(const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *reference_points,
                             const void *sampling_offsets, const void *attention_logits,
                             const uint8_t *key_padding_mask, void *out, int B, int S, int M, int D,
                             int L, int Nq, int P, int ref_dim, int dtype, rdetr_stream_t stream);
int rdetr_msda_fused_backward(const void *value, const int64_t *spatial_shapes,
                              const int64_t *level_start_index, const float *reference_points,
                              const void *sampling_offsets, const void *attention_logits,
                              const uint8_t *key_padding_mask, const void *grad_out, void *grad_value,
                              void *grad_offsets, void *grad_logits, int B, int S, int M, int D, int L,
                              int Nq, int P, int ref_dim, int dtype, void *workspace,
                              size_t workspace_bytes, rdetr_stream_t stream);

/*
 * Fused position-relation embedding, forward:
 *   out[b,h,i,j] = relu(bias[h] + sum_n weight[h,n] * f_n(src[b,i], tgt[b,j]))
 * with f = sin/cos encoding of the 4 log-ratio box features
 * (relation_transformer.py:481-490, 520-532; position_encoding.py:131-138).  The [B,N1,N2,64]
 * intermediate is never materialised.
 *
 *   src_boxes  [B, N1, 4] fp32 cxcywh     tgt_boxes [B, N2, 4] fp32 (may alias src_boxes)
 *   weight     [H, 64]    fp32 (= pos_proj.0.weight viewed 2-D)     bias [H] fp32
 *   dim_t      [8]        fp32 DEVICE pointer, temperature ** (2k/16) as torch computed it
 *   attn_mask  NULL or [N1, N2] bytes (torch.bool); non-zero => out = -inf (the decoder's
 *              masked_fill_, relation_transformer.py:372-374, fused)
 *   out        [B, H, N1, N2] fp32
 *   relu_bits  NULL or [B, N1, ceil(N2/32), H] uint32 (16-byte aligned); bit (j%32) of word
 *              [b, i, j/32, h] = pre-activation of out[b,h,i,j] > 0.  Needed by the backward because
 *              the caller mutates `out` in place.
 *   workspace  rdetr_relation_workspace_bytes() bytes of device scratch (per-box sin/cos tables of
 *              the FAST mode; 0 bytes / may be NULL in EXACT mode); contents undefined on return.
 * Supported: H == 8, 64 input features (num_pos_feats 16 x 4 box features).
 */
size_t rdetr_relation_workspace_bytes(int B, int N1, int N2, int flags);
int rdetr_relation_forward(const float *src_boxes, const float *tgt_boxes, const float *weight,
                           const float *bias, const float *dim_t, float scale, float eps,
                           const uint8_t *attn_mask, float *out, uint32_t *relu_bits, int B, int N1,
                           int N2, int H, int flags, void *workspace, size_t workspace_bytes,
                           rdetr_stream_t stream);

/*
 * Fused position-relation embedding, backward (parameters only; boxes carry no gradient because
 * the reference computes the geometry under no_grad, relation_transformer.py:527-529):
 *   G = grad_out * relu_bits;  grad_weight[h,n] = sum_{b,i,j} G * f_n;  grad_bias[h] = sum G
 *
 *   grad_out    [B, H, N1, N2] fp32
 *   relu_bits   as written by the forward (required)
 *   grad_weight [H, 64] fp32, grad_bias [H] fp32: ZEROED INSIDE this call, then accumulated
 *   workspace   as for the forward (same size; the tables are recomputed)
 */
int rdetr_relation_backward(const float *src_boxes, const float *tgt_boxes, const float *dim_t,
                            float scale, float eps, const float *grad_out,
                            const uint32_t *relu_bits, float *grad_weight, float *grad_bias, int B,
                            int N1, int N2, int H, int flags, void *workspace,
                            size_t workspace_bytes, rdetr_stream_t stream);

/*
 * Decoder self-attention with the position-relation bias generated on the fly.
 * Replaces, between nn.MultiheadAttention's input and output projections, the chain
 *   PositionRelationEmbedding(src, tgt).flatten(0, 1) -> masked_fill_(attn_mask, -inf) -> attn_mask of the attention
 * (upstream models/bricks/relation_transformer.py:369-374 producer, :453-459 consumer): the [B, H, N, N] bias and its
 * gradient never exist in HBM on the forward path.
 *   out[b,h,i,:] = sum_j softmax_j(q[b,h,i,:].k[b,h,j,:] / sqrt(D) + relu(weight[h,:].f(src_i, tgt_j) + bias[h])) v[b,h,j,:]
 *
 *   q, k, v     [B, H, N, D] fp32, contiguous, 16-byte aligned     out [B, H, N, D] fp32
 *   src_boxes   [B, N, 4] boxes of the query rows, tgt_boxes [B, N, 4] boxes of the keys (fp32 cxcywh)
 *   weight [H, 64], bias [H], dim_t [8] (device), scale, eps: as rdetr_relation_forward (FAST arithmetic only)
 *   attn_mask   NULL or [N, N] bytes (torch.bool), non-zero = key j blocked for query i
 *   lse         [B, H, N] fp32: log-sum-exp of every row, the only thing the backward needs besides the inputs and out
 *   grad_q / grad_k / grad_v like q; grad_weight [H, 64], grad_bias [H]: all ZEROED INSIDE the backward call
 *   workspace   rdetr_relation_attention_workspace_bytes(B, N, H, backward) bytes: per-box tables, plus -- backward only --
 *               one [B, H, N, N] fp32 buffer through which the ReLU-gated score gradient reaches the relation backward,
 *               plus -- forward only, small grids -- the partial (o, m, l) of the key splits (B * ceil(N/32) CTAs would
 *               not fill the 148 SMs at the training shape B = 2; up to 8 CTAs then share the keys of a row block)
 * Supported: H == 8, D == 32.  A row whose keys are all blocked yields NaN, as torch's softmax does.
 */
size_t rdetr_relation_attention_workspace_bytes(int B, int N, int H, int backward);
int rdetr_relation_attention_forward(const float *q, const float *k, const float *v, const float *src_boxes,
                                     const float *tgt_boxes, const float *weight, const float *bias,
                                     const float *dim_t, float scale, float eps, const uint8_t *attn_mask,
                                     float *out, float *lse, int B, int N, int H, int D, void *workspace,
                                     size_t workspace_bytes, rdetr_stream_t stream);
int rdetr_relation_attention_backward(const float *q, const float *k, const float *v, const float *src_boxes,
                                      const float *tgt_boxes, const float *weight, const float *bias,
                                      const float *dim_t, float scale, float eps, const uint8_t *attn_mask,
                                      const float *out, const float *lse, const float *grad_out, float *grad_q,
                                      float *grad_k, float *grad_v, float *grad_weight, float *grad_bias, int B,
                                      int N, int H, int D, void *workspace, size_t workspace_bytes,
                                      rdetr_stream_t stream);

/*
 * The encoder's memory_fusion input Linear without the concatenation.
 * Replaces torch.cat(queries, -1) -> Linear((L+1)*C, N) [-> ReLU] (upstream models/bricks/relation_transformer.py:168-173,
 * 203-204): out = relu(sum_t sources[t] @ weight[:, t*C:(t+1)*C]^T + bias), the K loop walking the sources in place.
 * Dense contraction => tcgen05.mma (kind::tf32, fp32 accumulation in TMEM), operands fetched by TMA.
 *
 *   sources  HOST array of nsrc DEVICE pointers, each [M, C] fp32 row-major, 16-byte aligned
 *   weight   [N, nsrc*C] fp32 (nn.Linear layout)     bias [N] fp32     out [M, N] fp32
 * Supported: N == 256, nsrc <= 8, C % 32 == 0.  TF32 products: use where the reference runs this GEMM in bf16 / tf32.
 */
int rdetr_memory_fusion_forward(const float *const *sources, int nsrc, const float *weight, const float *bias,
                                float *out, long long M, int C, int N, int relu, rdetr_stream_t stream);

/*
 * Two-stage query selection (second half).  Replaces, in RelationTransformer.forward
 * (upstream models/bricks/relation_transformer.py:90-96 and :104-111),
 *   enc_outputs_coord = (bbox_head(output_memory) + output_proposals).sigmoid()
 *   topk_index = torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)[1]
 *   enc_outputs_class.gather(1, topk_index...) ; enc_outputs_coord.gather(1, topk_index...)
 * by three launches: row maxima, one-CTA-per-image radix select + sort, row gather (sigmoid applied to the K selected
 * boxes only).  Index work: topk_index is torch.topk's (scores descending, NaN first); EQUAL scores are returned in
 * ascending index order, a deterministic refinement of torch's unspecified tie order.
 *
 *   class_logits [B, S, C] fp32     coord [B, S, 4] fp32 or NULL (then topk_coord must be NULL)
 *   apply_sigmoid  non-zero: coord holds pre-sigmoid boxes, topk_coord = sigmoid(selected rows), torch's arithmetic
 *   topk_class [B, K, C], topk_coord [B, K, 4], topk_index [B, K] int64; K <= S, K <= 4096
 *   workspace  rdetr_two_stage_workspace_bytes(B, S) bytes (the row maxima)
 * rdetr_topk_rows is the middle launch alone: torch.topk(scores, K, dim=1) for scores [B, S] (values may be NULL).
 * Backward: grad_class [B, S, C] / grad_coord [B, S, 4] are ZEROED INSIDE the call, then the K selected rows are
 * written (with sigmoid' = (1 - y) * y from topk_coord when apply_sigmoid); either gradient may be NULL.
 */
size_t rdetr_two_stage_workspace_bytes(int B, int S);
int rdetr_topk_rows(const float *scores, int B, int S, int K, int64_t *indices, float *values, rdetr_stream_t stream);
int rdetr_two_stage_select(const float *class_logits, const float *coord, int B, int S, int C, int K, int apply_sigmoid,
                           float *topk_class, float *topk_coord, int64_t *topk_index, void *workspace,
                           size_t workspace_bytes, rdetr_stream_t stream);
int rdetr_two_stage_select_backward(const float *grad_topk_class, const float *grad_topk_coord, const float *topk_coord,
                                    const int64_t *topk_index, int B, int S, int C, int K, int apply_sigmoid,
                                    float *grad_class, float *grad_coord, rdetr_stream_t stream);

/*
 * Batched rectangular linear-sum-assignment.
 * Replaces scipy.optimize.linear_sum_assignment(c.cpu()) at models/matcher/hungarian_matcher.py:80 and :87
 * (one device->host copy + host solve per image and per decoder layer).  All problems of one call are
 * solved by one launch (one CTA per problem; 64 problems per launch) on `stream`; nothing synchronises.
 *
 *   cost[p]     device, float32 [n_rows[p], n_cols[p]] row-major (the matcher's cost matrix, queries x boxes)
 *   row_ind[p]  device, int64 [min(n_rows[p], n_cols[p])]: input rows of the pairs, ascending
 *   col_ind[p]  device, int64, same length: the column matched to each of those rows
 *   status      device, int32 [n_problems]: 0 ok, 1 infeasible, 2 NaN / -inf entry (SciPy raises ValueError
 *               for both); on 1 and 2 the index buffers are filled with -1
 *   cost, n_rows, n_cols, row_ind, col_ind are HOST arrays of n_problems entries (pointers / extents).
 *
 * The optimum returned is the one SciPy returns, ties included: the kernel restates SciPy's shortest-
 * augmenting-path solver (Crouse 2016) with its scan order, in double precision on the float32 costs.
 * Limits: rows*cols < 2^31 and 29*max(rows,cols) + 13*min(rows,cols) + 16 <= 204800 bytes per problem
 * (state in shared memory; 1500 x 600 needs 51 KB) -> RDETR_ERR_UNSUPPORTED.
 * Workspace: rdetr_lsap_workspace_bytes() bytes of device memory (a transposed float32 copy of every
 * problem with more rows than columns), 256-byte aligned.
 */
size_t rdetr_lsap_workspace_bytes(const int64_t *n_rows, const int64_t *n_cols, int n_problems);
int rdetr_lsap_solve(const float *const *cost, const int64_t *n_rows, const int64_t *n_cols,
                     int64_t *const *row_ind, int64_t *const *col_ind, int32_t *status, int n_problems,
                     void *workspace, size_t workspace_bytes, rdetr_stream_t stream);

/*
 * The matcher's cost matrices for a batch of problems in one launch (row N3, with rdetr_lsap_solve).
 * Replaces HungarianMatcher.calculate_cost, models/matcher/hungarian_matcher.py:40-72 (about 30 eager
 * kernels per problem): cost[q, g] = w_bbox * L1(pred_boxes[q], gt_boxes[g])
 *                                  + w_class * (pos - neg)(sigmoid(pred_logits[q, gt_labels[g]]))
 *                                  + w_giou * -GIoU(xyxy(pred_boxes[q]), xyxy(gt_boxes[g]))
 * evaluated with the same single-precision operations in the same order as the eager chain, so the
 * values are the same BITS (tests/test_lsap_gpu.py: torch.equal on the GPU) and the assignment that
 * follows is the reference's.
 *   pred_boxes[p] float32 [n_queries[p], 4] cxcywh (16-byte aligned), pred_logits[p] float32
 *   [n_queries[p], num_classes], gt_boxes[p] float32 [n_gt[p], 4], gt_labels[p] int64 [n_gt[p]],
 *   cost[p] float32 [n_queries[p], n_gt[p]] -- all device; the arrays of pointers / extents are HOST arrays.
 */
int rdetr_match_cost(const float *const *pred_boxes, const float *const *pred_logits,
                     const float *const *gt_boxes, const int64_t *const *gt_labels, float *const *cost,
                     const int64_t *n_queries, const int64_t *n_gt, int num_classes, float w_class,
                     float w_bbox, float w_giou, double focal_alpha, double focal_gamma, int n_problems,
                     rdetr_stream_t stream);

/*
 * Diagnostics (not on the product path): measure on the current device the two hardware rates that
 * bound the MSDA kernels -- random 128-byte row gathers (8 lanes x 16-byte read-only loads per row) and
 * 128-byte vector reductions (red.global.add.v4.f32) -- over a caller-provided table of nrows rows.
 * The caller times the launch with CUDA events; *rows_out (host memory) receives the rows touched.
 */
int rdetr_diag_gather_rows(const void *table, long long nrows, int iters, float *sink,
                           long long *rows_out, rdetr_stream_t stream);
int rdetr_diag_red_rows(void *table, long long nrows, int iters, long long *rows_out,
                        rdetr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RDETR_OPS_H_ */
