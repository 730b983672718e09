/*
 * tdnnf_nas_b200.h -- C ABI of the B200-native (sm_100a) training hot path of TDNN-F_NAS.
 *
 * This is the drop-in boundary: the entry points below are what the six reference nnet3
 * components (and the chain denominator) call IN PLACE OF the CuMatrix/CuVector methods
 * they call in the reference.  Plain pointers and sizes only; every function returns an
 * int status (0 = ok) and never throws; memory is caller-owned except the opaque handles;
 * nothing here synchronises the host with the device unless the comment says so.
 *
 * Conventions
 *   - All matrices are row-major fp32 (Kaldi BaseFloat) given as (device pointer, rows,
 *     cols, stride-in-elements), exactly a CuMatrixBase view.  stride >= cols.
 *   - Kernels run on the context's stream (tdnnf_ctx_set_stream).
 *   - Citations "ref: file:line" are into the reference tree (skhu101/TDNN-F_NAS):
 *       tdnn.cc   = src/nnet3/nnet-tdnn-component.cc
 *       simple.cc = src/nnet3/nnet-simple-component.cc
 *       norm.cc   = src/nnet3/nnet-normalize-component.cc
 *     "kaldi:" citations are into upstream kaldi-asr/kaldi (not shipped with the reference).
 */
#ifndef TDNNF_NAS_B200_H_
#define TDNNF_NAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDNNF_OK 0
#define TDNNF_ERR_INVALID 1  /* bad argument (the reference would KALDI_ASSERT / KALDI_ERR) */
#define TDNNF_ERR_CUDA 2     /* a CUDA runtime / driver call failed                         */
#define TDNNF_ERR_NOMEM 3
#define TDNNF_ERR_UNSUPPORTED 4

#define TDNNF_MAX_OFFSETS 16

typedef struct tdnnf_ctx tdnnf_ctx;
typedef struct tdnnf_den_graph tdnnf_den_graph;
typedef struct tdnnf_den_comp tdnnf_den_comp;

/* ------------------------------------------------------------------ context ---------- */
/* Message of the last failing call on this host thread (never NULL). */
const char* tdnnf_last_error(void);
/* ABI version: major*1000 + minor. */
int tdnnf_abi_version(void);

/* Create a context on CUDA device `device` (replaces Kaldi's CuDevice singleton for this
 * path).  Fails with TDNNF_ERR_CUDA when there is no sm_100 device: there is NO CPU path. */
int tdnnf_ctx_create(int device, tdnnf_ctx** out);
int tdnnf_ctx_destroy(tdnnf_ctx* ctx);
/* stream: a cudaStream_t passed as void* (NULL = legacy default stream). */
int tdnnf_ctx_set_stream(tdnnf_ctx* ctx, void* stream);
int tdnnf_ctx_get_stream(tdnnf_ctx* ctx, void** stream);
/* Operand precision of the tensor-core GEMMs behind tdnnf_darts_{propagate,backprop_data,backprop_params}:
 * every fp32 operand is split into `planes` bf16 planes.  2 (default): x = hi + lo to ~2^-17, three products
 * per K step, results within ~5e-6 of fp32.  3: hi + mid + lo to 2^-24, six products, fp32-level results at
 * twice the tensor-pipe time -- used by the natural-gradient update, whose eigen-problem amplifies rounding. */
int tdnnf_ctx_set_gemm_planes(tdnnf_ctx* ctx, int planes);
/* Parameter-gradient operands with at least `min_rows` rows are contracted MN-major over the row planes (no
 * transposed pre-pass); smaller ones, and the fp16 gradient mode, use transposed planes.  Default 512; 1 = always,
 * a negative value = never.  (Environment: TDNNF_WGRAD_MN=0 disables, TDNNF_WGRAD_MN_MIN_ROWS overrides.) */
int tdnnf_ctx_set_wgrad_mn_min_rows(tdnnf_ctx* ctx, int min_rows);
/* Precision of the parameter-gradient GEMM (north-star tolerance: gradients 1e-3, forward 1e-4).
 * fast = 0 (default): three bf16 products per K step (~5e-6).
 * fast = 1: tdnnf_darts_backprop_params runs ONE product: both operands as one fp16 plane, scaled into the fp16
 *           range by an exact power of two of their max magnitude (measured 2.9e-4).  The data gradient keeps three
 *           products (its error would accumulate down the layers); tdnnf_darts_propagate is never affected; neither
 *           is anything run with 3 planes (the natural-gradient update). */
int tdnnf_ctx_set_gradient_mode(tdnnf_ctx* ctx, int fast);
/* Operand-plane cache.  Between begin and end, the bf16 operand planes built from one of the `num_sources` (<= 8)
 * registered device matrices (matched by base pointer, shape and stride) are kept and reused by later
 * tdnnf_darts_* calls that need the same matrix in the same form.  TdnnDARTSV3Component::Backprop (ref: tdnn.cc:335-431
 * and 457-626) reads in_value and out_deriv 3-4 times each: data gradient, the two PreconditionDirections calls, the
 * parameter gradient.  The caller promises the registered matrices are not written inside the scope.  Scopes do not nest. */
int tdnnf_ctx_operand_cache_begin(tdnnf_ctx* ctx, const float* const* sources, int num_sources);
int tdnnf_ctx_operand_cache_end(tdnnf_ctx* ctx);
int tdnnf_ctx_operand_cache_stats(const tdnnf_ctx* ctx, uint64_t* hits, uint64_t* misses);
/* Operand planes with a life of their own.  A GEMM operand (fp32 matrix) is split into bf16 hi/lo row planes before every
 * tensor-core product; tdnnf_planes is one such split kept by the caller:
 *   tdnnf_planes_acquire   the planes of (src, rows, cols, stride) for a component with `row_stride` (the reorder_t_in of
 *                          its PrecomputedIndexes): an attached handle for the same matrix is returned with one more
 *                          reference, otherwise the split runs once (one pass over src, also leaving the per-row sums of
 *                          squares the natural gradient needs);
 *   tdnnf_ctx_planes_attach / _detach   while attached, every tdnnf_darts_* call whose operand is that very matrix (same
 *                          pointer, shape, stride, row stride) uses the planes instead of splitting it again.  The caller
 *                          promises the matrix does not change while its planes are attached;
 *   tdnnf_planes_release   drops a reference; the memory returns to the context's pool.
 * TdnnDARTSV3Component::Propagate keeps the planes of its input in the memo and Backprop attaches them again (the
 * reference's Backprop re-reads in_value: tdnn.cc:476-539); the fused tail kernels below hand over planes of what they
 * have just written (tdnnf_relu_scale_offset_bypass_{fwd,bwd}_planes). */
typedef struct tdnnf_planes tdnnf_planes;
int tdnnf_planes_acquire(tdnnf_ctx* ctx, const float* src, int rows, int cols, int stride, int row_stride, tdnnf_planes** out);
int tdnnf_planes_release(tdnnf_planes* p);
int tdnnf_ctx_planes_attach(tdnnf_ctx* ctx, tdnnf_planes* p);
int tdnnf_ctx_planes_detach(tdnnf_ctx* ctx, tdnnf_planes* p);
/* 1 if p holds the planes of exactly this matrix */
int tdnnf_planes_matches(const tdnnf_planes* p, const float* src, int rows, int cols, int stride);
/* Pre-size the internal scratch arena (bf16 operand planes) so later calls never grow it. */
int tdnnf_ctx_reserve(tdnnf_ctx* ctx, uint64_t bytes);
/* Number of kernels this context has launched so far (for gpu_launches accounting). */
uint64_t tdnnf_ctx_launch_count(const tdnnf_ctx* ctx);
/* Roofline instrumentation: while enabled, every tensor-core GEMM launch is bracketed by CUDA events on
 * the context's stream and its algorithmic FLOPs (2*M*N*K over the un-padded operands, one pass: the
 * bf16 hi/lo split is NOT counted) are recorded.  tdnnf_ctx_gemm_timing_read synchronises the stream and
 * returns the sums since enabling, then clears them. */
int tdnnf_ctx_gemm_timing_enable(tdnnf_ctx* ctx, int enable);
int tdnnf_ctx_gemm_timing_read(tdnnf_ctx* ctx, double* total_ms, double* total_flops, uint64_t* launches);
/* The same, split by size: launches with at least `min_flops` algorithmic FLOPs go into total_* (tensor_pipe_flops =
 * sum of FLOPs x products per K step actually issued: 1, 2, 3 or 6), the rest (skinny natural-gradient products)
 * into other_*. */
int tdnnf_ctx_gemm_timing_read_ex(tdnnf_ctx* ctx, double min_flops, double* total_ms, double* total_flops,
                                  double* tensor_pipe_flops, uint64_t* launches, double* other_ms, uint64_t* other_launches);

/* ------------------------------------------------------------------ TdnnDARTSV3 ------- */
/* mode flags of TdnnDARTSV3Component (ref: conv.h:243-257) */
#define TDNNF_DARTS_USE_GUMBEL 1
#define TDNNF_DARTS_FREE_SELECT 2
#define TDNNF_DARTS_UNIFORM_SAMPLE 4
#define TDNNF_DARTS_USE_ENTROPY 8
#define TDNNF_DARTS_UPDATE_ALPHA 16

/* Mixing coefficients (ref: tdnn.cc:250-289; replaces SetRandUniform/ApplyLog/Scale/AddVec/
 * ApplySoftMax/ApplyFloor/ApplyExp/InvertElements and the per-element .Max() host syncs).
 *   alpha        device, n log-weights (bias_params_[0..n))
 *   u_gumbel     HOST, n uniforms in (0,1) for the Gumbel noise (used iff USE_GUMBEL)
 *   u_uniform    the single uniform draw of uniform-sample mode (used iff UNIFORM_SAMPLE)
 *   share_index  offset slot that is always passed with weight 1 (ref: tdnn.cc:230-241)
 *   coef         device out, n: the memo the reference returns from Propagate
 *   weff         device out, n: effective GEMM weight per offset (0 => offset skipped)
 * Randomness is injected by the caller so that every data-parallel rank draws identical
 * noise; nothing is copied back to the host. */
int tdnnf_darts_coef(tdnnf_ctx* ctx, const float* alpha, int n, int flags, float temperature,
                     const float* u_gumbel, float u_uniform, int share_index, float* coef, float* weff);

/* Effective weights from an existing memo (Backprop recomputes them, ref: tdnn.cc:352-364). */
int tdnnf_darts_weff_from_coef(tdnnf_ctx* ctx, const float* coef, int n, int flags, int share_index,
                               float* weff);

/* Propagate GEMMs (ref: tdnn.cc:230-241, 292-328; replaces CopyRowsFromVec/SetZero and the n
 * AddMatMat calls on GetInputPart views, tdnn.cc:806-820).
 *   out[k,:] = (bias_tail or 0 or out[k,:]) + sum_i weff[i] * in[row_offsets[i]+k*row_stride,:] * W_i^T
 *   W: out_dim x (n*in_dim), W_i = columns [i*in_dim,(i+1)*in_dim)
 *   bias_mode: 0 = out is accumulated into (kPropagateAdds), 1 = out overwritten starting from
 *              zero, 2 = out overwritten starting from the broadcast `bias` (out_dim, device). */
int tdnnf_darts_propagate(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride,
                          float* out, int out_rows, int out_dim, int out_stride, const float* W,
                          int w_stride, const float* bias, int bias_mode, const float* weff, int n,
                          const int32_t* row_offsets, int row_stride);

/* out (out_rows x rank) = [w_1 X_1 | ... | w_n X_n] W^T (+ bias[rank]) for a skinny W (rank x n*in_dim): the same
 * result as tdnnf_darts_propagate(..., out_dim = rank), organised for small rank: one un-spliced GEMM over the input
 * rows, then a gather-sum over the offsets.  Used for H = X W_t^T of OnlineNaturalGradient on the spliced input
 * (the reference materialises that input: in_value_temp, nnet-tdnn-component.cc:476-514). */
int tdnnf_darts_project(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride, float* out, int out_rows,
                        int rank, int out_stride, const float* W, int w_stride, const float* bias, const float* weff, int n,
                        const int32_t* row_offsets, int row_stride);
/* Backprop to the input (ref: tdnn.cc:366-416; kBackpropAdds):
 *   in_deriv[row_offsets[i]+k*row_stride,:] += weff[i] * out_deriv[k,:] * W_i   for all i,k */
int tdnnf_darts_backprop_data(tdnnf_ctx* ctx, const float* out_deriv, int out_rows, int out_dim,
                              int od_stride, float* in_deriv, int in_rows, int in_dim, int id_stride,
                              const float* W, int w_stride, const float* weff, int n,
                              const int32_t* row_offsets, int row_stride);

/* Parameter gradients (ref: tdnn.cc:482-539, 607-624 without the two PreconditionDirections
 * calls, i.e. in_scale = out_scale = 1; replaces the R x (n*D_in+1) in_value_temp, the n extra
 * AddMatMat+AddMatMatElements+Sum() and the final AddMatMat / AddMatVec):
 *   dW_i   += lr * weff[i] * out_deriv^T * X_i            (X_i = the offset-i input view)
 *   dbias  += lr * colsum(out_deriv)                      (dbias may be NULL)
 *   s[i]    = sum((X_i W_i^T) .* out_deriv) = <out_deriv^T X_i, W_i>   (s may be NULL; overwritten)
 *   W_model: the model's weights (for s), dW: the delta component's weights. */
int tdnnf_darts_backprop_params(tdnnf_ctx* ctx, const float* in_value, int in_rows, int in_dim,
                                int in_stride, const float* out_deriv, int out_rows, int out_dim,
                                int od_stride, const float* W_model, int w_stride, float* dW,
                                int dw_stride, float* dbias, const float* weff, int n,
                                const int32_t* row_offsets, int row_stride, float lr, float* s);

/* Architecture-weight update (ref: tdnn.cc:541-590): accumulates the softmax / sigmoid
 * Jacobian products of s into dalpha, then applies the x5 / xlr / x10000 scalings to the WHOLE
 * dalpha range exactly as the reference does.  All pointers device, n entries. */
int tdnnf_darts_alpha_update(tdnnf_ctx* ctx, const float* s, const float* coef, int n, int flags,
                             float temperature, int share_index, float lr, float* dalpha);

/* ------------------------------------------------------------------ mixing components - */
/* {Gumbel}SoftmaxFlopsComponent::Propagate (ref: simple.cc:10088-10113, 9968-9981):
 *   out = max(softmax_row((in + G) * inv_temp), 1e-20),  G_j = -log(-log u[j]) (u HOST, cols
 *   entries, one draw per COLUMN shared by all rows) or G = 0 when u == NULL. */
int tdnnf_softmax_flops_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                            int out_stride, const float* u, float inv_temp);
/* {Gumbel}SoftmaxFlopsComponent::Backprop (ref: simple.cc:10116-10158, 9984-10020):
 *   e = out_deriv + penalty * f,  f = (-25,-50,-80,-100,-120,-160,-200,-240) on columns 0..7
 *   in_deriv = (p .* e - p * (p . e)) * inv_temp;  penalty = scale / (rows_for_norm * cols).
 *   write_back_e != 0 reproduces the reference's mutation of out_deriv (it writes through a
 *   const reference); in_deriv may alias out_deriv (kBackpropInPlace).  cols must be >= 8. */
int tdnnf_softmax_flops_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, float* out_deriv,
                            int od_stride, float* in_deriv, int id_stride, int rows, int cols, float penalty,
                            float inv_temp, int write_back_e);
/* CopyNComponent (ref: simple.cc:4843-4867; AddMatBlocks): fwd out[:, b*in_cols + j] += scale*in[:, j];
 * bwd in_deriv[:, j] += scale * sum_b out_deriv[:, b*in_cols + j]. */
int tdnnf_copyn_fwd(tdnnf_ctx* ctx, const float* in, int rows, int in_cols, int in_stride, float* out,
                    int out_cols, int out_stride, float scale);
int tdnnf_copyn_bwd(tdnnf_ctx* ctx, const float* out_deriv, int rows, int out_cols, int od_stride,
                    float* in_deriv, int in_cols, int id_stride, float scale);
/* OnehotFunctionComponent::Propagate (ref: simple.cc:9504-9519): every row of out becomes the
 * one-hot of the slot i with i/dim <= u < (i+1)/dim (float compares); u injected by the caller. */
int tdnnf_onehot_fwd(tdnnf_ctx* ctx, float* out, int rows, int dim, int out_stride, float u);
/* vec[j] += scale * sum_rows mat[:, j]   (AddRowSumMat; ref: simple.cc:9544-9548, tdnn.cc:614) */
int tdnnf_add_row_sum(tdnnf_ctx* ctx, const float* mat, int rows, int cols, int stride, float scale,
                      float* vec);
/* BatchNormTestComponent (ref: norm.cc:843-877, 879-922; CopyFromMat+MulColsVec+AddVecToRows):
 *   fwd: out = in .* scale + offset ; bwd: in_deriv = out_deriv .* scale.  In-place allowed.
 *   scale/offset: device vectors of `cols` entries; offset NULL => none. */
int tdnnf_scale_offset_rows(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                            int out_stride, const float* scale, const float* offset);
/* ElementwiseProductComponent (ref: simple.cc:256-299): in is rows x (2*out_cols);
 * fwd: out = in[:, :D] .* in[:, D:]; bwd: in_deriv[:, :D] = od .* in[:, D:], in_deriv[:, D:] = od .* in[:, :D] */
int tdnnf_elementwise_product_fwd(tdnnf_ctx* ctx, const float* in, int rows, int out_cols, int in_stride,
                                  float* out, int out_stride);
int tdnnf_elementwise_product_bwd(tdnnf_ctx* ctx, const float* in, int in_stride, const float* out_deriv,
                                  int od_stride, float* in_deriv, int id_stride, int rows, int out_cols);

/* The shared-candidate mask of the bottleneck-dimension search block (ref: local/chain_NAS/scripts/
 * generate_bottleneckCB8share_onehottrain_config.py:11-90): the descriptors dim-range / Sum(p_j, .., p_{nb-1}), the nb
 * CopyNComponents (1 -> widths[j], `scale`), Append(copyn_j, linear_j), the nb ElementwiseProductComponents and the final
 * Append evaluated in one pass instead of ~5 nb matrix commands:
 *   fwd: out[r, c] = lin[r, c] * scale * sum_{k >= j(c)} p[r, k]       j(c) = the block column c belongs to
 *   bwd: d_lin[r, c] = d_out[r, c] * scale * sum_{k >= j(c)} p[r, k]    (d_lin may be NULL)
 *        d_p[r, k]   = scale * sum_{j <= k} sum_{c in block j} d_out[r, c] * lin[r, c]    (overwritten)
 * Equal to running the components one by one (tests/test_gpu_bottleneck_block.py).  widths: HOST, nb <= 8 entries. */
int tdnnf_shared_mask_fwd(tdnnf_ctx* ctx, const float* p, int rows, int nb, int p_stride, const float* lin, int cols,
                          int lin_stride, float* out, int out_stride, const int32_t* widths, float scale);
int tdnnf_shared_mask_bwd(tdnnf_ctx* ctx, const float* p, int p_stride, const float* lin, int lin_stride,
                          const float* d_out, int do_stride, float* d_lin, int dl_stride, float* d_p, int dp_stride,
                          int rows, int cols, int nb, const int32_t* widths, float scale);

/* ------------------------------------------------------------------ whole-parameter ops -- */
/* The CuMatrix/CuVector calls inside Scale / Add / DotProduct / PerturbParams / Vectorize of the
 * updatable components (ref: tdnn.cc:907-979, simple.cc:9606-9642): strided fp32 matrices; a
 * vector is a 1 x n matrix. */
int tdnnf_mat_set(tdnnf_ctx* ctx, float* a, int rows, int cols, int stride, float value);
int tdnnf_mat_scale(tdnnf_ctx* ctx, float* a, int rows, int cols, int stride, float scale);
/* out[r, :] = vec for every row (CopyRowsFromVec; ref: simple.cc:2606, 9518, tdnn.cc:234). */
int tdnnf_copy_rows_from_vec(tdnnf_ctx* ctx, const float* vec, float* out, int rows, int cols, int stride);
/* dst += alpha * src */
int tdnnf_mat_axpy(tdnnf_ctx* ctx, float alpha, const float* src, int src_stride, float* dst, int dst_stride,
                   int rows, int cols);
/* *result (HOST) = sum(a .* b); synchronises the stream (TraceMatMat(a, b, kTrans) / VecVec). */
int tdnnf_mat_dot(tdnnf_ctx* ctx, const float* a, int a_stride, const float* b, int b_stride, int rows, int cols,
                  float* result);
/* Same, but the result stays on the device (double accumulator): result_dev[0] += sum(a .* b). */
int tdnnf_mat_dot_dev(tdnnf_ctx* ctx, const float* a, int a_stride, const float* b, int b_stride, int rows,
                      int cols, double* result_dev);

/* The parameter step over MANY buffers in two launches (UpdateNnetWithMaxChange, ref nnet-utils.cc:2085-2175, needs
 * the squared norm of every component's delta, then model += factor * delta and delta = 0):
 *   tdnnf_multi_sumsq:     out_dev[groups[i]] += sum(bufs[i]^2)   (device doubles; groups NULL = one slot per buffer)
 *   tdnnf_multi_axpy_zero: dst[i] += factors[i] * src[i];  src[i] = 0
 * All arrays are HOST arrays of n <= TDNNF_MULTI_MAX entries; the pointers in them are device pointers. */
#define TDNNF_MULTI_MAX 96
int tdnnf_multi_sumsq(tdnnf_ctx* ctx, int n, const float* const* bufs, const int32_t* rows, const int32_t* cols,
                      const int32_t* strides, const int32_t* groups, double* out_dev);
int tdnnf_multi_axpy_zero(tdnnf_ctx* ctx, int n, float* const* dst, const int32_t* dst_strides, float* const* src,
                          const int32_t* src_strides, const int32_t* rows, const int32_t* cols, const float* factors);

/* UpdateNnetWithMaxChange followed by ScaleNnet(momentum, delta_nnet) (ref: nnet-utils.cc:2085-2175; kaldi:
 * NnetChainTrainer::TrainInternal) over the n parameter buffers of a network, as ONE call for a C++ host:
 *   dot_g = sum of squares of the delta buffers of updatable component g = groups[i]   (one launch, ONE read-back);
 *   per-component factor = max_change[g] * max_change_scale / (sqrt(dot_g) |scale|) where that is < 1 (max_change 0 = off),
 *   then the global max_param_change on sqrt(sum factor_g^2 dot_g) |scale|, all in BaseFloat as the reference;
 *   model_i += factor * scale * delta_i;  delta_i *= momentum (0 stores exact zeros)                (one launch).
 * A non-finite parameter change sets *applied = 0 ("Infinite parameter change, will not apply."): the model is left
 * untouched (the reference returns false) and the delta is still scaled by momentum.
 * Host arrays: model / delta (device pointers), strides, rows, cols, groups [n]; max_change [num_groups];
 * scale_factors_out [num_groups] (optional: factor * scale actually applied), num_max_change_per_component_applied
 * [num_groups] and num_max_change_global_applied (optional counters, incremented).  dots_dev: DEVICE double[num_groups]. */
int tdnnf_update_with_max_change(tdnnf_ctx* ctx, int n, float* const* model, const int32_t* model_strides,
                                 float* const* delta, const int32_t* delta_strides, const int32_t* rows,
                                 const int32_t* cols, const int32_t* groups, int num_groups, const float* max_change,
                                 float max_param_change, float max_change_scale, float scale, float momentum,
                                 double* dots_dev, float* scale_factors_out, int32_t* num_max_change_per_component_applied,
                                 int32_t* num_max_change_global_applied, int* applied);
/* ApplyL2Regularization (ref: nnet-utils.cc:2223-2245): delta_i += -2 * l2_regularize_scale * lrate[g] * l2_regularize[g]
 * * model_i for every buffer of an updatable component g with a non-zero product; one launch. */
int tdnnf_apply_l2_regularization(tdnnf_ctx* ctx, int n, float* const* model, const int32_t* model_strides,
                                  float* const* delta, const int32_t* delta_strides, const int32_t* rows,
                                  const int32_t* cols, const int32_t* groups, int num_groups, const float* lrate,
                                  const float* l2_regularize, float l2_regularize_scale);

/* ------------------------------------------------------------------ stock TDNN-F neighbours - */
/* The stock components that sit between the NAS components in a TDNN-F block, so
 * that a whole supernet step runs on this library: RectifiedLinearComponent without self-repair,
 * the bypass Sum(Scale(s, a), b), and BatchNormComponent in training mode
 * (ref: norm.cc:401-465, 467-549, 551-589). */
int tdnnf_relu_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out, int out_stride);
/* in_deriv = out_deriv .* (out_value > 0) */
int tdnnf_relu_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv, int od_stride,
                   float* in_deriv, int id_stride, int rows, int cols);
/* dst[r, :] = src[row_map[r], :] (row_map device int32, -1 => zero row): CuMatrix::CopyRows, the
 * row gather nnet3 inserts where ReorderIndexes changed a component's input order. */
int tdnnf_copy_rows(tdnnf_ctx* ctx, const float* src, int src_stride, float* dst, int dst_stride, int dst_rows, int cols,
                    const int32_t* row_map);
/* dst[row_map[r], :] += alpha * src[r, :] for r < src_rows (row_map entries unique or -1): the transpose of
 * tdnnf_copy_rows (CuMatrix::AddToRows). */
int tdnnf_add_to_rows(tdnnf_ctx* ctx, float alpha, const float* src, int src_stride, int src_rows, int cols, float* dst,
                      int dst_stride, const int32_t* row_map);
/* out = a * alpha + b * beta   (out may alias a or b) */
int tdnnf_add_scaled(tdnnf_ctx* ctx, const float* a, int a_stride, float alpha, const float* b, int b_stride,
                     float beta, float* out, int out_stride, int rows, int cols);
/* Fused tail of a TDNN-F block in the search stage: ReLU -> BatchNormTest (scale/offset) -> bypass sum, one
 * pass instead of three (and no intermediate matrices):
 *   fwd: out = relu(x) .* scale + offset + bypass_scale * prev
 *   bwd: d_x = (x > 0) ? d_out .* scale : 0 ;  d_prev = bypass_scale * d_out   (d_prev overwritten)
 * Equivalent to RectifiedLinearComponent + BatchNormTestComponent (ref: norm.cc:843-922) +
 * Sum(Scale(bypass_scale, prev), .) evaluated separately. */
int tdnnf_relu_scale_offset_bypass_fwd(tdnnf_ctx* ctx, const float* x, int rows, int cols, int x_stride,
                                       const float* scale, const float* offset, const float* prev, int prev_stride,
                                       float bypass_scale, float* out, int out_stride);
int tdnnf_relu_scale_offset_bypass_bwd(tdnnf_ctx* ctx, const float* d_out, int do_stride, const float* x, int x_stride,
                                       const float* scale, float bypass_scale, float* d_x, int dx_stride, float* d_prev,
                                       int dp_stride, int rows, int cols);
/* The same two kernels as producers of the next GEMM operand: besides out / d_x they write its bf16 hi/lo planes, the
 * per-row sums of squares and (backward) the column sums of d_x, i.e. the bias gradient; *planes is a new handle
 * (tdnnf_ctx_planes_attach it around the component call that reads the matrix, then tdnnf_planes_release).
 * cols <= 4096. */
int tdnnf_relu_scale_offset_bypass_fwd_planes(tdnnf_ctx* ctx, const float* x, int rows, int cols, int x_stride,
                                              const float* scale, const float* offset, const float* prev, int prev_stride,
                                              float bypass_scale, float* out, int out_stride, tdnnf_planes** planes);
int tdnnf_relu_scale_offset_bypass_bwd_planes(tdnnf_ctx* ctx, const float* d_out, int do_stride, const float* x, int x_stride,
                                              const float* scale, float bypass_scale, float* d_x, int dx_stride, float* d_prev,
                                              int dp_stride, int rows, int cols, tdnnf_planes** planes);
/* BatchNorm, training mode.  memo: device, 5 x cols (rows: mean, uvar, scale, -, -) as in the reference.
 *   fwd:  mean/var over rows; scale = target_rms * (var + eps)^-0.5; out = (in - mean) .* scale
 *   bwd:  x' = scale .* (z' - mean(z')) + z .* var_deriv_mod,
 *         var_deriv_mod = -1/target_rms^2 * mean(z' .* z) .* scale      (ref: norm.cc:392-398) */
int tdnnf_batchnorm_train_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                              int out_stride, float epsilon, float target_rms, float* memo);
int tdnnf_batchnorm_train_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv,
                              int od_stride, float* in_deriv, int id_stride, int rows, int cols, float target_rms,
                              const float* memo);
/* NonlinearComponent statistics (ref: nnet-component-itf.cc:433-481), device doubles as the reference's CuVector<double>:
 *   store_stats:          value_sum[c] += sum_r out[r][c];  deriv_sum[c] += #{r: out[r][c] > 0}  (deriv_sum may be NULL)
 *   store_backprop_stats: oderiv_sumsq[c] += sum_r out_deriv[r][c]^2
 * RectifiedLinearComponent::RepairGradients (ref: nnet-simple-component.cc:990-1074), after the thresholds were scaled by
 * the frame count: in_deriv[r][c] += -scale * ((s[c] > lower) + (s[c] > upper) - 1), s = deriv_sum averaged over the
 * blocks_per_row blocks of the full dimension; *num_dims_repaired += #{c: term != 0}. */
int tdnnf_nonlinear_store_stats(tdnnf_ctx* ctx, const float* out_value, int rows, int cols, int stride, double* value_sum,
                                double* deriv_sum);
int tdnnf_nonlinear_store_backprop_stats(tdnnf_ctx* ctx, const float* out_deriv, int rows, int cols, int stride,
                                         double* oderiv_sumsq);
int tdnnf_relu_repair_gradients(tdnnf_ctx* ctx, float* in_deriv, int rows, int block_dim, int stride, int blocks_per_row,
                                const double* deriv_sum, float lower_threshold, float upper_threshold, float scale,
                                double* num_dims_repaired);
/* BatchNormComponent::StoreStats (ref: norm.cc:551-589): stats[0:cols] += num_frames * mean, stats[cols:2cols] +=
 * num_frames * uvar (device doubles, as the reference's CuVector<double>), mean / uvar = rows 0 / 1 of memo. */
int tdnnf_batchnorm_accumulate_stats(tdnnf_ctx* ctx, const float* memo, int cols, float num_frames, double* stats);

/* ------------------------------------------------------------------ natural gradient --- */
/* Device-side pieces of OnlineNaturalGradient::PreconditionDirections (kaldi: nnet3/natural-gradient-online.cc;
 * called at ref tdnn.cc:598-599 and simple.cc:9542) that are not GEMMs.  The GEMM-shaped pieces (H = X W^T,
 * J = H^T X, L = H^T H, K = J J^T, W_{t+1} = A B) go through tdnnf_darts_{propagate,backprop_data,backprop_params}
 * (with one offset they are plain products), so the spliced input [w_1 X_1 | ... | w_n X_n | 1] of
 * tdnn.cc:476-514 is never materialised.
 *
 * sumsq[i] (device double[n], overwritten) = sum over the rows of view i of ||row||^2, view i = rows
 * row_offsets[i] + k*row_stride, k < out_rows (GetInputPart, ref tdnn.cc:806-820); n = 1, offset 0, stride 1
 * gives ||in||_F^2.  This is TraceMatMat(X, X, kTrans) of the spliced matrix once weighted by w_i^2. */
int tdnnf_darts_view_sumsq(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride, int out_rows,
                           int n, const int32_t* row_offsets, int row_stride, double* sumsq);
/* out3 (device float[3]) = { tr(X X^T), tr(Xhat Xhat^T), sqrt(ratio) } for Xhat = X - (X W^T) W, from
 * sumsq (above), the block weights weff (device [n] or NULL = 1), ones_rows (= row count when X carries the
 * appended column of ones, else 0), L = (X W^T)^T (X W^T) and W W^T (rank x rank).  No host sync. */
int tdnnf_ng_scale(tdnnf_ctx* ctx, const double* sumsq, const float* weff, int n, float ones_rows, const float* L,
                   int l_stride, const float* WWt, int w_stride, int rank, float* out3);
/* Inside an operand-cache scope: *rowsq = device array of per-row sums of squares of the registered matrix `source`
 * (`rows` rows), produced as a by-product of its first row-major operand split in this scope, or NULL if there was none. */
int tdnnf_ctx_operand_rowsq(tdnnf_ctx* ctx, const float* source, int rows, const float** rowsq);
/* The device half of one OnlineNaturalGradient step after H = X W^T, in two launches and one pass over H (rank <= 128):
 *   L (rank x rank, overwritten) = H^T H in fp32 FMAs (kaldi: L_t = H_t^T H_t), and
 *   out3 = { tr(X X^T), tr(Xhat Xhat^T), scale } as tdnnf_ng_scale.
 * tr(X X^T): with rowsq != NULL (tdnnf_ctx_operand_rowsq) from the per-view sums of rowsq (view i = rows row_offsets[i] +
 * k*row_stride, k < rows; sumsq is not used); with rowsq == NULL from sumsq[n] as tdnnf_darts_view_sumsq computes it. */
int tdnnf_ng_gram_scale(tdnnf_ctx* ctx, const float* H, int rows, int rank, int h_stride, float* L, int l_stride,
                        const float* WWt, int w_stride, const float* rowsq, double* sumsq, int in_rows, int n,
                        const int32_t* row_offsets, int row_stride, const float* weff, float ones_rows, float* out3);
/* G <- (I - Wo^T Wo) G (I - Wi^T Wi): the rank-r projections of OnlineNaturalGradient::PreconditionDirections applied to
 * the gradient G = out_deriv^T X (rows = D_out, cols = n D_in + 1) instead of to its operands: X_hat = X (I - Wi^T Wi) on
 * the input side and out_deriv_hat = out_deriv (I - Wo^T Wo) on the output side (ref: tdnn.cc:598-624).  Wi: ri x cols,
 * Wo: ro x rows (device, row-major; either may be NULL = no projection on that side), ranks <= 128.  In place, fp32. */
int tdnnf_ng_project_gradient(tdnnf_ctx* ctx, float* G, int rows, int cols, int g_stride, const float* Wi, int ri,
                              int wi_stride, const float* Wo, int ro, int wo_stride);
/* n <= 4 strided 2-D copies dst_k[r][c] = src_k[r][c] in one launch.  Either side may be page-locked host memory
 * (cudaMallocHost / cudaHostAlloc: device-accessible under unified addressing), so small matrices reach a host thread --
 * after an event recorded behind the call -- without the copy engine. */
int tdnnf_copy_blocks(tdnnf_ctx* ctx, int n, const float* const* src, const int32_t* src_strides, float* const* dst,
                      const int32_t* dst_strides, const int32_t* rows, const int32_t* cols);
/* W_next (rank x dim, overwritten) = A J + AC W  with A, AC rank x rank (rank <= 128) and J, W rank x dim: the update
 * W_{t+1} = A_t (J_t + diag(c) W_t) of OnlineNaturalGradient (kaldi: ComputeWt1), in exact fp32 FMAs. */
int tdnnf_ng_w_update(tdnnf_ctx* ctx, const float* A, int a_stride, const float* AC, int ac_stride, const float* J, int j_stride,
                      const float* W, int w_stride, int rank, int dim, float* W_next, int out_stride);
/* dst += alpha * (*factor1_dev) * (*factor2_dev) * src   (device scalars, either may be NULL = 1) */
int tdnnf_mat_axpy_dev(tdnnf_ctx* ctx, float alpha, const float* factor1_dev, const float* factor2_dev,
                       const float* src, int src_stride, float* dst, int dst_stride, int rows, int cols);
/* The same, and src = 0 afterwards (the gradient scratch of UpdateNaturalGradient is consumed exactly once). */
int tdnnf_mat_axpy_dev_zero(tdnnf_ctx* ctx, float alpha, const float* factor1_dev, const float* factor2_dev, float* src,
                            int src_stride, float* dst, int dst_stride, int rows, int cols);

/* ------------------------------------------------------------------ chain denominator - */
/* DenominatorGraph (kaldi: chain/chain-den-graph.{h,cc}).  Host arrays, copied to the device.
 *   fwd_ranges / bwd_ranges: num_states pairs [begin,end) into `transitions` (forward list
 *   first, then backward list), transitions: {prob, pdf_id, hmm_state} as three parallel arrays. */
int tdnnf_den_graph_create(tdnnf_ctx* ctx, int num_states, int num_pdfs, int num_transitions,
                           const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* trans_prob,
                           const int32_t* trans_pdf, const int32_t* trans_state, const float* initial_probs,
                           tdnnf_den_graph** out);
int tdnnf_den_graph_destroy(tdnnf_den_graph* g);

/* DenominatorComputation (kaldi: chain/chain-denominator.{h,cc}).  nnet_output is
 * (frames_per_seq*num_seqs) x num_pdfs with row = t*num_seqs + s. */
int tdnnf_den_create(tdnnf_ctx* ctx, const tdnnf_den_graph* g, int num_seqs, int frames_per_seq,
                     float leaky_hmm_coefficient, tdnnf_den_comp** out);
int tdnnf_den_destroy(tdnnf_den_comp* c);
/* Which kernels this computation runs: *path = 0 one launch per frame and direction (the default), 2 = sequence-slice
 * clusters (opt-in, TDNNF_DEN_PATH=slices; num_seqs a multiple of 8 and 8 x num_pdfs floats within shared memory: one launch
 * per direction, a cluster of *cluster CTAs per slice of 8 sequences, E(t) staged in shared memory in *parts pdf ranges,
 * *ctas CTAs in all; measured slower, see den_slices.cu), 1 = the resident experiment (TDNNF_DEN_RESIDENT=1).
 * Environment: TDNNF_DEN_PATH=slices, TDNNF_DEN_PARTS=1|2, TDNNF_DEN_CLUSTER=1..16. */
int tdnnf_den_describe(const tdnnf_den_comp* c, int* path, int* cluster, int* parts, int* ctas);
/* Forward(): returns the total log-prob summed over sequences in *logprob (HOST; this call
 * synchronises the stream -- the reference returns the scalar the same way). */
int tdnnf_den_forward(tdnnf_den_comp* c, const float* nnet_output, int stride, float* logprob);
/* Backward(deriv_weight, &nnet_output_deriv): nnet_output_deriv += deriv_weight * posterior.
 * *ok (HOST) = 0 when the t=0 alpha.beta check fails (the reference's return value).  Syncs. */
int tdnnf_den_backward(tdnnf_den_comp* c, float deriv_weight, float* nnet_output_deriv, int stride, int* ok);

/* LogSoftmaxComponent (ref: nnet-simple-component.cc:3607-3632; the `output-xent` branch of the chain recipes):
 *   fwd  out = in - log(sum_j exp(in_j)) per row (ApplyLogSoftMaxPerRow);
 *   bwd  in_deriv = out_deriv - exp(out_value) * rowsum(out_deriv) (DiffLogSoftmaxPerRow; in_deriv may alias out_deriv). */
int tdnnf_log_softmax_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out, int out_stride);
int tdnnf_log_softmax_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv, int od_stride,
                          float* in_deriv, int id_stride, int rows, int cols);

/* ------------------------------------------------------------------ dropout ------------ */
/* GeneralDropoutComponent (kaldi: nnet3/nnet-general-component.cc; the `dropout` of every tdnnf-layer and
 * relu-batchnorm-dropout-layer, driven by the `set-dropout-proportion` directive, ref: nnet-utils.cc:1297-1330).
 * tdnnf_dropout_mask = GetMemo: element e of the mask uses draw number counter + e of the component layer's counter
 * hash (seed, counter: tdnnf_nnet3_set_rand_seed / _get_rand_counter), u in (0,1):
 *   continuous == 0: mask = (u > p) / (1 - p);   continuous != 0: mask = 1 - 2p + 4p u.
 * tdnnf_mul_rows_indexed = CuMatrixBase::MulRows: out[r,:] = in[r,:] .* mask[row_index[r],:] (in may alias out;
 * row_index: DEVICE int32[rows]); the same call is the Backprop. */
int tdnnf_dropout_mask(tdnnf_ctx* ctx, unsigned long long seed, unsigned long long counter, float* mask, int rows,
                       int cols, int stride, float proportion, int continuous);
int tdnnf_mul_rows_indexed(tdnnf_ctx* ctx, const float* in, int in_stride, float* out, int out_stride, int rows,
                           int cols, const float* mask, int mask_stride, const int32_t* row_index_dev);

/* ------------------------------------------------------------------ orthonormal constraint - */
/* ConstrainOrthonormalInternal (ref: nnet-utils.cc:914-1035; called for LinearComponent / AffineComponent /
 * TdnnComponent parameters by ConstrainOrthonormal, nnet-utils.cc:1037-1077, i.e. for the `linear` half of every
 * TDNN-F layer of the manual and derived systems; it does NOT cover TdnnDARTSV3Component).
 *   P = A A^T with A = M when rows <= cols and A = M^T otherwise (the reference's transposed copy, :1067-1074);
 *   scale < 0 ("floating"): scale^2 = tr(P P^T) / tr(P), update_speed 0.125 halved for ratio > 1.02 and again > 1.1;
 *   A <- A - 4 (update_speed / scale^2) (P - scale^2 I) A,   in place, no host synchronisation.
 * info_dev (optional, DEVICE float[4]) <- {scale used, ratio (0 when the scale is fixed), update_speed,
 * ||P - scale^2 I||_F}.  Where the reference asserts ratio > 0.999 the update is skipped and info_dev[1] shows why.
 * min(rows, cols) <= 512. */
int tdnnf_constrain_orthonormal(tdnnf_ctx* ctx, float* M, int rows, int cols, int stride, float scale,
                                float* info_dev);

/* ------------------------------------------------------------------ chain numerator ---- */
/* GenericNumeratorComputation (kaldi: chain/chain-generic-numerator.{h,cc}; SURVEY "next" row N3): log-domain
 * forward-backward over one small FST per sequence (the unconstrained / e2e supervision the recipes train
 * with, `--constrained false`).  Every arc consumes one frame and carries (pdf-id, log transition prob).
 * Host arrays, copied to the device:
 *   state_offsets[num_seqs+1]        states of sequence s are [state_offsets[s], state_offsets[s+1]); the first is the start state
 *   fwd_ranges / bwd_ranges          per global state, [begin,end) into the arc arrays (arcs leaving / entering the state)
 *   arc_state                        the OTHER end of the arc as a global state index (destination in the forward list, source in the backward list)
 *   final_logprob[num_states]        log final probability, -inf (<= -1e30) when not final */
typedef struct tdnnf_num_graph tdnnf_num_graph;
int tdnnf_num_graph_create(tdnnf_ctx* ctx, int num_seqs, const int32_t* state_offsets, int num_arcs,
                           const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* arc_logprob,
                           const int32_t* arc_pdf, const int32_t* arc_state, const float* final_logprob,
                           tdnnf_num_graph** out);
int tdnnf_num_graph_destroy(tdnnf_num_graph* g);
/* The next minibatch's numerator FSTs into the same handle (every NnetChainExample carries its own supervision): a pinned
 * staging buffer and asynchronous copies on the context's stream; no allocation or synchronisation while the new graphs
 * fit the capacity of the largest seen so far.  */
int tdnnf_num_graph_update(tdnnf_num_graph* g, int num_seqs, const int32_t* state_offsets, int num_arcs,
                           const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* arc_logprob,
                           const int32_t* arc_pdf, const int32_t* arc_state, const float* final_logprob);
/* total log-prob summed over sequences in *logprob (HOST; synchronises).  If nnet_output_deriv != NULL:
 * nnet_output_deriv[t*num_seqs+s, pdf] += deriv_weight * posterior.  *ok = 0 if any sequence has no path. */
int tdnnf_num_forward_backward(tdnnf_ctx* ctx, const tdnnf_num_graph* g, const float* nnet_output, int stride,
                               int frames_per_seq, float deriv_weight, float* nnet_output_deriv, int deriv_stride,
                               float* logprob, int* ok);

/* ------------------------------------------------------------------ chain objective --- */
/* ComputeChainObjfAndDeriv (kaldi: chain/chain-training.cc; called by NnetChainTrainer::ProcessOutputs) for the
 * unconstrained-egs supervision the recipes train with:
 *   weight = w * num_seqs * frames_per_seq;  deriv = 0;
 *   den = w * Denominator.Forward();  Denominator.Backward(-w, &deriv)                     (the denominator first)
 *   num = w * GenericNumerator: posteriors (x w) into xent_output_deriv when given (then added to deriv), else into deriv
 *   objf = num - den;  non-finite objf or a failed denominator / numerator check: derivs zeroed, objf = -10 * weight
 *   l2_term = -0.5 * w * l2_regularize * ||nnet_output||^2;  deriv += -w * l2_regularize * nnet_output   (when numerator ok)
 *   out-of-range penalty: tdnnf_penalize_out_of_range with limit 30 and scale 2 * out_of_range_regularize * oor_row_step
 *   on the rows oor_row_offset + k * oor_row_step (upstream penalises a sub-sampled set of rows; the caller draws the
 *   offset).  out_of_range_regularize = 0 switches it off.
 * nnet_output_deriv / xent_output_deriv may be NULL (objective only).  objf, l2_term, weight: HOST.  Synchronises. */
int tdnnf_chain_objf_and_deriv(tdnnf_ctx* ctx, tdnnf_den_comp* den, const tdnnf_num_graph* num, const float* nnet_output,
                               int stride, int num_seqs, int frames_per_seq, int num_pdfs, float supervision_weight,
                               float l2_regularize, float out_of_range_regularize, int oor_row_step, int oor_row_offset,
                               float* nnet_output_deriv, int deriv_stride, float* xent_output_deriv, int xent_stride,
                               float* objf, float* l2_term, float* weight);
/* PenalizeOutOfRange (kaldi: chain/chain-training.cc): on the rows r = row_offset + k * row_step (r < rows),
 *   deriv[r][c] -= scale * (x - limit) where x > limit,  deriv[r][c] -= scale * (x + limit) where x < -limit. */
int tdnnf_penalize_out_of_range(tdnnf_ctx* ctx, const float* nnet_output, int rows, int cols, int stride, float limit,
                                float scale, int row_step, int row_offset, float* deriv, int deriv_stride);

/* ------------------------------------------------------------------ graph readers (host only) - */
/* den.fst and the per-sequence numerator FSTs of an unconstrained Supervision in AT&T FSM text form (`fstprint`: one arc
 * per line "src dst ilabel olabel [weight]", one line "state [weight]" per final state, start state = source of the
 * first line, tropical weights = -log probability, ilabel = pdf-id + 1).  No CUDA device is needed to parse.
 * tdnnf_den_graph_parse_fst_text = DenominatorGraph::SetTransitions + SetInitialProbs (kaldi: chain/chain-den-graph.cc;
 * SURVEY.md B.2): forward transitions grouped by source state then backward transitions grouped by destination, and the
 * initial probabilities as the average over 100 steps of the state distribution from the start state (every state's
 * outgoing mass normalised together with its final probability, renormalised after each step). */
typedef struct tdnnf_host_graph tdnnf_host_graph;
typedef struct tdnnf_host_num_graph tdnnf_host_num_graph;
int tdnnf_den_graph_parse_fst_text(const char* text, uint64_t len, int num_pdfs, tdnnf_host_graph** out);
int tdnnf_host_graph_dims(const tdnnf_host_graph* g, int* num_states, int* num_pdfs, int* num_transitions);
int tdnnf_host_graph_arrays(const tdnnf_host_graph* g, const int32_t** fwd_ranges, const int32_t** bwd_ranges, const float** prob,
                            const int32_t** pdf, const int32_t** state, const float** initial_probs);
int tdnnf_host_graph_free(tdnnf_host_graph* g);
int tdnnf_den_graph_create_from_host(tdnnf_ctx* ctx, const tdnnf_host_graph* g, tdnnf_den_graph** out);
/* One FSM text per sequence -> the arrays of tdnnf_num_graph_create (each sequence's start state first; final weights
 * become log final probabilities, absent = not final). */
int tdnnf_num_graph_parse_fst_texts(const char* const* texts, const uint64_t* lens, int num_seqs, int num_pdfs,
                                    tdnnf_host_num_graph** out);
int tdnnf_host_num_graph_arrays(const tdnnf_host_num_graph* g, int* num_seqs, int* num_arcs, const int32_t** state_offsets,
                                const int32_t** fwd_ranges, const int32_t** bwd_ranges, const float** arc_logprob,
                                const int32_t** arc_pdf, const int32_t** arc_state, const float** final_logprob);
int tdnnf_host_num_graph_free(tdnnf_host_num_graph* g);
int tdnnf_num_graph_create_from_host(tdnnf_ctx* ctx, const tdnnf_host_num_graph* g, tdnnf_num_graph** out);

/* den.fst in the OpenFst binary form chain-make-den-fst writes (VectorFst<StdArc>, "vector"/"standard"; a
 * "compact_acceptor" is read too): same result as tdnnf_den_graph_parse_fst_text on `fstprint den.fst`. */
int tdnnf_den_graph_parse_fst_binary(const char* buf, uint64_t len, int num_pdfs, tdnnf_host_graph** out);

/* Training examples: a Kaldi table archive of NnetChainExample (`ark:cegs.N.ark`, binary, or `ark,t:` text) -> the
 * minibatch the step takes.  Replaces, for this path's inputs, kaldi: nnet3/nnet-chain-example.cc
 * NnetChainExample::Read / MergeChainExamples, nnet3/nnet-example.cc NnetIo::Read, chain/chain-supervision.cc
 * Supervision::Read (called by nnet3-chain-merge-egs | nnet3-chain-train in the recipes,
 * ref: steps/nnet3/chain/train.py via run_TDNN_DARTSV3_fbk_stride_pretrain.sh:181-209, `--constrained false` at :195); formats in csrc/egs_io.cc.
 * Host only; the archive is parsed from memory (the caller reads or maps the file); every count in the data is checked
 * against the buffer, a malformed archive is TDNNF_ERR_INVALID with the example key and offset in tdnnf_last_error().
 *   inputs: name, indexes (n, t, x per row) and the matrix expanded to dense fp32 whatever its on-disk coding;
 *   supervision: weight, num_sequences, frames_per_seq, label_dim, the FST(s) (arcs: src, dst, ilabel = pdf-id + 1;
 *   tropical weights), alignment pdfs, derivative weights.
 * tdnnf_chain_egs_merge_*: examples [first, first + count) as ONE minibatch, laid out as the kernels take it: row =
 * (rank of t within the sequence) * num_seqs + sequence (nnet3's t-major order with n fastest), sequences numbered in
 * example order.  merge_input: out == NULL returns the dimensions only.  merge_supervision: derivative weights
 * [frames_per_seq * num_seqs] in that order (1 where an example has none), the mean supervision weight, and the
 * per-sequence numerator graph (unconstrained examples: their <Fsts>; constrained examples of one sequence: their FST). */
typedef struct tdnnf_chain_egs tdnnf_chain_egs;
int tdnnf_chain_egs_read_ark(const char* buf, uint64_t len, int max_examples /* <= 0: all */, tdnnf_chain_egs** out);
int tdnnf_chain_egs_free(tdnnf_chain_egs* e);
int tdnnf_chain_egs_count(const tdnnf_chain_egs* e, int* num_examples);
int tdnnf_chain_egs_example(const tdnnf_chain_egs* e, int i, const char** key, int* binary, int* num_inputs, int* num_outputs);
int tdnnf_chain_egs_input(const tdnnf_chain_egs* e, int i, int j, const char** name, int* rows, int* cols,
                          const int32_t** indexes /* rows x 3 */, const float** data /* rows x cols */);
int tdnnf_chain_egs_supervision(const tdnnf_chain_egs* e, int i, int j, const char** name, float* weight, int* num_sequences,
                                int* frames_per_seq, int* label_dim, int* e2e, int* num_fsts, const int32_t** indexes,
                                int* num_deriv_weights, const float** deriv_weights, int* num_alignment_pdfs,
                                const int32_t** alignment_pdfs);
int tdnnf_chain_egs_fst(const tdnnf_chain_egs* e, int i, int j, int k, int* start, int* num_states, int* num_arcs,
                        const int32_t** arcs /* num_arcs x 3: src dst ilabel */, const float** arc_weights, int* num_finals,
                        const int32_t** final_states, const float** final_weights);
int tdnnf_chain_egs_merge_input(const tdnnf_chain_egs* e, int first, int count, const char* name, float* out, int64_t out_floats,
                                int* num_t, int* num_seqs, int* cols, int* first_t);
int tdnnf_chain_egs_merge_supervision(const tdnnf_chain_egs* e, int first, int count, const char* name, int num_pdfs,
                                      float* deriv_weights, int deriv_weights_floats, int* num_seqs, int* frames_per_seq,
                                      float* weight, tdnnf_host_num_graph** num_graph /* NULL: not built */);

/* ------------------------------------------------------------------ data-parallel reduction - */
/* SURVEY 8b capability (8): the deltas of the ranks' minibatch shards are SUMMED over NCCL (NVLink / NVSwitch): the
 * synchronous replacement of the multi-job `nnet3-average` (ref: steps/libs/nnet3/train/common.py:144-164; learning
 * rate x num_jobs then averaging == summing the per-job deltas).  NCCL is bound at run time (the libnccl already in
 * the process, else libnccl.so.2).  One process per GPU:
 *   rank 0: tdnnf_dp_unique_id(id)  -> the host ships the 128 bytes to the other ranks (MPI, a file, torch.distributed)
 *   all:    tdnnf_dp_comm_create(ctx, nranks, rank, id, &comm)           (collective: ncclCommInitRank)
 *   or      tdnnf_dp_comm_adopt(ctx, existing ncclComm_t, ...)           (the host set NCCL up itself)
 *   step:   tdnnf_dp_allreduce_deltas(comm, n, bufs, counts)             in place, on the context's stream, one NCCL group
 *   or      tdnnf_dp_allreduce_bucket_async(comm, buf, count) per bucket as the backward pass completes it (a side
 *           stream that first waits for the work queued on the context's stream) + tdnnf_dp_allreduce_wait(comm)
 *           before the parameter step. */
typedef struct tdnnf_dp_comm tdnnf_dp_comm;
int tdnnf_dp_unique_id(char* id_out, int id_bytes /* >= 128 */);
int tdnnf_dp_comm_create(tdnnf_ctx* ctx, int nranks, int rank, const char* unique_id, tdnnf_dp_comm** out);
int tdnnf_dp_comm_adopt(tdnnf_ctx* ctx, void* nccl_comm, int nranks, int rank, tdnnf_dp_comm** out);
int tdnnf_dp_comm_destroy(tdnnf_dp_comm* c);
int tdnnf_dp_allreduce_deltas(tdnnf_dp_comm* c, int n, float* const* bufs, const int64_t* counts);
int tdnnf_dp_allreduce_bucket_async(tdnnf_dp_comm* c, float* buf, int64_t count);
int tdnnf_dp_allreduce_wait(tdnnf_dp_comm* c);
int tdnnf_dp_nccl_version(int* version);

#ifdef __cplusplus
}
#endif
#endif /* TDNNF_NAS_B200_H_ */
