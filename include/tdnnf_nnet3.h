/*
 * tdnnf_nnet3.h -- C handle API over the nnet3 component mirror (tdnn-f_nas_b200/csrc/nnet3).
 *
 * What a non-C++ host needs to drive the reference's components the way NnetComputer does
 * (ref: SURVEY.md section 8b; kaldi nnet3/nnet-computation.cc kPropagate / kBackprop):
 * create from a config line or a model stream, PrecomputeIndexes, Propagate -> memo,
 * Backprop(memo, to_update), Write, the UpdatableComponent whole-parameter ops and the
 * `nnet3-copy --edits` directives.  A C++ host uses csrc/nnet3/components.h directly.
 *
 * Every function returns 0 on success; on failure the KALDI_ERR / KALDI_ASSERT / kernel message is
 * available from tdnnf_nnet3_last_error().  Strings and index arrays returned through `char**` /
 * `int32_t**` are malloc'd: release them with tdnnf_nnet3_free().  Indexes are (n, t, x) triples.
 */
#ifndef TDNNF_NNET3_H_
#define TDNNF_NNET3_H_

#include <stdint.h>

#include "tdnnf_nas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* tdnnf_nnet3_last_error(void);
void tdnnf_nnet3_free(void* p);

/* The "CuDevice" of the mirror: the context the components launch on (per host thread). */
int tdnnf_nnet3_set_context(tdnnf_ctx* ctx);
/* Parameter arena: between begin and end every device allocation of the component layer (the parameters of the
 * components created, read or copied meanwhile) is carved in call order, 256-byte aligned, from the caller-owned device
 * range [base, base + bytes) (base 256-byte aligned); *used <- bytes consumed.  Creating the delta components of a network
 * inside one arena makes all deltas ONE contiguous range: tdnnf_dp_allreduce_deltas then needs a single buffer
 * (replaces the per-component nnet3-average of steps/libs/nnet3/train/common.py:144-164). */
int tdnnf_nnet3_arena_begin(void* base, uint64_t bytes);
int tdnnf_nnet3_arena_end(uint64_t* used);
/* Counter-based host RNG behind SetRandUniform / RandInt: same seed + counter => same draws on every rank. */
int tdnnf_nnet3_set_rand_seed(uint64_t seed);
int tdnnf_nnet3_set_rand_counter(uint64_t counter);
uint64_t tdnnf_nnet3_get_rand_counter(void);
float tdnnf_nnet3_rand_uniform(void);
int tdnnf_nnet3_rand_int(int lo, int hi); /* the RandInt the components use (advances the same counter) */
/* Data-parallel world size: the FLOPs penalty is normalised by rows * world_size. */
int tdnnf_nnet3_set_dp_world_size(int world_size);
/* TdnnDARTSV3Component::Propagate keeps the bf16 operand planes of its input in the memo and Backprop reuses them (default
 * 1): the input is split once per minibatch, not twice.  0 restores a split per call. */
int tdnnf_nnet3_set_keep_planes(int enable);
/* Re-enable the reference's per-minibatch "log_alpha" stdout print (ref: tdnn.cc:571, simple.cc:2640). */
int tdnnf_nnet3_set_print_log_alpha(int enable);
/* Parameter-gradient GEMM of TdnnDARTSV3Component::Backprop with one fp16 product (tdnnf_ctx_set_gradient_mode).
 * Default 0: the natural-gradient projection amplifies its 2.9e-4 error beyond the 1e-3 tolerance (measured). */
int tdnnf_nnet3_set_fast_gradients(int enable);

/* Component::NewComponentOfType + InitFromConfig (ref: itf.cc:126-293, tdnn.cc:109-212, ...). */
int tdnnf_nnet3_component_new(const char* type, const char* config_line, void** out);
int tdnnf_nnet3_tdnn_darts_for_indexing(const int32_t* time_offsets, int n, void** out);
/* Component::ReadNew / Write on memory buffers, text or binary (ref: itf.cc:106-124, tdnn.cc:659-761, ...). */
int tdnnf_nnet3_component_read(const char* data, uint64_t len, int binary, void** out);
/* tdnnf_nnet3_component_read that also reports the bytes consumed (to walk the component list of a raw nnet3 model). */
int tdnnf_nnet3_component_read_ex(const char* data, uint64_t len, int binary, void** out, uint64_t* consumed);
int tdnnf_nnet3_component_write(const void* comp, int binary, char** out, uint64_t* len);
int tdnnf_nnet3_component_copy(const void* comp, void** out);
int tdnnf_nnet3_component_delete(void* comp);
int tdnnf_nnet3_component_info(const void* comp, char** out);
int tdnnf_nnet3_component_type(const void* comp, char** out);
int tdnnf_nnet3_component_dims(const void* comp, int* input_dim, int* output_dim, int* properties);

/* Index methods (ref: tdnn.cc:628-657, 763-905). */
int tdnnf_nnet3_precompute_indexes(const void* comp, const int32_t* in_idx, int n_in, const int32_t* out_idx, int n_out,
                                   int need_backprop, void** out);
int tdnnf_nnet3_indexes_delete(void* idx);
int tdnnf_nnet3_indexes_write(const void* idx, int binary, char** out, uint64_t* len);
int tdnnf_nnet3_indexes_read(const char* data, uint64_t len, int binary, void** out);
int tdnnf_nnet3_reorder_indexes(const void* comp, const int32_t* in_idx, int n_in, const int32_t* out_idx, int n_out,
                                int32_t** new_in, int* new_n_in, int32_t** new_out, int* new_n_out);
int tdnnf_nnet3_get_input_indexes(const void* comp, int n, int t, int x, int32_t** out, int* n_out);
int tdnnf_nnet3_is_computable(const void* comp, int n, int t, int x, const int32_t* avail, int n_avail, int* result);

/* Propagate / Backprop / DeleteMemo on device matrices given as (pointer, rows, cols, stride). */
int tdnnf_nnet3_propagate(const void* comp, const void* indexes, const float* in, int in_rows, int in_cols, int in_stride,
                          float* out, int out_rows, int out_cols, int out_stride, void** memo);
int tdnnf_nnet3_backprop(const void* comp, const void* indexes, const float* in_value, int in_rows, int in_cols,
                         int in_stride, const float* out_value, int ov_stride, const float* out_deriv, int out_rows,
                         int out_cols, int od_stride, void* memo, void* to_update, float* in_deriv, int id_stride);
int tdnnf_nnet3_delete_memo(const void* comp, void* memo);
/* Component::StoreStats / ZeroStats (components with kStoresStats: BatchNormComponent in training mode, ref
 * nnet-normalize-component.cc:551-589, 668-678); in_value may be NULL.  tdnnf_nnet3_bn_count: its frame count. */
int tdnnf_nnet3_store_stats(void* comp, const float* in_value, int in_rows, int in_cols, int in_stride, const float* out_value,
                            int out_rows, int out_cols, int ov_stride, void* memo);
int tdnnf_nnet3_zero_stats(void* comp);
int tdnnf_nnet3_bn_count(const void* comp, double* count);

/* UpdatableComponent surface (ref: tdnn.cc:907-979, simple.cc:9606-9681, itf.cc:313-431). */
int tdnnf_nnet3_scale(void* comp, float scale);
int tdnnf_nnet3_add(void* comp, float alpha, const void* other);
int tdnnf_nnet3_dot_product(void* comp, const void* other, float* result);
int tdnnf_nnet3_num_parameters(void* comp, int* n);
int tdnnf_nnet3_vectorize(void* comp, float* params, int n);
int tdnnf_nnet3_unvectorize(void* comp, const float* params, int n);
int tdnnf_nnet3_perturb_params(void* comp, float stddev);
int tdnnf_nnet3_set_learning_rate(void* comp, float underlying_lrate);
int tdnnf_nnet3_set_actual_learning_rate(void* comp, float lrate);
int tdnnf_nnet3_get_learning_rate(void* comp, float* lrate);
int tdnnf_nnet3_set_test_mode(void* comp, int test_mode);
int tdnnf_nnet3_temp_proportion(const void* comp, float* value);
int tdnnf_nnet3_dropout_proportion(const void* comp, float* value); /* GeneralDropoutComponent */
/* ConstrainOrthonormal(Nnet*) (ref: nnet-utils.cc:1037-1077), to be called after every minibatch: each component of
 * the list that is a TdnnComponent with orthonormal-constraint != 0 is updated with probability 1/4 (RandInt(0,3),
 * one draw per constrained component, list order); *num_updated (optional) <- how many were.  Other types are skipped. */
int tdnnf_nnet3_constrain_orthonormal(void* const* comps, int n, int* num_updated);
int tdnnf_nnet3_orthonormal_constraint(const void* comp, float* value);
/* Device parameter buffers (<= 2) of an updatable component, for the all-reduce of the deltas. */
int tdnnf_nnet3_param_buffers(void* comp, float** ptrs, int* rows, int* cols, int* strides, int* count);
int tdnnf_nnet3_bn_test_set_stats(void* comp, int dim, int block_dim, float epsilon, float target_rms, double count,
                                  const double* sum, const double* sumsq);

/* Device pointers of BatchNormTestComponent's derived scale_ / offset_ vectors (ref: norm.cc:680-713). */
int tdnnf_nnet3_bn_test_scale_offset(const void* comp, const float** scale, const float** offset, int* dim);

/* OnlineNaturalGradient (kaldi: nnet3/natural-gradient-online.h) as a standalone object: what the components
 * own as preconditioner_in_ / preconditioner_out_ / preconditioner_ (ref: conv.h:324-327, simple.h:2790). */
int tdnnf_nnet3_ng_new(int rank, int update_period, float num_samples_history, float alpha, void** out);
int tdnnf_nnet3_ng_delete(void* ng);
int tdnnf_nnet3_ng_freeze(void* ng, int frozen);
/* PreconditionDirections(&X, &scale): X (device, rows x cols) is overwritten; *scale on the host (one sync). */
int tdnnf_nnet3_ng_precondition(void* ng, float* x, int rows, int cols, int stride, float* scale);
/* State read-back: t, rank, dim, rho; d (host, rank floats) and W (host, rank x dim, dense) may be NULL. */
int tdnnf_nnet3_ng_state(void* ng, int* t, int* rank, int* dim, float* rho, float* d, float* W, int* num_reorthogonalized);
/* The preconditioners of a TdnnDARTSV3Component (which = 0: preconditioner_in_, 1: preconditioner_out_) or of an
 * Onehot/ConstantFunction component (which = 0); the pointer is owned by the component. */
int tdnnf_nnet3_component_ng(void* comp, int which, void** ng);
/* Diagnostic switch: with 1, every PreconditionDirections call is the identity with scale 1, i.e. the components
 * accumulate the un-preconditioned gradient (what the first-level parity tests pin). */
int tdnnf_nnet3_set_ng_identity(int enable);
/* Host only (no GPU): the symmetric eigen-solver used by OnlineNaturalGradient's update (Householder tridiagonalisation +
 * implicit QL, double).  a: n x n row-major symmetric; vals[n]; vecs n x n row-major with eigenvector k in COLUMN k. */
int tdnnf_nnet3_symmetric_eigen(const double* a, int n, double* vals, double* vecs);

/* ReadEditConfig subset: set-temperature-proportion, set-learning-rate{,-factor} (ref: utils.cc:1166-1415). */
int tdnnf_nnet3_apply_edits(const char* edits, const char** names, void** comps, int n);

#ifdef __cplusplus
}
#endif
#endif /* TDNNF_NNET3_H_ */
