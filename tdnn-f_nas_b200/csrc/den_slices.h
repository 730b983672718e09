// Internal interface of the sequence-slice denominator kernels (den_slices.cu), used by den.cu.
#pragma once
#include <vector>

#include "context.h"

struct tdnnf_den_slices;

namespace tdnnf {

// *out == nullptr (and TDNNF_OK) when the slice path does not apply to this shape: num_seqs not a multiple of 8,
// 8 x num_pdfs floats beyond the shared memory of an SM, too few slices to fill the chip, or TDNNF_DEN_PATH=frames.
int den_slices_create(tdnnf_ctx* ctx, int N, int P, int S, int T, const std::vector<int>& fwd_ranges,
                      const std::vector<int>& bwd_ranges, const std::vector<float>& prob, const std::vector<int>& pdf,
                      const std::vector<int>& state, const std::vector<float>& init, tdnnf_den_slices** out);
void den_slices_destroy(tdnnf_den_slices* s);
void den_slices_describe(const tdnnf_den_slices* s, int* cluster, int* parts, int* ctas);
// exp + layout change, alpha(0), the whole forward recursion; fills tot[(T+1)][S] (zero-initialised by the caller is not needed).
int den_slices_forward(tdnnf_ctx* ctx, tdnnf_den_slices* s, const float* nnet_output, int stride, const float* init,
                       float init_sum, int N, int P, int S, int T, float leaky, float* tot);
// the whole backward recursion + nnet_output_deriv += deriv_weight * posteriors; *check (device double, zeroed by the caller)
// receives sum_h,s alpha'(0,h,s) betad(0,h,s).
int den_slices_backward(tdnnf_ctx* ctx, tdnnf_den_slices* s, const float* init, float init_sum, int N, int P, int S, int T,
                        float leaky, const float* tot, const float* tot_prob, float deriv_weight, float* nnet_output_deriv,
                        int stride, double* check);

}  // namespace tdnnf
