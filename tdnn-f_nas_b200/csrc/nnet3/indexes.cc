// time_height_convolution::GetComputationIo / GetIndexesForComputation (kaldi: nnet3/convolution.cc;
// upstream Kaldi, not shipped with the reference; semantics per SURVEY.md Appendix B.5).  Integer
// work: results must be bit-exact against the independent Python restatement in tests/.
#include <algorithm>

#include "shim.h"

namespace tdnnf {
namespace nnet3 {
namespace time_height_convolution {

// Sorted unique (n, x) pairs.
static void GetNxList(const std::vector<Index>& indexes, std::vector<std::pair<int32, int32> >* pairs) {
  pairs->clear();
  pairs->reserve(indexes.size());
  for (const Index& i : indexes) pairs->push_back(std::make_pair(i.n, i.x));
  std::sort(pairs->begin(), pairs->end());
  pairs->erase(std::unique(pairs->begin(), pairs->end()), pairs->end());
}

// Sorted unique t values, kNoTime excluded.
static void GetTList(const std::vector<Index>& indexes, std::vector<int32>* t_values) {
  t_values->clear();
  for (const Index& i : indexes)
    if (i.t != kNoTime) t_values->push_back(i.t);
  std::sort(t_values->begin(), t_values->end());
  t_values->erase(std::unique(t_values->begin(), t_values->end()), t_values->end());
}

static int32 Gcd(int32 m, int32 n) {
  if (m == 0 || n == 0) return m == 0 ? (n > 0 ? n : -n) : (m > 0 ? m : -m);
  while (true) {
    m %= n;
    if (m == 0) return n > 0 ? n : -n;
    n %= m;
    if (n == 0) return m > 0 ? m : -m;
  }
}

// (step, num) such that the t values lie on start + k*step, k < num.  step == 0 iff one value.
static void RegularizeTList(const std::vector<int32>& t_values, int32* step, int32* num_t_values) {
  KALDI_ASSERT(!t_values.empty());
  int32 gcd = 0;
  for (size_t i = 1; i < t_values.size(); ++i) gcd = Gcd(gcd, t_values[i] - t_values[i - 1]);
  *step = gcd;
  if (gcd == 0) {
    *num_t_values = (int32)t_values.size();
  } else {
    const int32 t_span = t_values.back() - t_values.front();
    *num_t_values = 1 + t_span / gcd;
  }
}

void GetComputationIo(const std::vector<Index>& input_indexes, const std::vector<Index>& output_indexes,
                      ConvolutionComputationIo* io) {
  std::vector<std::pair<int32, int32> > n_x_pairs;
  GetNxList(input_indexes, &n_x_pairs);
  KALDI_ASSERT(!n_x_pairs.empty());
  io->num_images = (int32)n_x_pairs.size();
  std::vector<int32> t_values;
  GetTList(input_indexes, &t_values);
  RegularizeTList(t_values, &(io->t_step_in), &(io->num_t_in));
  io->start_t_in = t_values[0];
  t_values.clear();
  GetTList(output_indexes, &t_values);
  RegularizeTList(t_values, &(io->t_step_out), &(io->num_t_out));
  io->start_t_out = t_values[0];
  io->reorder_t_in = 1;
}

// t-major blocks of `reorder_t` consecutive t's; inside a block the (n,x) pairs have stride reorder_t
// and t is fastest.
static void CreateIndexes(const std::vector<std::pair<int32, int32> >& n_x_pairs, int32 start_t, int32 t_stride,
                          int32 num_t_values, int32 reorder_t, std::vector<Index>* indexes) {
  KALDI_ASSERT(reorder_t >= 1 && num_t_values % reorder_t == 0 && t_stride >= 0);
  const int32 num_n_x_pairs = (int32)n_x_pairs.size();
  indexes->clear();
  indexes->reserve((size_t)num_n_x_pairs * num_t_values);
  if (t_stride == 0) {  // a single t value (num_t_values == reorder_t == 1)
    KALDI_ASSERT(num_t_values == 1);
    for (int32 nx = 0; nx < num_n_x_pairs; ++nx) indexes->push_back(Index(n_x_pairs[nx].first, start_t, n_x_pairs[nx].second));
    return;
  }
  const int32 outer_t_stride = t_stride * reorder_t;
  const int32 t_end = start_t + num_t_values * t_stride;
  Index index;
  for (int32 t_block = start_t; t_block < t_end; t_block += outer_t_stride) {
    for (int32 nx = 0; nx < num_n_x_pairs; ++nx) {
      index.n = n_x_pairs[nx].first;
      index.x = n_x_pairs[nx].second;
      for (int32 t = t_block; t < t_block + outer_t_stride; t += t_stride) {
        index.t = t;
        indexes->push_back(index);
      }
    }
  }
}

static void SetSomeIndexesBlank(const std::vector<Index>& ref_indexes, std::vector<Index>* indexes) {
  std::unordered_set<Index, IndexHasher> ref_set(ref_indexes.begin(), ref_indexes.end());
  for (Index& i : *indexes)
    if (ref_set.count(i) == 0) i.t = kNoTime;
}

void GetIndexesForComputation(const ConvolutionComputationIo& io, const std::vector<Index>& orig_input_indexes,
                              const std::vector<Index>& orig_output_indexes, std::vector<Index>* input_indexes,
                              std::vector<Index>* output_indexes) {
  std::vector<std::pair<int32, int32> > n_x_pairs;
  GetNxList(orig_input_indexes, &n_x_pairs);
  KALDI_ASSERT((int32)n_x_pairs.size() == io.num_images);
  CreateIndexes(n_x_pairs, io.start_t_in, io.t_step_in, io.num_t_in, io.reorder_t_in, input_indexes);
  SetSomeIndexesBlank(orig_input_indexes, input_indexes);
  CreateIndexes(n_x_pairs, io.start_t_out, io.t_step_out, io.num_t_out, 1, output_indexes);
  SetSomeIndexesBlank(orig_output_indexes, output_indexes);
}

}  // namespace time_height_convolution
}  // namespace nnet3
}  // namespace tdnnf
