// extern "C" handle API over the nnet3 component mirror, so that non-C++ hosts (the Python parity
// tests and bench.py) can drive the components exactly like NnetComputer would: create from a
// config line / read from a model stream, PrecomputeIndexes, Propagate, Backprop(to_update), Write.
// Errors (KALDI_ERR / KALDI_ASSERT / kernel failures) are caught at the boundary and reported as a
// non-zero status plus tdnnf_nnet3_last_error().
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "components.h"
#include "tdnnf_nnet3.h"

using namespace tdnnf::nnet3;

static thread_local std::string g_err;

#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const std::exception& e) {               \
    g_err = e.what();                             \
    return 1;                                     \
  }                                               \
  catch (...) {                                   \
    g_err = "unknown exception";                  \
    return 1;                                     \
  }                                               \
  return 0;

static char* DupString(const std::string& s, uint64_t* len) {
  char* p = static_cast<char*>(std::malloc(s.size() + 1));
  std::memcpy(p, s.data(), s.size());
  p[s.size()] = '\0';
  if (len) *len = s.size();
  return p;
}
static CuMatrixBase<BaseFloat> View(const float* p, int rows, int cols, int stride) {
  return CuMatrixBase<BaseFloat>(const_cast<float*>(p), rows, cols, stride);
}
static std::vector<Index> ToIndexes(const int32_t* nt_x, int n) {
  std::vector<Index> v(n);
  for (int i = 0; i < n; ++i) v[i] = Index(nt_x[3 * i], nt_x[3 * i + 1], nt_x[3 * i + 2]);
  return v;
}

extern "C" {

const char* tdnnf_nnet3_last_error(void) { return g_err.c_str(); }
void tdnnf_nnet3_free(void* p) { std::free(p); }

int tdnnf_nnet3_arena_begin(void* base, uint64_t bytes) { API_BEGIN DeviceArenaBegin(base, (size_t)bytes); API_END }
int tdnnf_nnet3_arena_end(uint64_t* used) {
  API_BEGIN
  const size_t u = DeviceArenaEnd();
  if (used) *used = u;
  API_END
}
int tdnnf_nnet3_set_context(tdnnf_ctx* ctx) { API_BEGIN SetCurrentContext(ctx); API_END }
int tdnnf_nnet3_set_rand_seed(uint64_t seed) { API_BEGIN SetRandSeed(seed); API_END }
int tdnnf_nnet3_set_rand_counter(uint64_t c) { API_BEGIN SetRandCounter(c); API_END }
uint64_t tdnnf_nnet3_get_rand_counter(void) { return GetRandCounter(); }
float tdnnf_nnet3_rand_uniform(void) { return RandUniformOpen(); }
int tdnnf_nnet3_rand_int(int lo, int hi) { return RandInt(lo, hi); }
int tdnnf_nnet3_set_dp_world_size(int g) { API_BEGIN SetDataParallelWorldSize(g); API_END }
int tdnnf_nnet3_set_keep_planes(int b) { API_BEGIN SetKeepPlanes(b != 0); API_END }
int tdnnf_nnet3_set_print_log_alpha(int b) { API_BEGIN SetPrintLogAlpha(b != 0); API_END }
int tdnnf_nnet3_set_fast_gradients(int b) { API_BEGIN SetFastGradients(b != 0); API_END }
int tdnnf_nnet3_set_ng_identity(int b) { API_BEGIN SetNaturalGradientIdentity(b != 0); API_END }
int tdnnf_nnet3_symmetric_eigen(const double* a, int n, double* vals, double* vecs) {
  API_BEGIN
  if (!SymmetricEigenForTest(a, n, vals, vecs)) KALDI_ERR << "symmetric eigen-solver did not converge";
  API_END
}

int tdnnf_nnet3_ng_new(int rank, int update_period, float num_samples_history, float alpha, void** out) {
  API_BEGIN
  OnlineNaturalGradient* ng = new OnlineNaturalGradient();
  ng->SetRank(rank);
  ng->SetUpdatePeriod(update_period);
  ng->SetNumSamplesHistory(num_samples_history);
  ng->SetAlpha(alpha);
  *out = ng;
  API_END
}
int tdnnf_nnet3_ng_delete(void* ng) { API_BEGIN delete static_cast<OnlineNaturalGradient*>(ng); API_END }
int tdnnf_nnet3_ng_freeze(void* ng, int frozen) { API_BEGIN static_cast<OnlineNaturalGradient*>(ng)->Freeze(frozen != 0); API_END }
int tdnnf_nnet3_ng_precondition(void* ng, float* x, int rows, int cols, int stride, float* scale) {
  API_BEGIN
  CuSubMatrix<BaseFloat> X(x, rows, cols, stride);
  BaseFloat s = 1.0;
  static_cast<OnlineNaturalGradient*>(ng)->PreconditionDirections(&X, &s);
  if (scale) *scale = s;
  API_END
}
int tdnnf_nnet3_ng_state(void* ng_in, int* t, int* rank, int* dim, float* rho, float* d, float* W, int* num_reorth) {
  API_BEGIN
  OnlineNaturalGradient* ng = static_cast<OnlineNaturalGradient*>(ng_in);
  int32 tt;
  BaseFloat r;
  std::vector<BaseFloat> dv;
  Matrix<BaseFloat> Wm;
  ng->GetState(&tt, &r, &dv, (W || dim) ? &Wm : NULL);
  if (t) *t = tt;
  if (rank) *rank = ng->GetRank();
  if (dim) *dim = Wm.cols;
  if (rho) *rho = r;
  if (d) std::copy(dv.begin(), dv.end(), d);
  if (W) std::copy(Wm.v.begin(), Wm.v.end(), W);
  if (num_reorth) *num_reorth = ng->NumReorthogonalized();
  API_END
}
int tdnnf_nnet3_component_ng(void* comp, int which, void** ng) {
  API_BEGIN
  Component* c = static_cast<Component*>(comp);
  if (TdnnDARTSV3Component* t = dynamic_cast<TdnnDARTSV3Component*>(c)) {
    *ng = which == 0 ? &t->PreconditionerIn() : &t->PreconditionerOut();
  } else if (VectorFunctionComponentBase* v = dynamic_cast<VectorFunctionComponentBase*>(c)) {
    *ng = &v->Preconditioner();
  } else {
    KALDI_ERR << "component " << c->Type() << " has no natural-gradient preconditioner";
  }
  API_END
}

// NewComponentOfType + InitFromConfig ("key=value ..." line as in an nnet3 config file)
int tdnnf_nnet3_component_new(const char* type, const char* config_line, void** out) {
  API_BEGIN
  Component* c = Component::NewComponentOfType(type);
  if (!c) KALDI_ERR << "Unknown component type " << type;
  std::unique_ptr<Component> holder(c);
  ConfigLine cfl;
  if (!cfl.ParseLine(std::string(type) + " " + config_line)) KALDI_ERR << "Invalid config line: " << config_line;
  c->InitFromConfig(&cfl);
  *out = holder.release();
  API_END
}
// Index-only TdnnDARTSV3Component (no parameters, no device): for ReorderIndexes / PrecomputeIndexes / ...
int tdnnf_nnet3_tdnn_darts_for_indexing(const int32_t* time_offsets, int n, void** out) {
  API_BEGIN
  *out = TdnnDARTSV3Component::NewForIndexing(std::vector<int32>(time_offsets, time_offsets + n));
  API_END
}
// Component::ReadNew from a memory buffer (text or binary model fragment starting at the <Type> token)
int tdnnf_nnet3_component_read(const char* data, uint64_t len, int binary, void** out) {
  API_BEGIN
  std::istringstream is(std::string(data, len));
  *out = Component::ReadNew(is, binary != 0);
  API_END
}
// The same, reporting how many bytes of the buffer the component occupied: lets a host walk the component list of a
// raw nnet3 model (Nnet::Read: "<ComponentName> name <Type> ... </Type>" repeated).
int tdnnf_nnet3_component_read_ex(const char* data, uint64_t len, int binary, void** out, uint64_t* consumed) {
  API_BEGIN
  std::istringstream is(std::string(data, len));
  *out = Component::ReadNew(is, binary != 0);
  if (consumed) {
    const std::streampos pos = is.tellg();
    *consumed = (pos == std::streampos(-1)) ? len : (uint64_t)pos;
  }
  API_END
}
int tdnnf_nnet3_component_write(const void* comp, int binary, char** out, uint64_t* len) {
  API_BEGIN
  std::ostringstream os;
  static_cast<const Component*>(comp)->Write(os, binary != 0);
  *out = DupString(os.str(), len);
  API_END
}
int tdnnf_nnet3_component_copy(const void* comp, void** out) { API_BEGIN *out = static_cast<const Component*>(comp)->Copy(); API_END }
int tdnnf_nnet3_component_delete(void* comp) { API_BEGIN delete static_cast<Component*>(comp); API_END }
int tdnnf_nnet3_component_info(const void* comp, char** out) {
  API_BEGIN *out = DupString(static_cast<const Component*>(comp)->Info(), nullptr); API_END
}
int tdnnf_nnet3_component_type(const void* comp, char** out) {
  API_BEGIN *out = DupString(static_cast<const Component*>(comp)->Type(), nullptr); API_END
}
int tdnnf_nnet3_component_dims(const void* comp, int* input_dim, int* output_dim, int* properties) {
  API_BEGIN
  const Component* c = static_cast<const Component*>(comp);
  *input_dim = c->InputDim();
  *output_dim = c->OutputDim();
  *properties = c->Properties();
  API_END
}

// indexes are (n, t, x) triples
int tdnnf_nnet3_precompute_indexes(const void* comp, const int32_t* in_idx, int n_in, const int32_t* out_idx, int n_out,
                                   int need_backprop, void** out) {
  API_BEGIN
  MiscComputationInfo misc;
  *out = static_cast<const Component*>(comp)->PrecomputeIndexes(misc, ToIndexes(in_idx, n_in), ToIndexes(out_idx, n_out),
                                                                need_backprop != 0);
  API_END
}
int tdnnf_nnet3_indexes_delete(void* idx) { API_BEGIN delete static_cast<ComponentPrecomputedIndexes*>(idx); API_END }
int tdnnf_nnet3_indexes_write(const void* idx, int binary, char** out, uint64_t* len) {
  API_BEGIN
  std::ostringstream os;
  static_cast<const ComponentPrecomputedIndexes*>(idx)->Write(os, binary != 0);
  *out = DupString(os.str(), len);
  API_END
}
int tdnnf_nnet3_indexes_read(const char* data, uint64_t len, int binary, void** out) {
  API_BEGIN
  std::istringstream is(std::string(data, len));
  *out = ComponentPrecomputedIndexes::ReadNew(is, binary != 0);
  API_END
}
// ReorderIndexes: arrays are resized by the callee (malloc'd, free with tdnnf_nnet3_free)
int tdnnf_nnet3_reorder_indexes(const void* comp, const int32_t* in_idx, int n_in, const int32_t* out_idx, int n_out,
                                int32_t** new_in, int* new_n_in, int32_t** new_out, int* new_n_out) {
  API_BEGIN
  std::vector<Index> in = ToIndexes(in_idx, n_in), out = ToIndexes(out_idx, n_out);
  static_cast<const Component*>(comp)->ReorderIndexes(&in, &out);
  auto dump = [](const std::vector<Index>& v, int32_t** p, int* n) {
    *n = (int)v.size();
    *p = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 3 * (v.size() + 1)));
    for (size_t i = 0; i < v.size(); ++i) { (*p)[3 * i] = v[i].n; (*p)[3 * i + 1] = v[i].t; (*p)[3 * i + 2] = v[i].x; }
  };
  dump(in, new_in, new_n_in);
  dump(out, new_out, new_n_out);
  API_END
}
int tdnnf_nnet3_get_input_indexes(const void* comp, int n, int t, int x, int32_t** out, int* n_out) {
  API_BEGIN
  std::vector<Index> v;
  MiscComputationInfo misc;
  static_cast<const Component*>(comp)->GetInputIndexes(misc, Index(n, t, x), &v);
  *n_out = (int)v.size();
  *out = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 3 * (v.size() + 1)));
  for (size_t i = 0; i < v.size(); ++i) { (*out)[3 * i] = v[i].n; (*out)[3 * i + 1] = v[i].t; (*out)[3 * i + 2] = v[i].x; }
  API_END
}
int tdnnf_nnet3_is_computable(const void* comp, int n, int t, int x, const int32_t* avail, int n_avail, int* result) {
  API_BEGIN
  IndexSet set(ToIndexes(avail, n_avail));
  MiscComputationInfo misc;
  std::vector<Index> used;
  *result = static_cast<const Component*>(comp)->IsComputable(misc, Index(n, t, x), set, &used) ? 1 : 0;
  API_END
}

int tdnnf_nnet3_propagate(const void* comp, const void* indexes, const float* in, int in_rows, int in_cols, int in_stride,
                          float* out, int out_rows, int out_cols, int out_stride, void** memo) {
  API_BEGIN
  CuMatrixBase<BaseFloat> in_v = View(in, in_rows, in_cols, in_stride), out_v = View(out, out_rows, out_cols, out_stride);
  void* m = static_cast<const Component*>(comp)->Propagate(static_cast<const ComponentPrecomputedIndexes*>(indexes), in_v, &out_v);
  if (memo) *memo = m;
  else static_cast<const Component*>(comp)->DeleteMemo(m);
  API_END
}
// Any of in_value / out_value / in_deriv may be NULL when the component's Properties() say it is not needed.
int tdnnf_nnet3_backprop(const void* comp, const void* indexes, const float* in_value, int in_rows, int in_cols,
                         int in_stride, const float* out_value, int ov_stride, const float* out_deriv, int out_rows,
                         int out_cols, int od_stride, void* memo, void* to_update, float* in_deriv, int id_stride) {
  API_BEGIN
  CuMatrixBase<BaseFloat> in_v = View(in_value, in_value ? in_rows : 0, in_value ? in_cols : 0, in_stride),
                          ov_v = View(out_value, out_value ? out_rows : 0, out_value ? out_cols : 0, ov_stride),
                          od_v = View(out_deriv, out_rows, out_cols, od_stride),
                          id_v = View(in_deriv, in_rows, in_cols, id_stride);
  static_cast<const Component*>(comp)->Backprop("", static_cast<const ComponentPrecomputedIndexes*>(indexes), in_v, ov_v, od_v,
                                                memo, static_cast<Component*>(to_update), in_deriv ? &id_v : NULL);
  API_END
}
int tdnnf_nnet3_delete_memo(const void* comp, void* memo) { API_BEGIN static_cast<const Component*>(comp)->DeleteMemo(memo); API_END }
// Component::StoreStats / ZeroStats (kStoresStats components: BatchNormComponent in training mode)
int tdnnf_nnet3_store_stats(void* comp, const float* in_value, int in_rows, int in_cols, int in_stride, const float* out_value,
                            int out_rows, int out_cols, int ov_stride, void* memo) {
  API_BEGIN
  CuMatrixBase<BaseFloat> in_v = View(in_value, in_value ? in_rows : 0, in_value ? in_cols : 0, in_stride),
                          ov_v = View(out_value, out_rows, out_cols, ov_stride);
  static_cast<Component*>(comp)->StoreStats(in_v, ov_v, memo);
  API_END
}
int tdnnf_nnet3_zero_stats(void* comp) { API_BEGIN static_cast<Component*>(comp)->ZeroStats(); API_END }
int tdnnf_nnet3_bn_count(const void* comp, double* count) {
  API_BEGIN
  const BatchNormComponent* b = dynamic_cast<const BatchNormComponent*>(static_cast<const Component*>(comp));
  if (!b) KALDI_ERR << "not a BatchNormComponent";
  *count = b->Count();
  API_END
}

// ---- UpdatableComponent surface
static UpdatableComponent* Upd(void* comp) {
  UpdatableComponent* u = dynamic_cast<UpdatableComponent*>(static_cast<Component*>(comp));
  if (!u) KALDI_ERR << "component is not updatable";
  return u;
}
int tdnnf_nnet3_scale(void* comp, float scale) { API_BEGIN static_cast<Component*>(comp)->Scale(scale); API_END }
int tdnnf_nnet3_add(void* comp, float alpha, const void* other) {
  API_BEGIN static_cast<Component*>(comp)->Add(alpha, *static_cast<const Component*>(other)); API_END
}
int tdnnf_nnet3_dot_product(void* comp, const void* other, float* result) {
  API_BEGIN
  const UpdatableComponent* o = dynamic_cast<const UpdatableComponent*>(static_cast<const Component*>(other));
  if (!o) KALDI_ERR << "component is not updatable";
  *result = Upd(comp)->DotProduct(*o);
  API_END
}
int tdnnf_nnet3_num_parameters(void* comp, int* n) { API_BEGIN *n = Upd(comp)->NumParameters(); API_END }
int tdnnf_nnet3_vectorize(void* comp, float* params, int n) {
  API_BEGIN
  std::vector<BaseFloat> v;
  Upd(comp)->Vectorize(&v);
  if ((int)v.size() != n) KALDI_ERR << "Vectorize: expected " << v.size() << " parameters, buffer has " << n;
  std::memcpy(params, v.data(), sizeof(float) * n);
  API_END
}
int tdnnf_nnet3_unvectorize(void* comp, const float* params, int n) {
  API_BEGIN Upd(comp)->UnVectorize(std::vector<BaseFloat>(params, params + n)); API_END
}
int tdnnf_nnet3_perturb_params(void* comp, float stddev) { API_BEGIN Upd(comp)->PerturbParams(stddev); API_END }
int tdnnf_nnet3_set_learning_rate(void* comp, float underlying_lrate) { API_BEGIN Upd(comp)->SetUnderlyingLearningRate(underlying_lrate); API_END }
int tdnnf_nnet3_set_actual_learning_rate(void* comp, float lrate) { API_BEGIN Upd(comp)->SetActualLearningRate(lrate); API_END }
int tdnnf_nnet3_get_learning_rate(void* comp, float* lrate) { API_BEGIN *lrate = Upd(comp)->LearningRate(); API_END }
int tdnnf_nnet3_set_test_mode(void* comp, int test_mode) {
  API_BEGIN
  Component* c = static_cast<Component*>(comp);
  if (BatchNormComponent* bn = dynamic_cast<BatchNormComponent*>(c)) bn->SetTestMode(test_mode != 0);
  else if (BatchNormTestComponent* b = dynamic_cast<BatchNormTestComponent*>(c)) b->SetTestMode(test_mode != 0);
  else if (RandomComponent* r = dynamic_cast<RandomComponent*>(c)) r->SetTestMode(test_mode != 0);
  else if (TdnnDARTSV3Component* t = dynamic_cast<TdnnDARTSV3Component*>(c)) t->SetTestMode(test_mode != 0);
  API_END
}
// Device pointers of the parameters (for the data-parallel all-reduce of the deltas): up to 2 buffers.
int tdnnf_nnet3_param_buffers(void* comp, float** ptrs, int* rows, int* cols, int* strides, int* count) {
  API_BEGIN
  Component* c = static_cast<Component*>(comp);
  *count = 0;
  if (TdnnDARTSV3Component* t = dynamic_cast<TdnnDARTSV3Component*>(c)) {
    ptrs[0] = t->LinearParams().Data(); rows[0] = t->LinearParams().NumRows(); cols[0] = t->LinearParams().NumCols(); strides[0] = t->LinearParams().Stride();
    *count = 1;
    if (t->BiasParams().Dim() != 0) {  // use-bias=false (the `linear` half of a stock tdnnf-layer): one buffer
      ptrs[1] = t->BiasParams().Data(); rows[1] = 1; cols[1] = t->BiasParams().Dim(); strides[1] = t->BiasParams().Dim();
      *count = 2;
    }
  } else if (VectorFunctionComponentBase* v = dynamic_cast<VectorFunctionComponentBase*>(c)) {
    ptrs[0] = v->Output().Data(); rows[0] = 1; cols[0] = v->Output().Dim(); strides[0] = v->Output().Dim();
    *count = 1;
  }
  API_END
}
// ConstrainOrthonormal(Nnet*) (utils.cc:1037-1077) over the given components, in order; *num_updated (optional) <- how many
// parameter matrices the 1-in-4 draw selected this time.
int tdnnf_nnet3_constrain_orthonormal(void* const* comps, int n, int* num_updated) {
  API_BEGIN
  std::vector<Component*> list;
  for (int i = 0; i < n; ++i) list.push_back(static_cast<Component*>(comps[i]));
  const int32 k = ConstrainOrthonormal(list);
  if (num_updated) *num_updated = k;
  API_END
}
int tdnnf_nnet3_orthonormal_constraint(const void* comp, float* value) {
  API_BEGIN
  const TdnnDARTSV3Component* t = dynamic_cast<const TdnnDARTSV3Component*>(static_cast<const Component*>(comp));
  if (!t) KALDI_ERR << "component has no orthonormal-constraint";
  *value = t->OrthonormalConstraint();
  API_END
}
int tdnnf_nnet3_dropout_proportion(const void* comp, float* value) {
  API_BEGIN
  const GeneralDropoutComponent* g = dynamic_cast<const GeneralDropoutComponent*>(static_cast<const Component*>(comp));
  if (!g) KALDI_ERR << "not a GeneralDropoutComponent";
  *value = g->DropoutProportion();
  API_END
}
int tdnnf_nnet3_temp_proportion(const void* comp, float* value) {
  API_BEGIN
  const Component* c = static_cast<const Component*>(comp);
  if (const TdnnDARTSV3Component* t = dynamic_cast<const TdnnDARTSV3Component*>(c)) *value = t->TempProportion();
  else if (const GumbelSoftmaxFlopsComponent* g = dynamic_cast<const GumbelSoftmaxFlopsComponent*>(c)) *value = g->TempProportion();
  else KALDI_ERR << "component has no temperature";
  API_END
}
// BatchNormTest test hook (the reference only creates instances through Read()).
int tdnnf_nnet3_bn_test_set_stats(void* comp, int dim, int block_dim, float epsilon, float target_rms, double count,
                                  const double* sum, const double* sumsq) {
  API_BEGIN
  BatchNormTestComponent* b = dynamic_cast<BatchNormTestComponent*>(static_cast<Component*>(comp));
  if (!b) KALDI_ERR << "not a BatchNormTestComponent";
  b->SetStats(dim, block_dim, epsilon, target_rms, count, std::vector<double>(sum, sum + block_dim),
              std::vector<double>(sumsq, sumsq + block_dim));
  API_END
}

// Device pointers of a BatchNormTestComponent's derived scale_ / offset_ (block_dim entries each).
int tdnnf_nnet3_bn_test_scale_offset(const void* comp, const float** scale, const float** offset, int* dim) {
  API_BEGIN
  const BatchNormTestComponent* b = dynamic_cast<const BatchNormTestComponent*>(static_cast<const Component*>(comp));
  if (!b) KALDI_ERR << "not a BatchNormTestComponent";
  *scale = b->ScaleVec().Data();
  *offset = b->Offset().Data();
  *dim = b->ScaleVec().Dim();
  API_END
}

// ReadEditConfig over a list of (name, component): the `nnet3-copy --edits=...` step of train.py:524-532.
int tdnnf_nnet3_apply_edits(const char* edits, const char** names, void** comps, int n) {
  API_BEGIN
  std::vector<std::string> nm(names, names + n);
  std::vector<Component*> cs(n);
  for (int i = 0; i < n; ++i) cs[i] = static_cast<Component*>(comps[i]);
  std::istringstream is(edits);
  ReadEditConfig(is, nm, cs);
  API_END
}

}  // extern "C"
