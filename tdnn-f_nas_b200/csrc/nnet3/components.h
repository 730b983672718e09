// The reference's NAS components with their nnet3 Component interface intact (same class names,
// virtuals, Properties() flags, config keys and on-disk tokens), re-hosted on the C ABI of
// include/tdnnf_nas_b200.h: every CuMatrix/CuVector call of the reference's method bodies is
// replaced by one fused sm_100a kernel call.  Citations are into /root/reference:
//   conv.h = src/nnet3/nnet-convolutional-component.h, tdnn.cc = src/nnet3/nnet-tdnn-component.cc,
//   simple.{h,cc} = src/nnet3/nnet-simple-component.{h,cc}, norm.{h,cc} = src/nnet3/nnet-normalize-component.{h,cc}
#pragma once
#include "shim.h"

namespace tdnnf {
namespace nnet3 {

// X as a preconditioner sees it, WITHOUT materialising it: the splice [w_1 X_1 | ... | w_n X_n | 1] of a device
// matrix (views as in GetInputPart, ref tdnn.cc:806-820), or a plain matrix (n = 1, no weights, no ones).
struct NgOperand {
  const BaseFloat* data;      // device matrix the views are taken from
  int32 rows, cols, stride;
  int32 num_rows;             // N: rows of X (= rows of every view)
  int32 n;                    // number of views
  const int32* row_offsets;   // host, n entries
  int32 row_stride;
  const BaseFloat* weff;      // device, n block weights (may be NULL only when n == 1: weight 1)
  bool ones_col;              // X carries an appended column of ones
  int32 Dim() const { return n * cols + (ones_col ? 1 : 0); }
  static NgOperand Plain(const CuMatrixBase<BaseFloat>& m);
};

// What a preconditioning call hands back: X_hat = X - H W and the device-resident scale.
struct NgProjection {
  bool identity;               // dimension 1 / rank 0: X_hat = X, scale 1
  int32 rank;
  const CuMatrix* W;           // W_t (rank x Dim) the projection was taken with
  const CuMatrix* H;           // X W_t^T (N x rank)
  const BaseFloat* scale_dev;  // device scalar sqrt(tr(X X^T) / tr(X_hat X_hat^T))
};

// OnlineNaturalGradient (kaldi: nnet3/natural-gradient-online.h); see natural_gradient.cc.
class OnlineNaturalGradient {
 public:
  OnlineNaturalGradient();
  OnlineNaturalGradient(const OnlineNaturalGradient& other);
  OnlineNaturalGradient& operator=(const OnlineNaturalGradient& other);
  ~OnlineNaturalGradient();
  // The setters first complete a refresh whose host half may still be running on a worker thread (it reads rank_,
  // alpha_, d_t_, rho_t_): Read() / an edit config may call them between two minibatches.
  void SetRank(int32 rank) { FinishPendingUpdate(); rank_ = rank; }
  void SetUpdatePeriod(int32 update_period) { FinishPendingUpdate(); update_period_ = update_period; }
  void SetNumSamplesHistory(BaseFloat h) { FinishPendingUpdate(); num_samples_history_ = h; }
  void SetAlpha(BaseFloat alpha) { FinishPendingUpdate(); alpha_ = alpha; }
  int32 GetRank() const { return rank_; }
  int32 GetUpdatePeriod() const { return update_period_; }
  BaseFloat GetNumSamplesHistory() const { return num_samples_history_; }
  BaseFloat GetAlpha() const { return alpha_; }
  void Freeze(bool frozen) { FinishPendingUpdate(); frozen_ = frozen; }
  void Swap(OnlineNaturalGradient* other);
  // The upstream call: X_t <- X_hat_t, *scale on the host (synchronises the stream once).
  void PreconditionDirections(CuMatrixBase<BaseFloat>* X_t, BaseFloat* scale);
  // The same update of the Fisher estimate with X (and X_hat) left implicit and no host sync.
  void PreconditionImplicit(const NgOperand& X, NgProjection* out);
  // State read-back (finishes a pending update first): for tests and diagnostics.
  void GetState(int32* t, BaseFloat* rho, std::vector<BaseFloat>* d, Matrix<BaseFloat>* W);
  int32 NumReorthogonalized() const { return num_reorthogonalized_; }
  void FreeScratch();

 private:
  struct Pending;
  BaseFloat Eta(int32 N) const;
  bool Updating() const;
  void InitDefault(int32 D);
  void Init(const NgOperand& X);
  void Step(const NgOperand& X, bool updating);
  void FinishPendingUpdate();
  void HostHalfOfUpdate(int32 D);
  void Reorthogonalize();
  void RefreshDerived();
  void EnsureConsts();

  int32 rank_, update_period_;
  BaseFloat num_samples_history_, alpha_, epsilon_, delta_;
  bool frozen_;
  int32 t_;
  BaseFloat rho_t_;
  std::vector<BaseFloat> d_t_;
  int32 num_reorthogonalized_;
  CuMatrix W_t_;     // rank x D
  CuMatrix WWt_;     // rank x rank
  CuVector w_last_;  // last column of W_t (weights of the column of ones), contiguous
  CuVector consts_;  // {1, -1} on the device
  // scratch, sized on first use
  CuMatrix H_, J_, L_, K_, A_, AC_, W_next_;
  CuVector scal_, tmp_r_;
  double* sumsq_ = nullptr;
  Pending* pending_;
};

// Data-parallel world size used to normalise the FLOPs penalty by the GLOBAL row count
// (the reference divides by its local rows, simple.cc:10154; summed over G ranks that
// would count the penalty G times).  Default 1 = the reference's arithmetic.
void SetDataParallelWorldSize(int32 g);
int32 GetDataParallelWorldSize();
// The reference prints "log_alpha [ ... ]" to stdout every minibatch per component (tdnn.cc:571,
// simple.cc:2640), which costs a device->host sync each time; off by default here.
void SetPrintLogAlpha(bool b);
// TdnnDARTSV3Component::Propagate keeps the operand planes of its input in the memo for Backprop (default on).
void SetKeepPlanes(bool b);
// One fp16 tensor-core product in the parameter-gradient GEMM of TdnnDARTSV3Component (default: off, see components.cc).
void SetFastGradients(bool b);
bool FastGradients();
// Diagnostic: PreconditionDirections == identity with scale 1 (the un-preconditioned gradient).
void SetNaturalGradientIdentity(bool b);
// The symmetric eigen-solver of the host half of the update (Householder + implicit QL), exposed for host tests.
bool SymmetricEigenForTest(const double* a, int n, double* vals, double* vecs);
bool NaturalGradientIdentity();

// The six mode booleans of TdnnDARTSV3Component in their on-disk order (conv.h:243-257).
struct TdnnDARTSV3ModeFlags {
  bool use_gumbel, use_entropy, free_select, update_alpha, update_theta, uniform_sample;
};

// ------------------------------------------------------------------ TdnnDARTSV3Component (conv.h:112-332)
class TdnnDARTSV3Component : public UpdatableComponent {
 public:
  TdnnDARTSV3Component();
  TdnnDARTSV3Component(const TdnnDARTSV3Component& other) : TdnnDARTSV3Component(other, true) {}
  virtual int32 InputDim() const { return linear_params_.NumCols() / static_cast<int32>(time_offsets_.size()); }
  virtual int32 OutputDim() const { return linear_params_.NumRows(); }
  virtual std::string Info() const;
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual std::string Type() const { return "TdnnDARTSV3Component"; }
  virtual int32 Properties() const {
    return kUpdatableComponent | kReordersIndexes | kBackpropAdds | (bias_params_.Dim() == 0 ? kPropagateAdds : 0) |
           kBackpropNeedsInput | kUsesMemo;
  }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void DeleteMemo(void* memo) const;
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual Component* Copy() const { return new TdnnDARTSV3Component(*this); }
  virtual void ReorderIndexes(std::vector<Index>* input_indexes, std::vector<Index>* output_indexes) const;
  virtual void GetInputIndexes(const MiscComputationInfo& misc_info, const Index& output_index,
                               std::vector<Index>* desired_indexes) const;
  virtual bool IsComputable(const MiscComputationInfo& misc_info, const Index& output_index,
                            const IndexSet& input_index_set, std::vector<Index>* used_inputs) const;
  virtual ComponentPrecomputedIndexes* PrecomputeIndexes(const MiscComputationInfo& misc_info,
                                                         const std::vector<Index>& input_indexes,
                                                         const std::vector<Index>& output_indexes,
                                                         bool need_backprop) const;
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const Component& other);
  virtual void PerturbParams(BaseFloat stddev);
  virtual BaseFloat DotProduct(const UpdatableComponent& other) const;
  virtual int32 NumParameters() const;
  virtual void Vectorize(std::vector<BaseFloat>* params) const;
  virtual void UnVectorize(const std::vector<BaseFloat>& params);
  virtual void FreezeNaturalGradient(bool freeze);

  class PrecomputedIndexes : public ComponentPrecomputedIndexes {
   public:
    PrecomputedIndexes() {}
    PrecomputedIndexes(const PrecomputedIndexes& other) : row_stride(other.row_stride), row_offsets(other.row_offsets) {}
    virtual PrecomputedIndexes* Copy() const;
    virtual void Write(std::ostream& os, bool binary) const;
    virtual void Read(std::istream& os, bool binary);
    virtual std::string Type() const { return "TdnnDARTSV3ComponentPrecomputedIndexes"; }
    virtual ~PrecomputedIndexes() {}
    int32 row_stride;
    std::vector<int32> row_offsets;
  };

  CuMatrix& LinearParams() { return linear_params_; }
  CuVector& BiasParams() { return bias_params_; }
  const CuMatrix& LinearParams() const { return linear_params_; }
  const CuVector& BiasParams() const { return bias_params_; }
  BaseFloat OrthonormalConstraint() const { return orthonormal_constraint_; }
  OnlineNaturalGradient& PreconditionerIn() { return preconditioner_in_; }
  OnlineNaturalGradient& PreconditionerOut() { return preconditioner_out_; }
  void ConsolidateMemory();
  void SetTempProportion(BaseFloat p) { temp_proportion_ = p; }
  BaseFloat TempProportion() const { return temp_proportion_; }
  void SetTestMode(bool test_mode) { test_mode_ = test_mode; }
  bool test_mode_;  // public, unused -- as in the reference (conv.h:241)
  // A parameter-less instance that only knows its time offsets: enough for the index methods
  // (ReorderIndexes, PrecomputeIndexes, GetInputIndexes, IsComputable), which need no device.
  static TdnnDARTSV3Component* NewForIndexing(const std::vector<int32>& time_offsets);

  // The memo returned by Propagate: the mixing coefficients (what the reference keeps) plus the
  // effective GEMM weights derived from them; both stay on the device.
  // in_planes: the bf16 operand planes of Propagate's input (tdnnf_planes_acquire), attached again in Backprop, whose
  // natural-gradient projection and parameter gradient re-read in_value (tdnn.cc:476-539): the matrix is split once per
  // minibatch instead of twice.  nnet3 keeps in_value unchanged between the two calls (kBackpropNeedsInput).
  struct Memo {
    CuVector coef, weff;
    tdnnf_planes* in_planes = nullptr;
    ~Memo() { tdnnf_planes_release(in_planes); }
  };

 protected:
  // check == false: a derived class runs its own Check() (NumAlphaSlots() is virtual, hence not usable from this ctor)
  TdnnDARTSV3Component(const TdnnDARTSV3Component& other, bool check);
  int32 Flags() const;
  // share_offset_index of tdnn.cc:227-241; KALDI_ERR where the reference reads it uninitialised.
  int32 ShareOffsetIndex() const;
  static void ModifyComputationIo(time_height_convolution::ConvolutionComputationIo* io);
  // Entries of bias_params_ in front of the real bias: the n architecture weights here (tdnn.cc:172-176), none in the
  // stock TdnnComponent below.
  virtual int32 NumAlphaSlots() const { return static_cast<int32>(time_offsets_.size()); }
  void Check() const;
  // The part of UpdateNaturalGradient that the stock TdnnComponent shares (tdnn.cc:592-624): both
  // PreconditionDirections calls on the implicit operands, the raw gradient G = out_deriv^T [w_1 X_1 | .. | w_n X_n | 1]
  // (+ the inner products s_i when `s` is given), the rank-r corrections and the scaled accumulation into
  // linear_params_ / the bias part of bias_params_.  `this` is the delta component.
  void PreconditionedUpdate(const PrecomputedIndexes& indexes, const CuMatrixBase<BaseFloat>& in_value,
                            const CuMatrixBase<BaseFloat>& out_deriv, const CuMatrix* model_linear_params,
                            const BaseFloat* weff_dev, CuVector* s);
  void UpdateNaturalGradient(const PrecomputedIndexes& indexes, const CuMatrixBase<BaseFloat>& in_value,
                             const CuMatrixBase<BaseFloat>& out_deriv, const CuMatrix& linear_params_temp_,
                             const CuVector& bias_params_temp_, const Memo& memo, int32 share_offset_index_temp_,
                             int32 model_flags,
                             BaseFloat temp_proportion_temp_);
  void UpdateSimple(const PrecomputedIndexes& indexes, const CuMatrixBase<BaseFloat>& in_value,
                    const CuMatrixBase<BaseFloat>& out_deriv);

  bool use_gumbel_, use_entropy_, free_select_, update_alpha_, update_theta_, uniform_sample_;
  BaseFloat temp_proportion_;
  std::vector<int32> time_offsets_;
  CuMatrix linear_params_;
  CuVector bias_params_;
  BaseFloat orthonormal_constraint_;
  bool use_natural_gradient_;
  OnlineNaturalGradient preconditioner_in_, preconditioner_out_;
  // scratch of UpdateNaturalGradient (sized on first use; not part of the model)
  CuMatrix ng_grad_, ng_g1_, ng_t_;
  CuVector ng_colsum_, ng_consts_;
};

// ------------------------------------------------------------------ TdnnComponent (upstream kaldi, nnet-convolutional-component.h)
// The stock class TdnnDARTSV3Component was forked from (conv.h:112-332 is its declaration plus the mode flags), and
// what the manual TDNN-F system (NAS/run_tdnn_7q_fbk_40_manual.sh, BASELINE configs[1]) and every architecture the
// search emits (NAS/scripts/generate_top_list.py) are built from: out = bias + sum_i X_i W_i^T, no architecture weights,
// no memo; ConstrainOrthonormal covers it (utils.cc:1062-1066).  It is implemented ON the DARTS class: the same GEMM
// kernels with every w_i = 1, the same index methods and whole-parameter ops; bias_params_ has dimension D_out.
class TdnnComponent : public TdnnDARTSV3Component {
 public:
  TdnnComponent();
  TdnnComponent(const TdnnComponent& other);
  virtual std::string Type() const { return "TdnnComponent"; }
  virtual int32 Properties() const {
    return kUpdatableComponent | kReordersIndexes | kBackpropAdds | (bias_params_.Dim() == 0 ? kPropagateAdds : 0) |
           kBackpropNeedsInput;
  }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void DeleteMemo(void*) const {}
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual Component* Copy() const { return new TdnnComponent(*this); }
  virtual ComponentPrecomputedIndexes* PrecomputeIndexes(const MiscComputationInfo& misc_info,
                                                         const std::vector<Index>& input_indexes,
                                                         const std::vector<Index>& output_indexes,
                                                         bool need_backprop) const;
  // Same content as the DARTS indexes (row_stride, row_offsets); the stock on-disk tokens.
  class PrecomputedIndexes : public TdnnDARTSV3Component::PrecomputedIndexes {
   public:
    PrecomputedIndexes() {}
    explicit PrecomputedIndexes(const TdnnDARTSV3Component::PrecomputedIndexes& other)
        : TdnnDARTSV3Component::PrecomputedIndexes(other) {}
    virtual PrecomputedIndexes* Copy() const { return new PrecomputedIndexes(*this); }
    virtual void Write(std::ostream& os, bool binary) const;
    virtual void Read(std::istream& is, bool binary);
    virtual std::string Type() const { return "TdnnComponentPrecomputedIndexes"; }
  };

 protected:
  virtual int32 NumAlphaSlots() const { return 0; }

 private:
  const BaseFloat* Ones() const;  // n device ones: the effective weights of the shared GEMM kernels
  void UpdateSimple(const PrecomputedIndexes& indexes, const CuMatrixBase<BaseFloat>& in_value,
                    const CuMatrixBase<BaseFloat>& out_deriv);
  mutable CuVector ones_;
};

// ConstrainOrthonormal (utils.cc:1037-1077) over a list of components: every TdnnComponent with a non-zero
// orthonormal-constraint is updated with probability 1/4 (RandInt(0, 3) == 0, one draw per constrained component in
// list order).  Returns how many were updated.  (LinearComponent / AffineComponent are not part of this library;
// TdnnDARTSV3Component is deliberately not covered, utils.cc:1047-1066 -- TdnnComponent* casts only.)
int32 ConstrainOrthonormal(const std::vector<Component*>& components);

// ------------------------------------------------------------------ {Gumbel}SoftmaxFlopsComponent (simple.h:2924-3040)
class SoftmaxFlopsComponent : public RandomComponent {
 public:
  SoftmaxFlopsComponent() : dim_(0), scale_(1.0) {}
  SoftmaxFlopsComponent(const SoftmaxFlopsComponent& other) : RandomComponent(other), dim_(other.dim_), scale_(other.scale_) {}
  void Init(int32 dim, BaseFloat scale) { scale_ = scale; dim_ = dim; }
  virtual int32 Properties() const {
    return kBackpropInPlace | kSimpleComponent | kBackpropNeedsInput | kBackpropNeedsOutput | kRandomComponent;
  }
  virtual std::string Type() const { return "SoftmaxFlopsComponent"; }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual Component* Copy() const { return new SoftmaxFlopsComponent(*this); }
  virtual std::string Info() const;
 private:
  int32 dim_;
  BaseFloat scale_;
};

class GumbelSoftmaxFlopsComponent : public RandomComponent {
 public:
  GumbelSoftmaxFlopsComponent() : dim_(0), scale_(1.0), temp_proportion_(0.0) {}
  GumbelSoftmaxFlopsComponent(const GumbelSoftmaxFlopsComponent& other)
      : RandomComponent(other), dim_(other.dim_), scale_(other.scale_), temp_proportion_(other.temp_proportion_) {}
  void Init(int32 dim, BaseFloat scale, BaseFloat temp_proportion) { temp_proportion_ = temp_proportion; scale_ = scale; dim_ = dim; }
  virtual int32 Properties() const {
    return kBackpropInPlace | kSimpleComponent | kBackpropNeedsInput | kBackpropNeedsOutput | kRandomComponent;
  }
  virtual std::string Type() const { return "GumbelSoftmaxFlopsComponent"; }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual Component* Copy() const { return new GumbelSoftmaxFlopsComponent(*this); }
  virtual std::string Info() const;
  void SetTempProportion(BaseFloat p) { temp_proportion_ = p; }
  BaseFloat TempProportion() const { return temp_proportion_; }
 private:
  int32 dim_;
  BaseFloat scale_;
  BaseFloat temp_proportion_;
};

// ------------------------------------------------------------------ CopyNComponent (simple.h:2119-2150)
class CopyNComponent : public Component {
 public:
  CopyNComponent() : input_dim_(0), output_dim_(0), scale_(1.0) {}
  CopyNComponent(const CopyNComponent& other) : input_dim_(other.input_dim_), output_dim_(other.output_dim_), scale_(other.scale_) {}
  virtual int32 Properties() const { return kSimpleComponent | kPropagateAdds | kBackpropAdds; }
  virtual std::string Type() const { return "CopyNComponent"; }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const { return output_dim_; }
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual Component* Copy() const { return new CopyNComponent(*this); }
  virtual std::string Info() const;
 private:
  int32 input_dim_, output_dim_;
  BaseFloat scale_;
};

// ------------------------------------------------------------------ Onehot / Constant function (simple.h:2734-2794)
// OnehotFunctionComponent and the reference's modified ConstantFunctionComponent share everything
// except Propagate, the non-natural-gradient learning-rate factor (x5 for Constant, simple.cc:2636) and
// the log_alpha print (simple.cc:2640).
class VectorFunctionComponentBase : public UpdatableComponent {
 public:
  VectorFunctionComponentBase() : UpdatableComponent(), input_dim_(-1), is_updatable_(true), use_natural_gradient_(true) {}
  VectorFunctionComponentBase(const VectorFunctionComponentBase& other)
      : UpdatableComponent(other), input_dim_(other.input_dim_), output_(other.output_),
        is_updatable_(other.is_updatable_), use_natural_gradient_(other.use_natural_gradient_),
        preconditioner_(other.preconditioner_) {}
  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const { return output_.Dim(); }
  virtual std::string Info() const;
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual int32 Properties() const {
    return kSimpleComponent | (is_updatable_ ? kUpdatableComponent : 0) |
           (InputDim() == OutputDim() ? kPropagateInPlace : 0) | kBackpropAdds;
  }
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const Component& other);
  virtual void PerturbParams(BaseFloat stddev);
  virtual BaseFloat DotProduct(const UpdatableComponent& other) const;
  virtual int32 NumParameters() const;
  virtual void Vectorize(std::vector<BaseFloat>* params) const;
  virtual void UnVectorize(const std::vector<BaseFloat>& params);
  virtual void ConsolidateMemory();
  CuVector& Output() { return output_; }
  const CuVector& Output() const { return output_; }
  OnlineNaturalGradient& Preconditioner() { return preconditioner_; }
 protected:
  virtual BaseFloat PlainUpdateFactor() const = 0;  // multiplies learning_rate_ when NG is off
  virtual bool PrintsLogAlpha() const = 0;
  int32 input_dim_;
  CuVector output_;
  bool is_updatable_;
  bool use_natural_gradient_;
  OnlineNaturalGradient preconditioner_;
  CuMatrix out_deriv_copy_;  // scratch of Backprop (the reference's out_deriv_copy)
};

class OnehotFunctionComponent : public VectorFunctionComponentBase {
 public:
  OnehotFunctionComponent() {}
  OnehotFunctionComponent(const OnehotFunctionComponent& other) : VectorFunctionComponentBase(other) {}
  virtual std::string Type() const { return "OnehotFunctionComponent"; }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual Component* Copy() const { return new OnehotFunctionComponent(*this); }
 protected:
  virtual BaseFloat PlainUpdateFactor() const { return 1.0; }  // simple.cc:9547
  virtual bool PrintsLogAlpha() const { return false; }
};

class ConstantFunctionComponent : public VectorFunctionComponentBase {
 public:
  ConstantFunctionComponent() {}
  ConstantFunctionComponent(const ConstantFunctionComponent& other) : VectorFunctionComponentBase(other) {}
  virtual std::string Type() const { return "ConstantFunctionComponent"; }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual Component* Copy() const { return new ConstantFunctionComponent(*this); }
 protected:
  virtual BaseFloat PlainUpdateFactor() const { return 5.0; }  // simple.cc:2636 (the reference's silent modification)
  virtual bool PrintsLogAlpha() const { return true; }         // simple.cc:2640
};

// ------------------------------------------------------------------ ElementwiseProductComponent (simple.cc:237-299)
class ElementwiseProductComponent : public Component {
 public:
  ElementwiseProductComponent() : input_dim_(0), output_dim_(0) {}
  void Init(int32 input_dim, int32 output_dim);
  virtual int32 Properties() const { return kSimpleComponent | kBackpropNeedsInput; }
  virtual std::string Type() const { return "ElementwiseProductComponent"; }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const { return output_dim_; }
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual Component* Copy() const { return new ElementwiseProductComponent(*this); }
 private:
  int32 input_dim_, output_dim_;
};

// ------------------------------------------------------------------ BatchNormTestComponent (norm.h:336-471)
class BatchNormTestComponent : public Component {
 public:
  BatchNormTestComponent() : dim_(0), block_dim_(0), epsilon_(1.0e-03), target_rms_(1.0), test_mode_(false), count_(0) {}
  BatchNormTestComponent(const BatchNormTestComponent& other);
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual std::string Info() const;
  virtual void InitFromConfig(ConfigLine* cfl);  // empty in the reference (norm.cc:757-759)
  virtual std::string Type() const { return "BatchNormTestComponent"; }
  virtual int32 Properties() const {
    return kSimpleComponent | kBackpropNeedsOutput | kPropagateInPlace | kBackpropInPlace |
           (block_dim_ < dim_ ? kInputContiguous | kOutputContiguous : 0);
  }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual Component* Copy() const { return new BatchNormTestComponent(*this); }
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const Component& other);
  virtual void ZeroStats() {}                                                                  // norm.cc:1008-1010
  virtual void StoreStats(const CuMatrixBase<BaseFloat>&, const CuMatrixBase<BaseFloat>&, void*) {}  // norm.cc:924-929
  void SetTestMode(bool test_mode);
  const CuVector& Offset() const { return offset_; }
  const CuVector& ScaleVec() const { return scale_; }
  // test hook: install statistics directly (the reference only gets them through Read()).
  void SetStats(int32 dim, int32 block_dim, BaseFloat epsilon, BaseFloat target_rms, double count,
                const std::vector<double>& sum, const std::vector<double>& sumsq);
 protected:
  virtual const char* Token() const { return "BatchNormTestComponent"; }  // on-disk name (Read / Write)
  void Check() const;
  void ComputeDerived();
  int32 dim_, block_dim_;
  BaseFloat epsilon_, target_rms_;
  bool test_mode_;
  double count_;
  std::vector<double> stats_sum_, stats_sumsq_;  // CuVector<double> in the reference; tiny, host side here
  CuVector offset_, scale_;
};

// ------------------------------------------------------------------ BatchNormComponent (norm.h:150-262, norm.cc:209-680)
// The stock component of the supernet PRETRAIN stage (the search stage `sed`s it into BatchNormTestComponent):
// training mode normalises with the minibatch statistics (Memo 5 x block_dim: mean, uvar, scale, var_deriv_mod,
// temp), StoreStats accumulates them; test mode is the same affine map as BatchNormTestComponent.  Shares the
// statistics, derived vectors and on-disk layout with BatchNormTestComponent (only the token differs).
class BatchNormComponent : public BatchNormTestComponent {
 public:
  struct Memo {
    int32 num_frames;
    CuVector mean_uvar_scale;  // 5 x block_dim, rows contiguous: mean, uvar, scale, var_deriv_mod, temp
  };
  BatchNormComponent();
  BatchNormComponent(const BatchNormComponent& other);
  virtual ~BatchNormComponent();
  virtual std::string Info() const;
  virtual void InitFromConfig(ConfigLine* cfl);  // norm.cc:289-317
  virtual std::string Type() const { return "BatchNormComponent"; }
  virtual int32 Properties() const {
    return kSimpleComponent | kBackpropNeedsOutput | kPropagateInPlace | kBackpropInPlace |
           (block_dim_ < dim_ ? kInputContiguous | kOutputContiguous : 0) | (test_mode_ ? 0 : kUsesMemo | kStoresStats);
  }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void StoreStats(const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value, void* memo);
  virtual void ZeroStats();
  virtual void DeleteMemo(void* memo) const { delete static_cast<Memo*>(memo); }
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual Component* Copy() const { return new BatchNormComponent(*this); }
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const Component& other);
  void SetTestMode(bool test_mode);
  double Count() const;

 protected:
  virtual const char* Token() const { return "BatchNormComponent"; }
  void ComputeDerivedBn();  // norm.cc:209-247: empty vectors outside test mode
  // StoreStats accumulates on the device (no host sync per minibatch); the host copies are brought up to date
  // whenever they are read.
  void FlushStats() const;
  mutable double* d_stats_;        // device [2 * block_dim]: sum, sumsq accumulated since the last flush
  mutable double pending_count_;
};

// ------------------------------------------------------------------ NonlinearComponent / RectifiedLinearComponent
// The ReLU of every TDNN-F block with its activation statistics and self-repair (nnet-component-itf.h NonlinearComponent,
// nnet-component-itf.cc:433-725; nnet-simple-component.h:344-376, nnet-simple-component.cc:958-1094).  The statistics
// live in device doubles (the reference's CuVector<double>) and are mirrored on the host only for I/O, Scale and Add.
class NonlinearComponent : public Component {
 public:
  NonlinearComponent();
  NonlinearComponent(const NonlinearComponent& other);
  virtual ~NonlinearComponent();
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual std::string Info() const;
  virtual void ZeroStats();
  virtual void Scale(BaseFloat scale);
  virtual void Add(BaseFloat alpha, const Component& other);
  double Count() const { return count_; }
  double OderivCount() const { return oderiv_count_; }
  double NumDimsProcessed() const { return num_dims_processed_; }
  double NumDimsSelfRepaired() const;

 protected:
  struct HostStats {
    std::vector<double> value_sum, deriv_sum, oderiv_sumsq;  // empty = "Dim() == 0" in the reference
    double num_dims_self_repaired;
  };
  void Pull(HostStats* h) const;
  void Push(const HostStats& h);
  void EnsureDevice() const;
  double* ValueSum() const { return stats_dev_; }
  double* DerivSum() const { return stats_dev_ + dim_; }
  double* OderivSumsq() const { return stats_dev_ + 2 * dim_; }
  double* Repaired() const { return stats_dev_ + 3 * dim_; }
  void StoreStatsInternal(const CuMatrixBase<BaseFloat>& out_value, bool with_deriv);  // deriv = Heaviside(out_value)
  void StoreBackpropStats(const CuMatrixBase<BaseFloat>& out_deriv);
  static constexpr BaseFloat kUnsetThreshold = -1000.0f;
  int32 dim_, block_dim_;
  mutable double* stats_dev_;  // [3 * dim + 1]: value_sum, deriv_sum, oderiv_sumsq, num_dims_self_repaired
  bool has_value_, has_deriv_, has_oderiv_;
  double count_, oderiv_count_, num_dims_processed_;
  BaseFloat self_repair_lower_threshold_, self_repair_upper_threshold_, self_repair_scale_;
};

class RectifiedLinearComponent : public NonlinearComponent {
 public:
  RectifiedLinearComponent() {}
  RectifiedLinearComponent(const RectifiedLinearComponent& other) : NonlinearComponent(other) {}
  virtual std::string Type() const { return "RectifiedLinearComponent"; }
  virtual Component* Copy() const { return new RectifiedLinearComponent(*this); }
  virtual int32 Properties() const {
    return kSimpleComponent | kBackpropNeedsOutput | kPropagateInPlace | kStoresStats | (block_dim_ != dim_ ? kInputContiguous : 0);
  }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void StoreStats(const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value, void* memo);

 private:
  void RepairGradients(CuMatrixBase<BaseFloat>* in_deriv, RectifiedLinearComponent* to_update) const;
};

// LogSoftmaxComponent (simple.h:~715-740, simple.cc:3607-3632): the non-linearity of the `output-xent` branch.
class LogSoftmaxComponent : public NonlinearComponent {
 public:
  LogSoftmaxComponent() {}
  LogSoftmaxComponent(const LogSoftmaxComponent& other) : NonlinearComponent(other) {}
  virtual std::string Type() const { return "LogSoftmaxComponent"; }
  virtual Component* Copy() const { return new LogSoftmaxComponent(*this); }
  virtual int32 Properties() const { return kSimpleComponent | kBackpropNeedsOutput | kStoresStats; }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
};

// ------------------------------------------------------------------ edit directives (utils.cc:1166-1415)
// The subset of ReadEditConfig this path needs: set-temperature-proportion (utils.cc:1352-1405)
// plus set-learning-rate / set-learning-rate-factor for the recipes' model surgery.
// `names[i]` is the component name of components[i]; name patterns use '*' wildcards as in Kaldi.
void ReadEditConfig(std::istream& config_file, const std::vector<std::string>& names,
                    const std::vector<Component*>& components);
bool NameMatchesPattern(const char* name, const char* pattern);

// ------------------------------------------------------------------ GeneralDropoutComponent (upstream kaldi, nnet-general-component.h)
// The `dropout` node of every tdnnf-layer / relu-batchnorm-dropout-layer of the recipes (composite_layers.py; configured
// `dropout-proportion=0.0 continuous=true`, driven per iteration by `set-dropout-proportion`, utils.cc:1297-1330, from the
// schedule 0,0@0.20,0.5@0.50,0).  One mask row per sequence n (per block of time-period frames when time-period != 0),
// shared by all its frames: out[r,:] = in[r,:] .* mask[indexes[r],:].  Proportion 0 / test mode: a copy, no memo.
// SpecAugment masks (specaugment-max-proportion != 0) are not built: KALDI_ERR.
class GeneralDropoutComponent : public RandomComponent {
 public:
  GeneralDropoutComponent();
  virtual int32 InputDim() const { return dim_; }
  virtual int32 OutputDim() const { return dim_; }
  virtual std::string Info() const;
  virtual void InitFromConfig(ConfigLine* cfl);
  virtual std::string Type() const { return "GeneralDropoutComponent"; }
  virtual int32 Properties() const {
    return kRandomComponent | kPropagateInPlace | kBackpropInPlace | kUsesMemo |
           (block_dim_ != dim_ ? (kInputContiguous | kOutputContiguous) : 0);
  }
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const;
  virtual void DeleteMemo(void* memo) const { delete static_cast<CuMatrix*>(memo); }
  virtual ComponentPrecomputedIndexes* PrecomputeIndexes(const MiscComputationInfo& misc_info,
                                                         const std::vector<Index>& input_indexes,
                                                         const std::vector<Index>& output_indexes,
                                                         bool need_backprop) const;
  virtual void Read(std::istream& is, bool binary);
  virtual void Write(std::ostream& os, bool binary) const;
  virtual Component* Copy() const { return new GeneralDropoutComponent(*this); }
  void SetDropoutProportion(BaseFloat p) { dropout_proportion_ = p; }
  BaseFloat DropoutProportion() const { return dropout_proportion_; }

  class PrecomputedIndexes : public ComponentPrecomputedIndexes {
   public:
    PrecomputedIndexes() : num_mask_rows(0) {}
    PrecomputedIndexes(const PrecomputedIndexes& other) : num_mask_rows(other.num_mask_rows), indexes(other.indexes) {}
    virtual PrecomputedIndexes* Copy() const { return new PrecomputedIndexes(*this); }
    virtual void Write(std::ostream& os, bool binary) const;
    virtual void Read(std::istream& is, bool binary);
    virtual std::string Type() const { return "GeneralDropoutComponentPrecomputedIndexes"; }
    const int32* DeviceIndexes() const;  // uploaded on first use (CuArray<int32> upstream)
    int32 num_mask_rows;
    std::vector<int32> indexes;  // per (reshaped) row of the input: its mask row

   private:
    mutable CuVector dev_;  // the int32 indexes, bit-copied into device words
  };

 private:
  void MulRows(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out, const CuMatrix& mask,
               const PrecomputedIndexes& indexes) const;
  int32 dim_, block_dim_, time_period_;
  BaseFloat dropout_proportion_, specaugment_max_proportion_;
  bool continuous_;
};

}  // namespace nnet3
}  // namespace tdnnf
