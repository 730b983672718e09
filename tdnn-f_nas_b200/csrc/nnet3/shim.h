// Minimal stand-ins for the parts of Kaldi's base/, matrix/, cudamatrix/, util/ and nnet3/ that
// the six NAS components touch: typedefs, error macros, a CuMatrixBase-like
// VIEW (device pointer + rows/cols/stride, no arithmetic -- all device work goes through the C ABI
// in include/tdnnf_nas_b200.h), owning CuVector / CuMatrix buffers for parameters, Kaldi's token
// I/O (text + binary), ConfigLine, Index / IndexSet, and the Component base classes with the same
// virtuals, property flags and Read/Write helpers as the reference (itf.cc:300-431).
//
// In a real Kaldi tree none of this file is needed: the component method bodies in components.cc
// compile against Kaldi's own headers with `view(mat)` adapters.
#pragma once
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "tdnnf_nas_b200.h"

namespace tdnnf {
namespace nnet3 {

typedef float BaseFloat;
typedef int32_t int32;
typedef int64_t int64;

// ------------------------------------------------------------------ errors
// KALDI_ERR throws std::runtime_error (as in Kaldi); KALDI_ASSERT aborts in Kaldi -- here it throws
// AssertionFailure so that a host program (and the parity tests) can observe it.
struct AssertionFailure : public std::logic_error {
  using std::logic_error::logic_error;
};

class ErrStream {
 public:
  ErrStream(const char* file, int line) { ss_ << "ERROR (" << file << ":" << line << ") "; }
  template <class T>
  ErrStream& operator<<(const T& v) { ss_ << v; return *this; }
  [[noreturn]] ~ErrStream() noexcept(false) { throw std::runtime_error(ss_.str()); }
 private:
  std::ostringstream ss_;
};
#define KALDI_ERR ::tdnnf::nnet3::ErrStream(__FILE__, __LINE__)
#define KALDI_ASSERT(cond)                                                                              \
  do {                                                                                                  \
    if (!(cond))                                                                                        \
      throw ::tdnnf::nnet3::AssertionFailure(std::string("KALDI_ASSERT failed: ") + #cond + " at " +    \
                                             __FILE__ + ":" + std::to_string(__LINE__));                \
  } while (0)
void KaldiLog(const std::string& msg);   // KALDI_LOG: stderr
void KaldiWarn(const std::string& msg);  // KALDI_WARN: stderr
int32 GetVerboseLevel();
void SetVerboseLevel(int32 v);

// ------------------------------------------------------------------ device plumbing
// The "CuDevice" of this mirror: the tdnnf context the components launch on (thread-local).
tdnnf_ctx* CurrentContext();          // throws if none is set
void SetCurrentContext(tdnnf_ctx* c);
void CheckStatus(int rc);             // throws std::runtime_error(tdnnf_last_error()) on failure

// A view, exactly the information a CuMatrixBase<BaseFloat> carries.
template <typename Real>
class CuMatrixBase {
 public:
  CuMatrixBase() : data_(nullptr), num_rows_(0), num_cols_(0), stride_(0) {}
  CuMatrixBase(Real* data, int32 rows, int32 cols, int32 stride)
      : data_(data), num_rows_(rows), num_cols_(cols), stride_(stride) {}
  int32 NumRows() const { return num_rows_; }
  int32 NumCols() const { return num_cols_; }
  int32 Stride() const { return stride_; }
  const Real* Data() const { return data_; }
  Real* Data() { return data_; }
 protected:
  Real* data_;
  int32 num_rows_, num_cols_, stride_;
};
template <typename Real>
using CuSubMatrix = CuMatrixBase<Real>;
template <typename Real>
inline bool SameDim(const CuMatrixBase<Real>& a, const CuMatrixBase<Real>& b) {
  return a.NumRows() == b.NumRows() && a.NumCols() == b.NumCols();
}

// Host vector / matrix (Kaldi Vector<> / Matrix<>): used for I/O and initialisation.
template <typename Real>
struct Vector {
  std::vector<Real> v;
  Vector() {}
  explicit Vector(int32 dim) : v(dim, Real(0)) {}
  int32 Dim() const { return (int32)v.size(); }
  void Resize(int32 d) { v.assign(d, Real(0)); }
  Real& operator()(int32 i) { return v[i]; }
  const Real& operator()(int32 i) const { return v[i]; }
  void Write(std::ostream& os, bool binary) const;
  void Read(std::istream& is, bool binary);
};
template <typename Real>
struct Matrix {
  std::vector<Real> v;
  int32 rows = 0, cols = 0;
  Matrix() {}
  Matrix(int32 r, int32 c) : v((size_t)r * c, Real(0)), rows(r), cols(c) {}
  void Resize(int32 r, int32 c) { rows = r; cols = c; v.assign((size_t)r * c, Real(0)); }
  Real& operator()(int32 r, int32 c) { return v[(size_t)r * cols + c]; }
  const Real& operator()(int32 r, int32 c) const { return v[(size_t)r * cols + c]; }
  void Write(std::ostream& os, bool binary) const;
  void Read(std::istream& is, bool binary);
};

// Owning device vector (CuVector<BaseFloat>): parameters live here.
class CuVector {
 public:
  CuVector() {}
  explicit CuVector(int32 dim) { Resize(dim); }
  CuVector(const CuVector& o);
  CuVector& operator=(const CuVector& o);
  ~CuVector();
  void Resize(int32 dim);          // zero-filled
  int32 Dim() const { return dim_; }
  BaseFloat* Data() { return data_; }
  const BaseFloat* Data() const { return data_; }
  void CopyFromHost(const std::vector<BaseFloat>& h);
  std::vector<BaseFloat> ToHost() const;
  void SetZero();
  void Scale(BaseFloat s);
  void AddVec(BaseFloat alpha, const CuVector& o);
  void Write(std::ostream& os, bool binary) const;
  void Read(std::istream& is, bool binary);
 private:
  BaseFloat* data_ = nullptr;
  int32 dim_ = 0;
};

// Owning device matrix (CuMatrix<BaseFloat>), pitch-aligned rows like cudaMallocPitch.
class CuMatrix : public CuMatrixBase<BaseFloat> {
 public:
  CuMatrix() {}
  CuMatrix(int32 rows, int32 cols) { Resize(rows, cols); }
  CuMatrix(const CuMatrix& o);
  CuMatrix& operator=(const CuMatrix& o);
  ~CuMatrix();
  void Resize(int32 rows, int32 cols);  // zero-filled
  void Swap(CuMatrix* o) {
    std::swap(data_, o->data_);
    std::swap(num_rows_, o->num_rows_);
    std::swap(num_cols_, o->num_cols_);
    std::swap(stride_, o->stride_);
  }
  void CopyFromHost(const Matrix<BaseFloat>& h);
  Matrix<BaseFloat> ToHost() const;
  void SetZero();
  void Scale(BaseFloat s);
  void AddMat(BaseFloat alpha, const CuMatrix& o);
  void Write(std::ostream& os, bool binary) const;
  void Read(std::istream& is, bool binary);
};
BaseFloat VecVec(const CuVector& a, const CuVector& b);
// Between Begin and End every CuVector / CuMatrix allocation is carved, in call order and 256-byte aligned, from the
// caller-owned device range [base, base + bytes); End returns the bytes used.  See DevicePool in shim.cc.
void DeviceArenaBegin(void* base, size_t bytes);
size_t DeviceArenaEnd();
BaseFloat TraceMatMatTrans(const CuMatrix& a, const CuMatrix& b);  // TraceMatMat(a, b, kTrans)

// ------------------------------------------------------------------ Kaldi token I/O (base/io-funcs.h)
void WriteToken(std::ostream& os, bool binary, const std::string& token);
void ReadToken(std::istream& is, bool binary, std::string* token);
void ExpectToken(std::istream& is, bool binary, const std::string& token);
int PeekToken(std::istream& is, bool binary);
void ExpectOneOrTwoTokens(std::istream& is, bool binary, const std::string& token1, const std::string& token2);
void WriteBasicType(std::ostream& os, bool binary, bool v);
void WriteBasicType(std::ostream& os, bool binary, int32 v);
void WriteBasicType(std::ostream& os, bool binary, float v);
void WriteBasicType(std::ostream& os, bool binary, double v);
void ReadBasicType(std::istream& is, bool binary, bool* v);
void ReadBasicType(std::istream& is, bool binary, int32* v);
void ReadBasicType(std::istream& is, bool binary, float* v);
void ReadBasicType(std::istream& is, bool binary, double* v);
void WriteIntegerVector(std::ostream& os, bool binary, const std::vector<int32>& v);
void ReadIntegerVector(std::istream& is, bool binary, std::vector<int32>* v);
bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32>* out);

// ------------------------------------------------------------------ ConfigLine (util/text-utils.h)
class ConfigLine {
 public:
  // Parses "first-token key1=value1 key2=value2 ..." (values may be quoted); returns false on syntax error.
  bool ParseLine(const std::string& line);
  bool GetValue(const std::string& key, std::string* value);
  bool GetValue(const std::string& key, BaseFloat* value);
  bool GetValue(const std::string& key, int32* value);
  bool GetValue(const std::string& key, std::vector<int32>* value);
  bool GetValue(const std::string& key, bool* value);
  bool HasUnusedValues() const;
  std::string UnusedValues() const;
  const std::string& FirstToken() const { return first_token_; }
  const std::string& WholeLine() const { return whole_line_; }
 private:
  std::string whole_line_, first_token_;
  std::map<std::string, std::pair<std::string, bool> > data_;
};

// ------------------------------------------------------------------ indexes (nnet3/nnet-common.h)
const int32 kNoTime = -32768;
struct Index {
  int32 n, t, x;
  Index() : n(0), t(0), x(0) {}
  Index(int32 n_, int32 t_, int32 x_ = 0) : n(n_), t(t_), x(x_) {}
  bool operator==(const Index& o) const { return n == o.n && t == o.t && x == o.x; }
  bool operator!=(const Index& o) const { return !(*this == o); }
  bool operator<(const Index& o) const {  // Kaldi order: t, then x, then n
    if (t != o.t) return t < o.t;
    if (x != o.x) return x < o.x;
    return n < o.n;
  }
};
struct IndexHasher {
  size_t operator()(const Index& i) const noexcept { return (size_t)i.n + 1619u * (size_t)i.t + 15649u * (size_t)i.x; }
};
// IndexSet: in Kaldi an interface onto the ComputationGraph; here backed by a hash set.
class IndexSet {
 public:
  IndexSet() {}
  explicit IndexSet(const std::vector<Index>& v) : set_(v.begin(), v.end()) {}
  bool operator()(const Index& i) const { return set_.count(i) != 0; }
  void Insert(const Index& i) { set_.insert(i); }
 private:
  std::unordered_set<Index, IndexHasher> set_;
};
struct MiscComputationInfo {};

namespace time_height_convolution {
// nnet3/convolution.h: ConvolutionComputationIo and the index helpers TdnnDARTSV3Component uses.
struct ConvolutionComputationIo {
  int32 num_images = 0;
  int32 start_t_in = 0, t_step_in = 0, num_t_in = 0;
  int32 start_t_out = 0, t_step_out = 0, num_t_out = 0;
  int32 reorder_t_in = 1;
};
void GetComputationIo(const std::vector<Index>& input_indexes, const std::vector<Index>& output_indexes,
                      ConvolutionComputationIo* io);
void GetIndexesForComputation(const ConvolutionComputationIo& io, const std::vector<Index>& orig_input_indexes,
                              const std::vector<Index>& orig_output_indexes, std::vector<Index>* input_indexes,
                              std::vector<Index>* output_indexes);
}  // namespace time_height_convolution

// ------------------------------------------------------------------ randomness
// Kaldi draws from the process-global CuRand / rand().  Here every component draw comes from a
// counter-based generator keyed by (global seed, draw counter), so that all data-parallel ranks --
// which must apply identical Gumbel noise / one-hot choices per minibatch -- agree.
void SetRandSeed(uint64_t seed);
uint64_t GetRandSeed();
void SetRandCounter(uint64_t counter);
uint64_t GetRandCounter();
float RandUniformOpen();   // in (0,1): next draw
int32 RandInt(int32 lo, int32 hi);

// ------------------------------------------------------------------ Component interface
enum ComponentProperties {  // nnet3/nnet-component-itf.h
  kSimpleComponent = 0x001,
  kUpdatableComponent = 0x002,
  kPropagateInPlace = 0x004,
  kPropagateAdds = 0x008,
  kReordersIndexes = 0x010,
  kBackpropAdds = 0x020,
  kBackpropNeedsInput = 0x040,
  kBackpropNeedsOutput = 0x080,
  kBackpropInPlace = 0x100,
  kStoresStats = 0x200,
  kInputContiguous = 0x400,
  kOutputContiguous = 0x800,
  kUsesMemo = 0x1000,
  kRandomComponent = 0x2000
};

class ComponentPrecomputedIndexes {
 public:
  virtual ComponentPrecomputedIndexes* Copy() const = 0;
  virtual void Write(std::ostream& os, bool binary) const = 0;
  virtual void Read(std::istream& os, bool binary) = 0;
  virtual std::string Type() const = 0;
  static ComponentPrecomputedIndexes* ReadNew(std::istream& is, bool binary);
  static ComponentPrecomputedIndexes* NewComponentPrecomputedIndexesOfType(const std::string& cpi_type);
  virtual ~ComponentPrecomputedIndexes() {}
};

class Component {
 public:
  virtual void* Propagate(const ComponentPrecomputedIndexes* indexes, const CuMatrixBase<BaseFloat>& in,
                          CuMatrixBase<BaseFloat>* out) const = 0;
  virtual void Backprop(const std::string& debug_info, const ComponentPrecomputedIndexes* indexes,
                        const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                        const CuMatrixBase<BaseFloat>& out_deriv, void* memo, Component* to_update,
                        CuMatrixBase<BaseFloat>* in_deriv) const = 0;
  virtual void StoreStats(const CuMatrixBase<BaseFloat>& in_value, const CuMatrixBase<BaseFloat>& out_value,
                          void* memo) {}
  virtual void ZeroStats() {}
  virtual void GetInputIndexes(const MiscComputationInfo& misc_info, const Index& output_index,
                               std::vector<Index>* desired_indexes) const;
  virtual bool IsComputable(const MiscComputationInfo& misc_info, const Index& output_index,
                            const IndexSet& input_index_set, std::vector<Index>* used_inputs) const;
  virtual void ReorderIndexes(std::vector<Index>* input_indexes, std::vector<Index>* output_indexes) const {}
  virtual ComponentPrecomputedIndexes* PrecomputeIndexes(const MiscComputationInfo& misc_info,
                                                         const std::vector<Index>& input_indexes,
                                                         const std::vector<Index>& output_indexes,
                                                         bool need_backprop) const { return NULL; }
  virtual std::string Type() const = 0;
  virtual void InitFromConfig(ConfigLine* cfl) = 0;
  virtual int32 InputDim() const = 0;
  virtual int32 OutputDim() const = 0;
  virtual int32 Properties() const = 0;
  static Component* ReadNew(std::istream& is, bool binary);
  virtual Component* Copy() const = 0;
  static Component* NewComponentOfType(const std::string& type);
  virtual void Read(std::istream& is, bool binary) = 0;
  virtual void Write(std::ostream& os, bool binary) const = 0;
  virtual std::string Info() const;
  virtual void Scale(BaseFloat scale) {}
  virtual void Add(BaseFloat alpha, const Component& other) {}
  virtual void DeleteMemo(void* memo) const {}
  virtual void ConsolidateMemory() {}
  Component() {}
  virtual ~Component() {}
};

class RandomComponent : public Component {
 public:
  RandomComponent() : test_mode_(false) {}
  RandomComponent(const RandomComponent& other) : test_mode_(other.test_mode_) {}
  void SetTestMode(bool test_mode) { test_mode_ = test_mode; }
 protected:
  bool test_mode_;
};

class UpdatableComponent : public Component {
 public:
  UpdatableComponent(const UpdatableComponent& other);
  UpdatableComponent() : learning_rate_(0.001), learning_rate_factor_(1.0), l2_regularize_(0.0),
                         is_gradient_(false), max_change_(0.0) {}
  virtual ~UpdatableComponent() {}
  virtual BaseFloat DotProduct(const UpdatableComponent& other) const = 0;
  virtual void PerturbParams(BaseFloat stddev) = 0;
  virtual void SetUnderlyingLearningRate(BaseFloat lrate) { learning_rate_ = lrate * learning_rate_factor_; }
  virtual void SetActualLearningRate(BaseFloat lrate) { learning_rate_ = lrate; }
  virtual void SetAsGradient() { learning_rate_ = 1.0; is_gradient_ = true; }
  virtual BaseFloat LearningRateFactor() { return learning_rate_factor_; }
  virtual void SetLearningRateFactor(BaseFloat lrate_factor) { learning_rate_factor_ = lrate_factor; }
  void SetUpdatableConfigs(const UpdatableComponent& other);
  virtual void FreezeNaturalGradient(bool freeze) {}
  BaseFloat LearningRate() const { return learning_rate_; }
  BaseFloat MaxChange() const { return max_change_; }
  void SetMaxChange(BaseFloat max_change) { max_change_ = max_change; }
  BaseFloat L2Regularization() const { return l2_regularize_; }
  void SetL2Regularization(BaseFloat a) { l2_regularize_ = a; }
  virtual std::string Info() const;
  virtual int32 NumParameters() const { KALDI_ASSERT(0); return 0; }
  virtual void Vectorize(std::vector<BaseFloat>* params) const { KALDI_ASSERT(0); }
  virtual void UnVectorize(const std::vector<BaseFloat>& params) { KALDI_ASSERT(0); }
 protected:
  void InitLearningRatesFromConfig(ConfigLine* cfl);
  std::string ReadUpdatableCommon(std::istream& is, bool binary);
  void WriteUpdatableCommon(std::ostream& is, bool binary) const;
  BaseFloat learning_rate_, learning_rate_factor_, l2_regularize_;
  bool is_gradient_;
  BaseFloat max_change_;
};

// Summaries used by Info() (nnet3/nnet-parse.h: PrintParameterStats / SummarizeVector), host side.
std::string SummarizeVector(const std::vector<BaseFloat>& v);
void PrintParameterStats(std::ostringstream& os, const std::string& name, const CuVector& params,
                         bool include_mean = false);
void PrintParameterStats(std::ostringstream& os, const std::string& name, const CuMatrix& params,
                         bool include_mean = false);

}  // namespace nnet3
}  // namespace tdnnf
