// Implementation of the Kaldi stand-ins declared in shim.h: token I/O in Kaldi's text and binary
// formats (base/io-funcs{,-inl}.h, matrix/kaldi-{vector,matrix}.cc), ConfigLine
// (util/text-utils.cc), owning device buffers, the UpdatableComponent common code (itf.cc:313-431).
#include "shim.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cerrno>
#include <iomanip>
#include <cctype>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace tdnnf {
namespace nnet3 {

// ------------------------------------------------------------------ logging / context
static int32 g_verbose = 0;
int32 GetVerboseLevel() { return g_verbose; }
void SetVerboseLevel(int32 v) { g_verbose = v; }
void KaldiLog(const std::string& msg) { std::cerr << "LOG (tdnnf-nnet3) " << msg << std::endl; }
void KaldiWarn(const std::string& msg) { std::cerr << "WARNING (tdnnf-nnet3) " << msg << std::endl; }

static thread_local tdnnf_ctx* g_ctx = nullptr;
tdnnf_ctx* CurrentContext() {
  if (!g_ctx) KALDI_ERR << "no tdnnf context selected (call SetCurrentContext / tdnnf_nnet3_set_context first)";
  return g_ctx;
}
void SetCurrentContext(tdnnf_ctx* c) { g_ctx = c; }
static tdnnf_ctx* CurrentContextOrNull() { return g_ctx; }
void CheckStatus(int rc) {
  if (rc != TDNNF_OK) KALDI_ERR << "tdnnf kernel call failed (" << rc << "): " << tdnnf_last_error();
}
static void CudaOk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) KALDI_ERR << what << ": " << cudaGetErrorString(e);
}

// ------------------------------------------------------------------ randomness (splitmix64 counter hash)
static uint64_t g_seed = 0, g_counter = 0;
void SetRandSeed(uint64_t seed) { g_seed = seed; g_counter = 0; }
uint64_t GetRandSeed() { return g_seed; }
void SetRandCounter(uint64_t c) { g_counter = c; }
uint64_t GetRandCounter() { return g_counter; }
static uint64_t Mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
float RandUniformOpen() {
  const uint64_t r = Mix(Mix(g_seed) ^ (g_counter++ * 0xD1342543DE82EF95ull));
  // 24 random bits -> (0,1): (k + 0.5) / 2^24 never returns 0 or 1
  return ((float)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
}
int32 RandInt(int32 lo, int32 hi) {
  const uint64_t r = Mix(Mix(g_seed ^ 0xABCDEFull) ^ (g_counter++ * 0xD1342543DE82EF95ull));
  return lo + (int32)(r % (uint64_t)(hi - lo + 1));
}
static double RandGauss() {
  const double u1 = RandUniformOpen(), u2 = RandUniformOpen();
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
}

// ------------------------------------------------------------------ token I/O
void WriteToken(std::ostream& os, bool binary, const std::string& token) {
  KALDI_ASSERT(!token.empty());
  os << token << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteToken.";
}
void ReadToken(std::istream& is, bool binary, std::string* str) {
  if (!binary) is >> std::ws;
  is >> *str;
  if (is.fail()) KALDI_ERR << "ReadToken, failed to read token at file position " << is.tellg();
  if (!isspace(is.peek())) KALDI_ERR << "ReadToken, expected space after token, saw instead " << (char)is.peek();
  is.get();
}
int PeekToken(std::istream& is, bool binary) {
  if (!binary) is >> std::ws;
  bool read_bracket;
  if ((char)is.peek() == '<') { read_bracket = true; is.get(); } else { read_bracket = false; }
  int ans = is.peek();
  if (read_bracket) {
    if (!is.unget()) is.clear();
  }
  return ans;
}
void ExpectToken(std::istream& is, bool binary, const std::string& token) {
  int pos_at_start = is.tellg();
  if (!binary) is >> std::ws;
  std::string str;
  is >> str;
  is.get();
  if (is.fail()) KALDI_ERR << "Failed to read token [started at file position " << pos_at_start << "], expected " << token;
  if (str != token) KALDI_ERR << "Expected token \"" << token << "\", got instead \"" << str << "\".";
}
void ExpectOneOrTwoTokens(std::istream& is, bool binary, const std::string& token1, const std::string& token2) {
  KALDI_ASSERT(token1 != token2);
  std::string temp;
  ReadToken(is, binary, &temp);
  if (temp == token1) ExpectToken(is, binary, token2);
  else if (temp != token2) KALDI_ERR << "Expecting token " << token1 << " or " << token2 << " but got " << temp;
}

void WriteBasicType(std::ostream& os, bool binary, bool b) {
  os << (b ? "T" : "F");
  if (!binary) os << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteBasicType<bool>";
}
void ReadBasicType(std::istream& is, bool binary, bool* b) {
  if (!binary) is >> std::ws;
  char c = is.peek();
  if (c == 'T') { *b = true; is.get(); }
  else if (c == 'F') { *b = false; is.get(); }
  else KALDI_ERR << "Read failure in ReadBasicType<bool>, file position is " << is.tellg() << ", next char is " << (int)c;
}
template <class T>
static void WriteBin(std::ostream& os, T v) {
  os.put((char)sizeof(T));
  os.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
template <class T>
static void ReadBin(std::istream& is, T* v, const char* what) {
  int len = is.get();
  if (len != (int)sizeof(T)) KALDI_ERR << "ReadBasicType: expected " << what << " of size " << sizeof(T) << ", saw size byte " << len;
  is.read(reinterpret_cast<char*>(v), sizeof(T));
  if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType (" << what << ")";
}
void WriteBasicType(std::ostream& os, bool binary, int32 v) {
  if (binary) WriteBin(os, v); else os << v << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteBasicType.";
}
void WriteBasicType(std::ostream& os, bool binary, float v) {
  if (binary) WriteBin(os, v); else os << v << " ";
}
void WriteBasicType(std::ostream& os, bool binary, double v) {
  if (binary) WriteBin(os, v); else os << v << " ";
}
void ReadBasicType(std::istream& is, bool binary, int32* v) {
  if (binary) { ReadBin(is, v, "int32"); return; }
  is >> *v;
  if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType<int32>, file position is " << is.tellg();
}
template <class T>
static void ReadFloatText(std::istream& is, T* f) {
  // Kaldi accepts inf / nan spellings in text mode.
  is >> std::ws;
  std::string tok;
  is >> tok;
  if (is.fail()) KALDI_ERR << "ReadBasicType: failed to read floating-point value, file position " << is.tellg();
  std::string low(tok);
  std::transform(low.begin(), low.end(), low.begin(), ::tolower);
  if (low == "inf" || low == "infinity" || low == "+inf") { *f = std::numeric_limits<T>::infinity(); return; }
  if (low == "-inf" || low == "-infinity") { *f = -std::numeric_limits<T>::infinity(); return; }
  if (low == "nan" || low == "-nan") { *f = std::numeric_limits<T>::quiet_NaN(); return; }
  char* end = nullptr;
  const double d = std::strtod(tok.c_str(), &end);
  if (end == tok.c_str() || *end != '\0') KALDI_ERR << "ReadBasicType: expected a floating-point value, got \"" << tok << "\"";
  *f = (T)d;
}
void ReadBasicType(std::istream& is, bool binary, float* f) {
  if (binary) {
    int c = is.peek();
    if (c == (int)sizeof(float)) { ReadBin(is, f, "float"); }
    else if (c == (int)sizeof(double)) { double d; ReadBin(is, &d, "double"); *f = (float)d; }
    else KALDI_ERR << "ReadBasicType: expected float, saw " << c << ", at file position " << is.tellg();
  } else {
    ReadFloatText(is, f);
  }
}
void ReadBasicType(std::istream& is, bool binary, double* d) {
  if (binary) {
    int c = is.peek();
    if (c == (int)sizeof(double)) { ReadBin(is, d, "double"); }
    else if (c == (int)sizeof(float)) { float f; ReadBin(is, &f, "float"); *d = f; }
    else KALDI_ERR << "ReadBasicType: expected double, saw " << c << ", at file position " << is.tellg();
  } else {
    ReadFloatText(is, d);
  }
}
void WriteIntegerVector(std::ostream& os, bool binary, const std::vector<int32>& v) {
  if (binary) {
    char sz = sizeof(int32);
    os.write(&sz, 1);
    int32 vecsz = (int32)v.size();
    os.write(reinterpret_cast<const char*>(&vecsz), sizeof(vecsz));
    if (vecsz != 0) os.write(reinterpret_cast<const char*>(v.data()), sizeof(int32) * vecsz);
  } else {
    os << "[ ";
    for (int32 x : v) os << x << " ";
    os << "]\n";
  }
  if (os.fail()) KALDI_ERR << "Write failure in WriteIntegerVector.";
}
void ReadIntegerVector(std::istream& is, bool binary, std::vector<int32>* v) {
  if (binary) {
    int sz = is.peek();
    if (sz == (int)sizeof(int32)) is.get();
    else KALDI_ERR << "ReadIntegerVector: expected to see type of size " << sizeof(int32) << ", saw instead " << sz;
    int32 vecsz;
    is.read(reinterpret_cast<char*>(&vecsz), sizeof(vecsz));
    if (is.fail() || vecsz < 0) KALDI_ERR << "ReadIntegerVector: read failure";
    v->resize(vecsz);
    if (vecsz > 0) is.read(reinterpret_cast<char*>(v->data()), sizeof(int32) * vecsz);
  } else {
    std::vector<int32> tmp;
    is >> std::ws;
    if (is.peek() != (int)'[') KALDI_ERR << "ReadIntegerVector: expected to see [, saw " << (char)is.peek();
    is.get();
    is >> std::ws;
    while (is.peek() != (int)']') {
      int32 next;
      is >> next >> std::ws;
      if (is.fail()) KALDI_ERR << "ReadIntegerVector: read failure";
      tmp.push_back(next);
    }
    is.get();
    *v = tmp;
  }
  if (is.fail()) KALDI_ERR << "ReadIntegerVector: read failure at file position " << is.tellg();
}
bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32>* out) {
  out->clear();
  if (full.empty()) return true;
  size_t start = 0;
  while (true) {
    size_t end = full.find_first_of(delim, start);
    std::string piece = full.substr(start, end == std::string::npos ? std::string::npos : end - start);
    if (!(omit_empty && piece.empty())) {
      char* e = nullptr;
      errno = 0;
      long v = std::strtol(piece.c_str(), &e, 10);
      if (piece.empty() || e == piece.c_str() || *e != '\0' || errno != 0) { out->clear(); return false; }
      out->push_back((int32)v);
    }
    if (end == std::string::npos) break;
    start = end + 1;
  }
  return true;
}

// ------------------------------------------------------------------ Vector / Matrix I/O
template <typename Real>
static const char* TypeToken(bool matrix) {
  return sizeof(Real) == 4 ? (matrix ? "FM" : "FV") : (matrix ? "DM" : "DV");
}
template <typename Real>
void Vector<Real>::Write(std::ostream& os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, TypeToken<Real>(false));
    WriteBasicType(os, binary, Dim());
    os.write(reinterpret_cast<const char*>(v.data()), sizeof(Real) * v.size());
  } else {
    os << " [ ";
    for (Real x : v) os << x << " ";
    os << "]\n";
  }
  if (os.fail()) KALDI_ERR << "Failed to write vector to stream";
}
template <typename Real, typename Other>
static void ReadRaw(std::istream& is, size_t n, std::vector<Real>* out) {
  std::vector<Other> tmp(n);
  is.read(reinterpret_cast<char*>(tmp.data()), sizeof(Other) * n);
  out->resize(n);
  for (size_t i = 0; i < n; ++i) (*out)[i] = (Real)tmp[i];
}
template <typename Real>
void Vector<Real>::Read(std::istream& is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    int32 size;
    if (tok != "FV" && tok != "DV") KALDI_ERR << "Vector::Read: expected token FV or DV, got " << tok;
    ReadBasicType(is, binary, &size);
    if (size < 0) KALDI_ERR << "Vector::Read: negative size";
    if (tok == "FV") ReadRaw<Real, float>(is, size, &v); else ReadRaw<Real, double>(is, size, &v);
    if (is.fail()) KALDI_ERR << "Vector::Read: read failure";
    return;
  }
  std::string s;
  is >> s;
  if (is.fail() || s != "[") KALDI_ERR << "Failed to read vector from stream. Expected \"[\" but got " << s;
  v.clear();
  while (true) {
    is >> std::ws;
    int c = is.peek();
    if (c == ']') { is.get(); break; }
    if (c == EOF || is.fail()) KALDI_ERR << "Failed to read vector from stream: EOF";
    Real x;
    ReadFloatText(is, &x);
    // ReadFloatText consumes a whole whitespace-delimited token; "1.0]" style is not produced by Kaldi
    v.push_back(x);
  }
  // Kaldi consumes the rest of the line after "]"
  if (is.peek() == '\r') is.get();
  if (is.peek() == '\n') is.get();
}
template <typename Real>
void Matrix<Real>::Write(std::ostream& os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, TypeToken<Real>(true));
    WriteBasicType(os, binary, rows);
    WriteBasicType(os, binary, cols);
    os.write(reinterpret_cast<const char*>(v.data()), sizeof(Real) * v.size());
  } else {
    if (cols == 0) {
      os << " [ ]\n";
    } else {
      os << " [";
      for (int32 i = 0; i < rows; ++i) {
        os << "\n  ";
        for (int32 j = 0; j < cols; ++j) os << (*this)(i, j) << " ";
      }
      os << "]\n";
    }
  }
  if (os.fail()) KALDI_ERR << "Failed to write matrix to stream";
}
template <typename Real>
void Matrix<Real>::Read(std::istream& is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    if (tok != "FM" && tok != "DM") KALDI_ERR << "Matrix::Read: expected token FM or DM, got " << tok << " (compressed matrices are not supported)";
    int32 r, c;
    ReadBasicType(is, binary, &r);
    ReadBasicType(is, binary, &c);
    if (r < 0 || c < 0) KALDI_ERR << "Matrix::Read: negative size";
    rows = r; cols = c;
    if (tok == "FM") ReadRaw<Real, float>(is, (size_t)r * c, &v); else ReadRaw<Real, double>(is, (size_t)r * c, &v);
    if (is.fail()) KALDI_ERR << "Matrix::Read: read failure";
    return;
  }
  std::string s;
  is >> s;
  if (is.fail() || s != "[") KALDI_ERR << "Failed to read matrix from stream. Expected \"[\" but got " << s;
  std::vector<std::vector<Real> > data;
  std::vector<Real> cur;
  while (true) {
    int c = is.peek();
    if (c == EOF || is.fail()) KALDI_ERR << "Failed to read matrix from stream: EOF";
    if (c == ']') {
      is.get();
      if (!cur.empty()) data.push_back(cur);
      break;
    } else if (c == '\n' || c == ';') {
      is.get();
      if (!cur.empty()) { data.push_back(cur); cur.clear(); }
    } else if (isspace(c)) {
      is.get();
    } else {
      Real x;
      // read one number without swallowing a following newline
      std::string tok;
      while (true) {
        int d = is.peek();
        if (d == EOF || isspace(d) || d == ']' || d == ';') break;
        tok.push_back((char)is.get());
      }
      std::istringstream ts(tok + " ");
      ReadFloatText(ts, &x);
      cur.push_back(x);
    }
  }
  if (is.peek() == '\r') is.get();
  if (is.peek() == '\n') is.get();
  rows = (int32)data.size();
  cols = rows ? (int32)data[0].size() : 0;
  v.assign((size_t)rows * cols, Real(0));
  for (int32 i = 0; i < rows; ++i) {
    if ((int32)data[i].size() != cols) KALDI_ERR << "Matrix::Read: rows of differing length";
    for (int32 j = 0; j < cols; ++j) (*this)(i, j) = data[i][j];
  }
}
template struct Vector<float>;
template struct Vector<double>;
template struct Matrix<float>;
template struct Matrix<double>;

// ------------------------------------------------------------------ device buffers
// Size-bucketed free lists in front of cudaMalloc (the role of Kaldi's CuAllocator).  The components create small
// device vectors every minibatch (the memo of Propagate, the s_i of UpdateNaturalGradient): a cudaFree per object
// is a device-wide synchronisation -- 84 pipeline drains per training step of the bench supernet -- and cudaMalloc is
// slow.  A freed block goes back to its bucket and is handed out again; reuse is ordered by the stream all the work
// runs on.  Nothing is returned to the driver before process exit.
namespace {
class DevicePool {
 public:
  static DevicePool& Get() {
    static DevicePool* p = new DevicePool();  // leaked on purpose: outlives every static CuVector
    return *p;
  }
  // Arena scope: between ArenaBegin and ArenaEnd every allocation of this thread is carved (256-byte aligned, in call
  // order) from a caller-owned device range, so that e.g. the parameters of all the delta components of a network
  // form ONE contiguous range for the data-parallel all-reduce.  Arena blocks are never recycled by the pool.
  void ArenaBegin(void* base, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu_);
    arena_base_ = static_cast<char*>(base);
    arena_bytes_ = bytes;
    arena_off_ = 0;
    arenas_.push_back(std::make_pair(arena_base_, bytes));
  }
  size_t ArenaEnd() {
    std::lock_guard<std::mutex> lk(mu_);
    const size_t used = arena_off_;
    arena_base_ = nullptr;
    arena_bytes_ = arena_off_ = 0;
    return used;
  }
  void* Alloc(size_t bytes) {
    const Key k = MakeKey(bytes);
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (arena_base_ != nullptr) {
        const size_t need = (bytes + 255) & ~(size_t)255;
        if (arena_off_ + need > arena_bytes_) KALDI_ERR << "device arena exhausted (" << arena_bytes_ << " bytes)";
        void* p = arena_base_ + arena_off_;
        arena_off_ += need;
        return p;
      }
      auto it = free_.find(k);
      if (it != free_.end() && !it->second.empty()) {
        void* p = it->second.back();
        it->second.pop_back();
        return p;
      }
    }
    void* p = nullptr;
    CudaOk(cudaMalloc(&p, k.second), "cudaMalloc");
    return p;
  }
  void Free(void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu_);
    for (const auto& a : arenas_)
      if (static_cast<char*>(p) >= a.first && static_cast<char*>(p) < a.first + a.second) return;  // caller-owned
    free_[MakeKey(bytes)].push_back(p);
  }

 private:
  typedef std::pair<int, size_t> Key;  // (device, rounded bytes)
  static Key MakeKey(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    size_t r = 256;
    while (r < bytes) r <<= 1;                           // powers of two up to 1 MiB, then 1 MiB steps
    if (bytes > (1u << 20)) r = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    return Key(dev, r);
  }
  std::mutex mu_;
  std::map<Key, std::vector<void*>> free_;
  char* arena_base_ = nullptr;
  size_t arena_bytes_ = 0, arena_off_ = 0;
  std::vector<std::pair<char*, size_t>> arenas_;
};

// The stream the kernels run on (the legacy default stream when no context is selected): allocation zero-fills and
// the copies below are ordered on it, so that a block handed out again by the pool is never touched out of order.
cudaStream_t WorkStream() {
  void* st = nullptr;
  tdnnf_ctx* ctx = CurrentContextOrNull();
  if (ctx != nullptr && tdnnf_ctx_get_stream(ctx, &st) == TDNNF_OK) return static_cast<cudaStream_t>(st);
  return nullptr;
}
void ZeroFill(void* p, size_t bytes) { CudaOk(cudaMemsetAsync(p, 0, bytes, WorkStream()), "cudaMemsetAsync"); }
}  // namespace

void DeviceArenaBegin(void* base, size_t bytes) {
  KALDI_ASSERT(base != nullptr && (reinterpret_cast<uintptr_t>(base) & 255) == 0);
  DevicePool::Get().ArenaBegin(base, bytes);
}
size_t DeviceArenaEnd() { return DevicePool::Get().ArenaEnd(); }

CuVector::~CuVector() { DevicePool::Get().Free(data_, sizeof(BaseFloat) * (size_t)dim_); }
void CuVector::Resize(int32 dim) {
  DevicePool::Get().Free(data_, sizeof(BaseFloat) * (size_t)dim_);
  data_ = nullptr;
  dim_ = dim;
  if (dim > 0) {
    data_ = static_cast<BaseFloat*>(DevicePool::Get().Alloc(sizeof(BaseFloat) * (size_t)dim));
    ZeroFill(data_, sizeof(BaseFloat) * (size_t)dim);
  }
}
CuVector::CuVector(const CuVector& o) { *this = o; }
CuVector& CuVector::operator=(const CuVector& o) {
  if (this == &o) return *this;
  Resize(o.dim_);
  if (dim_ > 0) CudaOk(cudaMemcpyAsync(data_, o.data_, sizeof(BaseFloat) * dim_, cudaMemcpyDeviceToDevice, WorkStream()), "cudaMemcpyAsync");
  return *this;
}
void CuVector::CopyFromHost(const std::vector<BaseFloat>& h) {
  if ((int32)h.size() != dim_) Resize((int32)h.size());
  if (dim_ > 0) {
    CudaOk(cudaMemcpyAsync(data_, h.data(), sizeof(BaseFloat) * dim_, cudaMemcpyHostToDevice, WorkStream()), "cudaMemcpyAsync");
    CudaOk(cudaStreamSynchronize(WorkStream()), "cudaStreamSynchronize");  // h may be a temporary
  }
}
std::vector<BaseFloat> CuVector::ToHost() const {
  std::vector<BaseFloat> h(dim_);
  if (dim_ > 0) {
    CudaOk(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    CudaOk(cudaMemcpy(h.data(), data_, sizeof(BaseFloat) * dim_, cudaMemcpyDeviceToHost), "cudaMemcpy");
  }
  return h;
}
void CuVector::SetZero() { if (dim_) CheckStatus(tdnnf_mat_set(CurrentContext(), data_, 1, dim_, dim_, 0.f)); }
void CuVector::Scale(BaseFloat s) { if (dim_) CheckStatus(tdnnf_mat_scale(CurrentContext(), data_, 1, dim_, dim_, s)); }
void CuVector::AddVec(BaseFloat alpha, const CuVector& o) {
  KALDI_ASSERT(o.dim_ == dim_);
  if (dim_) CheckStatus(tdnnf_mat_axpy(CurrentContext(), alpha, o.data_, dim_, data_, dim_, 1, dim_));
}
void CuVector::Write(std::ostream& os, bool binary) const {
  Vector<BaseFloat> h;
  h.v = ToHost();
  h.Write(os, binary);
}
void CuVector::Read(std::istream& is, bool binary) {
  Vector<BaseFloat> h;
  h.Read(is, binary);
  CopyFromHost(h.v);
}
BaseFloat VecVec(const CuVector& a, const CuVector& b) {
  KALDI_ASSERT(a.Dim() == b.Dim());
  float r = 0.f;
  if (a.Dim()) CheckStatus(tdnnf_mat_dot(CurrentContext(), a.Data(), a.Dim(), b.Data(), b.Dim(), 1, a.Dim(), &r));
  return r;
}

CuMatrix::~CuMatrix() { DevicePool::Get().Free(data_, sizeof(BaseFloat) * (size_t)num_rows_ * stride_); }
void CuMatrix::Resize(int32 rows, int32 cols) {
  DevicePool::Get().Free(data_, sizeof(BaseFloat) * (size_t)num_rows_ * stride_);
  data_ = nullptr;
  num_rows_ = rows;
  num_cols_ = cols;
  stride_ = (cols + 63) / 64 * 64;  // 256-byte pitch like cudaMallocPitch
  if (rows > 0 && cols > 0) {
    data_ = static_cast<BaseFloat*>(DevicePool::Get().Alloc(sizeof(BaseFloat) * (size_t)rows * stride_));
    ZeroFill(data_, sizeof(BaseFloat) * (size_t)rows * stride_);
  }
}
CuMatrix::CuMatrix(const CuMatrix& o) : CuMatrixBase<BaseFloat>() { *this = o; }
CuMatrix& CuMatrix::operator=(const CuMatrix& o) {
  if (this == &o) return *this;
  Resize(o.num_rows_, o.num_cols_);
  if (data_) CudaOk(cudaMemcpy2DAsync(data_, sizeof(BaseFloat) * stride_, o.data_, sizeof(BaseFloat) * o.stride_,
                                      sizeof(BaseFloat) * num_cols_, num_rows_, cudaMemcpyDeviceToDevice, WorkStream()), "cudaMemcpy2DAsync");
  return *this;
}
void CuMatrix::CopyFromHost(const Matrix<BaseFloat>& h) {
  if (h.rows != num_rows_ || h.cols != num_cols_) Resize(h.rows, h.cols);
  if (data_) {
    CudaOk(cudaMemcpy2DAsync(data_, sizeof(BaseFloat) * stride_, h.v.data(), sizeof(BaseFloat) * h.cols,
                             sizeof(BaseFloat) * h.cols, h.rows, cudaMemcpyHostToDevice, WorkStream()), "cudaMemcpy2DAsync");
    CudaOk(cudaStreamSynchronize(WorkStream()), "cudaStreamSynchronize");  // h may be a temporary
  }
}
Matrix<BaseFloat> CuMatrix::ToHost() const {
  Matrix<BaseFloat> h(num_rows_, num_cols_);
  if (data_) {
    CudaOk(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    CudaOk(cudaMemcpy2D(h.v.data(), sizeof(BaseFloat) * h.cols, data_, sizeof(BaseFloat) * stride_,
                        sizeof(BaseFloat) * h.cols, h.rows, cudaMemcpyDeviceToHost), "cudaMemcpy2D");
  }
  return h;
}
void CuMatrix::SetZero() { if (data_) CheckStatus(tdnnf_mat_set(CurrentContext(), data_, num_rows_, num_cols_, stride_, 0.f)); }
void CuMatrix::Scale(BaseFloat s) { if (data_) CheckStatus(tdnnf_mat_scale(CurrentContext(), data_, num_rows_, num_cols_, stride_, s)); }
void CuMatrix::AddMat(BaseFloat alpha, const CuMatrix& o) {
  KALDI_ASSERT(SameDim(*this, o));
  if (data_) CheckStatus(tdnnf_mat_axpy(CurrentContext(), alpha, o.Data(), o.Stride(), data_, stride_, num_rows_, num_cols_));
}
void CuMatrix::Write(std::ostream& os, bool binary) const { ToHost().Write(os, binary); }
void CuMatrix::Read(std::istream& is, bool binary) {
  Matrix<BaseFloat> h;
  h.Read(is, binary);
  CopyFromHost(h);
}
BaseFloat TraceMatMatTrans(const CuMatrix& a, const CuMatrix& b) {
  KALDI_ASSERT(SameDim(a, b));
  float r = 0.f;
  if (a.NumRows()) CheckStatus(tdnnf_mat_dot(CurrentContext(), a.Data(), a.Stride(), b.Data(), b.Stride(), a.NumRows(), a.NumCols(), &r));
  return r;
}

// ------------------------------------------------------------------ ConfigLine
static bool IsValidName(const std::string& name) {
  if (name.empty()) return false;
  for (size_t i = 0; i < name.size(); ++i) {
    const char c = name[i];
    if (i == 0 && !isalpha(c) && c != '_') return false;
    if (!isalnum(c) && c != '_' && c != '-' && c != '.') return false;
  }
  return true;
}
bool ConfigLine::ParseLine(const std::string& line) {
  data_.clear();
  whole_line_ = line;
  if (line.empty()) return false;
  size_t pos = 0, size = line.size();
  while (pos < size && isspace(line[pos])) pos++;
  if (pos == size) return false;
  size_t first_token_start = pos;
  while (pos < size && !isspace(line[pos]) && line[pos] != '=') pos++;
  if (pos < size && line[pos] == '=') {
    pos = first_token_start;  // no first token: the line starts with key=value
    first_token_ = "";
  } else {
    first_token_ = line.substr(first_token_start, pos - first_token_start);
    if (!IsValidName(first_token_)) return false;
  }
  while (pos < size) {
    while (pos < size && isspace(line[pos])) pos++;
    if (pos == size) break;
    size_t key_start = pos;
    while (pos < size && line[pos] != '=' && !isspace(line[pos])) pos++;
    if (pos == size || line[pos] != '=') return false;
    std::string key = line.substr(key_start, pos - key_start);
    if (!IsValidName(key)) return false;
    pos++;  // '='
    std::string value;
    if (pos < size && (line[pos] == '"' || line[pos] == '\'')) {
      const char q = line[pos++];
      size_t vstart = pos;
      while (pos < size && line[pos] != q) pos++;
      if (pos == size) return false;
      value = line.substr(vstart, pos - vstart);
      pos++;
    } else {
      size_t vstart = pos;
      while (pos < size && !isspace(line[pos])) pos++;
      value = line.substr(vstart, pos - vstart);
    }
    if (data_.count(key)) return false;  // repeated key
    data_[key] = std::make_pair(value, false);
  }
  return true;
}
bool ConfigLine::GetValue(const std::string& key, std::string* value) {
  auto it = data_.find(key);
  if (it == data_.end()) return false;
  *value = it->second.first;
  it->second.second = true;
  return true;
}
bool ConfigLine::GetValue(const std::string& key, BaseFloat* value) {
  auto it = data_.find(key);
  if (it == data_.end()) return false;
  char* e = nullptr;
  const double d = std::strtod(it->second.first.c_str(), &e);
  if (e == it->second.first.c_str() || *e != '\0') return false;
  *value = (BaseFloat)d;
  it->second.second = true;
  return true;
}
bool ConfigLine::GetValue(const std::string& key, int32* value) {
  auto it = data_.find(key);
  if (it == data_.end()) return false;
  char* e = nullptr;
  const long v = std::strtol(it->second.first.c_str(), &e, 10);
  if (e == it->second.first.c_str() || *e != '\0') return false;
  *value = (int32)v;
  it->second.second = true;
  return true;
}
bool ConfigLine::GetValue(const std::string& key, std::vector<int32>* value) {
  auto it = data_.find(key);
  if (it == data_.end()) return false;
  if (!SplitStringToIntegers(it->second.first, ":,", true, value)) return false;
  it->second.second = true;
  return true;
}
bool ConfigLine::GetValue(const std::string& key, bool* value) {
  auto it = data_.find(key);
  if (it == data_.end()) return false;
  const std::string& s = it->second.first;
  if (s.empty()) return false;
  if (s[0] == 't' || s[0] == 'T') *value = true;
  else if (s[0] == 'f' || s[0] == 'F') *value = false;
  else return false;
  it->second.second = true;
  return true;
}
bool ConfigLine::HasUnusedValues() const {
  for (auto& kv : data_) if (!kv.second.second) return true;
  return false;
}
std::string ConfigLine::UnusedValues() const {
  std::string unused;
  for (auto& kv : data_)
    if (!kv.second.second) {
      if (!unused.empty()) unused += " ";
      unused += kv.first + "=" + kv.second.first;
    }
  return unused;
}

// ------------------------------------------------------------------ Component / UpdatableComponent common code
std::string Component::Info() const {
  std::stringstream stream;
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim();
  return stream.str();
}
void Component::GetInputIndexes(const MiscComputationInfo&, const Index& output_index, std::vector<Index>* input_indexes) const {
  input_indexes->resize(1);
  (*input_indexes)[0] = output_index;
}
bool Component::IsComputable(const MiscComputationInfo&, const Index& output_index, const IndexSet& input_index_set,
                             std::vector<Index>* used_inputs) const {  // itf.cc:296-310
  if (!input_index_set(output_index)) return false;
  if (used_inputs) {
    used_inputs->clear();
    used_inputs->push_back(output_index);
  }
  return true;
}
Component* Component::ReadNew(std::istream& is, bool binary) {  // itf.cc:106-124
  std::string token;
  ReadToken(is, binary, &token);
  if (token.size() < 3) KALDI_ERR << "Invalid token " << token;
  token.erase(0, 1);
  token.erase(token.length() - 1);
  Component* ans = NewComponentOfType(token);
  if (!ans) KALDI_ERR << "Unknown component type " << token;
  ans->Read(is, binary);
  return ans;
}
ComponentPrecomputedIndexes* ComponentPrecomputedIndexes::ReadNew(std::istream& is, bool binary) {  // itf.cc:38-53
  std::string token;
  ReadToken(is, binary, &token);
  token.erase(0, 1);
  token.erase(token.length() - 1);
  ComponentPrecomputedIndexes* ans = NewComponentPrecomputedIndexesOfType(token);
  if (!ans) KALDI_ERR << "Unknown ComponentPrecomputedIndexes type " << token;
  ans->Read(is, binary);
  return ans;
}

UpdatableComponent::UpdatableComponent(const UpdatableComponent& other)
    : learning_rate_(other.learning_rate_), learning_rate_factor_(other.learning_rate_factor_),
      l2_regularize_(other.l2_regularize_), is_gradient_(other.is_gradient_), max_change_(other.max_change_) {}
void UpdatableComponent::SetUpdatableConfigs(const UpdatableComponent& other) {
  learning_rate_ = other.learning_rate_;
  learning_rate_factor_ = other.learning_rate_factor_;
  l2_regularize_ = other.l2_regularize_;
  is_gradient_ = other.is_gradient_;
  max_change_ = other.max_change_;
}
void UpdatableComponent::InitLearningRatesFromConfig(ConfigLine* cfl) {  // itf.cc:330-344
  learning_rate_ = 0.001;
  cfl->GetValue("learning-rate", &learning_rate_);
  learning_rate_factor_ = 1.0;
  cfl->GetValue("learning-rate-factor", &learning_rate_factor_);
  max_change_ = 0.0;
  cfl->GetValue("max-change", &max_change_);
  l2_regularize_ = 0.0;
  cfl->GetValue("l2-regularize", &l2_regularize_);
  if (learning_rate_ < 0.0 || learning_rate_factor_ < 0.0 || max_change_ < 0.0 || l2_regularize_ < 0.0)
    KALDI_ERR << "Bad initializer " << cfl->WholeLine();
}
std::string UpdatableComponent::ReadUpdatableCommon(std::istream& is, bool binary) {  // itf.cc:347-388
  std::ostringstream opening_tag;
  opening_tag << '<' << this->Type() << '>';
  std::string token;
  ReadToken(is, binary, &token);
  if (token == opening_tag.str()) ReadToken(is, binary, &token);
  if (token == "<LearningRateFactor>") { ReadBasicType(is, binary, &learning_rate_factor_); ReadToken(is, binary, &token); }
  else learning_rate_factor_ = 1.0;
  if (token == "<IsGradient>") { ReadBasicType(is, binary, &is_gradient_); ReadToken(is, binary, &token); }
  else is_gradient_ = false;
  if (token == "<MaxChange>") { ReadBasicType(is, binary, &max_change_); ReadToken(is, binary, &token); }
  else max_change_ = 0.0;
  if (token == "<L2Regularize>") { ReadBasicType(is, binary, &l2_regularize_); ReadToken(is, binary, &token); }
  else l2_regularize_ = 0.0;
  if (token == "<LearningRate>") { ReadBasicType(is, binary, &learning_rate_); return ""; }
  return token;
}
void UpdatableComponent::WriteUpdatableCommon(std::ostream& os, bool binary) const {  // itf.cc:390-414
  std::ostringstream opening_tag;
  opening_tag << '<' << this->Type() << '>';
  WriteToken(os, binary, opening_tag.str());
  if (learning_rate_factor_ != 1.0) { WriteToken(os, binary, "<LearningRateFactor>"); WriteBasicType(os, binary, learning_rate_factor_); }
  if (is_gradient_) { WriteToken(os, binary, "<IsGradient>"); WriteBasicType(os, binary, is_gradient_); }
  if (max_change_ > 0.0) { WriteToken(os, binary, "<MaxChange>"); WriteBasicType(os, binary, max_change_); }
  if (l2_regularize_ > 0.0) { WriteToken(os, binary, "<L2Regularize>"); WriteBasicType(os, binary, l2_regularize_); }
  WriteToken(os, binary, "<LearningRate>");
  WriteBasicType(os, binary, learning_rate_);
}
std::string UpdatableComponent::Info() const {  // itf.cc:417-431
  std::stringstream stream;
  stream << Type() << ", input-dim=" << InputDim() << ", output-dim=" << OutputDim() << ", learning-rate=" << LearningRate();
  if (is_gradient_) stream << ", is-gradient=true";
  if (l2_regularize_ != 0.0) stream << ", l2-regularize=" << l2_regularize_;
  if (learning_rate_factor_ != 1.0) stream << ", learning-rate-factor=" << learning_rate_factor_;
  if (max_change_ > 0.0) stream << ", max-change=" << max_change_;
  return stream.str();
}

// ------------------------------------------------------------------ Info() helpers
std::string SummarizeVector(const std::vector<BaseFloat>& v) {
  std::ostringstream os;
  if (v.size() <= 10) {
    os << "[ ";
    for (BaseFloat x : v) os << x << " ";
    os << "]";
  } else {
    std::vector<BaseFloat> s(v);
    std::sort(s.begin(), s.end());
    double mean = 0, sq = 0;
    for (BaseFloat x : v) { mean += x; sq += (double)x * x; }
    mean /= v.size();
    const double stddev = std::sqrt(std::max(0.0, sq / v.size() - mean * mean));
    auto pct = [&](int p) { return s[std::min(s.size() - 1, (size_t)((p * s.size()) / 100))]; };
    os << "[percentiles(0,1,2,5 10,20,50,80,90 95,98,99,100)=(" << s.front() << " " << pct(1) << " " << pct(2) << " " << pct(5)
       << "  " << pct(10) << " " << pct(20) << " " << pct(50) << " " << pct(80) << " " << pct(90) << "  " << pct(95) << " "
       << pct(98) << " " << pct(99) << " " << s.back() << "), mean=" << mean << ", stddev=" << stddev << "]";
  }
  return os.str();
}
void PrintParameterStats(std::ostringstream& os, const std::string& name, const CuVector& params, bool include_mean) {
  const std::vector<BaseFloat> h = params.ToHost();
  os << std::setprecision(4);
  os << ", " << name << '-';
  double sum = 0, sq = 0;
  for (BaseFloat x : h) { sum += x; sq += (double)x * x; }
  const double n = std::max<size_t>(h.size(), 1);
  if (include_mean) {
    const double mean = sum / n;
    os << "{mean,stddev}=" << mean << ',' << std::sqrt(std::max(0.0, sq / n - mean * mean));
  } else {
    os << "rms=" << std::sqrt(sq / n);
  }
  os << std::setprecision(6);
}
void PrintParameterStats(std::ostringstream& os, const std::string& name, const CuMatrix& params, bool include_mean) {
  const Matrix<BaseFloat> h = params.ToHost();
  os << std::setprecision(4);
  os << ", " << name << '-';
  double sum = 0, sq = 0;
  for (BaseFloat x : h.v) { sum += x; sq += (double)x * x; }
  const double n = std::max<size_t>(h.v.size(), 1);
  if (include_mean) {
    const double mean = sum / n;
    os << "{mean,stddev}=" << mean << ',' << std::sqrt(std::max(0.0, sq / n - mean * mean));
  } else {
    os << "rms=" << std::sqrt(sq / n);
  }
  os << std::setprecision(6);
}

// exposed for components.cc (InitFromConfig's SetRandn)
double ShimRandGauss() { return RandGauss(); }

}  // namespace nnet3
}  // namespace tdnnf
