// OnlineNaturalGradient (kaldi: nnet3/natural-gradient-online.{h,cc}; called by the reference at
// tdnn.cc:598-599 and simple.cc:9542) -- SURVEY.md row N1.
//
// The method keeps a rank-R estimate of the Fisher matrix of the rows of X,
//     F_t = R_t^T D_t R_t + rho_t I,          W_t = E_t^{1/2} R_t,
// preconditions  X_hat = X - (X W_t^T) W_t  and returns scale = sqrt(tr(X X^T) / tr(X_hat X_hat^T)).
// On "updating" calls (the first ten, then every update_period-th) it refreshes (W, d, rho) from
//     H = X W^T,  J = H^T X,  L = H^T H,  K = J J^T      and an R x R symmetric eigenproblem on the host.
//
// B200 form.  Kaldi materialises X (for TdnnDARTSV3 that is the R_out x (n*D_in+1) spliced input, up to
// 860 MB) and overwrites it.  Here X stays IMPLICIT: an NgOperand describes it as the splice of a device
// matrix (views, per-view weights, optional column of ones) and every N-sized product is one call of the
// tcgen05 splice GEMM through the C ABI (with one offset these are plain products):
//     H          = tdnnf_darts_propagate        (W_t as the weights, its last column as the bias)
//     J          = tdnnf_darts_backprop_params  (H as the "output derivative")
//     L, K, WW^T = the same two calls on the small matrices
//     W_{t+1}    = A_t J + (A_t diag(c)) W_t :   tdnnf_darts_backprop_data twice
// X_hat itself is never formed: callers get (W_t, H, scale) and fold the projection into their own update
// (TdnnDARTSV3Component::UpdateNaturalGradient applies it to the D_out x D gradient, components.cc).
// tr(X_hat X_hat^T) = tr(XX^T) - 2 tr(L) + <L, W W^T> needs only R x R quantities, so "scale" is computed on
// the device and never visits the host.  The eigen-update needs L, K and tr(XX^T) on the host (as upstream,
// which runs it on the CPU): they are copied to pinned memory asynchronously and the update is FINISHED LAZILY
// at the next call (W_{t+1} is not needed before then), so steady-state training has no host sync here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <future>
#include <mutex>
#include <thread>

#include "components.h"

namespace tdnnf {
namespace nnet3 {

namespace {

void CudaCheck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) KALDI_ERR << what << ": " << cudaGetErrorString(e);
}

// A few long-lived host threads for the eigen-updates (one task per preconditioner per refresh: 56 per step of the
// bench supernet, ~1-3 ms each).  Long-lived rather than one std::async thread per task: each new thread that touches
// CUDA pays a per-thread runtime initialisation, and tools that instrument CUDA calls see a stable set of threads.
class HostWorkers {
 public:
  static HostWorkers& Get() {
    static HostWorkers* w = new HostWorkers();  // leaked on purpose: no join during static destruction
    return *w;
  }
  std::future<void> Run(std::function<void()> fn) {
    std::packaged_task<void()> task(std::move(fn));
    std::future<void> fut = task.get_future();
    if (threads_.empty()) {  // TDNNF_NG_ASYNC=0: on the calling thread, now
      task();
      return fut;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      queue_.push_back(std::move(task));
    }
    cv_.notify_one();
    return fut;
  }

 private:
  HostWorkers() {
    const char* e = getenv("TDNNF_NG_ASYNC");
    int n = 4;
    if (e) n = atoi(e);
    n = std::max(0, std::min(n, 16));
    for (int i = 0; i < n; ++i) threads_.emplace_back([this] { Loop(); });
    for (std::thread& t : threads_) t.detach();
  }
  void Loop() {
    for (;;) {
      std::packaged_task<void()> task;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return !queue_.empty(); });
        task = std::move(queue_.front());
        queue_.pop_front();
      }
      task();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::packaged_task<void()>> queue_;
  std::vector<std::thread> threads_;
};

cudaStream_t Stream() {
  void* s = nullptr;
  CheckStatus(tdnnf_ctx_get_stream(CurrentContext(), &s));
  return static_cast<cudaStream_t>(s);
}

void EnsureSize(CuMatrix* m, int32 rows, int32 cols) {
  if (m->NumRows() != rows || m->NumCols() != cols) m->Resize(rows, cols);
}

const int32 kZeroOffset[1] = {0};

// fp32-level GEMMs (three bf16 planes per operand, six products) while in scope: everything that feeds the
// eigen-update.  Measured: with the default two planes (~5e-6) the update's cancellations (captured vs total
// energy in rho_{t+1}, the 1/sqrt(c) factors of A_t) amplify the rounding to 1e-4..1e-2 in X_hat.
struct FullPrecisionGemms {
  explicit FullPrecisionGemms(bool on) : on_(on) {
    if (on_ && depth()++ == 0) CheckStatus(tdnnf_ctx_set_gemm_planes(CurrentContext(), 3));
  }
  ~FullPrecisionGemms() {
    if (on_ && --depth() == 0) tdnnf_ctx_set_gemm_planes(CurrentContext(), 2);
  }
  static int& depth() {
    static thread_local int d = 0;
    return d;
  }
  bool on_;
};

// out (rows x w.rows) = in * w^T (+ bias)        -- plain product through the splice GEMM with one offset
void ProductABt(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& w, int32 w_cols, const BaseFloat* one,
                CuMatrixBase<BaseFloat>* out) {
  KALDI_ASSERT(in.NumCols() == w_cols && out->NumRows() == in.NumRows() && out->NumCols() == w.NumRows());
  CheckStatus(tdnnf_darts_propagate(CurrentContext(), in.Data(), in.NumRows(), in.NumCols(), in.Stride(), out->Data(),
                                    out->NumRows(), out->NumCols(), out->Stride(), w.Data(), w.Stride(), NULL, 1, one, 1,
                                    kZeroOffset, 1));
}

// out (a.cols x b.cols) += lr * a^T b      (a, b: same number of rows)
void ProductAtB(const CuMatrixBase<BaseFloat>& a, const CuMatrixBase<BaseFloat>& b, BaseFloat lr, const BaseFloat* one,
                CuMatrixBase<BaseFloat>* out) {
  KALDI_ASSERT(a.NumRows() == b.NumRows() && out->NumRows() == a.NumCols() && out->NumCols() == b.NumCols());
  CheckStatus(tdnnf_darts_backprop_params(CurrentContext(), b.Data(), b.NumRows(), b.NumCols(), b.Stride(), a.Data(),
                                          a.NumRows(), a.NumCols(), a.Stride(), NULL, 0, out->Data(), out->Stride(), NULL,
                                          one, 1, kZeroOffset, 1, lr, NULL));
}

// out (a.rows x b.cols) += sign * a b,  sign = *sign_dev
void ProductAB(const CuMatrixBase<BaseFloat>& a, const CuMatrixBase<BaseFloat>& b, const BaseFloat* sign_dev,
               CuMatrixBase<BaseFloat>* out) {
  KALDI_ASSERT(a.NumCols() == b.NumRows() && out->NumRows() == a.NumRows() && out->NumCols() == b.NumCols());
  CheckStatus(tdnnf_darts_backprop_data(CurrentContext(), a.Data(), a.NumRows(), a.NumCols(), a.Stride(), out->Data(),
                                        out->NumRows(), out->NumCols(), out->Stride(), b.Data(), b.Stride(), sign_dev, 1,
                                        kZeroOffset, 1));
}

// Symmetric eigenproblem of an n x n matrix in double: Householder reduction to tridiagonal form with the
// transformation accumulated, then QL iterations with implicit shifts (the classical EISPACK tred2 / tql2 scheme).
// ~(4/3 + 3) n^3 flops: 0.4 ms at n = 80 where cyclic Jacobi took 17 ms.  vals[k] goes with the COLUMN vecs[:, k].
// Returns false if an eigenvalue did not converge in 60 iterations (never observed).
bool SymmetricEigen(std::vector<double> a, int n, std::vector<double>* vals, std::vector<double>* vecs) {
  std::vector<double> d(n, 0.0), e(n, 0.0);
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  // ---- Householder tridiagonalisation; on exit `a` holds the orthogonal matrix Q with Q^T A Q tridiagonal
  for (int i = n - 1; i >= 1; --i) {
    const int l = i - 1;
    double h = 0.0;
    if (l > 0) {
      double scale = 0.0;
      for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          A(i, k) /= scale;
          h += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          A(j, i) = A(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
          e[j] = g / h;
          f += e[j] * A(i, j);
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
        }
      }
    } else {
      e[i] = A(i, l);
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
  }
  // ---- QL with implicit shifts on (d, e), rotating the columns of Q
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  bool ok = true;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) + dd == dd) break;
      }
      if (m != l) {
        if (iter++ == 60) {
          ok = false;
          break;
        }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = r = std::hypot(f, g);
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = A(k, i + 1);
            A(k, i + 1) = s * A(k, i) + c * f;
            A(k, i) = c * A(k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  *vals = d;
  *vecs = a;
  return ok;
}

// Fallback only (QL not converging): cyclic Jacobi rotations in double, ~40x slower at n = 80.
// Symmetric eigenproblem of an n x n matrix (cyclic Jacobi rotations in double).  vals[k] with vecs[:, k].
void JacobiEigen(std::vector<double> a, int n, std::vector<double>* vals, std::vector<double>* vecs) {
  std::vector<double>& v = *vecs;
  v.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
  auto at = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < n; ++i) {
      dia += at(i, i) * at(i, i);
      for (int j = i + 1; j < n; ++j) off += 2.0 * at(i, j) * at(i, j);
    }
    if (off <= 1e-30 * dia || off == 0.0) break;
    for (int p = 0; p + 1 < n; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double g = at(p, q);
        if (g == 0.0) continue;
        const double tau = (at(q, q) - at(p, p)) / (2.0 * g);
        const double t = std::copysign(1.0, tau) / (std::fabs(tau) + std::hypot(1.0, tau));
        const double c = 1.0 / std::hypot(1.0, t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double x = at(k, p), y = at(k, q);
          at(k, p) = c * x - s * y;
          at(k, q) = s * x + c * y;
        }
        for (int k = 0; k < n; ++k) {
          const double x = at(p, k), y = at(q, k);
          at(p, k) = c * x - s * y;
          at(q, k) = s * x + c * y;
        }
        for (int k = 0; k < n; ++k) {
          const double x = v[(size_t)k * n + p], y = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * x - s * y;
          v[(size_t)k * n + q] = s * x + c * y;
        }
      }
    }
  }
  vals->resize(n);
  for (int i = 0; i < n; ++i) (*vals)[i] = at(i, i);
}

// e_i = 1 / (beta / d_i + 1)   (ComputeEt)
void FisherE(const std::vector<BaseFloat>& d, BaseFloat beta, std::vector<BaseFloat>* sqrt_e,
             std::vector<BaseFloat>* inv_sqrt_e) {
  sqrt_e->resize(d.size());
  inv_sqrt_e->resize(d.size());
  for (size_t i = 0; i < d.size(); ++i) {
    const BaseFloat e = 1.0f / (beta / d[i] + 1.0f);
    (*sqrt_e)[i] = std::sqrt(e);
    (*inv_sqrt_e)[i] = 1.0f / (*sqrt_e)[i];
  }
}

BaseFloat Sum(const std::vector<BaseFloat>& v) {
  BaseFloat s = 0.f;
  for (BaseFloat x : v) s += x;
  return s;
}

}  // namespace

// Host-only entry for tests: the eigen-solver of the natural-gradient update (handle_api.cc).
bool SymmetricEigenForTest(const double* a, int n, double* vals, double* vecs) {
  std::vector<double> m(a, a + (size_t)n * n), va, ve;
  const bool ok = SymmetricEigen(m, n, &va, &ve);
  std::copy(va.begin(), va.end(), vals);
  std::copy(ve.begin(), ve.end(), vecs);
  return ok;
}

static bool g_ng_identity = false;
void SetNaturalGradientIdentity(bool b) { g_ng_identity = b; }
bool NaturalGradientIdentity() { return g_ng_identity; }

NgOperand NgOperand::Plain(const CuMatrixBase<BaseFloat>& m) {
  NgOperand x;
  x.data = m.Data();
  x.rows = m.NumRows();
  x.cols = m.NumCols();
  x.stride = m.Stride();
  x.num_rows = m.NumRows();
  x.n = 1;
  x.row_offsets = kZeroOffset;
  x.row_stride = 1;
  x.weff = NULL;
  x.ones_col = false;
  return x;
}

struct OnlineNaturalGradient::Pending {
  bool active = false;
  int32 N = 0;
  float* host = nullptr;  // pinned: L (R*R), K (R*R), {tr(XX^T), tr(X^X^^T), scale, -}, A (R*R), A diag(c) (R*R)
  size_t host_floats = 0;
  cudaEvent_t ready = nullptr;     // L, K and the traces have arrived
  cudaEvent_t uploaded = nullptr;  // the device has consumed A, A diag(c)
  bool upload_in_flight = false;
  // The host half of the update runs on a worker thread (it waits for `ready`, then solves the R x R
  // eigenproblem) while the caller keeps launching the rest of the backward pass; FinishPendingUpdate joins it.
  std::future<void> host_half;
  std::vector<BaseFloat> d_t1;
  BaseFloat rho_t1 = 0.f;
  bool must_reorthogonalize = false;
  void Join() {
    if (host_half.valid()) host_half.get();
  }
  ~Pending() {
    if (host_half.valid()) {
      try { host_half.get(); } catch (...) {}
    }
    if (host) cudaFreeHost(host);
    if (ready) cudaEventDestroy(ready);
    if (uploaded) cudaEventDestroy(uploaded);
  }
};

OnlineNaturalGradient::OnlineNaturalGradient()
    : rank_(40), update_period_(1), num_samples_history_(2000.0), alpha_(4.0), epsilon_(1.0e-10), delta_(5.0e-04),
      frozen_(false), t_(0), rho_t_(0.0), num_reorthogonalized_(0), pending_(new Pending()) {}

OnlineNaturalGradient::~OnlineNaturalGradient() {
  delete pending_;
  if (sumsq_) cudaFree(sumsq_);
}

OnlineNaturalGradient::OnlineNaturalGradient(const OnlineNaturalGradient& other) : pending_(new Pending()) { *this = other; }

OnlineNaturalGradient& OnlineNaturalGradient::operator=(const OnlineNaturalGradient& other) {
  if (this == &other) return *this;
  const_cast<OnlineNaturalGradient&>(other).FinishPendingUpdate();
  pending_->Join();  // a worker of ours may still read the state overwritten below
  pending_->active = false;
  rank_ = other.rank_;
  update_period_ = other.update_period_;
  num_samples_history_ = other.num_samples_history_;
  alpha_ = other.alpha_;
  epsilon_ = other.epsilon_;
  delta_ = other.delta_;
  frozen_ = other.frozen_;
  t_ = other.t_;
  rho_t_ = other.rho_t_;
  d_t_ = other.d_t_;
  num_reorthogonalized_ = other.num_reorthogonalized_;
  W_t_ = other.W_t_;
  WWt_ = other.WWt_;
  consts_ = other.consts_;
  w_last_ = other.w_last_;
  // scratch (H_, J_, L_, K_, ...) is not state: left empty, sized on first use
  return *this;
}

void OnlineNaturalGradient::Swap(OnlineNaturalGradient* other) {
  OnlineNaturalGradient tmp(*other);
  *other = *this;
  *this = tmp;
}

BaseFloat OnlineNaturalGradient::Eta(int32 N) const {
  KALDI_ASSERT(num_samples_history_ > 0.0);
  BaseFloat ans = 1.0f - std::exp(-(BaseFloat)N / num_samples_history_);
  return ans > 0.9f ? 0.9f : ans;  // "Don't let eta approach 1"
}

bool OnlineNaturalGradient::Updating() const {
  const int32 num_initial_updates = 10;  // must exceed the 3 initialisation passes
  return !frozen_ && (t_ <= num_initial_updates || (t_ - num_initial_updates) % update_period_ == 0);
}

void OnlineNaturalGradient::EnsureConsts() {
  if (consts_.Dim() == 0) {
    consts_.Resize(2);
    consts_.CopyFromHost(std::vector<BaseFloat>{1.0f, -1.0f});
  }
}

// InitDefault: W_0 = E^{1/2} R_0 with R_0 the "special" orthonormal matrix (1.1 then 1's, row r at columns
// r, r+R, r+2R, ...), d = rho = epsilon.
void OnlineNaturalGradient::InitDefault(int32 D) {
  if (rank_ >= D) {
    KaldiWarn("Rank of online preconditioner is >= dim, reducing it to dim - 1");
    rank_ = D - 1;
  }
  if (rank_ == 0) return;
  KALDI_ASSERT(num_samples_history_ > 0.0 && num_samples_history_ <= 1.0e+06 && alpha_ >= 0.0);
  const int32 R = rank_;
  rho_t_ = epsilon_;
  d_t_.assign(R, epsilon_);
  Matrix<BaseFloat> w(R, D);
  const BaseFloat first_elem = 1.1f;
  const BaseFloat e_tii = 1.0f / (2.0f + (D + R) * alpha_ / D);
  for (int32 r = 0; r < R; r++) {
    const int32 count = (D - r + R - 1) / R;
    const BaseFloat normalizer = 1.0f / std::sqrt(first_elem * first_elem + count - 1);
    for (int32 c = r, i = 0; c < D; c += R, i++) w(r, c) = std::sqrt(e_tii) * normalizer * (i == 0 ? first_elem : 1.0f);
  }
  W_t_.CopyFromHost(w);
  t_ = 0;
  RefreshDerived();
}

// W W^T (needed by the device-side scale) and the contiguous copy of W's last column (the weights of the
// appended column of ones, handed to the GEMM as its bias).
void OnlineNaturalGradient::RefreshDerived() {
  EnsureConsts();
  FullPrecisionGemms full(true);
  const int32 R = W_t_.NumRows(), D = W_t_.NumCols();
  EnsureSize(&WWt_, R, R);
  ProductABt(W_t_, W_t_, D, consts_.Data(), &WWt_);
  if (w_last_.Dim() != R) w_last_.Resize(R);
  w_last_.SetZero();
  CheckStatus(tdnnf_mat_axpy(CurrentContext(), 1.0f, W_t_.Data() + (D - 1), W_t_.Stride(), w_last_.Data(), 1, R, 1));
}

void OnlineNaturalGradient::Init(const NgOperand& X) {
  const int32 D = X.Dim();
  InitDefault(D);
  if (rank_ == 0) return;
  // three passes over the first minibatch from the default start ("faster than an eigendecomposition"),
  // only when it has more rows than the rank
  const int32 num_init_iters = (X.num_rows <= rank_) ? 0 : 3;
  const bool frozen = frozen_;
  frozen_ = false;
  t_ = 1;
  for (int32 i = 0; i < num_init_iters; i++) {
    Step(X, true);
    FinishPendingUpdate();
    t_ += 1;
  }
  t_ = 0;
  frozen_ = frozen;
}

void OnlineNaturalGradient::Step(const NgOperand& X, bool updating) {
  EnsureConsts();
  tdnnf_ctx* ctx = CurrentContext();
  const int32 R = rank_, D = X.Dim(), N = X.num_rows, spliced = X.n * X.cols;
  KALDI_ASSERT(R > 0 && R < D && W_t_.NumRows() == R && W_t_.NumCols() == D);
  const BaseFloat* one = consts_.Data();
  const BaseFloat* weff = X.weff ? X.weff : one;
  KALDI_ASSERT(X.weff != NULL || X.n == 1);
  FullPrecisionGemms full(updating);
  // H_t = X_t W_t^T
  EnsureSize(&H_, N, R);
  if (X.n > 1) {  // spliced operand, skinny W: one un-spliced GEMM + gather-sum over the offsets
    CheckStatus(tdnnf_darts_project(ctx, X.data, X.rows, X.cols, X.stride, H_.Data(), N, R, H_.Stride(), W_t_.Data(),
                                    W_t_.Stride(), X.ones_col ? w_last_.Data() : NULL, weff, X.n, X.row_offsets, X.row_stride));
  } else {
    CheckStatus(tdnnf_darts_propagate(ctx, X.data, X.rows, X.cols, X.stride, H_.Data(), N, R, H_.Stride(), W_t_.Data(),
                                      W_t_.Stride(), X.ones_col ? w_last_.Data() : NULL, X.ones_col ? 2 : 1, weff, X.n,
                                      X.row_offsets, X.row_stride));
  }
  // L_t = H_t^T H_t and tr(X X^T).  Inside Backprop's operand-cache scope the per-row sums of squares of X came with
  // the operand split of the H product above, and one launch does both (fp32 FMAs, one pass over H); otherwise the
  // trace takes one pass over X, and ranks beyond 128 go through the tensor-core GEMM.
  EnsureSize(&L_, R, R);
  if (sumsq_ == nullptr) CudaCheck(cudaMalloc(reinterpret_cast<void**>(&sumsq_), sizeof(double) * TDNNF_MAX_OFFSETS), "cudaMalloc");
  if (scal_.Dim() != 4) scal_.Resize(4);
  const float* rowsq = NULL;
  CheckStatus(tdnnf_ctx_operand_rowsq(ctx, X.data, X.rows, &rowsq));
  if (rowsq == NULL || R > 128)
    CheckStatus(tdnnf_darts_view_sumsq(ctx, X.data, X.rows, X.cols, X.stride, N, X.n, X.row_offsets, X.row_stride, sumsq_));
  if (R > 128) {
    L_.SetZero();
    ProductAtB(H_, H_, 1.0f, one, &L_);
    CheckStatus(tdnnf_ng_scale(ctx, sumsq_, X.weff, X.n, X.ones_col ? (float)N : 0.f, L_.Data(), L_.Stride(), WWt_.Data(),
                               WWt_.Stride(), R, scal_.Data()));
  } else {
    CheckStatus(tdnnf_ng_gram_scale(ctx, H_.Data(), N, R, H_.Stride(), L_.Data(), L_.Stride(), WWt_.Data(), WWt_.Stride(),
                                    rowsq, sumsq_, X.rows, X.n, X.row_offsets, X.row_stride, X.weff,
                                    X.ones_col ? (float)N : 0.f, scal_.Data()));
  }
  if (!updating) return;
  // J_t = H_t^T X_t   (block i scaled by w_i, last column = column sums of H)
  EnsureSize(&J_, R, D);
  J_.SetZero();
  if (X.ones_col) {
    if (tmp_r_.Dim() != R) tmp_r_.Resize(R);
    tmp_r_.SetZero();
  }
  CheckStatus(tdnnf_darts_backprop_params(ctx, X.data, X.rows, X.cols, X.stride, H_.Data(), N, R, H_.Stride(), NULL, 0,
                                          J_.Data(), J_.Stride(), X.ones_col ? tmp_r_.Data() : NULL, weff, X.n,
                                          X.row_offsets, X.row_stride, 1.0f, NULL));
  if (X.ones_col)
    CheckStatus(tdnnf_mat_axpy(ctx, 1.0f, tmp_r_.Data(), 1, J_.Data() + spliced, J_.Stride(), R, 1));
  // K_t = J_t J_t^T
  EnsureSize(&K_, R, R);
  ProductABt(J_, J_, D, one, &K_);
  // L, K and the traces to pinned host memory; the eigen-update is finished lazily
  const size_t need = (size_t)4 * R * R + 4;
  if (pending_->host_floats < need) {
    if (pending_->upload_in_flight) CudaCheck(cudaEventSynchronize(pending_->uploaded), "cudaEventSynchronize");
    pending_->upload_in_flight = false;
    if (pending_->host) cudaFreeHost(pending_->host);
    CudaCheck(cudaMallocHost(reinterpret_cast<void**>(&pending_->host), sizeof(float) * need), "cudaMallocHost");
    pending_->host_floats = need;
  }
  if (!pending_->ready) CudaCheck(cudaEventCreateWithFlags(&pending_->ready, cudaEventDisableTiming), "cudaEventCreate");
  cudaStream_t st = Stream();
  {
    const float* src[3] = {L_.Data(), K_.Data(), scal_.Data()};
    float* dst[3] = {pending_->host, pending_->host + (size_t)R * R, pending_->host + (size_t)2 * R * R};
    const int32_t ss[3] = {L_.Stride(), K_.Stride(), 4}, ds[3] = {R, R, 4}, rr[3] = {R, R, 1}, cc[3] = {R, R, 4};
    CheckStatus(tdnnf_copy_blocks(ctx, 3, src, ss, dst, ds, rr, cc));
  }
  CudaCheck(cudaEventRecord(pending_->ready, st), "cudaEventRecord");
  pending_->active = true;
  pending_->N = N;
  int device = 0;
  CudaCheck(cudaGetDevice(&device), "cudaGetDevice");
  const int32 D_total = D;
  // worker threads (TDNNF_NG_ASYNC=<n>, default 4; 0 = on the calling thread, synchronously)
  pending_->host_half = HostWorkers::Get().Run([this, device, D_total] {
    CudaCheck(cudaSetDevice(device), "cudaSetDevice");
    HostHalfOfUpdate(D_total);
  });
}

// The host half of PreconditionDirectionsInternal (ComputeZt, the eigenproblem, the floors, rho_{t+1},
// D_{t+1}, ComputeWt1's coefficient matrices): worker thread, reads only the pinned buffer and the (t)-state,
// writes A_t, A_t diag(c) into the pinned buffer and (d_{t+1}, rho_{t+1}) into Pending.
void OnlineNaturalGradient::HostHalfOfUpdate(int32 D) {
  CudaCheck(cudaEventSynchronize(pending_->ready), "cudaEventSynchronize");
  const int32 R = rank_, N = pending_->N;
  const float* Lh = pending_->host;
  const float* Kh = pending_->host + (size_t)R * R;
  const BaseFloat tr_X_Xt = pending_->host[(size_t)2 * R * R];
  const BaseFloat eta = Eta(N), rho_t = rho_t_;
  const std::vector<BaseFloat>& d_t = d_t_;
  const BaseFloat d_sum = Sum(d_t);
  const BaseFloat beta_t = rho_t * (1.0f + alpha_) + alpha_ * d_sum / D;
  std::vector<BaseFloat> sqrt_e_t, inv_sqrt_e_t;
  FisherE(d_t, beta_t, &sqrt_e_t, &inv_sqrt_e_t);
  // Z_t = (eta/N)^2 E^-.5 K E^-.5 + (eta/N)(1-eta) [E^-.5 L E^-.5 (D+rho I) + (D+rho I) E^-.5 L E^-.5] + (1-eta)^2 (D+rho I)^2
  std::vector<double> Z((size_t)R * R);
  const double etaN = (double)eta / N, eta1 = 1.0 - (double)eta;
  for (int32 i = 0; i < R; i++) {
    const double ei = inv_sqrt_e_t[i], di = (double)(d_t[i] + rho_t);
    for (int32 j = 0; j < R; j++) {
      const double ej = inv_sqrt_e_t[j], dj = (double)(d_t[j] + rho_t);
      // symmetrise the fp32 products (upstream reads the lower triangle)
      const double Lij = i >= j ? Lh[(size_t)i * R + j] : Lh[(size_t)j * R + i];
      const double Kij = i >= j ? Kh[(size_t)i * R + j] : Kh[(size_t)j * R + i];
      Z[(size_t)i * R + j] = etaN * etaN * ei * Kij * ej + etaN * eta1 * ei * Lij * ej * (dj + di) +
                             (i == j ? eta1 * eta1 * di * di : 0.0);
    }
  }
  double trace = 0.0;
  for (int32 i = 0; i < R; i++) trace += Z[(size_t)i * R + i];
  const BaseFloat z_t_scale = (BaseFloat)std::max(1.0, trace);  // avoids overflow: Z ~ data^4
  for (double& z : Z) z = (double)(float)(z / z_t_scale);
  std::vector<double> vals, vecs;
  if (!SymmetricEigen(Z, R, &vals, &vecs)) JacobiEigen(Z, R, &vals, &vecs);
  std::vector<int32> order(R);
  for (int32 i = 0; i < R; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int32 a, int32 b) { return std::fabs(vals[a]) > std::fabs(vals[b]); });
  std::vector<BaseFloat> c_t(R);
  for (int32 i = 0; i < R; i++) c_t[i] = (BaseFloat)vals[order[i]] * z_t_scale;
  const BaseFloat condition_threshold = 1.0e+06f;
  bool must_reorthogonalize = c_t[0] > condition_threshold * c_t[R - 1];
  const BaseFloat c_t_floor = (rho_t * (1.0f - eta)) * (rho_t * (1.0f - eta));
  for (BaseFloat& c : c_t)
    if (c < c_t_floor) {
      c = c_t_floor;
      must_reorthogonalize = true;
    }
  std::vector<BaseFloat> sqrt_c_t(R);
  BaseFloat sqrt_c_sum = 0.f, sqrt_c_max = 0.f;
  for (int32 i = 0; i < R; i++) {
    sqrt_c_t[i] = std::sqrt(c_t[i]);
    sqrt_c_sum += sqrt_c_t[i];
    sqrt_c_max = std::max(sqrt_c_max, sqrt_c_t[i]);
  }
  // rho_{t+1} = 1/(D-R) (eta/N tr(X X^T) + (1-eta)(D rho_t + tr(D_t)) - tr(C_t^{1/2}))
  BaseFloat rho_t1 = 1.0f / (D - R) * (eta / N * tr_X_Xt + (1 - eta) * (D * rho_t + d_sum) - sqrt_c_sum);
  std::vector<BaseFloat> d_t1(R);
  for (int32 i = 0; i < R; i++) d_t1[i] = sqrt_c_t[i] - rho_t1;
  const BaseFloat floor_val = std::max(epsilon_, delta_ * sqrt_c_max);
  if (rho_t1 < floor_val) rho_t1 = floor_val;
  for (BaseFloat& v : d_t1)
    if (v < floor_val) v = floor_val;
  // A_t = (eta/N) E_{t+1}^{1/2} C_t^{-1/2} U_t^T E_t^{-1/2};  B_t = J_t + (1-eta)/(eta/N) (D_t + rho_t I) W_t
  const BaseFloat beta_t1 = rho_t1 * (1.0f + alpha_) + alpha_ * Sum(d_t1) / D;
  KALDI_ASSERT(beta_t1 > 0.0);
  std::vector<BaseFloat> sqrt_e_t1, inv_sqrt_e_t1;
  FisherE(d_t1, beta_t1, &sqrt_e_t1, &inv_sqrt_e_t1);
  // staged in pinned memory (the previous upload from this staging area must have been consumed)
  if (pending_->upload_in_flight) CudaCheck(cudaEventSynchronize(pending_->uploaded), "cudaEventSynchronize");
  float* A = pending_->host + (size_t)2 * R * R + 4;
  float* AC = A + (size_t)R * R;
  for (int32 i = 0; i < R; i++) {
    const BaseFloat i_factor = (eta / N) * sqrt_e_t1[i] / sqrt_c_t[i];
    for (int32 j = 0; j < R; j++) {
      const BaseFloat u_ji = (BaseFloat)vecs[(size_t)j * R + order[i]];
      A[(size_t)i * R + j] = u_ji * (i_factor * inv_sqrt_e_t[j]);
      AC[(size_t)i * R + j] = A[(size_t)i * R + j] * ((1.0f - eta) / (eta / N) * (d_t[j] + rho_t));
    }
  }
  pending_->d_t1 = d_t1;
  pending_->rho_t1 = rho_t1;
  pending_->must_reorthogonalize = must_reorthogonalize;
}

// The device half: W_{t+1} = A_t (J_t + diag(c) W_t), uploaded on the stream (no device-wide sync).
void OnlineNaturalGradient::FinishPendingUpdate() {
  if (!pending_->active) return;
  pending_->active = false;
  pending_->Join();
  FullPrecisionGemms full(true);
  const int32 R = rank_, D = W_t_.NumCols();
  const float* A = pending_->host + (size_t)2 * R * R + 4;
  const float* AC = A + (size_t)R * R;
  EnsureConsts();
  EnsureSize(&A_, R, R);
  EnsureSize(&AC_, R, R);
  {
    cudaStream_t st = Stream();
    const float* src[2] = {A, AC};
    float* dst[2] = {A_.Data(), AC_.Data()};
    const int32_t ss[2] = {R, R}, ds[2] = {A_.Stride(), AC_.Stride()}, rr[2] = {R, R}, cc[2] = {R, R};
    CheckStatus(tdnnf_copy_blocks(CurrentContext(), 2, src, ss, dst, ds, rr, cc));
    if (!pending_->uploaded) CudaCheck(cudaEventCreateWithFlags(&pending_->uploaded, cudaEventDisableTiming), "cudaEventCreate");
    CudaCheck(cudaEventRecord(pending_->uploaded, st), "cudaEventRecord");
    pending_->upload_in_flight = true;
  }
  EnsureSize(&W_next_, R, D);
  if (R <= 128) {
    CheckStatus(tdnnf_ng_w_update(CurrentContext(), A_.Data(), A_.Stride(), AC_.Data(), AC_.Stride(), J_.Data(), J_.Stride(),
                                  W_t_.Data(), W_t_.Stride(), R, D, W_next_.Data(), W_next_.Stride()));
  } else {
    W_next_.SetZero();
    ProductAB(A_, J_, consts_.Data(), &W_next_);
    ProductAB(AC_, W_t_, consts_.Data(), &W_next_);
  }
  W_t_.Swap(&W_next_);
  d_t_ = pending_->d_t1;
  rho_t_ = pending_->rho_t1;
  RefreshDerived();
  if (pending_->must_reorthogonalize) Reorthogonalize();
}

// ReorthogonalizeRt1: O = E^{-1/2} W W^T E^{-1/2} should be the unit matrix; if not, W <- E^{1/2} C^{-1} E^{-1/2} W
// with O = C C^T (Cholesky), or Gram-Schmidt on the host when the Cholesky factor is out of range.
void OnlineNaturalGradient::Reorthogonalize() {
  const BaseFloat threshold = 1.0e-03f;
  const int32 R = rank_, D = W_t_.NumCols();
  const BaseFloat beta = rho_t_ * (1.0f + alpha_) + alpha_ * Sum(d_t_) / D;
  std::vector<BaseFloat> sqrt_e, inv_sqrt_e;
  FisherE(d_t_, beta, &sqrt_e, &inv_sqrt_e);
  Matrix<BaseFloat> Oh = WWt_.ToHost();  // synchronises (rare path)
  std::vector<double> O((size_t)R * R);
  bool is_unit = true;
  for (int32 i = 0; i < R; i++)
    for (int32 j = 0; j < R; j++) {
      const double o = (double)(i >= j ? Oh(i, j) : Oh(j, i)) * inv_sqrt_e[i] * inv_sqrt_e[j];
      O[(size_t)i * R + j] = o;
      if (std::fabs(o - (i == j ? 1.0 : 0.0)) > threshold) is_unit = false;
    }
  if (is_unit) return;
  num_reorthogonalized_++;
  std::vector<double> C((size_t)R * R, 0.0), Ci((size_t)R * R, 0.0);
  bool ok = true;
  for (int32 i = 0; i < R && ok; i++) {
    for (int32 j = 0; j <= i; j++) {
      double s = O[(size_t)i * R + j];
      for (int32 k = 0; k < j; k++) s -= C[(size_t)i * R + k] * C[(size_t)j * R + k];
      if (i == j) {
        if (!(s > 0.0)) { ok = false; break; }
        C[(size_t)i * R + i] = std::sqrt(s);
      } else {
        C[(size_t)i * R + j] = s / C[(size_t)j * R + j];
      }
    }
  }
  if (ok) {
    double mx = -1e300;
    for (int32 i = 0; i < R; i++) {
      Ci[(size_t)i * R + i] = 1.0 / C[(size_t)i * R + i];
      for (int32 j = 0; j < i; j++) {
        double s = 0.0;
        for (int32 k = j; k < i; k++) s += C[(size_t)i * R + k] * Ci[(size_t)k * R + j];
        Ci[(size_t)i * R + j] = -s / C[(size_t)i * R + i];
      }
      for (int32 j = 0; j <= i; j++) mx = std::max(mx, Ci[(size_t)i * R + j]);
    }
    if (!(mx < 100.0)) ok = false;
  }
  if (!ok) {
    KaldiWarn("Cholesky out of expected range, reorthogonalizing with Gram-Schmidt");
    Matrix<BaseFloat> w = W_t_.ToHost();
    for (int32 i = 0; i < R; i++) {
      for (int32 j = 0; j < i; j++) {
        double p = 0.0;
        for (int32 k = 0; k < D; k++) p += (double)w(i, k) * w(j, k);
        for (int32 k = 0; k < D; k++) w(i, k) -= (BaseFloat)p * w(j, k);
      }
      double nn = 0.0;
      for (int32 k = 0; k < D; k++) nn += (double)w(i, k) * w(i, k);
      const BaseFloat inv = (BaseFloat)(1.0 / std::sqrt(nn));
      for (int32 k = 0; k < D; k++) w(i, k) *= inv;
    }
    for (int32 i = 0; i < R; i++)
      for (int32 k = 0; k < D; k++) w(i, k) *= sqrt_e[i];
    W_t_.CopyFromHost(w);
    RefreshDerived();
    return;
  }
  Matrix<BaseFloat> M(R, R);
  for (int32 i = 0; i < R; i++)
    for (int32 j = 0; j <= i; j++) M(i, j) = (BaseFloat)(Ci[(size_t)i * R + j] * (i == j ? 1.0 : (double)sqrt_e[i] * inv_sqrt_e[j]));
  A_.CopyFromHost(M);
  EnsureSize(&W_next_, R, D);
  W_next_.SetZero();
  ProductAB(A_, W_t_, consts_.Data(), &W_next_);
  W_t_.Swap(&W_next_);
  RefreshDerived();
}

void OnlineNaturalGradient::PreconditionImplicit(const NgOperand& X, NgProjection* out) {
  out->identity = true;
  out->rank = 0;
  out->W = NULL;
  out->H = NULL;
  out->scale_dev = NULL;
  if (X.Dim() == 1 || g_ng_identity) return;  // "our natural gradient update with rescaling becomes a no-op"
  if (t_ == 0) Init(X);
  if (rank_ == 0) return;
  FinishPendingUpdate();
  KALDI_ASSERT(W_t_.NumCols() == X.Dim());
  Step(X, Updating());
  t_ += 1;
  out->identity = false;
  out->rank = rank_;
  out->W = &W_t_;
  out->H = &H_;
  out->scale_dev = scal_.Data() + 2;
}

// The upstream signature: X_t is overwritten by X_hat_t and *scale returned on the host (one sync).
void OnlineNaturalGradient::PreconditionDirections(CuMatrixBase<BaseFloat>* X_t, BaseFloat* scale) {
  NgProjection p;
  PreconditionImplicit(NgOperand::Plain(*X_t), &p);
  if (p.identity) {
    if (scale) *scale = 1.0;
    return;
  }
  ProductAB(*p.H, *p.W, consts_.Data() + 1, X_t);  // X_hat = X - H W
  if (scale) {
    float s = 1.0f;
    cudaStream_t st = Stream();
    CudaCheck(cudaMemcpyAsync(&s, p.scale_dev, sizeof(float), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    CudaCheck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    *scale = s;
  }
}

void OnlineNaturalGradient::GetState(int32* t, BaseFloat* rho, std::vector<BaseFloat>* d, Matrix<BaseFloat>* W) {
  FinishPendingUpdate();
  if (t) *t = t_;
  if (rho) *rho = rho_t_;
  if (d) *d = d_t_;
  if (W) *W = W_t_.ToHost();
}

void OnlineNaturalGradient::FreeScratch() {
  H_.Resize(0, 0);
  J_.Resize(0, 0);
  W_next_.Resize(0, 0);
}

}  // namespace nnet3
}  // namespace tdnnf
