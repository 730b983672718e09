// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product: nothing under tdnn-f_nas_b200/
// may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs load it, and only as the checker / the timed CPU baseline.
//
// A plain fp32 CPU restatement of the reference's algorithms for the training hot path,
// written method by method after the reference (skhu101/TDNN-F_NAS) sources:
//   tdnn.cc   = /root/reference/src/nnet3/nnet-tdnn-component.cc
//   simple.cc = /root/reference/src/nnet3/nnet-simple-component.cc
//   norm.cc   = /root/reference/src/nnet3/nnet-normalize-component.cc
// The reference cannot be compiled here (it is a patch set on upstream Kaldi, which is absent:
// no cudamatrix/, matrix/, chain/, OpenFst, BLAS), and it ships no tests or golden vectors, so
//     PARITY IS UNPINNED
// against the reference's own binaries; the restatement is instead cross-checked against an
// independent float64 numpy restatement, finite differences and the denominator invariants
// (tests/test_oracle_*.py).  The denominator follows upstream kaldi chain-denominator.cc
// (CPU code path) as summarised in SURVEY.md Appendix B.
//
// The two OnlineNaturalGradient::PreconditionDirections calls (tdnn.cc:598-599) follow upstream Kaldi's
// natural-gradient-online.cc as restated in oracle_ng.inc (orc_tdnn_backprop_ng); orc_tdnn_backprop is
// the same with both preconditioners = identity (the "raw gradient" path of BASELINE.md section 3).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#include <dlfcn.h>

namespace {

// Optional BLAS behind AddMatMat (orc_use_blas): a Kaldi CPU build calls cblas_sgemm for every AddMatMat, so the timed
// CPU baseline should too.  No BLAS is installed system-wide in the image; numpy ships OpenBLAS (ILP64, symbols
// scipy_cblas_sgemm64_), bound at run time.  Without it the plain OpenMP loops below are used (the parity tests).
typedef void (*sgemm64_fn)(int order, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                           const float* B, int64_t ldb, float beta, float* C, int64_t ldc);
typedef void (*sgemm32_fn)(int order, int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                           int ldb, float beta, float* C, int ldc);
sgemm64_fn g_sgemm64 = nullptr;
sgemm32_fn g_sgemm32 = nullptr;
void (*g_blas_set_threads)(int) = nullptr;

typedef float BaseFloat;

// A CuMatrixBase-like view on host memory.
struct Mat {
  BaseFloat* data;
  int rows, cols, stride;
  BaseFloat& operator()(int r, int c) { return data[(size_t)r * stride + c]; }
  const BaseFloat& operator()(int r, int c) const { return data[(size_t)r * stride + c]; }
  Mat Range(int r0, int nr, int c0, int nc) const { return Mat{data + (size_t)r0 * stride + c0, nr, nc, stride}; }
};

struct OwnedMat {
  std::vector<BaseFloat> buf;
  Mat m;
  OwnedMat(int rows, int cols) : buf((size_t)rows * cols, 0.f) { m = Mat{buf.data(), rows, cols, cols}; }
};

enum Trans { kNoTrans, kTrans };

// C = alpha * op(A) * op(B) + beta * C   (CuMatrixBase::AddMatMat)
void AddMatMat(Mat C, BaseFloat alpha, const Mat& A, Trans ta, const Mat& B, Trans tb, BaseFloat beta) {
  const int M = C.rows, N = C.cols;
  const int K = (ta == kNoTrans) ? A.cols : A.rows;
  if ((g_sgemm64 || g_sgemm32) && M > 0 && N > 0 && K > 0 && !(ta == kTrans && tb == kTrans)) {
    // row-major: CblasRowMajor = 101, CblasNoTrans = 111, CblasTrans = 112
    const int cta = ta == kNoTrans ? 111 : 112, ctb = tb == kNoTrans ? 111 : 112;
    if (g_sgemm64) g_sgemm64(101, cta, ctb, M, N, K, alpha, A.data, A.stride, B.data, B.stride, beta, C.data, C.stride);
    else g_sgemm32(101, cta, ctb, M, N, K, alpha, A.data, A.stride, B.data, B.stride, beta, C.data, C.stride);
    return;
  }
  if (ta == kNoTrans && tb == kTrans) {
    // C[m,n] = sum_k A[m,k] B[n,k]
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
      const BaseFloat* a = &A(m, 0);
      for (int n = 0; n < N; ++n) {
        const BaseFloat* b = &B(n, 0);
        BaseFloat acc = 0.f;
#pragma omp simd reduction(+ : acc)
        for (int k = 0; k < K; ++k) acc += a[k] * b[k];
        C(m, n) = alpha * acc + (beta == 0.f ? 0.f : beta * C(m, n));
      }
    }
  } else if (ta == kNoTrans && tb == kNoTrans) {
    // C[m,:] = sum_k A[m,k] B[k,:]
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
      std::vector<BaseFloat> row(N, 0.f);
      for (int k = 0; k < K; ++k) {
        const BaseFloat a = A(m, k);
        const BaseFloat* b = &B(k, 0);
#pragma omp simd
        for (int n = 0; n < N; ++n) row[n] += a * b[n];
      }
      for (int n = 0; n < N; ++n) C(m, n) = alpha * row[n] + (beta == 0.f ? 0.f : beta * C(m, n));
    }
  } else if (ta == kTrans && tb == kNoTrans) {
    // C[m,:] = sum_k A[k,m] B[k,:]   (K = rows of A): parallel over m-blocks, each thread owns its C rows
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
      std::vector<BaseFloat> row(N, 0.f);
      for (int k = 0; k < K; ++k) {
        const BaseFloat a = A(k, m);
        const BaseFloat* b = &B(k, 0);
#pragma omp simd
        for (int n = 0; n < N; ++n) row[n] += a * b[n];
      }
      for (int n = 0; n < N; ++n) C(m, n) = alpha * row[n] + (beta == 0.f ? 0.f : beta * C(m, n));
    }
  } else {
    fprintf(stderr, "oracle: AddMatMat(kTrans,kTrans) not needed\n");
    abort();
  }
}

// GetInputPart (tdnn.cc:806-820): rows row_offset, row_offset+row_stride, ...
Mat GetInputPart(const Mat& input, int num_output_rows, int row_stride, int row_offset) {
  return Mat{input.data + (size_t)input.stride * row_offset, num_output_rows, input.cols, input.stride * row_stride};
}

}  // namespace

#include "oracle_ng.inc"

extern "C" {

#define ORC_USE_GUMBEL 1
#define ORC_FREE_SELECT 2
#define ORC_UNIFORM_SAMPLE 4
#define ORC_USE_ENTROPY 8
#define ORC_UPDATE_ALPHA 16

// Binds cblas_sgemm from the shared library at `path` (NULL: back to the plain loops).  Returns 1 if bound.
int orc_use_blas(const char* path) {
  g_sgemm64 = nullptr;
  g_sgemm32 = nullptr;
  g_blas_set_threads = nullptr;
  if (path == nullptr) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return 0;
  g_sgemm64 = reinterpret_cast<sgemm64_fn>(dlsym(h, "scipy_cblas_sgemm64_"));
  if (!g_sgemm64) g_sgemm64 = reinterpret_cast<sgemm64_fn>(dlsym(h, "cblas_sgemm64_"));
  if (!g_sgemm64) g_sgemm32 = reinterpret_cast<sgemm32_fn>(dlsym(h, "cblas_sgemm"));
  for (const char* nm : {"scipy_openblas_set_num_threads64_", "openblas_set_num_threads64_", "openblas_set_num_threads"})
    if (!g_blas_set_threads) g_blas_set_threads = reinterpret_cast<void (*)(int)>(dlsym(h, nm));
#ifdef _OPENMP
  if (g_blas_set_threads) g_blas_set_threads(omp_get_max_threads());
#endif
  return (g_sgemm64 || g_sgemm32) ? 1 : 0;
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// share_offset_index selection (tdnn.cc:227-241).  Returns -1 where the reference leaves the
// variable uninitialised (time_offsets[1] == 0 or n == 1).
int orc_share_index(const int* time_offsets, int n) {
  if (n < 2) return -1;
  if (time_offsets[1] > 0) return 0;
  if (time_offsets[1] < 0) return n - 1;
  return -1;
}

// Mixing coefficients, tdnn.cc:250-289.
void orc_darts_coef(const float* log_alpha, int n, int flags, float temp_proportion, const float* u_gumbel,
                    float u_uniform, float* coef) {
  std::vector<BaseFloat> c(log_alpha, log_alpha + n);            // coef_.CopyFromVec(bias_params_.Range(0, n))
  if (flags & ORC_USE_GUMBEL) {
    std::vector<BaseFloat> rand_(u_gumbel, u_gumbel + n);        // rand_.SetRandUniform()
    for (int i = 0; i < n; ++i) rand_[i] = logf(rand_[i]);       // ApplyLog
    for (int i = 0; i < n; ++i) rand_[i] *= -1;                  // Scale(-1)
    for (int i = 0; i < n; ++i) rand_[i] = logf(rand_[i]);       // ApplyLog
    for (int i = 0; i < n; ++i) rand_[i] *= -1;                  // Scale(-1)
    for (int i = 0; i < n; ++i) c[i] += 1.0f * rand_[i];         // AddVec
    const BaseFloat sc = 1.0f / temp_proportion;
    for (int i = 0; i < n; ++i) c[i] *= sc;                      // Scale(1/T)
    BaseFloat mx = -FLT_MAX;                                     // ApplySoftMax
    for (int i = 0; i < n; ++i) mx = std::max(mx, c[i]);
    BaseFloat sum = 0.f;
    for (int i = 0; i < n; ++i) { c[i] = expf(c[i] - mx); sum += c[i]; }
    for (int i = 0; i < n; ++i) c[i] /= sum;
    for (int i = 0; i < n; ++i) c[i] = std::max(c[i], 1.0e-20f);  // ApplyFloor
  } else if (flags & ORC_FREE_SELECT) {
    for (int i = 0; i < n; ++i) c[i] *= -1;                      // Scale(-1)
    for (int i = 0; i < n; ++i) c[i] = expf(c[i]);               // ApplyExp
    for (int i = 0; i < n; ++i) c[i] += 1.0f;                    // Add(1.0)
    for (int i = 0; i < n; ++i) c[i] = 1.0f / c[i];              // InvertElements
  } else {
    BaseFloat mx = -FLT_MAX;
    for (int i = 0; i < n; ++i) mx = std::max(mx, c[i]);
    BaseFloat sum = 0.f;
    for (int i = 0; i < n; ++i) { c[i] = expf(c[i] - mx); sum += c[i]; }
    for (int i = 0; i < n; ++i) c[i] /= sum;
    for (int i = 0; i < n; ++i) c[i] = std::max(c[i], 1.0e-20f);
  }
  if (flags & ORC_UNIFORM_SAMPLE) {                              // tdnn.cc:280-289
    for (int i = 0; i < n; ++i) c[i] = 0.f;
    for (int i = 0; i < n; ++i)
      if (u_uniform >= (float)(i) / n && u_uniform < (float)(i + 1) / n) c[i] = 1.0f;
  }
  for (int i = 0; i < n; ++i) coef[i] = c[i];
}

// TdnnDARTSV3Component::Propagate, tdnn.cc:214-333.  bias_params: n + out_dim entries
// (log-alpha then the real bias) or NULL (kPropagateAdds: out is added to).
// Returns -1 if share_offset_index would be uninitialised in the reference.
int orc_tdnn_propagate(const int* time_offsets, int n, int flags, float temp_proportion, const float* W, int w_stride,
                       const float* bias_params, const float* in, int in_rows, int in_dim, int in_stride, float* out,
                       int out_rows, int out_dim, int out_stride, const int* row_offsets, int row_stride,
                       const float* u_gumbel, float u_uniform, float* coef_memo) {
  Mat in_m{const_cast<float*>(in), in_rows, in_dim, in_stride};
  Mat out_m{out, out_rows, out_dim, out_stride};
  Mat lin{const_cast<float*>(W), out_dim, n * in_dim, w_stride};
  int share = orc_share_index(time_offsets, n);
  if (bias_params != nullptr) {
    if (n >= 2 && time_offsets[1] > 0) {                         // out->CopyRowsFromVec(bias tail)
      for (int r = 0; r < out_rows; ++r)
        for (int c = 0; c < out_dim; ++c) out_m(r, c) = bias_params[n + c];
    } else if (n >= 2 && time_offsets[1] < 0) {                  // out->SetZero(): bias NOT added (quirk Q2)
      for (int r = 0; r < out_rows; ++r)
        for (int c = 0; c < out_dim; ++c) out_m(r, c) = 0.f;
    }
  }
  if (share < 0) return -1;
  std::vector<float> coef(n);
  // note: the reference reads bias_params_.Range(0, n) unconditionally (quirk Q3: use-bias=false cannot work)
  if (bias_params == nullptr) return -2;
  orc_darts_coef(bias_params, n, flags, temp_proportion, u_gumbel, u_uniform, coef.data());
  for (int i = 0; i < n; ++i) {
    Mat in_part = GetInputPart(in_m, out_rows, row_stride, row_offsets[i]);
    Mat lin_part = lin.Range(0, out_dim, i * in_dim, in_dim);
    if (flags & ORC_UNIFORM_SAMPLE) {
      if (i == share || coef[i] == 1)
        AddMatMat(out_m, 1.0f, in_part, kNoTrans, lin_part, kTrans, 1.0f);
    } else if (flags & ORC_FREE_SELECT) {
      AddMatMat(out_m, coef[i], in_part, kNoTrans, lin_part, kTrans, 1.0f);
    } else if (i != share) {
      AddMatMat(out_m, coef[i], in_part, kNoTrans, lin_part, kTrans, 1.0f);
    } else {
      AddMatMat(out_m, 1.0f, in_part, kNoTrans, lin_part, kTrans, 1.0f);
    }
  }
  for (int i = 0; i < n; ++i) coef_memo[i] = coef[i];
  return 0;
}

// TdnnDARTSV3Component::Backprop + UpdateNaturalGradient, tdnn.cc:335-431, 457-626.
//   in_deriv   may be NULL; it is ADDED to (kBackpropAdds)
//   dW / dbias are the delta component's linear_params_ / bias_params_ (n + out_dim); NULL => no update
//   s_out      (optional, n): the raw inner products out_temp.Sum() per offset (0 where not computed)
//   ng_in / ng_out  (optional, orc_ng_create handles): preconditioner_in_ / preconditioner_out_ of the delta
//              component; NULL = the identity with scale 1 (the un-preconditioned gradient).
int orc_tdnn_backprop_ng(const int* time_offsets, int n, int flags, float temp_proportion, const float* W, int w_stride,
                         const float* in_value, int in_rows, int in_dim, int in_stride, const float* out_deriv,
                         int out_rows, int out_dim, int od_stride, const float* coef_memo, const int* row_offsets,
                         int row_stride, float* in_deriv, int id_stride, float learning_rate, float* dW, int dw_stride,
                         float* dbias, float* s_out, void* ng_in, void* ng_out, float* scales_out) {
  Mat in_m{const_cast<float*>(in_value), in_rows, in_dim, in_stride};
  Mat od{const_cast<float*>(out_deriv), out_rows, out_dim, od_stride};
  Mat lin{const_cast<float*>(W), out_dim, n * in_dim, w_stride};
  const int share = orc_share_index(time_offsets, n);
  if (share < 0) return -1;
  const float* coef = coef_memo;
  if (in_deriv != nullptr) {                                     // tdnn.cc:366-416
    Mat id{in_deriv, in_rows, in_dim, id_stride};
    for (int i = 0; i < n; ++i) {
      Mat id_part = GetInputPart(id, out_rows, row_stride, row_offsets[i]);
      Mat lin_part = lin.Range(0, out_dim, i * in_dim, in_dim);
      if (flags & ORC_UNIFORM_SAMPLE) {
        if (i == share || coef[i] == 1) AddMatMat(id_part, 1.0f, od, kNoTrans, lin_part, kNoTrans, 1.0f);
      } else if (flags & ORC_FREE_SELECT) {
        AddMatMat(id_part, coef[i], od, kNoTrans, lin_part, kNoTrans, 1.0f);
      } else if (i != share) {
        AddMatMat(id_part, coef[i], od, kNoTrans, lin_part, kNoTrans, 1.0f);
      } else {
        AddMatMat(id_part, 1.0f, od, kNoTrans, lin_part, kNoTrans, 1.0f);
      }
    }
  }
  if (s_out) for (int i = 0; i < n; ++i) s_out[i] = 0.f;
  if (dW == nullptr) return 0;
  if (learning_rate == 0.0f) return 0;                           // tdnn.cc:423-424

  // ---- UpdateNaturalGradient, tdnn.cc:457-626
  const int spliced = n * in_dim, augmented = spliced + 1;       // bias always present (Q3)
  OwnedMat in_value_temp(out_rows, augmented);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < out_rows; ++r) in_value_temp.m(r, spliced) = 1.0f;
  const bool gumbel = flags & ORC_USE_GUMBEL, uniform = flags & ORC_UNIFORM_SAMPLE, freesel = flags & ORC_FREE_SELECT;
  for (int i = 0; i < n; ++i) {
    Mat tpart = in_value_temp.m.Range(0, out_rows, i * in_dim, in_dim);
    Mat in_part = GetInputPart(in_m, out_rows, row_stride, row_offsets[i]);
    Mat lin_part = lin.Range(0, out_dim, i * in_dim, in_dim);
    if (uniform) {
      if (i == share || coef[i] == 1) {
        for (int r = 0; r < out_rows; ++r) memcpy(&tpart(r, 0), &in_part(r, 0), sizeof(float) * in_dim);
        // (out_temp is computed and discarded by the reference: no alpha gradient in this mode)
      } else {
        for (int r = 0; r < out_rows; ++r)
          for (int c = 0; c < in_dim; ++c) tpart(r, c) *= 0.0f;
      }
      continue;
    }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < out_rows; ++r) memcpy(&tpart(r, 0), &in_part(r, 0), sizeof(float) * in_dim);
    if (freesel || i != share) {
#pragma omp parallel for schedule(static)
      for (int r = 0; r < out_rows; ++r)
        for (int c = 0; c < in_dim; ++c) tpart(r, c) *= coef[i];
    }
    OwnedMat out_temp(out_rows, out_dim);
    AddMatMat(out_temp.m, 1.0f, in_part, kNoTrans, lin_part, kTrans, 0.0f);
    // out_temp.AddMatMatElements(1.0, out_temp, out_deriv, 0.0); out_temp.Sum()
    double sum = 0.0;  // CuMatrix::Sum() reduces in BaseFloat on the GPU; the order is unspecified
#pragma omp parallel for schedule(static) reduction(+ : sum)
    for (int r = 0; r < out_rows; ++r)
      for (int c = 0; c < out_dim; ++c) sum += (double)(out_temp.m(r, c) * od(r, c));
    const BaseFloat s = (BaseFloat)sum;
    if (s_out) s_out[i] = s;
    if (freesel) {
      dbias[i] += s * coef[i];                                   // AddVec(sum, coef_i)
      dbias[i] += (-1.0f * s) * coef[i] * coef[i];               // AddVecVec(-sum, coef_i, coef_i)
    } else if (i != share) {
      for (int j = 0; j < n; ++j) {
        if (gumbel) dbias[j] += (BaseFloat)(-1.0 * s / temp_proportion) * coef[i] * coef[j];
        else dbias[j] += (BaseFloat)(-1.0 * s) * coef[i] * coef[j];
      }
      if (gumbel) dbias[i] += (s / temp_proportion) * coef[i];
      else dbias[i] += s * coef[i];
    }
  }
  if (flags & ORC_USE_ENTROPY)                                   // tdnn.cc:565-569
    for (int i = 0; i < n; ++i) dbias[i] *= 5;
  if (freesel) { for (int i = 0; i < n; ++i) dbias[i] *= 5 * learning_rate; }       // tdnn.cc:574-586
  else if (gumbel) { for (int i = 0; i < n; ++i) dbias[i] *= learning_rate; }
  else { for (int i = 0; i < n; ++i) dbias[i] *= 5 * learning_rate; }
  if (flags & ORC_UPDATE_ALPHA)                                  // tdnn.cc:588-590
    for (int i = 0; i < n; ++i) dbias[i] *= 10000;

  // CuMatrix<BaseFloat> out_deriv_temp(out_deriv); the two PreconditionDirections calls   tdnn.cc:592-604
  OwnedMat out_deriv_temp(out_rows, out_dim);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < out_rows; ++r) memcpy(&out_deriv_temp.m(r, 0), &od(r, 0), sizeof(float) * out_dim);
  BaseFloat in_scale = 1.0f, out_scale = 1.0f;
  if (ng_in) NgPrecondition(static_cast<OrcNG*>(ng_in), in_value_temp.m, &in_scale);
  if (ng_out) NgPrecondition(static_cast<OrcNG*>(ng_out), out_deriv_temp.m, &out_scale);
  if (scales_out) { scales_out[0] = in_scale; scales_out[1] = out_scale; }
  const BaseFloat scale = in_scale * out_scale, local_lrate = scale * learning_rate;
  // bias tail: AddMatVec(local_lrate, out_deriv_temp, kTrans, precon_ones, 1.0)   tdnn.cc:607-617
#pragma omp parallel for schedule(static)
  for (int c = 0; c < out_dim; ++c) {
    double sum = 0.0;
    for (int r = 0; r < out_rows; ++r) sum += (double)out_deriv_temp.m(r, c) * (double)in_value_temp.m(r, spliced);
    dbias[n + c] += local_lrate * (BaseFloat)sum;
  }
  // linear_params_.AddMatMat(local_lrate, out_deriv_temp, kTrans, in_value_precon_part, kNoTrans, 1.0)   tdnn.cc:619-624
  Mat dlin{dW, out_dim, spliced, dw_stride};
  Mat precon = in_value_temp.m.Range(0, out_rows, 0, spliced);
  AddMatMat(dlin, local_lrate, out_deriv_temp.m, kTrans, precon, kNoTrans, 1.0f);
  return 0;
}

// The same with both preconditioners = identity (the raw-gradient form the first parity tests pin).
int orc_tdnn_backprop(const int* time_offsets, int n, int flags, float temp_proportion, const float* W, int w_stride,
                      const float* in_value, int in_rows, int in_dim, int in_stride, const float* out_deriv,
                      int out_rows, int out_dim, int od_stride, const float* coef_memo, const int* row_offsets,
                      int row_stride, float* in_deriv, int id_stride, float learning_rate, float* dW, int dw_stride,
                      float* dbias, float* s_out) {
  return orc_tdnn_backprop_ng(time_offsets, n, flags, temp_proportion, W, w_stride, in_value, in_rows, in_dim, in_stride,
                              out_deriv, out_rows, out_dim, od_stride, coef_memo, row_offsets, row_stride, in_deriv,
                              id_stride, learning_rate, dW, dw_stride, dbias, s_out, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------ stock TdnnComponent (BASELINE configs[1])
// The manual / derived TDNN-F systems (NAS/run_tdnn_7q_fbk_40_manual.sh and the models generate_top_list.py emits)
// are built from upstream Kaldi's TdnnComponent, the class TdnnDARTSV3Component was forked from: the reference's
// method bodies minus the architecture weights (every w_i = 1), bias_params_ of dimension D_out (no alpha slots)
// which Propagate always adds (tdnn.cc:230-241 keeps only that branch), and no memo.  upstream: kaldi
// src/nnet3/nnet-tdnn-component.cc (absent here, unpinned); structure as in tdnn.cc:214-333, 335-431, 457-626.
int orc_plain_tdnn_propagate(int n, const float* W, int w_stride, const float* bias, const float* in, int in_rows, int in_dim,
                             int in_stride, float* out, int out_rows, int out_dim, int out_stride, const int* row_offsets,
                             int row_stride) {
  Mat in_m{const_cast<float*>(in), in_rows, in_dim, in_stride};
  Mat out_m{out, out_rows, out_dim, out_stride};
  Mat lin{const_cast<float*>(W), out_dim, n * in_dim, w_stride};
  if (bias != nullptr)                                           // out->CopyRowsFromVec(bias_params_); else kPropagateAdds
    for (int r = 0; r < out_rows; ++r)
      for (int c = 0; c < out_dim; ++c) out_m(r, c) = bias[c];
  for (int i = 0; i < n; ++i) {
    Mat in_part = GetInputPart(in_m, out_rows, row_stride, row_offsets[i]);
    AddMatMat(out_m, 1.0f, in_part, kNoTrans, lin.Range(0, out_dim, i * in_dim, in_dim), kTrans, 1.0f);
  }
  return 0;
}

// Backprop + UpdateSimple (natural_gradient == 0) / UpdateNaturalGradient.  dbias (D_out) may be NULL (use-bias=false).
int orc_plain_tdnn_backprop(int n, const float* W, int w_stride, const float* in_value, int in_rows, int in_dim, int in_stride,
                            const float* out_deriv, int out_rows, int out_dim, int od_stride, const int* row_offsets,
                            int row_stride, float* in_deriv, int id_stride, float learning_rate, float* dW, int dw_stride,
                            float* dbias, int natural_gradient, void* ng_in, void* ng_out, float* scales_out) {
  Mat in_m{const_cast<float*>(in_value), in_rows, in_dim, in_stride};
  Mat od{const_cast<float*>(out_deriv), out_rows, out_dim, od_stride};
  Mat lin{const_cast<float*>(W), out_dim, n * in_dim, w_stride};
  if (in_deriv != nullptr) {
    Mat id{in_deriv, in_rows, in_dim, id_stride};
    for (int i = 0; i < n; ++i) {
      Mat id_part = GetInputPart(id, out_rows, row_stride, row_offsets[i]);
      AddMatMat(id_part, 1.0f, od, kNoTrans, lin.Range(0, out_dim, i * in_dim, in_dim), kNoTrans, 1.0f);
    }
  }
  if (dW == nullptr || learning_rate == 0.0f) return 0;
  const int spliced = n * in_dim;
  Mat dlin{dW, out_dim, spliced, dw_stride};
  if (!natural_gradient) {                                       // UpdateSimple (the shape of tdnn.cc:433-455)
    if (dbias != nullptr)
      for (int c = 0; c < out_dim; ++c) {
        double sum = 0.0;
        for (int r = 0; r < out_rows; ++r) sum += (double)od(r, c);
        dbias[c] += learning_rate * (BaseFloat)sum;
      }
    for (int i = 0; i < n; ++i) {
      Mat in_part = GetInputPart(in_m, out_rows, row_stride, row_offsets[i]);
      AddMatMat(dlin.Range(0, out_dim, i * in_dim, in_dim), learning_rate, od, kTrans, in_part, kNoTrans, 1.0f);
    }
    return 0;
  }
  // [X_1 | ... | X_n | 1]: the column of ones only when there is a bias (tdnn.cc:477-478: bias_params_.Dim() != 0)
  const bool has_bias = dbias != nullptr;
  OwnedMat in_value_temp(out_rows, spliced + (has_bias ? 1 : 0));
  if (has_bias)
    for (int r = 0; r < out_rows; ++r) in_value_temp.m(r, spliced) = 1.0f;
  for (int i = 0; i < n; ++i) {
    Mat in_part = GetInputPart(in_m, out_rows, row_stride, row_offsets[i]);
    for (int r = 0; r < out_rows; ++r) memcpy(&in_value_temp.m(r, i * in_dim), &in_part(r, 0), sizeof(float) * in_dim);
  }
  OwnedMat out_deriv_temp(out_rows, out_dim);
  for (int r = 0; r < out_rows; ++r) memcpy(&out_deriv_temp.m(r, 0), &od(r, 0), sizeof(float) * out_dim);
  BaseFloat in_scale = 1.0f, out_scale = 1.0f;
  if (ng_in) NgPrecondition(static_cast<OrcNG*>(ng_in), in_value_temp.m, &in_scale);
  if (ng_out) NgPrecondition(static_cast<OrcNG*>(ng_out), out_deriv_temp.m, &out_scale);
  if (scales_out) { scales_out[0] = in_scale; scales_out[1] = out_scale; }
  const BaseFloat local_lrate = in_scale * out_scale * learning_rate;
  if (dbias != nullptr)
    for (int c = 0; c < out_dim; ++c) {
      double sum = 0.0;
      for (int r = 0; r < out_rows; ++r) sum += (double)out_deriv_temp.m(r, c) * (double)in_value_temp.m(r, spliced);
      dbias[c] += local_lrate * (BaseFloat)sum;
    }
  AddMatMat(dlin, local_lrate, out_deriv_temp.m, kTrans, in_value_temp.m.Range(0, out_rows, 0, spliced), kNoTrans, 1.0f);
  return 0;
}

// ------------------------------------------------------------------ ConstrainOrthonormalInternal (utils.cc:914-1035)
// M (rows x cols, rows <= cols: the caller transposes otherwise, utils.cc:1067-1074) <- M - 4 alpha (M M^T - scale^2 I) M.
// scale < 0: the "floating" scale^2 = tr(P P^T) / tr(P) with the ratio-driven slow-down of update_speed.
// info (optional, 4): {scale used, ratio (0 if scale fixed), update_speed, Frobenius norm of P - scale^2 I}.
// Returns -1 where the reference asserts (scale == 0, ratio <= 0.999).
int orc_constrain_orthonormal(float scale, float* M_data, int rows, int cols, int stride, float* info) {
  if (scale == 0.0f) return -1;
  Mat M{M_data, rows, cols, stride};
  OwnedMat P(rows, rows), M_update(rows, cols);
  AddMatMat(P.m, 1.0f, M, kNoTrans, M, kTrans, 0.0f);            // SymAddMat2 + CopyLowerToUpper
  BaseFloat update_speed = 0.125f, ratio = 0.0f;
  if (scale < 0.0f) {
    double tr = 0.0, tr2 = 0.0;
    for (int i = 0; i < rows; ++i) {
      tr += P.m(i, i);
      for (int j = 0; j < rows; ++j) tr2 += (double)P.m(i, j) * P.m(i, j);
    }
    const BaseFloat trace_P = (BaseFloat)tr, trace_P_P = (BaseFloat)tr2;
    scale = std::sqrt(trace_P_P / trace_P);
    ratio = trace_P_P * rows / (trace_P * trace_P);
    if (!(ratio > 0.999f)) return -1;
    if (ratio > 1.02f) {
      update_speed *= 0.5f;
      if (ratio > 1.1f) update_speed *= 0.5f;
    }
  }
  for (int i = 0; i < rows; ++i) P.m(i, i) -= scale * scale;     // P.AddToDiag(-scale^2)
  if (info) {
    double e = 0.0;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < rows; ++j) e += (double)P.m(i, j) * P.m(i, j);
    info[0] = scale; info[1] = ratio; info[2] = update_speed; info[3] = (float)std::sqrt(e);
  }
  const BaseFloat alpha = update_speed / (scale * scale);
  AddMatMat(M_update.m, -4.0f * alpha, P.m, kNoTrans, M, kNoTrans, 0.0f);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) M(r, c) += M_update.m(r, c);
  return 0;
}

// ------------------------------------------------------------------ {Gumbel}SoftmaxFlops
// Propagate: simple.cc:10088-10113 (Gumbel, u != NULL, inv_temp = 1/T) and 9968-9981 (plain).
void orc_softmax_flops_fwd(const float* in, int rows, int cols, int in_stride, float* out, int out_stride,
                           const float* u, float temp_proportion) {
  std::vector<BaseFloat> rand_(cols, 0.f);
  if (u) {
    for (int j = 0; j < cols; ++j) rand_[j] = -1 * logf(-1 * logf(u[j]));
  }
  for (int r = 0; r < rows; ++r) {
    std::vector<BaseFloat> x(cols);
    for (int j = 0; j < cols; ++j) {
      x[j] = in[(size_t)r * in_stride + j];
      if (u) { x[j] += 1.0f * rand_[j]; x[j] *= (1.0f / temp_proportion); }
    }
    BaseFloat mx = -FLT_MAX;
    for (int j = 0; j < cols; ++j) mx = std::max(mx, x[j]);
    BaseFloat sum = 0.f;
    for (int j = 0; j < cols; ++j) { x[j] = expf(x[j] - mx); sum += x[j]; }
    for (int j = 0; j < cols; ++j) out[(size_t)r * out_stride + j] = std::max(x[j] / sum, 1.0e-20f);
  }
}

// Backprop: simple.cc:10116-10158 / 9984-10020.  MUTATES out_deriv (the reference writes through
// a const reference) unless in_deriv aliases it, in which case only in_deriv's final value is visible.
void orc_softmax_flops_bwd(const float* out_value, int ov_stride, float* out_deriv, int od_stride, float* in_deriv,
                           int id_stride, int rows, int cols, float scale, int is_gumbel, float temp_proportion) {
  std::vector<BaseFloat> flops_(cols, 0.f);
  const float f[8] = {-25, -50, -80, -100, -120, -160, -200, -240};
  for (int j = 0; j < 8 && j < cols; ++j) flops_[j] = f[j];
  const BaseFloat a = scale / rows / cols;                       // scale_/NumRows()/NumCols()
  for (int r = 0; r < rows; ++r)
    for (int j = 0; j < cols; ++j) out_deriv[(size_t)r * od_stride + j] += a * flops_[j];
  for (int r = 0; r < rows; ++r) {                               // DiffSoftmaxPerRow(out_value, out_deriv)
    BaseFloat pe = 0.f;
    for (int j = 0; j < cols; ++j) pe += out_value[(size_t)r * ov_stride + j] * out_deriv[(size_t)r * od_stride + j];
    for (int j = 0; j < cols; ++j) {
      const BaseFloat p = out_value[(size_t)r * ov_stride + j], e = out_deriv[(size_t)r * od_stride + j];
      BaseFloat d = p * e - p * pe;
      if (is_gumbel) d *= (1.0f / temp_proportion);              // in_deriv->Scale(1/T)
      in_deriv[(size_t)r * id_stride + j] = d;
    }
  }
}

// CopyNComponent, simple.cc:4843-4867 (AddMatBlocks).
void orc_copyn_fwd(const float* in, int rows, int in_cols, int in_stride, float* out, int out_cols, int out_stride,
                   float scale) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < out_cols; ++c) out[(size_t)r * out_stride + c] += scale * in[(size_t)r * in_stride + c % in_cols];
}
void orc_copyn_bwd(const float* od, int rows, int out_cols, int od_stride, float* id, int in_cols, int id_stride,
                   float scale) {
  for (int r = 0; r < rows; ++r)
    for (int j = 0; j < in_cols; ++j) {
      BaseFloat s = 0.f;
      for (int b = 0; b < out_cols / in_cols; ++b) s += od[(size_t)r * od_stride + b * in_cols + j];
      id[(size_t)r * id_stride + j] += scale * s;
    }
}

// OnehotFunctionComponent::Propagate, simple.cc:9504-9519.
void orc_onehot_fwd(float* out, int rows, int dim, int stride, float u) {
  std::vector<BaseFloat> onehot(dim, 0.f);
  for (int i = 0; i < dim; ++i)
    if (u >= (float)(i) / dim && u < (float)(i + 1) / dim) onehot[i] = 1.0f;
  for (int r = 0; r < rows; ++r)
    for (int i = 0; i < dim; ++i) out[(size_t)r * stride + i] = onehot[i];
}

// AddRowSumMat(scale, mat, 1.0): simple.cc:9544-9548.
void orc_add_row_sum(const float* mat, int rows, int cols, int stride, float scale, float* vec) {
  for (int c = 0; c < cols; ++c) {
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += mat[(size_t)r * stride + c];
    vec[c] += scale * (BaseFloat)s;
  }
}

// BatchNormTestComponent::ComputeDerived, norm.cc:680-713 (stats in double as in the reference's CuVector<double>).
void orc_bn_test_derived(const double* stats_sum, const double* stats_sumsq, double count, int dim, float epsilon,
                         float target_rms, float* scale, float* offset) {
  for (int i = 0; i < dim; ++i) {
    BaseFloat off = (BaseFloat)stats_sum[i];     // offset_.CopyFromVec(stats_sum_) (double -> float)
    off *= (BaseFloat)(-1.0 / count);
    BaseFloat sc = (BaseFloat)stats_sumsq[i];
    sc *= (BaseFloat)(1.0 / count);
    sc += -1.0f * off * off;
    sc = std::max(sc, 0.0f);
    sc += epsilon;
    sc = powf(sc, -0.5f);
    sc *= target_rms;
    off *= sc;
    scale[i] = sc;
    offset[i] = off;
  }
}
// Propagate / Backprop in test mode, norm.cc:868-876, 915-921.
void orc_scale_offset_rows(const float* in, int rows, int cols, int in_stride, float* out, int out_stride,
                           const float* scale, const float* offset) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      BaseFloat y = in[(size_t)r * in_stride + c];               // CopyFromMat
      y *= scale[c];                                             // MulColsVec
      if (offset) y += 1.0f * offset[c];                         // AddVecToRows
      out[(size_t)r * out_stride + c] = y;
    }
}

// ElementwiseProductComponent, simple.cc:256-299.
void orc_ewprod_fwd(const float* in, int rows, int D, int in_stride, float* out, int out_stride) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < D; ++c) out[(size_t)r * out_stride + c] = in[(size_t)r * in_stride + c] * in[(size_t)r * in_stride + D + c];
}
void orc_ewprod_bwd(const float* in, int in_stride, const float* od, int od_stride, float* id, int id_stride, int rows,
                    int D) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < D; ++c) {
      const BaseFloat g = od[(size_t)r * od_stride + c];
      id[(size_t)r * id_stride + c] = g * in[(size_t)r * in_stride + D + c];
      id[(size_t)r * id_stride + D + c] = g * in[(size_t)r * in_stride + c];
    }
}

// ------------------------------------------------------------------ chain denominator
// kaldi chain-denominator.cc, CPU code path.  alpha: (T+1) x (N*S + S), beta: 2 x (N*S + S),
// element (t, h*S + s); the S trailing columns hold the per-sequence alpha sums.
// Transitions: fwd_ranges/bwd_ranges [N][2] into (prob, pdf, state) arrays.
// Returns the total log-prob; if nnet_output_deriv != NULL runs Backward(deriv_weight) and sets *ok.
float orc_den_forward_backward(int N, int P, const int* fwd_ranges, const int* bwd_ranges, const float* tprob,
                               const int* tpdf, const int* tstate, const float* init, int S, int T, float leaky,
                               const float* nnet_output, int no_stride, float deriv_weight, float* nnet_output_deriv,
                               int nd_stride, int* ok) {
  const size_t asz = (size_t)N * S;
  // exp_nnet_output_transposed_: P x (T*S), ApplyExpLimited(-30, 30)
  std::vector<BaseFloat> E((size_t)P * T * S);
#pragma omp parallel for schedule(static)
  for (int p = 0; p < P; ++p)
    for (int ts = 0; ts < T * S; ++ts) {
      BaseFloat x = nnet_output[(size_t)ts * no_stride + p];
      x = std::min(std::max(x, -30.0f), 30.0f);
      E[(size_t)p * T * S + ts] = expf(x);
    }
  std::vector<BaseFloat> alpha((size_t)(T + 1) * (asz + S), 0.f);
  auto arow = [&](int t) { return alpha.data() + (size_t)t * (asz + S); };
  auto alpha_dash = [&](int t) {  // AlphaDash(t)
    BaseFloat* a = arow(t);
    BaseFloat* sum = a + asz;
    for (int s = 0; s < S; ++s) {
      double acc = 0.0;  // AddRowSumMat
      for (int h = 0; h < N; ++h) acc += a[(size_t)h * S + s];
      sum[s] = (BaseFloat)acc;
    }
    for (int h = 0; h < N; ++h)  // alpha_mat.AddVecVec(leaky, initial_probs, alpha_sum_vec)
      for (int s = 0; s < S; ++s) a[(size_t)h * S + s] += leaky * init[h] * sum[s];
  };
  // AlphaFirstFrame
  for (int h = 0; h < N; ++h)
    for (int s = 0; s < S; ++s) arow(0)[(size_t)h * S + s] = init[h];
  alpha_dash(0);
  for (int t = 1; t <= T; ++t) {  // AlphaGeneralFrame(t)
    const BaseFloat* prev = arow(t - 1);
    BaseFloat* cur = arow(t);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < N; ++h) {
      for (int s = 0; s < S; ++s) {
        double this_tot_alpha = 0.0;
        for (int a = bwd_ranges[2 * h]; a < bwd_ranges[2 * h + 1]; ++a) {
          const BaseFloat prob = E[(size_t)tpdf[a] * T * S + (size_t)(t - 1) * S + s];
          this_tot_alpha += prev[(size_t)tstate[a] * S + s] * tprob[a] * prob;
        }
        const BaseFloat arbitrary_scale = 1.0f / prev[asz + s];
        cur[(size_t)h * S + s] = (BaseFloat)this_tot_alpha * arbitrary_scale;
      }
    }
    alpha_dash(t);
  }
  // ComputeTotLogLike
  std::vector<BaseFloat> tot_prob(S);
  double tot_log_prob = 0.0;
  for (int s = 0; s < S; ++s) {
    double acc = 0.0;
    for (int h = 0; h < N; ++h) acc += arow(T)[(size_t)h * S + s];
    tot_prob[s] = (BaseFloat)acc;
    tot_log_prob += logf(tot_prob[s]);
  }
  double log_inv_scales = 0.0;
  for (int t = 0; t < T; ++t)
    for (int s = 0; s < S; ++s) log_inv_scales += logf(arow(t)[asz + s]);
  const float logprob = (float)(tot_log_prob + log_inv_scales);
  if (nnet_output_deriv == nullptr) return logprob;

  // Backward
  bool ok_ = true;
  std::vector<BaseFloat> beta((size_t)2 * (asz + S), 0.f);
  auto brow = [&](int t) { return beta.data() + (size_t)(t % 2) * (asz + S); };
  std::vector<BaseFloat> gamma((size_t)P * T * S, 0.f);  // nnet_output_deriv_transposed_ over all frames
  {  // BetaDashLastFrame
    BaseFloat* b = brow(T);
    for (int h = 0; h < N; ++h)
      for (int s = 0; s < S; ++s) b[(size_t)h * S + s] = 1.0f / tot_prob[s];
  }
  auto beta_fn = [&](int t) {  // Beta(t): beta = beta_dash + leaky * (init . beta_dash)
    BaseFloat* b = brow(t);
    for (int s = 0; s < S; ++s) {
      double acc = 0.0;
      for (int h = 0; h < N; ++h) acc += (double)init[h] * b[(size_t)h * S + s];
      b[asz + s] = (BaseFloat)acc;
    }
    for (int h = 0; h < N; ++h)
      for (int s = 0; s < S; ++s) b[(size_t)h * S + s] += leaky * b[asz + s];
  };
  beta_fn(T);
  for (int t = T - 1; t >= 0; --t) {  // BetaDashGeneralFrame(t)
    const BaseFloat* this_alpha_dash = arow(t);
    const BaseFloat* next_beta = brow(t + 1);
    BaseFloat* this_beta_dash = brow(t);
    // serial over h (the derivative accumulation collides across states)
    for (int h = 0; h < N; ++h) {
      for (int s = 0; s < S; ++s) {
        const BaseFloat this_alpha_dash_prob = this_alpha_dash[(size_t)h * S + s],
                        inv_arbitrary_scale = this_alpha_dash[asz + s];
        double tot_variable_factor = 0.0;
        const BaseFloat occupation_factor = this_alpha_dash_prob / inv_arbitrary_scale;
        for (int a = fwd_ranges[2 * h]; a < fwd_ranges[2 * h + 1]; ++a) {
          const size_t eidx = (size_t)tpdf[a] * T * S + (size_t)t * S + s;
          const BaseFloat variable_factor = tprob[a] * next_beta[(size_t)tstate[a] * S + s] * E[eidx];
          tot_variable_factor += variable_factor;
          gamma[eidx] += variable_factor * occupation_factor;
        }
        this_beta_dash[(size_t)h * S + s] = (BaseFloat)tot_variable_factor / inv_arbitrary_scale;
      }
    }
    if (t == 0) {  // BetaGeneralFrameDebug(0)
      double alpha_beta_product = 0.0;
      for (size_t i = 0; i < asz; ++i) alpha_beta_product += (double)this_alpha_dash[i] * this_beta_dash[i];
      if (!(fabs(alpha_beta_product - S) <= 2.0)) ok_ = false;
    }
    beta_fn(t);
  }
  for (int ts = 0; ts < T * S; ++ts)
    for (int p = 0; p < P; ++p)
      nnet_output_deriv[(size_t)ts * nd_stride + p] += deriv_weight * gamma[(size_t)p * T * S + ts];
  if (ok) *ok = ok_ ? 1 : 0;
  return logprob;
}

// ------------------------------------------------------------------ chain numerator (generic, per-sequence FST)
// kaldi chain-generic-numerator.cc (CPU, log domain), restated in float64: alpha(0,start)=0;
// alpha(t+1,dst) = logsum_arcs alpha(t,src) + w + x(t,pdf); total = logsum_h alpha(T,h) + final(h);
// posterior(arc,t) = exp(alpha(t,src) + w + x(t,pdf) + beta(t+1,dst) - total).
// Arrays as in tdnnf_num_graph_create (only the forward arc list is used here).  Returns the summed log-prob;
// *ok = 0 if some sequence has no complete path (its derivative is then left untouched).
double orc_num_forward_backward(int S, const int* state_offsets, const int* fwd_ranges, const float* arc_logprob,
                                const int* arc_pdf, const int* arc_state, const float* final_logprob,
                                const float* nnet_output, int no_stride, int T, float deriv_weight, float* deriv,
                                int d_stride, int* ok) {
  const double kZero = -1.0e30;
  auto log_add = [&](double a, double b) {
    if (a < b) std::swap(a, b);
    if (b <= kZero) return a;
    return a + log1p(exp(b - a));
  };
  double total_all = 0.0;
  *ok = 1;
  for (int s = 0; s < S; ++s) {
    const int s0 = state_offsets[s], ns = state_offsets[s + 1] - s0;
    std::vector<double> alpha((size_t)(T + 1) * ns, kZero), beta((size_t)(T + 1) * ns, kZero);
    alpha[0] = 0.0;
    for (int t = 0; t < T; ++t)
      for (int h = 0; h < ns; ++h) {
        const double a = alpha[(size_t)t * ns + h];
        if (a <= kZero) continue;
        for (int e = fwd_ranges[2 * (s0 + h)]; e < fwd_ranges[2 * (s0 + h) + 1]; ++e) {
          double& d = alpha[(size_t)(t + 1) * ns + (arc_state[e] - s0)];
          d = log_add(d, a + arc_logprob[e] + nnet_output[((size_t)t * S + s) * no_stride + arc_pdf[e]]);
        }
      }
    double tot = kZero;
    for (int h = 0; h < ns; ++h) {
      beta[(size_t)T * ns + h] = final_logprob[s0 + h] > kZero ? final_logprob[s0 + h] : kZero;
      if (final_logprob[s0 + h] > kZero && alpha[(size_t)T * ns + h] > kZero) tot = log_add(tot, alpha[(size_t)T * ns + h] + final_logprob[s0 + h]);
    }
    if (tot <= kZero) { *ok = 0; continue; }
    total_all += tot;
    if (!deriv) continue;
    for (int t = T - 1; t >= 0; --t)
      for (int h = 0; h < ns; ++h) {
        double acc = kZero;
        for (int e = fwd_ranges[2 * (s0 + h)]; e < fwd_ranges[2 * (s0 + h) + 1]; ++e) {
          const double b = beta[(size_t)(t + 1) * ns + (arc_state[e] - s0)];
          if (b <= kZero) continue;
          const double v = arc_logprob[e] + nnet_output[((size_t)t * S + s) * no_stride + arc_pdf[e]] + b;
          acc = log_add(acc, v);
          const double a = alpha[(size_t)t * ns + h];
          if (a > kZero) deriv[((size_t)t * S + s) * d_stride + arc_pdf[e]] += deriv_weight * (float)exp(a + v - tot);
        }
        beta[(size_t)t * ns + h] = acc;
      }
  }
  return total_all;
}

}  // extern "C"
