// Whole-parameter ops of the updatable components and the stock TDNN-F neighbours (ReLU, bypass
// sum, BatchNorm training mode).  All bandwidth-bound, one fused pass each.
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "context.h"
#include "ptx.cuh"

using namespace tdnnf;

namespace {

inline int grid_for(long long total, int threads, int num_sms) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define ELEMWISE_LOOP(total)                                                                      \
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (total);           \
       idx += (long long)gridDim.x * blockDim.x)

__global__ void mat_set_kernel(float* a, int rows, int cols, long long stride, float v) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) a[(idx / cols) * stride + idx % cols] = v;
}
__global__ void copy_rows_from_vec_kernel(const float* __restrict__ vec, float* __restrict__ out, int rows, int cols,
                                          long long stride) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) out[(idx / cols) * stride + idx % cols] = vec[idx % cols];
}
__global__ void copy_rows_kernel(const float* __restrict__ src, long long ss, float* __restrict__ dst, long long ds,
                                 int rows, int cols, const int32_t* __restrict__ map) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    const int m = map[r];
    dst[r * ds + c] = m >= 0 ? src[(long long)m * ss + c] : 0.f;
  }
}
__global__ void add_to_rows_kernel(float alpha, const float* __restrict__ src, long long ss, int rows, int cols,
                                   float* __restrict__ dst, long long ds, const int32_t* __restrict__ map) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    const int m = map[r];
    if (m >= 0) dst[(long long)m * ds + c] += alpha * src[r * ss + c];
  }
}
__global__ void mat_scale_kernel(float* a, int rows, int cols, long long stride, float s) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) a[(idx / cols) * stride + idx % cols] *= s;
}
__global__ void mat_axpy_kernel(float alpha, const float* __restrict__ src, long long ss, float* __restrict__ dst,
                                long long ds, int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    dst[r * ds + c] += alpha * src[r * ss + c];
  }
}
__global__ void mat_dot_kernel(const float* __restrict__ a, long long as, const float* __restrict__ b, long long bs,
                               int rows, int cols, double* __restrict__ result) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ double red[256];
  double acc = 0.0;
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    acc += (double)a[r * as + c] * (double)b[r * bs + c];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(result, red[0]);
}

__global__ void relu_fwd_kernel(const float* __restrict__ in, int rows, int cols, long long is, float* __restrict__ out,
                                long long os, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (vec) {
    const int c4 = cols >> 2;
    ELEMWISE_LOOP((long long)rows * c4) {
      const long long r = idx / c4;
      const int c = (int)(idx % c4) * 4;
      float4 x = *reinterpret_cast<const float4*>(in + r * is + c);
      x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
      *reinterpret_cast<float4*>(out + r * os + c) = x;
    }
  } else {
    ELEMWISE_LOOP((long long)rows * cols) {
      const long long r = idx / cols;
      const int c = (int)(idx % cols);
      out[r * os + c] = fmaxf(in[r * is + c], 0.f);
    }
  }
}
__global__ void relu_bwd_kernel(const float* __restrict__ ov, long long vs, const float* __restrict__ od, long long ds,
                                float* __restrict__ id, long long is, int rows, int cols, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (vec) {
    const int c4 = cols >> 2;
    ELEMWISE_LOOP((long long)rows * c4) {
      const long long r = idx / c4;
      const int c = (int)(idx % c4) * 4;
      const float4 v = *reinterpret_cast<const float4*>(ov + r * vs + c);
      float4 g = *reinterpret_cast<const float4*>(od + r * ds + c);
      g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f; g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
      *reinterpret_cast<float4*>(id + r * is + c) = g;
    }
  } else {
    ELEMWISE_LOOP((long long)rows * cols) {
      const long long r = idx / cols;
      const int c = (int)(idx % cols);
      id[r * is + c] = ov[r * vs + c] > 0.f ? od[r * ds + c] : 0.f;
    }
  }
}
__global__ void add_scaled_kernel(const float* a, long long as, float alpha, const float* b, long long bs, float beta,
                                  float* out, long long os, int rows, int cols, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (vec) {
    const int c4 = cols >> 2;
    ELEMWISE_LOOP((long long)rows * c4) {
      const long long r = idx / c4;
      const int c = (int)(idx % c4) * 4;
      const float4 x = *reinterpret_cast<const float4*>(a + r * as + c);
      const float4 y = *reinterpret_cast<const float4*>(b + r * bs + c);
      *reinterpret_cast<float4*>(out + r * os + c) =
          make_float4(alpha * x.x + beta * y.x, alpha * x.y + beta * y.y, alpha * x.z + beta * y.z, alpha * x.w + beta * y.w);
    }
  } else {
    ELEMWISE_LOOP((long long)rows * cols) {
      const long long r = idx / cols;
      const int c = (int)(idx % cols);
      out[r * os + c] = alpha * a[r * as + c] + beta * b[r * bs + c];
    }
  }
}

// Column statistics: sums[0][c] += sum_r f(r,c), sums[1][c] += sum_r g(r,c).  Block = (32 cols, 8 row lanes).
//   MODE 0: f = x, g = x*x                       (BatchNorm forward statistics)
//   MODE 1: f = z' (od), g = z' * z (od * ov)    (BatchNorm backward statistics)
template <int MODE>
__global__ void col_stats_kernel(const float* __restrict__ a, long long as, const float* __restrict__ b, long long bs,
                                 int rows, int cols, float* __restrict__ sums /* 2 x cols */) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red0[8][33], red1[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (col < cols) {
    for (long long r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      const float x = a[r * as + col];
      if (MODE == 0) { s0 += x; s1 += x * x; }
      else { s0 += x; s1 += x * b[r * bs + col]; }
    }
  }
  red0[threadIdx.y][threadIdx.x] = s0;
  red1[threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += red0[i][threadIdx.x]; t1 += red1[i][threadIdx.x]; }
    atomicAdd(sums + col, t0);
    atomicAdd(sums + cols + col, t1);
  }
}

// memo rows: 0 mean, 1 uvar, 2 scale ; rows 3,4 scratch sums
__global__ void bn_finalize_fwd_kernel(float* memo, int cols, int rows, float eps, float target_rms) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const float mean = memo[3 * cols + c] / rows;
  const float uvar = memo[4 * cols + c] / rows;  // row 1 = the UNCENTRED second moment, as the reference's memo
  float var = uvar - mean * mean;                //   (uvar.AddDiagMat2, norm.cc:432): StoreStats accumulates it
  var = fmaxf(var, 0.f);
  memo[c] = mean;
  memo[cols + c] = uvar;
  memo[2 * cols + c] = target_rms * powf(var + eps, -0.5f);
}
__global__ void bn_apply_fwd_kernel(const float* __restrict__ in, long long is, float* __restrict__ out, long long os,
                                    int rows, int cols, const float* __restrict__ memo) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    out[r * os + c] = (in[r * is + c] - memo[c]) * memo[2 * cols + c];
  }
}
__global__ void bn_apply_bwd_kernel(const float* __restrict__ ov, long long vs, const float* __restrict__ od,
                                    long long ds, float* __restrict__ id, long long is, int rows, int cols,
                                    float target_rms, const float* __restrict__ memo, const float* __restrict__ sums) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    const float scale = memo[2 * cols + c];
    const float mean_od = sums[c] / rows;
    const float var_deriv_mod = -1.0f / (target_rms * target_rms) * (sums[cols + c] / rows) * scale;
    id[r * is + c] = scale * (od[r * ds + c] - mean_od) + ov[r * vs + c] * var_deriv_mod;
  }
}

__global__ void tail_fwd_kernel(const float* __restrict__ x, long long xs, const float* __restrict__ scale,
                                const float* __restrict__ offset, const float* __restrict__ prev, long long ps, float bs,
                                float* __restrict__ out, long long os, int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int c4 = cols >> 2;
  ELEMWISE_LOOP((long long)rows * c4) {
    const long long r = idx / c4;
    const int c = (int)(idx % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * xs + c);
    const float4 s = *reinterpret_cast<const float4*>(scale + c);
    const float4 o = *reinterpret_cast<const float4*>(offset + c);
    const float4 p = *reinterpret_cast<const float4*>(prev + r * ps + c);
    float4 y;
    y.x = fmaxf(v.x, 0.f) * s.x + o.x + bs * p.x;
    y.y = fmaxf(v.y, 0.f) * s.y + o.y + bs * p.y;
    y.z = fmaxf(v.z, 0.f) * s.z + o.z + bs * p.z;
    y.w = fmaxf(v.w, 0.f) * s.w + o.w + bs * p.w;
    *reinterpret_cast<float4*>(out + r * os + c) = y;
  }
}
__global__ void tail_bwd_kernel(const float* __restrict__ dout, long long dos, const float* __restrict__ x, long long xs,
                                const float* __restrict__ scale, float bs, float* __restrict__ dx, long long dxs,
                                float* __restrict__ dprev, long long dps, int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int c4 = cols >> 2;
  ELEMWISE_LOOP((long long)rows * c4) {
    const long long r = idx / c4;
    const int c = (int)(idx % c4) * 4;
    const float4 g = *reinterpret_cast<const float4*>(dout + r * dos + c);
    const float4 v = *reinterpret_cast<const float4*>(x + r * xs + c);
    const float4 s = *reinterpret_cast<const float4*>(scale + c);
    *reinterpret_cast<float4*>(dx + r * dxs + c) =
        make_float4(v.x > 0.f ? g.x * s.x : 0.f, v.y > 0.f ? g.y * s.y : 0.f, v.z > 0.f ? g.z * s.z : 0.f, v.w > 0.f ? g.w * s.w : 0.f);
    *reinterpret_cast<float4*>(dprev + r * dps + c) = make_float4(bs * g.x, bs * g.y, bs * g.z, bs * g.w);
  }
}

// The fused tails as PRODUCERS of the next GEMM's operand: besides the fp32 matrix they write its bf16 hi/lo row planes
// (the layout of split_rows_kernel with one row group), the per-row sums of squares (tr(X X^T) of the natural gradient)
// and, in the backward form, the column sums (the bias gradient) -- the matrix is not read again by a split pass.
// One block per row at a time, thread = 4 consecutive columns (blockDim.x = Kpad / 4 <= 1024; threads beyond cols / 4 only
// zero the K padding of the planes).
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long o, float4 y) {
  __align__(8) __nv_bfloat16 h[4];
  __align__(8) __nv_bfloat16 l[4];
  const float v[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(v[j]);
    l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
  }
  *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l);
}
__device__ __forceinline__ float block_sum(float v, float* red /* [32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // red may still be read from the previous row
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}
__global__ void tail_fwd_planes_kernel(const float* __restrict__ x, long long xs, const float* __restrict__ scale,
                                       const float* __restrict__ offset, const float* __restrict__ prev, long long ps, float bs,
                                       float* __restrict__ out, long long os, int rows, int cols, int Kpad,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ rowsq) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red[32];
  const int c = threadIdx.x * 4;
  const bool live = c < cols;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), o = s;
  if (live) {
    s = *reinterpret_cast<const float4*>(scale + c);
    o = *reinterpret_cast<const float4*>(offset + c);
  }
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * xs + c);
      const float4 p = *reinterpret_cast<const float4*>(prev + r * ps + c);
      y.x = fmaxf(v.x, 0.f) * s.x + o.x + bs * p.x;
      y.y = fmaxf(v.y, 0.f) * s.y + o.y + bs * p.y;
      y.z = fmaxf(v.z, 0.f) * s.z + o.z + bs * p.z;
      y.w = fmaxf(v.w, 0.f) * s.w + o.w + bs * p.w;
      *reinterpret_cast<float4*>(out + r * os + c) = y;
    }
    store_planes4(hi, lo, r * Kpad + c, y);
    const float sq = block_sum(y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w, red);
    if (threadIdx.x == 0) rowsq[r] = sq;
  }
}
__global__ void tail_bwd_planes_kernel(const float* __restrict__ dout, long long dos, const float* __restrict__ x, long long xs,
                                       const float* __restrict__ scale, float bs, float* __restrict__ dx, long long dxs,
                                       float* __restrict__ dprev, long long dps, int rows, int cols, int Kpad,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ rowsq,
                                       float* __restrict__ colsum) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red[32];
  const int c = threadIdx.x * 4;
  const bool live = c < cols;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), cs = s;
  if (live) s = *reinterpret_cast<const float4*>(scale + c);
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const float4 g = *reinterpret_cast<const float4*>(dout + r * dos + c);
      const float4 v = *reinterpret_cast<const float4*>(x + r * xs + c);
      y = make_float4(v.x > 0.f ? g.x * s.x : 0.f, v.y > 0.f ? g.y * s.y : 0.f, v.z > 0.f ? g.z * s.z : 0.f, v.w > 0.f ? g.w * s.w : 0.f);
      *reinterpret_cast<float4*>(dx + r * dxs + c) = y;
      *reinterpret_cast<float4*>(dprev + r * dps + c) = make_float4(bs * g.x, bs * g.y, bs * g.z, bs * g.w);
      cs.x += y.x; cs.y += y.y; cs.z += y.z; cs.w += y.w;
    }
    store_planes4(hi, lo, r * Kpad + c, y);
    const float sq = block_sum(y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w, red);
    if (threadIdx.x == 0) rowsq[r] = sq;
  }
  if (live) {
    atomicAdd(colsum + c, cs.x);
    atomicAdd(colsum + c + 1, cs.y);
    atomicAdd(colsum + c + 2, cs.z);
    atomicAdd(colsum + c + 3, cs.w);
  }
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define LAUNCH_CHECK(ctx)              \
  do {                                 \
    (ctx)->launches++;                 \
    TDNNF_CUDA_OK(cudaGetLastError()); \
  } while (0)
#define PROLOGUE(cond, msg)                       \
  TDNNF_REQUIRE(ctx != nullptr, "null context");  \
  TDNNF_REQUIRE(cond, msg);                       \
  if (rows == 0 || cols == 0) return TDNNF_OK;    \
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device))

extern "C" int tdnnf_mat_set(tdnnf_ctx* ctx, float* a, int rows, int cols, int stride, float value) {
  PROLOGUE(a && rows >= 0 && cols >= 0 && stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(mat_set_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, a,
                           rows, cols, stride, value));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_copy_rows_from_vec(tdnnf_ctx* ctx, const float* vec, float* out, int rows, int cols, int stride) {
  PROLOGUE(vec && out && rows >= 0 && cols >= 0 && stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(copy_rows_from_vec_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, vec, out, rows, cols,
                                                                                                         stride));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_copy_rows(tdnnf_ctx* ctx, const float* src, int src_stride, float* dst, int dst_stride, int rows,
                               int cols, const int32_t* row_map) {
  PROLOGUE(src && dst && row_map && src_stride >= cols && dst_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(copy_rows_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, src, src_stride, dst, dst_stride,
                                                                                                 rows, cols, row_map));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_add_to_rows(tdnnf_ctx* ctx, float alpha, const float* src, int src_stride, int rows, int cols,
                                 float* dst, int dst_stride, const int32_t* row_map) {
  PROLOGUE(src && dst && row_map && src_stride >= cols && dst_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(add_to_rows_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, alpha, src, src_stride, rows,
                                                                                                   cols, dst, dst_stride, row_map));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_mat_scale(tdnnf_ctx* ctx, float* a, int rows, int cols, int stride, float scale) {
  PROLOGUE(a && rows >= 0 && cols >= 0 && stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(mat_scale_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, a, rows, cols, stride, scale));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_mat_axpy(tdnnf_ctx* ctx, float alpha, const float* src, int src_stride, float* dst,
                              int dst_stride, int rows, int cols) {
  PROLOGUE(src && dst && src_stride >= cols && dst_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(mat_axpy_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, alpha, src, src_stride, dst,
                                                                                                dst_stride, rows, cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_mat_dot_dev(tdnnf_ctx* ctx, const float* a, int a_stride, const float* b, int b_stride, int rows,
                                 int cols, double* result_dev) {
  PROLOGUE(a && b && result_dev && a_stride >= cols && b_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(mat_dot_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, a, a_stride, b, b_stride, rows,
                                                                                               cols, result_dev));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_mat_dot(tdnnf_ctx* ctx, const float* a, int a_stride, const float* b, int b_stride, int rows,
                             int cols, float* result) {
  TDNNF_REQUIRE(ctx && result, "null argument");
  *result = 0.f;
  if (rows == 0 || cols == 0) return TDNNF_OK;
  ctx->ws_reset();
  int rc = ctx->ws_reserve(64);
  if (rc) return rc;
  double* acc = static_cast<double*>(ctx->ws_alloc(sizeof(double)));
  if (!acc) return TDNNF_ERR_NOMEM;
  TDNNF_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), ctx->stream));
  rc = tdnnf_mat_dot_dev(ctx, a, a_stride, b, b_stride, rows, cols, acc);
  if (rc) return rc;
  double h = 0.0;
  TDNNF_CUDA_OK(cudaMemcpyAsync(&h, acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  *result = (float)h;
  return TDNNF_OK;
}

// ---- UpdateNnetWithMaxChange over many parameter buffers in two launches (utils.cc:2085-2175).  A network has
// ~70 parameter buffers; one kernel launch (and one host call) per buffer per operation made the parameter step
// host-bound (~200 calls).  The buffer table travels as a kernel argument.
struct MultiBufTable {
  const float* src[TDNNF_MULTI_MAX];
  float* dst[TDNNF_MULTI_MAX];
  int rows[TDNNF_MULTI_MAX], cols[TDNNF_MULTI_MAX], src_stride[TDNNF_MULTI_MAX], dst_stride[TDNNF_MULTI_MAX];
  int group[TDNNF_MULTI_MAX];
  float factor[TDNNF_MULTI_MAX];
  int first_block[TDNNF_MULTI_MAX + 1];  // blocks [first_block[i], first_block[i+1]) work on buffer i
  int n;
};

__device__ __forceinline__ int multi_find(const MultiBufTable& t, int block) {
  int lo = 0, hi = t.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.first_block[mid] <= block) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// out[group[i]] += sum of squares of buffer i
__global__ void __launch_bounds__(256) multi_sumsq_kernel(const __grid_constant__ MultiBufTable t, double* __restrict__ out) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int i = multi_find(t, blockIdx.x);
  const int nb = t.first_block[i + 1] - t.first_block[i], b = blockIdx.x - t.first_block[i];
  const long long total = (long long)t.rows[i] * t.cols[i];
  const float* src = t.src[i];
  const int cols = t.cols[i];
  const long long ld = t.src_stride[i];
  double acc = 0.0;
  for (long long idx = (long long)b * 256 + threadIdx.x; idx < total; idx += (long long)nb * 256) {
    const float v = src[(idx / cols) * ld + idx % cols];
    acc += (double)v * (double)v;
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    if (s != 0.0) atomicAdd(out + t.group[i], s);
  }
}

// dst_i += factor_i * src_i;  src_i = keep * src_i.  factor 0 adds nothing (0 * inf would be NaN: Kaldi's
// UpdateNnetWithMaxChange leaves the model untouched on a non-finite delta); keep 0 stores exact zeros
// (ScaleNnet(0.0) is SetZero), keep 1 leaves the source alone (ApplyL2Regularization reads the model).
__global__ void __launch_bounds__(256) multi_axpy_zero_kernel(const __grid_constant__ MultiBufTable t, float keep) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int i = multi_find(t, blockIdx.x);
  const int nb = t.first_block[i + 1] - t.first_block[i], b = blockIdx.x - t.first_block[i];
  const long long total = (long long)t.rows[i] * t.cols[i];
  float* src = const_cast<float*>(t.src[i]);
  float* dst = t.dst[i];
  const int cols = t.cols[i];
  const long long sld = t.src_stride[i], dld = t.dst_stride[i];
  const float f = t.factor[i];
  for (long long idx = (long long)b * 256 + threadIdx.x; idx < total; idx += (long long)nb * 256) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    if (f != 0.f) dst[r * dld + c] += f * src[r * sld + c];
    if (keep == 0.f) src[r * sld + c] = 0.f;
    else if (keep != 1.f) src[r * sld + c] *= keep;
  }
}

static int multi_table(tdnnf_ctx* ctx, int n, const float* const* src, float* const* dst, const int32_t* rows, const int32_t* cols,
                       const int32_t* src_strides, const int32_t* dst_strides, const int32_t* groups, const float* factors,
                       MultiBufTable* t, int* blocks) {
  TDNNF_REQUIRE(ctx && src && rows && cols && src_strides, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MULTI_MAX, "buffer count out of range");
  memset(t, 0, sizeof(*t));
  t->n = n;
  int nb = 0;
  for (int i = 0; i < n; ++i) {
    TDNNF_REQUIRE(src[i] && rows[i] > 0 && cols[i] > 0 && src_strides[i] >= cols[i], "bad buffer");
    t->src[i] = src[i];
    t->dst[i] = dst ? dst[i] : nullptr;
    t->rows[i] = rows[i];
    t->cols[i] = cols[i];
    t->src_stride[i] = src_strides[i];
    t->dst_stride[i] = dst_strides ? dst_strides[i] : 0;
    t->group[i] = groups ? groups[i] : i;
    t->factor[i] = factors ? factors[i] : 1.f;
    t->first_block[i] = nb;
    const long long total = (long long)rows[i] * cols[i];
    nb += (int)std::max<long long>(1, std::min<long long>((total + 4095) / 4096, 64));
  }
  t->first_block[n] = nb;
  *blocks = nb;
  return TDNNF_OK;
}

extern "C" int tdnnf_multi_sumsq(tdnnf_ctx* ctx, int n, const float* const* bufs, const int32_t* rows, const int32_t* cols,
                                 const int32_t* strides, const int32_t* groups, double* out_dev) {
  TDNNF_REQUIRE(out_dev != nullptr, "null argument");
  MultiBufTable t;
  int blocks = 0;
  int rc = multi_table(ctx, n, bufs, nullptr, rows, cols, strides, nullptr, groups, nullptr, &t, &blocks);
  if (rc) return rc;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(multi_sumsq_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, t, out_dev));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_multi_axpy_zero(tdnnf_ctx* ctx, int n, float* const* dst, const int32_t* dst_strides, float* const* src,
                                     const int32_t* src_strides, const int32_t* rows, const int32_t* cols, const float* factors) {
  TDNNF_REQUIRE(dst && dst_strides && factors, "null argument");
  MultiBufTable t;
  int blocks = 0;
  int rc = multi_table(ctx, n, src, dst, rows, cols, src_strides, dst_strides, nullptr, factors, &t, &blocks);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) TDNNF_REQUIRE(dst[i] && dst_strides[i] >= cols[i], "bad destination buffer");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(multi_axpy_zero_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, t, 0.f));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// UpdateNnetWithMaxChange + ScaleNnet(momentum, delta_nnet) (ref: nnet-utils.cc:2085-2175; kaldi: NnetChainTrainer::TrainInternal)
// as ONE entry: squared norms of every component's delta on the device (one launch), ONE read-back, the per-component /
// global factors on the host exactly as the reference computes them, then model += factor * delta and delta *= momentum
// (one launch).  *applied = 0 reproduces "Infinite parameter change, will not apply.": the model is untouched and the
// delta is still scaled by momentum (the trainer calls ScaleNnet either way).
extern "C" int tdnnf_update_with_max_change(tdnnf_ctx* ctx, int n, float* const* model, const int32_t* model_strides,
                                            float* const* delta, const int32_t* delta_strides, const int32_t* rows,
                                            const int32_t* cols, const int32_t* groups, int num_groups, const float* max_change,
                                            float max_param_change, float max_change_scale, float scale, float momentum,
                                            double* dots_dev, float* scale_factors_out, int32_t* num_max_change_per_component_applied,
                                            int32_t* num_max_change_global_applied, int* applied) {
  TDNNF_REQUIRE(ctx && model && model_strides && delta && delta_strides && rows && cols && groups && max_change && dots_dev && applied,
                "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MULTI_MAX, "buffer count out of range");  // per_buf below is indexed by buffer
  TDNNF_REQUIRE(num_groups >= 1 && num_groups <= TDNNF_MULTI_MAX, "group count out of range");
  for (int i = 0; i < n; ++i) TDNNF_REQUIRE(groups[i] >= 0 && groups[i] < num_groups, "group index out of range");
  for (int g = 0; g < num_groups; ++g) TDNNF_REQUIRE(max_change[g] >= 0.f, "max-change must be >= 0");  // KALDI_ASSERT :2112
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(cudaMemsetAsync(dots_dev, 0, sizeof(double) * num_groups, ctx->stream));
  int rc = tdnnf_multi_sumsq(ctx, n, delta, rows, cols, delta_strides, groups, dots_dev);
  if (rc) return rc;
  double dots[TDNNF_MULTI_MAX];
  TDNNF_CUDA_OK(cudaMemcpyAsync(dots, dots_dev, sizeof(double) * num_groups, cudaMemcpyDeviceToHost, ctx->stream));
  TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  float factors[TDNNF_MULTI_MAX];
  float param_delta_squared = 0.f;
  for (int g = 0; g < num_groups; ++g) {  // BaseFloat arithmetic as in the reference
    const float dot_prod = (float)dots[g];
    const float mc = max_change[g];
    if (mc != 0.f && std::sqrt(dot_prod) * std::fabs(scale) > mc * max_change_scale) {
      factors[g] = mc * max_change_scale / (std::sqrt(dot_prod) * std::fabs(scale));
      if (num_max_change_per_component_applied) num_max_change_per_component_applied[g]++;
    } else {
      factors[g] = 1.f;
    }
    param_delta_squared += factors[g] * factors[g] * dot_prod;
  }
  float param_delta = std::sqrt(param_delta_squared) * std::fabs(scale);
  *applied = 1;
  if (param_delta - param_delta != 0.f) {
    // "Infinite parameter change, will not apply."  The reference only gets here for +inf (nnet-utils.cc:2147-2150): when a
    // per-component clip has already turned the sum into NaN (factor 0 times an infinite dot product) its test
    // `param_delta > max` is false and it goes on to add 0 * inf = NaN into the model.  Both cases are refused here.
    *applied = 0;
  } else if (max_param_change != 0.f && param_delta > max_param_change * max_change_scale) {
    scale *= max_param_change * max_change_scale / param_delta;
    if (num_max_change_global_applied) (*num_max_change_global_applied)++;
  }
  float per_buf[TDNNF_MULTI_MAX];
  for (int g = 0; g < num_groups; ++g) {
    factors[g] = *applied ? factors[g] * scale : 0.f;
    if (scale_factors_out) scale_factors_out[g] = factors[g];
  }
  for (int i = 0; i < n; ++i) per_buf[i] = factors[groups[i]];
  MultiBufTable t;
  int blocks = 0;
  rc = multi_table(ctx, n, delta, model, rows, cols, delta_strides, model_strides, nullptr, per_buf, &t, &blocks);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) TDNNF_REQUIRE(model[i] && model_strides[i] >= cols[i], "bad model buffer");
  TDNNF_CUDA_OK(launch_pdl(multi_axpy_zero_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, t, momentum));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// ApplyL2Regularization (ref: nnet-utils.cc:2223-2245): delta_g += -2 * l2_regularize_scale * lrate_g * l2_g * model_g for
// every updatable component g with a non-zero product, in one launch.
extern "C" int tdnnf_apply_l2_regularization(tdnnf_ctx* ctx, int n, float* const* model, const int32_t* model_strides,
                                             float* const* delta, const int32_t* delta_strides, const int32_t* rows,
                                             const int32_t* cols, const int32_t* groups, int num_groups, const float* lrate,
                                             const float* l2_regularize, float l2_regularize_scale) {
  TDNNF_REQUIRE(ctx && model && model_strides && delta && delta_strides && rows && cols && groups && lrate && l2_regularize,
                "null argument");
  if (l2_regularize_scale == 0.f) return TDNNF_OK;
  float per_buf[TDNNF_MULTI_MAX];
  bool any = false;
  for (int i = 0; i < n && i < TDNNF_MULTI_MAX; ++i) {
    TDNNF_REQUIRE(groups[i] >= 0 && groups[i] < num_groups, "group index out of range");
    TDNNF_REQUIRE(lrate[groups[i]] >= 0.f && l2_regularize[groups[i]] >= 0.f, "lrate and l2-regularize must be >= 0");  // :2241
    per_buf[i] = -2.f * l2_regularize_scale * lrate[groups[i]] * l2_regularize[groups[i]];
    any = any || per_buf[i] != 0.f;
  }
  if (!any) return TDNNF_OK;
  MultiBufTable t;
  int blocks = 0;
  int rc = multi_table(ctx, n, model, delta, rows, cols, model_strides, delta_strides, nullptr, per_buf, &t, &blocks);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) TDNNF_REQUIRE(delta[i] && delta_strides[i] >= cols[i], "bad delta buffer");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(multi_axpy_zero_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, t, 1.f));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// PenalizeOutOfRange (kaldi: chain/chain-training.cc): on the rows row_offset + k * row_step,
//   deriv[r][c] -= scale * (x - limit) for x > limit,  deriv[r][c] -= scale * (x + limit) for x < -limit.
__global__ void penalize_out_of_range_kernel(const float* __restrict__ x, long long xs, int rows, int cols, float limit,
                                             float scale, int row_step, int row_offset, float* __restrict__ d, long long ds) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = (long long)row_offset + (i / cols) * row_step;
    const int c = (int)(i % cols);
    const float v = x[r * xs + c];
    if (v > limit) d[r * ds + c] -= scale * (v - limit);
    else if (v < -limit) d[r * ds + c] -= scale * (v + limit);
  }
}

extern "C" int tdnnf_penalize_out_of_range(tdnnf_ctx* ctx, const float* nnet_output, int rows, int cols, int stride, float limit,
                                           float scale, int row_step, int row_offset, float* deriv, int deriv_stride) {
  TDNNF_REQUIRE(ctx && nnet_output && deriv, "null argument");
  TDNNF_REQUIRE(limit > 0.f && scale >= 0.f && row_step >= 1 && row_offset >= 0 && row_offset < row_step, "bad argument");
  TDNNF_REQUIRE(stride >= cols && deriv_stride >= cols, "stride < cols");
  if (scale == 0.f || rows <= row_offset || cols == 0) return TDNNF_OK;
  const int sub_rows = (rows - row_offset + row_step - 1) / row_step;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)sub_rows * cols;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 8);
  TDNNF_CUDA_OK(launch_pdl(penalize_out_of_range_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, nnet_output, stride, sub_rows, cols, limit, scale, row_step,
                                                               row_offset, deriv, deriv_stride));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_relu_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                              int out_stride) {
  PROLOGUE(in && out && in_stride >= cols && out_stride >= cols, "bad matrix");
  const bool vec = cols % 4 == 0 && in_stride % 4 == 0 && out_stride % 4 == 0 && al16(in) && al16(out);
  TDNNF_CUDA_OK(launch_pdl(relu_fwd_kernel, dim3(grid_for((long long)rows * cols / (vec ? 4 : 1), 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in, rows, cols, in_stride, out, out_stride, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_relu_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv,
                              int od_stride, float* in_deriv, int id_stride, int rows, int cols) {
  PROLOGUE(out_value && out_deriv && in_deriv && ov_stride >= cols && od_stride >= cols && id_stride >= cols, "bad matrix");
  const bool vec = cols % 4 == 0 && ov_stride % 4 == 0 && od_stride % 4 == 0 && id_stride % 4 == 0 && al16(out_value) &&
                   al16(out_deriv) && al16(in_deriv);
  TDNNF_CUDA_OK(launch_pdl(relu_bwd_kernel, dim3(grid_for((long long)rows * cols / (vec ? 4 : 1), 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      out_value, ov_stride, out_deriv, od_stride, in_deriv, id_stride, rows, cols, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_add_scaled(tdnnf_ctx* ctx, const float* a, int a_stride, float alpha, const float* b, int b_stride,
                                float beta, float* out, int out_stride, int rows, int cols) {
  PROLOGUE(a && b && out && a_stride >= cols && b_stride >= cols && out_stride >= cols, "bad matrix");
  const bool vec = cols % 4 == 0 && a_stride % 4 == 0 && b_stride % 4 == 0 && out_stride % 4 == 0 && al16(a) && al16(b) &&
                   al16(out);
  TDNNF_CUDA_OK(launch_pdl(add_scaled_kernel, dim3(grid_for((long long)rows * cols / (vec ? 4 : 1), 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      a, a_stride, alpha, b, b_stride, beta, out, out_stride, rows, cols, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_batchnorm_train_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                                         int out_stride, float epsilon, float target_rms, float* memo) {
  PROLOGUE(in && out && memo && in_stride >= cols && out_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(cudaMemsetAsync(memo + 3 * (size_t)cols, 0, sizeof(float) * 2 * cols, ctx->stream));
  int gy = (rows + 255) / 256;
  if (gy > 128) gy = 128;
  TDNNF_CUDA_OK(launch_pdl(col_stats_kernel<0>, dim3((cols + 31) / 32, gy), dim3(32, 8), 0, ctx->stream, 1, in, in_stride, nullptr, 0, rows, cols,
                                                                                memo + 3 * (size_t)cols));
  LAUNCH_CHECK(ctx);
  TDNNF_CUDA_OK(launch_pdl(bn_finalize_fwd_kernel, dim3((cols + 255) / 256), dim3(256), 0, ctx->stream, 1, memo, cols, rows, epsilon, target_rms));
  LAUNCH_CHECK(ctx);
  TDNNF_CUDA_OK(launch_pdl(bn_apply_fwd_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, in, in_stride, out,
                                                                                                    out_stride, rows, cols, memo));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_batchnorm_train_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv,
                                         int od_stride, float* in_deriv, int id_stride, int rows, int cols,
                                         float target_rms, const float* memo) {
  PROLOGUE(out_value && out_deriv && in_deriv && memo && ov_stride >= cols && od_stride >= cols && id_stride >= cols,
           "bad matrix");
  ctx->ws_reset();
  int rc = ctx->ws_reserve(sizeof(float) * 2 * cols);
  if (rc) return rc;
  float* sums = static_cast<float*>(ctx->ws_alloc(sizeof(float) * 2 * cols));
  if (!sums) return TDNNF_ERR_NOMEM;
  TDNNF_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * cols, ctx->stream));
  int gy = (rows + 255) / 256;
  if (gy > 128) gy = 128;
  TDNNF_CUDA_OK(launch_pdl(col_stats_kernel<1>, dim3((cols + 31) / 32, gy), dim3(32, 8), 0, ctx->stream, 1, out_deriv, od_stride, out_value, ov_stride,
                                                                                rows, cols, sums));
  LAUNCH_CHECK(ctx);
  TDNNF_CUDA_OK(launch_pdl(bn_apply_bwd_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      out_value, ov_stride, out_deriv, od_stride, in_deriv, id_stride, rows, cols, target_rms, memo, sums));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// NonlinearComponent statistics (ref: nnet-component-itf.cc:433-481), accumulated in device doubles:
//   MODE 0: s0[c] += sum_r x[r][c],  s1[c] += sum_r (x[r][c] > 0)      (value sum, ReLU derivative sum; s1 may be null)
//   MODE 1: s0[c] += sum_r x[r][c]^2                                    (out_deriv sum of squares)
template <int MODE>
__global__ void nonlin_stats_kernel(const float* __restrict__ x, long long xs, int rows, int cols, double* __restrict__ s0,
                                    double* __restrict__ s1) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red0[8][33], red1[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  if (col < cols) {
    for (long long r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      const float v = x[r * xs + col];
      if (MODE == 0) { a0 += v; a1 += v > 0.f ? 1.f : 0.f; }
      else { a0 += v * v; }
    }
  }
  red0[threadIdx.y][threadIdx.x] = a0;
  red1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += red0[i][threadIdx.x]; t1 += red1[i][threadIdx.x]; }
    atomicAdd(s0 + col, (double)t0);
    if (MODE == 0 && s1 != nullptr) atomicAdd(s1 + col, (double)t1);
  }
}

// RectifiedLinearComponent::RepairGradients (ref: nnet-simple-component.cc:990-1074):
//   in_deriv[r][c] += -scale * ((stat[c] > lower) + (stat[c] > upper) - 1),  stat = deriv_sum averaged over the blocks;
//   *num_repaired += number of columns whose term is non-zero (block 0 of the grid only counts).
__global__ void relu_repair_kernel(float* __restrict__ in_deriv, long long ld, int rows, int block_dim, int num_blocks_of_dim,
                                   const double* __restrict__ deriv_sum, float lower, float upper, float scale,
                                   double* __restrict__ num_repaired) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)rows * block_dim;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % block_dim);
    const long long r = idx / block_dim;
    double st = 0.0;
    for (int b = 0; b < num_blocks_of_dim; ++b) st += deriv_sum[(long long)b * block_dim + c];
    const float stat = (float)(st / num_blocks_of_dim);
    const float t = (stat - lower > 0.f ? 1.f : 0.f) + (stat - upper > 0.f ? 1.f : 0.f) - 1.f;
    if (t != 0.f) in_deriv[r * ld + c] += -scale * t;
    if (r == 0 && t != 0.f) atomicAdd(num_repaired, 1.0);
  }
}

// stats[i] += num_frames * mean[i], stats[cols + i] += num_frames * uvar[i]   (BatchNormComponent::StoreStats, norm.cc:583-588)
__global__ void bn_accumulate_stats_kernel(const float* __restrict__ memo, int cols, float num_frames, double* __restrict__ stats) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cols) {
    stats[i] += (double)num_frames * (double)memo[i];
    stats[cols + i] += (double)num_frames * (double)memo[cols + i];
  }
}

extern "C" int tdnnf_nonlinear_store_stats(tdnnf_ctx* ctx, const float* out_value, int rows, int cols, int stride,
                                           double* value_sum, double* deriv_sum) {
  PROLOGUE(out_value && value_sum && stride >= cols, "bad argument");
  int gy = std::min(128, (rows + 255) / 256);
  TDNNF_CUDA_OK(launch_pdl(nonlin_stats_kernel<0>, dim3((cols + 31) / 32, std::max(gy, 1)), dim3(32, 8), 0, ctx->stream, 1, out_value, stride, rows, cols,
                                                                                                  value_sum, deriv_sum));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_nonlinear_store_backprop_stats(tdnnf_ctx* ctx, const float* out_deriv, int rows, int cols, int stride,
                                                    double* oderiv_sumsq) {
  PROLOGUE(out_deriv && oderiv_sumsq && stride >= cols, "bad argument");
  int gy = std::min(128, (rows + 255) / 256);
  TDNNF_CUDA_OK(launch_pdl(nonlin_stats_kernel<1>, dim3((cols + 31) / 32, std::max(gy, 1)), dim3(32, 8), 0, ctx->stream, 1, out_deriv, stride, rows, cols,
                                                                                                  oderiv_sumsq, nullptr));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_relu_repair_gradients(tdnnf_ctx* ctx, float* in_deriv, int rows, int block_dim, int stride, int blocks_per_row,
                                           const double* deriv_sum, float lower_threshold, float upper_threshold, float scale,
                                           double* num_dims_repaired) {
  const int cols = block_dim;
  PROLOGUE(in_deriv && deriv_sum && num_dims_repaired && stride >= block_dim && blocks_per_row >= 1, "bad argument");
  TDNNF_CUDA_OK(launch_pdl(relu_repair_kernel, dim3(grid_for((long long)rows * block_dim, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in_deriv, stride, rows, block_dim, blocks_per_row, deriv_sum, lower_threshold, upper_threshold, scale, num_dims_repaired));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_batchnorm_accumulate_stats(tdnnf_ctx* ctx, const float* memo, int cols, float num_frames, double* stats) {
  const int rows = 1;
  PROLOGUE(memo && stats && cols > 0, "bad argument");
  TDNNF_CUDA_OK(launch_pdl(bn_accumulate_stats_kernel, dim3((cols + 255) / 256), dim3(256), 0, ctx->stream, 1, memo, cols, num_frames, stats));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_relu_scale_offset_bypass_fwd(tdnnf_ctx* ctx, const float* x, int rows, int cols, int x_stride,
                                                  const float* scale, const float* offset, const float* prev,
                                                  int prev_stride, float bypass_scale, float* out, int out_stride) {
  PROLOGUE(x && scale && offset && prev && out && x_stride >= cols && prev_stride >= cols && out_stride >= cols, "bad matrix");
  TDNNF_REQUIRE(cols % 4 == 0 && x_stride % 4 == 0 && prev_stride % 4 == 0 && out_stride % 4 == 0 && al16(x) && al16(prev) &&
                    al16(out) && al16(scale) && al16(offset),
                "fused tail needs 16-byte aligned rows (cols and strides multiples of 4)");
  TDNNF_CUDA_OK(launch_pdl(tail_fwd_kernel, dim3(grid_for((long long)rows * cols / 4, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      x, x_stride, scale, offset, prev, prev_stride, bypass_scale, out, out_stride, rows, cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_relu_scale_offset_bypass_bwd(tdnnf_ctx* ctx, const float* d_out, int do_stride, const float* x,
                                                  int x_stride, const float* scale, float bypass_scale, float* d_x,
                                                  int dx_stride, float* d_prev, int dp_stride, int rows, int cols) {
  PROLOGUE(d_out && x && scale && d_x && d_prev && do_stride >= cols && x_stride >= cols && dx_stride >= cols && dp_stride >= cols,
           "bad matrix");
  TDNNF_REQUIRE(cols % 4 == 0 && do_stride % 4 == 0 && x_stride % 4 == 0 && dx_stride % 4 == 0 && dp_stride % 4 == 0 &&
                    al16(d_out) && al16(x) && al16(d_x) && al16(d_prev) && al16(scale),
                "fused tail needs 16-byte aligned rows (cols and strides multiples of 4)");
  TDNNF_CUDA_OK(launch_pdl(tail_bwd_kernel, dim3(grid_for((long long)rows * cols / 4, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      d_out, do_stride, x, x_stride, scale, bypass_scale, d_x, dx_stride, d_prev, dp_stride, rows, cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// The fused tails that also hand the next GEMM its operand planes (see tail_fwd_planes_kernel).  *planes: a new handle for
// `out` / `d_x` (tdnnf_planes_release when done); attach it to the context before the component call that consumes the matrix.
extern "C" int tdnnf_relu_scale_offset_bypass_fwd_planes(tdnnf_ctx* ctx, const float* x, int rows, int cols, int x_stride,
                                                         const float* scale, const float* offset, const float* prev, int prev_stride,
                                                         float bypass_scale, float* out, int out_stride, tdnnf_planes** planes) {
  PROLOGUE(x && scale && offset && prev && out && planes && x_stride >= cols && prev_stride >= cols && out_stride >= cols, "bad matrix");
  TDNNF_REQUIRE(cols % 4 == 0 && x_stride % 4 == 0 && prev_stride % 4 == 0 && out_stride % 4 == 0 && al16(x) && al16(prev) &&
                    al16(out) && al16(scale) && al16(offset),
                "fused tail needs 16-byte aligned rows (cols and strides multiples of 4)");
  tdnnf_planes* pl = nullptr;
  int rc = tdnnf::planes_alloc_for_producer(ctx, out, rows, cols, out_stride, &pl);
  if (rc) return rc;
  TDNNF_REQUIRE(pl->Kpad / 4 <= 1024, "fused tail with planes: at most 4096 columns");
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(pl->base);
  TDNNF_CUDA_OK(launch_pdl(tail_fwd_planes_kernel, dim3(std::min(rows, ctx->num_sms * 4)), dim3(pl->Kpad / 4), 0, ctx->stream, 1, 
      x, x_stride, scale, offset, prev, prev_stride, bypass_scale, out, out_stride, rows, cols, pl->Kpad, hi, hi + pl->plane_elems, pl->rowsq));
  LAUNCH_CHECK(ctx);
  *planes = pl;
  return TDNNF_OK;
}

extern "C" int tdnnf_relu_scale_offset_bypass_bwd_planes(tdnnf_ctx* ctx, const float* d_out, int do_stride, const float* x,
                                                         int x_stride, const float* scale, float bypass_scale, float* d_x,
                                                         int dx_stride, float* d_prev, int dp_stride, int rows, int cols,
                                                         tdnnf_planes** planes) {
  PROLOGUE(d_out && x && scale && d_x && d_prev && planes && do_stride >= cols && x_stride >= cols && dx_stride >= cols && dp_stride >= cols,
           "bad matrix");
  TDNNF_REQUIRE(cols % 4 == 0 && do_stride % 4 == 0 && x_stride % 4 == 0 && dx_stride % 4 == 0 && dp_stride % 4 == 0 &&
                    al16(d_out) && al16(x) && al16(d_x) && al16(d_prev) && al16(scale),
                "fused tail needs 16-byte aligned rows (cols and strides multiples of 4)");
  tdnnf_planes* pl = nullptr;
  int rc = tdnnf::planes_alloc_for_producer(ctx, d_x, rows, cols, dx_stride, &pl);
  if (rc) return rc;
  TDNNF_REQUIRE(pl->Kpad / 4 <= 1024, "fused tail with planes: at most 4096 columns");
  { int zrc = zero_async(ctx, pl->colsum, sizeof(float) * (size_t)cols); if (zrc) return zrc; }
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(pl->base);
  TDNNF_CUDA_OK(launch_pdl(tail_bwd_planes_kernel, dim3(std::min(rows, ctx->num_sms * 4)), dim3(pl->Kpad / 4), 0, ctx->stream, 1, 
      d_out, do_stride, x, x_stride, scale, bypass_scale, d_x, dx_stride, d_prev, dp_stride, rows, cols, pl->Kpad, hi,
      hi + pl->plane_elems, pl->rowsq, pl->colsum));
  LAUNCH_CHECK(ctx);
  pl->has_colsum = true;
  *planes = pl;
  return TDNNF_OK;
}

// ------------------------------------------------------------------ GeneralDropoutComponent (kaldi: nnet-general-component.cc)
namespace {
// The component layer's counter hash (csrc/nnet3/shim.cc: Mix / RandUniformOpen), evaluated on the device: element e of
// a mask is draw number counter0 + e, so a test can replay the mask with the same host calls.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// GetMemo: u ~ U(0,1); !continuous: mask = (u > p) / (1 - p); continuous: mask = 1 - 2p + 4p u  (expected value 1)
__global__ void dropout_mask_kernel(unsigned long long seed_mixed, unsigned long long counter0, float* __restrict__ mask,
                                    int rows, int cols, long long stride, float p, int continuous) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const unsigned long long r = mix64(seed_mixed ^ ((counter0 + (unsigned long long)idx) * 0xD1342543DE82EF95ull));
    const float u = ((float)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    float m;
    if (continuous) m = __fadd_rn(__fmul_rn(u, p * 4.0f), 1.0f - 2.0f * p);  // Scale(4p) then Add(1 - 2p): two roundings
    else m = (u - p > 0.0f ? 1.0f : 0.0f) * (1.0f / (1.0f - p));
    mask[(idx / cols) * stride + idx % cols] = m;
  }
}
// CuMatrixBase::MulRows(mask, indexes): out[r,:] = in[r,:] .* mask[indexes[r],:]   (in may alias out)
__global__ void mul_rows_indexed_kernel(const float* in, long long is, float* out, long long os, int rows, int cols4,
                                        const float* __restrict__ mask, long long ms, const int32_t* __restrict__ index) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols4) {
    const long long r = idx / cols4;
    const int c = (int)(idx % cols4) * 4;
    const int m = index[r];
    const float4 v = *reinterpret_cast<const float4*>(in + r * is + c);
    const float4 w = *reinterpret_cast<const float4*>(mask + (long long)m * ms + c);
    *reinterpret_cast<float4*>(out + r * os + c) = make_float4(v.x * w.x, v.y * w.y, v.z * w.z, v.w * w.w);
  }
}
__global__ void mul_rows_indexed_scalar_kernel(const float* in, long long is, float* out, long long os, int rows, int cols,
                                               const float* __restrict__ mask, long long ms, const int32_t* __restrict__ index) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  ELEMWISE_LOOP((long long)rows * cols) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    out[r * os + c] = in[r * is + c] * mask[(long long)index[r] * ms + c];
  }
}
}  // namespace

extern "C" int tdnnf_dropout_mask(tdnnf_ctx* ctx, unsigned long long seed, unsigned long long counter, float* mask, int rows,
                                  int cols, int stride, float proportion, int continuous) {
  PROLOGUE(mask && rows >= 0 && cols >= 0 && stride >= cols && proportion >= 0.0f && proportion <= 1.0f, "bad argument");
  // host side of mix64(seed): the same value for every element
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  TDNNF_CUDA_OK(launch_pdl(dropout_mask_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, z, counter, mask, rows, cols,
                                                                                                   stride, proportion, continuous));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_mul_rows_indexed(tdnnf_ctx* ctx, const float* in, int in_stride, float* out, int out_stride, int rows,
                                      int cols, const float* mask, int mask_stride, const int32_t* row_index_dev) {
  PROLOGUE(in && out && mask && row_index_dev && in_stride >= cols && out_stride >= cols && mask_stride >= cols, "bad matrix");
  if (cols % 4 == 0 && in_stride % 4 == 0 && out_stride % 4 == 0 && mask_stride % 4 == 0 && al16(in) && al16(out) && al16(mask))
    TDNNF_CUDA_OK(launch_pdl(mul_rows_indexed_kernel, dim3(grid_for((long long)rows * cols / 4, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
        in, in_stride, out, out_stride, rows, cols / 4, mask, mask_stride, row_index_dev));
  else
    TDNNF_CUDA_OK(launch_pdl(mul_rows_indexed_scalar_kernel, dim3(grid_for((long long)rows * cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
        in, in_stride, out, out_stride, rows, cols, mask, mask_stride, row_index_dev));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

// ------------------------------------------------------------------ LogSoftmaxComponent (ref: nnet-simple-component.cc:3607-3632)
namespace {
__device__ __forceinline__ float block_reduce_256(float v, bool is_max, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, other) : v + other;
  }
  __syncthreads();  // sh may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, sh[w]) : r + sh[w];
  return r;
}
// ApplyLogSoftMaxPerRow: out = x - max - log(sum exp(x - max)); one CTA of 256 threads per row (grid-stride over rows)
__global__ void __launch_bounds__(256) log_softmax_fwd_kernel(const float* __restrict__ in, long long is, float* __restrict__ out,
                                                              long long os, int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float sh[8];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* x = in + r * is;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, x[c]);
    m = block_reduce_256(m, true, sh);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) s += __expf(x[c] - m);
    s = block_reduce_256(s, false, sh);
    const float shift = m + __logf(s);
    for (int c = threadIdx.x; c < cols; c += 256) out[r * os + c] = x[c] - shift;
  }
}
// DiffLogSoftmaxPerRow: in_deriv = out_deriv - exp(out_value) * sum_row(out_deriv)   (in_deriv may alias out_deriv)
__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ ov, long long ovs, const float* od,
                                                              long long ods, float* id, long long ids, int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float sh[8];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) s += od[r * ods + c];
    s = block_reduce_256(s, false, sh);
    for (int c = threadIdx.x; c < cols; c += 256) id[r * ids + c] = od[r * ods + c] - __expf(ov[r * ovs + c]) * s;
  }
}
}  // namespace

extern "C" int tdnnf_log_softmax_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                                     int out_stride) {
  PROLOGUE(in && out && in_stride >= cols && out_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(log_softmax_fwd_kernel, dim3(std::min(rows, ctx->num_sms * 8)), dim3(256), 0, ctx->stream, 1, in, in_stride, out, out_stride, rows, cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
extern "C" int tdnnf_log_softmax_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, const float* out_deriv,
                                     int od_stride, float* in_deriv, int id_stride, int rows, int cols) {
  PROLOGUE(out_value && out_deriv && in_deriv && ov_stride >= cols && od_stride >= cols && id_stride >= cols, "bad matrix");
  TDNNF_CUDA_OK(launch_pdl(log_softmax_bwd_kernel, dim3(std::min(rows, ctx->num_sms * 8)), dim3(256), 0, ctx->stream, 1, out_value, ov_stride, out_deriv, od_stride,
                                                                                   in_deriv, id_stride, rows, cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
