// Bandwidth-bound kernels of the NAS components: mixing coefficients and their update,
// {Gumbel}SoftmaxFlops forward/backward, CopyN, Onehot, BatchNormTest scale/offset,
// ElementwiseProduct, AddRowSumMat.  One fused pass per reference method; fp32 throughout.
#include <cfloat>

#include "context.h"
#include "ptx.cuh"

using namespace tdnnf;

namespace {

constexpr float kFloor = 1.0e-20f;  // ref: tdnn.cc:268,277  simple.cc:9978,10110

struct Uniforms {
  float u[TDNNF_MAX_OFFSETS];
};

inline int grid_for(long long total, int threads, int num_sms) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ DARTS coefficients
// ref: tdnn.cc:250-289 (coef) and the weight selection inside the GEMM loops tdnn.cc:292-328.
__device__ void weff_from_coef(const float* coef, int n, int flags, int share, float* weff) {
  for (int i = 0; i < n; ++i) {
    float w;
    if (flags & TDNNF_DARTS_UNIFORM_SAMPLE) w = (i == share || coef[i] == 1.0f) ? 1.0f : 0.0f;
    else if (flags & TDNNF_DARTS_FREE_SELECT) w = coef[i];
    else w = (i == share) ? 1.0f : coef[i];
    weff[i] = w;
  }
}

__global__ void darts_coef_kernel(const float* __restrict__ alpha, int n, int flags, float temperature, Uniforms ug,
                                  float u_uniform, int share, float* __restrict__ coef, float* __restrict__ weff) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float c[TDNNF_MAX_OFFSETS];
  for (int i = 0; i < n; ++i) c[i] = alpha[i];
  if (flags & TDNNF_DARTS_USE_GUMBEL) {
    const float inv_t = 1.0f / temperature;
    float mx = -FLT_MAX;
    for (int i = 0; i < n; ++i) {
      const float g = -logf(-logf(ug.u[i]));  // G = -log(-log U)
      c[i] = (c[i] + g) * inv_t;
      mx = fmaxf(mx, c[i]);
    }
    float sum = 0.f;
    for (int i = 0; i < n; ++i) { c[i] = expf(c[i] - mx); sum += c[i]; }
    for (int i = 0; i < n; ++i) c[i] = fmaxf(c[i] / sum, kFloor);
  } else if (flags & TDNNF_DARTS_FREE_SELECT) {
    for (int i = 0; i < n; ++i) c[i] = 1.0f / (expf(-c[i]) + 1.0f);
  } else {
    float mx = -FLT_MAX;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, c[i]);
    float sum = 0.f;
    for (int i = 0; i < n; ++i) { c[i] = expf(c[i] - mx); sum += c[i]; }
    for (int i = 0; i < n; ++i) c[i] = fmaxf(c[i] / sum, kFloor);
  }
  if (flags & TDNNF_DARTS_UNIFORM_SAMPLE) {
    for (int i = 0; i < n; ++i)
      c[i] = (u_uniform >= (float)i / n && u_uniform < (float)(i + 1) / n) ? 1.0f : 0.0f;
  }
  for (int i = 0; i < n; ++i) coef[i] = c[i];
  weff_from_coef(c, n, flags, share, weff);
}

__global__ void darts_weff_kernel(const float* __restrict__ coef, int n, int flags, int share,
                                  float* __restrict__ weff) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float c[TDNNF_MAX_OFFSETS];
  for (int i = 0; i < n; ++i) c[i] = coef[i];
  weff_from_coef(c, n, flags, share, weff);
}

// ref: tdnn.cc:541-590.  Sequential on purpose: n <= 16 and the accumulation order of the
// reference (offset by offset into the delta) is kept.
__global__ void darts_alpha_update_kernel(const float* __restrict__ s, const float* __restrict__ coef, int n, int flags,
                                          float temperature, int share, float lr, float* __restrict__ dalpha) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float d[TDNNF_MAX_OFFSETS], c[TDNNF_MAX_OFFSETS];
  for (int i = 0; i < n; ++i) { d[i] = dalpha[i]; c[i] = coef[i]; }
  if (!(flags & TDNNF_DARTS_UNIFORM_SAMPLE)) {
    for (int i = 0; i < n; ++i) {
      const float si = s[i];
      if (flags & TDNNF_DARTS_FREE_SELECT) {
        d[i] += si * c[i];
        d[i] += (-1.0f * si) * c[i] * c[i];
      } else if (i != share) {
        const float g = (flags & TDNNF_DARTS_USE_GUMBEL) ? si / temperature : si;
        for (int j = 0; j < n; ++j) d[j] += (-1.0f * g) * c[i] * c[j];
        d[i] += g * c[i];
      }
    }
  }
  float mul = 1.0f;
  if (flags & TDNNF_DARTS_USE_ENTROPY) mul *= 5.0f;
  if (flags & TDNNF_DARTS_FREE_SELECT) mul *= 5.0f * lr;
  else if (flags & TDNNF_DARTS_USE_GUMBEL) mul *= lr;
  else mul *= 5.0f * lr;
  if (flags & TDNNF_DARTS_UPDATE_ALPHA) mul *= 10000.0f;
  for (int i = 0; i < n; ++i) dalpha[i] = d[i] * mul;
}

// ------------------------------------------------------------------ {Gumbel}SoftmaxFlops
// One thread per row for narrow matrices (the recipes use 8 columns): the row lives in
// registers, loads/stores are 2 x float4 per thread, a warp touches 1 KB contiguous.
template <int COLS>
__global__ void softmax_flops_fwd_small(const float* __restrict__ in, int rows, long long in_stride,
                                        float* __restrict__ out, long long out_stride, Uniforms g, float inv_temp,
                                        bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    float x[COLS];
    const float* src = in + r * in_stride;
    if (vec) {
#pragma unroll
      for (int j = 0; j < COLS; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < COLS; ++j) x[j] = src[j];
    }
    float mx = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < COLS; ++j) { x[j] = (x[j] + g.u[j]) * inv_temp; mx = fmaxf(mx, x[j]); }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < COLS; ++j) { x[j] = expf(x[j] - mx); sum += x[j]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < COLS; ++j) x[j] = fmaxf(x[j] * inv, kFloor);
    float* dst = out + r * out_stride;
    if (vec) {
#pragma unroll
      for (int j = 0; j < COLS; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < COLS; ++j) dst[j] = x[j];
    }
  }
}

// General width: one warp per row, shuffle reductions.  `noise` (device, cols) may be null.
__global__ void softmax_flops_fwd_warp(const float* __restrict__ in, int rows, int cols, long long in_stride,
                                       float* __restrict__ out, long long out_stride, const float* __restrict__ noise,
                                       float inv_temp) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* src = in + r * in_stride;
    float mx = -FLT_MAX;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, (src[j] + (noise ? noise[j] : 0.f)) * inv_temp);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < cols; j += 32) sum += expf((src[j] + (noise ? noise[j] : 0.f)) * inv_temp - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float* dst = out + r * out_stride;
    for (int j = lane; j < cols; j += 32)
      dst[j] = fmaxf(expf((src[j] + (noise ? noise[j] : 0.f)) * inv_temp - mx) * inv, kFloor);
  }
}

__constant__ float kFlops[8] = {-25.f, -50.f, -80.f, -100.f, -120.f, -160.f, -200.f, -240.f};  // ref: simple.cc:10145-10152

template <int COLS>
__global__ void softmax_flops_bwd_small(const float* __restrict__ out_value, long long ov_stride, float* out_deriv,
                                        long long od_stride, float* in_deriv, long long id_stride, int rows,
                                        float penalty, float inv_temp, int write_back_e, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    float p[COLS], e[COLS];
    const float* pv = out_value + r * ov_stride;
    float* od = out_deriv + r * od_stride;
    if (vec) {
#pragma unroll
      for (int j = 0; j < COLS; j += 4) {
        const float4 a = *reinterpret_cast<const float4*>(pv + j);
        const float4 b = *reinterpret_cast<const float4*>(od + j);
        p[j] = a.x; p[j + 1] = a.y; p[j + 2] = a.z; p[j + 3] = a.w;
        e[j] = b.x; e[j + 1] = b.y; e[j + 2] = b.z; e[j + 3] = b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < COLS; ++j) { p[j] = pv[j]; e[j] = od[j]; }
    }
    float pe = 0.f;
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
      if (j < 8) e[j] += penalty * kFlops[j];
      pe += p[j] * e[j];
    }
    float d[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) d[j] = (p[j] * e[j] - p[j] * pe) * inv_temp;
    float* id = in_deriv + r * id_stride;
    if (write_back_e && id != od) {
      if (vec) {
#pragma unroll
        for (int j = 0; j < COLS; j += 4) *reinterpret_cast<float4*>(od + j) = make_float4(e[j], e[j + 1], e[j + 2], e[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < COLS; ++j) od[j] = e[j];
      }
    }
    if (vec) {
#pragma unroll
      for (int j = 0; j < COLS; j += 4) *reinterpret_cast<float4*>(id + j) = make_float4(d[j], d[j + 1], d[j + 2], d[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < COLS; ++j) id[j] = d[j];
    }
  }
}

__global__ void softmax_flops_bwd_warp(const float* __restrict__ out_value, long long ov_stride, float* out_deriv,
                                       long long od_stride, float* in_deriv, long long id_stride, int rows, int cols,
                                       float penalty, float inv_temp, int write_back_e) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* pv = out_value + r * ov_stride;
    float* od = out_deriv + r * od_stride;
    float* id = in_deriv + r * id_stride;
    float pe = 0.f;
    for (int j = lane; j < cols; j += 32) {
      const float e = od[j] + (j < 8 ? penalty * kFlops[j] : 0.f);
      pe += pv[j] * e;
    }
    for (int o = 16; o > 0; o >>= 1) pe += __shfl_xor_sync(0xffffffffu, pe, o);
    for (int j = lane; j < cols; j += 32) {
      const float e = od[j] + (j < 8 ? penalty * kFlops[j] : 0.f);
      const float p = pv[j];
      const float d = (p * e - p * pe) * inv_temp;
      if (write_back_e && id != od) od[j] = e;
      id[j] = d;
    }
  }
}

// ------------------------------------------------------------------ CopyN (AddMatBlocks)
__global__ void copyn_fwd_kernel(const float* __restrict__ in, int rows, int in_cols, long long in_stride,
                                 float* __restrict__ out, int out_cols, long long out_stride, float scale) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * out_cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % out_cols);
    const long long r = idx / out_cols;
    out[r * out_stride + c] += scale * in[r * in_stride + (c % in_cols)];
  }
}

__global__ void copyn_bwd_kernel(const float* __restrict__ od, int rows, int out_cols, long long od_stride,
                                 float* __restrict__ id, int in_cols, long long id_stride, float scale) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * in_cols;
  const int nblocks = out_cols / in_cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % in_cols);
    const long long r = idx / in_cols;
    float sum = 0.f;
    for (int b = 0; b < nblocks; ++b) sum += od[r * od_stride + b * in_cols + j];
    id[r * id_stride + j] += scale * sum;
  }
}

// ------------------------------------------------------------------ Onehot
__global__ void onehot_fwd_kernel(float* __restrict__ out, int rows, int dim, long long stride, float u) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * dim;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % dim);
    const long long r = idx / dim;
    out[r * stride + i] = (u >= (float)i / dim && u < (float)(i + 1) / dim) ? 1.0f : 0.0f;
  }
}

// ------------------------------------------------------------------ AddRowSumMat
__global__ void add_row_sum_kernel(const float* __restrict__ mat, int rows, int cols, long long stride, float scale,
                                   float* __restrict__ vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float sum = 0.f;
  if (col < cols) {
    for (long long r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) sum += mat[r * stride + col];
  }
  red[threadIdx.y][threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(vec + col, scale * t);
  }
}

// ------------------------------------------------------------------ BatchNormTest: y = x .* scale + offset
__global__ void scale_offset_rows_kernel(const float* __restrict__ in, int rows, int cols, long long in_stride,
                                         float* __restrict__ out, long long out_stride, const float* __restrict__ scale,
                                         const float* __restrict__ offset, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  if (vec) {
    const int c4 = cols >> 2;
    const long long total = (long long)rows * c4;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(idx % c4) * 4;
      const long long r = idx / c4;
      const float4 x = *reinterpret_cast<const float4*>(in + r * in_stride + c);
      const float4 s = *reinterpret_cast<const float4*>(scale + c);
      float4 y = make_float4(x.x * s.x, x.y * s.y, x.z * s.z, x.w * s.w);
      if (offset) {
        const float4 o = *reinterpret_cast<const float4*>(offset + c);
        y.x += o.x; y.y += o.y; y.z += o.z; y.w += o.w;
      }
      *reinterpret_cast<float4*>(out + r * out_stride + c) = y;
    }
  } else {
    const long long total = (long long)rows * cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(idx % cols);
      const long long r = idx / cols;
      float y = in[r * in_stride + c] * scale[c];
      if (offset) y += offset[c];
      out[r * out_stride + c] = y;
    }
  }
}

// ------------------------------------------------------------------ ElementwiseProduct
__global__ void ewprod_fwd_kernel(const float* __restrict__ in, int rows, int D, long long in_stride,
                                  float* __restrict__ out, long long out_stride) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long r = idx / D;
    out[r * out_stride + c] = in[r * in_stride + c] * in[r * in_stride + D + c];
  }
}

__global__ void ewprod_bwd_kernel(const float* __restrict__ in, long long in_stride, const float* __restrict__ od,
                                  long long od_stride, float* __restrict__ id, long long id_stride, int rows, int D) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long r = idx / D;
    const float g = od[r * od_stride + c];
    const float a = in[r * in_stride + c], b = in[r * in_stride + D + c];
    id[r * id_stride + c] = g * b;
    id[r * id_stride + D + c] = g * a;
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// ------------------------------------------------------------------ shared-candidate bottleneck mask (configs[3])
// The descriptor sub-graph generate_bottleneckCB8share_onehottrain_config.py:11-90 builds around every tdnnf `linear`:
//   m_j = Sum(p_j .. p_{nb-1})  ->  CopyNComponent(1 -> b_j, scale)  ->  Append(copyn_j, linear_j)  ->
//   ElementwiseProductComponent  ->  Append of the nb blocks
// i.e. out[r, c] = lin[r, c] * scale * sum_{k >= j(c)} p[r, k]: candidate k keeps the first b_0 + .. + b_k bottleneck columns.
// One warp per row (the row's nb <= 8 probabilities and its suffix sums live in registers), float4 columns.
struct MaskBlocks {
  int nb;
  int end[8];  // exclusive end column of block j
};
__device__ __forceinline__ int mask_block_of(const MaskBlocks& b, int c) {
  int j = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) j += (k < b.nb - 1 && c >= b.end[k]) ? 1 : 0;
  return j;
}
__global__ void __launch_bounds__(256) shared_mask_fwd_kernel(const float* __restrict__ p, long long ps, const float* __restrict__ lin,
                                                              long long ls, float* __restrict__ out, long long os, int rows, int cols,
                                                              MaskBlocks blk, float scale, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    float m[8];
    float run = 0.f;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      if (k < blk.nb) run += p[r * ps + k];
      m[k] = run * scale;
    }
    const float* src = lin + r * ls;
    float* dst = out + r * os;
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(src + c);
        float4 o;
        o.x = v.x * m[mask_block_of(blk, c)];
        o.y = v.y * m[mask_block_of(blk, c + 1)];
        o.z = v.z * m[mask_block_of(blk, c + 2)];
        o.w = v.w * m[mask_block_of(blk, c + 3)];
        *reinterpret_cast<float4*>(dst + c) = o;
      }
    } else {
      for (int c = lane; c < cols; c += 32) dst[c] = src[c] * m[mask_block_of(blk, c)];
    }
  }
}
// d_lin[r, c] = d_out[r, c] * m_j(c);  d_p[r, k] = scale * sum_{j <= k} sum_{c in block j} d_out[r, c] lin[r, c]
// (the transposes of ElementwiseProduct, CopyN and the Sum descriptor); d_p is overwritten.
__global__ void __launch_bounds__(256) shared_mask_bwd_kernel(const float* __restrict__ p, long long ps, const float* __restrict__ lin,
                                                              long long ls, const float* __restrict__ d_out, long long dos,
                                                              float* __restrict__ d_lin, long long dls, float* __restrict__ d_p,
                                                              long long dps, int rows, int cols, MaskBlocks blk, float scale, bool vec) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    float m[8], dm[8];
    float run = 0.f;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      if (k < blk.nb) run += p[r * ps + k];
      m[k] = run * scale;
      dm[k] = 0.f;
    }
    const float* src = lin + r * ls;
    const float* dsrc = d_out + r * dos;
    float* dst = d_lin ? d_lin + r * dls : nullptr;
    auto one = [&](int c, float x, float d) -> float {
      const int j = mask_block_of(blk, c);
#pragma unroll
      for (int k = 0; k < 8; ++k) dm[k] += (k == j) ? d * x : 0.f;
      return d * m[j];
    };
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(src + c);
        const float4 d = *reinterpret_cast<const float4*>(dsrc + c);
        float4 o;
        o.x = one(c, v.x, d.x); o.y = one(c + 1, v.y, d.y); o.z = one(c + 2, v.z, d.z); o.w = one(c + 3, v.w, d.w);
        if (dst) *reinterpret_cast<float4*>(dst + c) = o;
      }
    } else {
      for (int c = lane; c < cols; c += 32) {
        const float o = one(c, src[c], dsrc[c]);
        if (dst) dst[c] = o;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dm[k] += __shfl_xor_sync(0xffffffffu, dm[k], o);
    if (lane == 0) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc += dm[k];
        if (k < blk.nb) d_p[r * dps + k] = acc * scale;
      }
    }
  }
}

#define LAUNCH_CHECK(ctx)          \
  do {                             \
    (ctx)->launches++;             \
    TDNNF_CUDA_OK(cudaGetLastError()); \
  } while (0)

extern "C" int tdnnf_darts_coef(tdnnf_ctx* ctx, const float* alpha, int n, int flags, float temperature,
                                const float* u_gumbel, float u_uniform, int share_index, float* coef, float* weff) {
  TDNNF_REQUIRE(ctx && alpha && coef && weff, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS, "number of time offsets must be in [1,16]");
  TDNNF_REQUIRE(!(flags & TDNNF_DARTS_USE_GUMBEL) || u_gumbel != nullptr, "use-gumbel needs n uniforms");
  TDNNF_REQUIRE(!(flags & TDNNF_DARTS_USE_GUMBEL) || temperature > 0.f, "temperature must be > 0");
  Uniforms ug;
  for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) ug.u[i] = (u_gumbel && i < n) ? u_gumbel[i] : 0.5f;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(darts_coef_kernel, dim3(1), dim3(32), 0, ctx->stream, 1, alpha, n, flags, temperature, ug, u_uniform, share_index, coef, weff));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_darts_weff_from_coef(tdnnf_ctx* ctx, const float* coef, int n, int flags, int share_index,
                                          float* weff) {
  TDNNF_REQUIRE(ctx && coef && weff, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS, "number of time offsets must be in [1,16]");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(darts_weff_kernel, dim3(1), dim3(32), 0, ctx->stream, 1, coef, n, flags, share_index, weff));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_darts_alpha_update(tdnnf_ctx* ctx, const float* s, const float* coef, int n, int flags,
                                        float temperature, int share_index, float lr, float* dalpha) {
  TDNNF_REQUIRE(ctx && s && coef && dalpha, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS, "number of time offsets must be in [1,16]");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(darts_alpha_update_kernel, dim3(1), dim3(32), 0, ctx->stream, 1, s, coef, n, flags, temperature, share_index, lr, dalpha));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_softmax_flops_fwd(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                                       int out_stride, const float* u, float inv_temp) {
  TDNNF_REQUIRE(ctx && in && out, "null argument");
  TDNNF_REQUIRE(rows >= 0 && cols > 0 && in_stride >= cols && out_stride >= cols, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  if (cols == 8 || cols == 16) {
    Uniforms g;
    for (int j = 0; j < TDNNF_MAX_OFFSETS; ++j) g.u[j] = (u && j < cols) ? -logf(-logf(u[j])) : 0.f;
    const bool vec = aligned16(in) && aligned16(out) && (in_stride % 4 == 0) && (out_stride % 4 == 0);
    const int grid = grid_for(rows, 128, ctx->num_sms);
    if (cols == 8)
      TDNNF_CUDA_OK(launch_pdl(softmax_flops_fwd_small<8>, dim3(grid), dim3(128), 0, ctx->stream, 1, in, rows, in_stride, out, out_stride, g, inv_temp, vec));
    else
      TDNNF_CUDA_OK(launch_pdl(softmax_flops_fwd_small<16>, dim3(grid), dim3(128), 0, ctx->stream, 1, in, rows, in_stride, out, out_stride, g, inv_temp, vec));
    LAUNCH_CHECK(ctx);
    return TDNNF_OK;
  }
  float* noise = nullptr;
  if (u) {
    // Gumbel noise for a wide matrix: computed on the host (cols floats) and staged through the arena.
    ctx->ws_reset();
    int rc = ctx->ws_reserve(sizeof(float) * cols);
    if (rc) return rc;
    noise = static_cast<float*>(ctx->ws_alloc(sizeof(float) * cols));
    if (!noise) return TDNNF_ERR_NOMEM;
    std::string tmp(sizeof(float) * cols, '\0');
    float* h = reinterpret_cast<float*>(&tmp[0]);
    for (int j = 0; j < cols; ++j) h[j] = -logf(-logf(u[j]));
    TDNNF_CUDA_OK(cudaMemcpyAsync(noise, h, sizeof(float) * cols, cudaMemcpyHostToDevice, ctx->stream));
    TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));  // h is a pageable temporary
  }
  const int grid = grid_for((long long)rows * 32, 256, ctx->num_sms);
  TDNNF_CUDA_OK(launch_pdl(softmax_flops_fwd_warp, dim3(grid), dim3(256), 0, ctx->stream, 1, in, rows, cols, in_stride, out, out_stride, noise, inv_temp));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_softmax_flops_bwd(tdnnf_ctx* ctx, const float* out_value, int ov_stride, float* out_deriv,
                                       int od_stride, float* in_deriv, int id_stride, int rows, int cols,
                                       float penalty, float inv_temp, int write_back_e) {
  TDNNF_REQUIRE(ctx && out_value && out_deriv && in_deriv, "null argument");
  // the reference writes flops_ entries 0..7 unconditionally (ref: simple.cc:10144-10152)
  TDNNF_REQUIRE(cols >= 8, "SoftmaxFlops backprop requires dim >= 8");
  TDNNF_REQUIRE(rows >= 0 && ov_stride >= cols && od_stride >= cols && id_stride >= cols, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  if (cols == 8 || cols == 16) {
    const bool vec = aligned16(out_value) && aligned16(out_deriv) && aligned16(in_deriv) && (ov_stride % 4 == 0) &&
                     (od_stride % 4 == 0) && (id_stride % 4 == 0);
    const int grid = grid_for(rows, 128, ctx->num_sms);
    if (cols == 8)
      TDNNF_CUDA_OK(launch_pdl(softmax_flops_bwd_small<8>, dim3(grid), dim3(128), 0, ctx->stream, 1, out_value, ov_stride, out_deriv, od_stride, in_deriv,
                                                                 id_stride, rows, penalty, inv_temp, write_back_e, vec));
    else
      TDNNF_CUDA_OK(launch_pdl(softmax_flops_bwd_small<16>, dim3(grid), dim3(128), 0, ctx->stream, 1, out_value, ov_stride, out_deriv, od_stride, in_deriv,
                                                                  id_stride, rows, penalty, inv_temp, write_back_e, vec));
  } else {
    const int grid = grid_for((long long)rows * 32, 256, ctx->num_sms);
    TDNNF_CUDA_OK(launch_pdl(softmax_flops_bwd_warp, dim3(grid), dim3(256), 0, ctx->stream, 1, out_value, ov_stride, out_deriv, od_stride, in_deriv,
                                                          id_stride, rows, cols, penalty, inv_temp, write_back_e));
  }
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_copyn_fwd(tdnnf_ctx* ctx, const float* in, int rows, int in_cols, int in_stride, float* out,
                               int out_cols, int out_stride, float scale) {
  TDNNF_REQUIRE(ctx && in && out, "null argument");
  TDNNF_REQUIRE(in_cols > 0 && out_cols % in_cols == 0 && in_stride >= in_cols && out_stride >= out_cols,
                "CopyN: output-dim must be a multiple of input-dim");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(copyn_fwd_kernel, dim3(grid_for((long long)rows * out_cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in, rows, in_cols, in_stride, out, out_cols, out_stride, scale));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_copyn_bwd(tdnnf_ctx* ctx, const float* out_deriv, int rows, int out_cols, int od_stride,
                               float* in_deriv, int in_cols, int id_stride, float scale) {
  TDNNF_REQUIRE(ctx && out_deriv && in_deriv, "null argument");
  TDNNF_REQUIRE(in_cols > 0 && out_cols % in_cols == 0 && id_stride >= in_cols && od_stride >= out_cols,
                "CopyN: output-dim must be a multiple of input-dim");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(copyn_bwd_kernel, dim3(grid_for((long long)rows * in_cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      out_deriv, rows, out_cols, od_stride, in_deriv, in_cols, id_stride, scale));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_onehot_fwd(tdnnf_ctx* ctx, float* out, int rows, int dim, int out_stride, float u) {
  TDNNF_REQUIRE(ctx && out, "null argument");
  TDNNF_REQUIRE(dim > 0 && out_stride >= dim, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(onehot_fwd_kernel, dim3(grid_for((long long)rows * dim, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, out, rows, dim,
                                                                                                 out_stride, u));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_add_row_sum(tdnnf_ctx* ctx, const float* mat, int rows, int cols, int stride, float scale,
                                 float* vec) {
  TDNNF_REQUIRE(ctx && mat && vec, "null argument");
  TDNNF_REQUIRE(cols > 0 && stride >= cols, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  int gy = (rows + 255) / 256;
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  TDNNF_CUDA_OK(launch_pdl(add_row_sum_kernel, dim3((cols + 31) / 32, gy), dim3(32, 8), 0, ctx->stream, 1, mat, rows, cols, stride, scale, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_scale_offset_rows(tdnnf_ctx* ctx, const float* in, int rows, int cols, int in_stride, float* out,
                                       int out_stride, const float* scale, const float* offset) {
  TDNNF_REQUIRE(ctx && in && out && scale, "null argument");
  TDNNF_REQUIRE(cols > 0 && in_stride >= cols && out_stride >= cols, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const bool vec = (cols % 4 == 0) && (in_stride % 4 == 0) && (out_stride % 4 == 0) && aligned16(in) &&
                   aligned16(out) && aligned16(scale) && (!offset || aligned16(offset));
  const long long total = vec ? (long long)rows * (cols / 4) : (long long)rows * cols;
  TDNNF_CUDA_OK(launch_pdl(scale_offset_rows_kernel, dim3(grid_for(total, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in, rows, cols, in_stride, out, out_stride, scale, offset, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_elementwise_product_fwd(tdnnf_ctx* ctx, const float* in, int rows, int out_cols, int in_stride,
                                             float* out, int out_stride) {
  TDNNF_REQUIRE(ctx && in && out, "null argument");
  TDNNF_REQUIRE(out_cols > 0 && in_stride >= 2 * out_cols && out_stride >= out_cols, "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(ewprod_fwd_kernel, dim3(grid_for((long long)rows * out_cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in, rows, out_cols, in_stride, out, out_stride));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_elementwise_product_bwd(tdnnf_ctx* ctx, const float* in, int in_stride, const float* out_deriv,
                                             int od_stride, float* in_deriv, int id_stride, int rows, int out_cols) {
  TDNNF_REQUIRE(ctx && in && out_deriv && in_deriv, "null argument");
  TDNNF_REQUIRE(out_cols > 0 && in_stride >= 2 * out_cols && id_stride >= 2 * out_cols && od_stride >= out_cols,
                "bad matrix shape");
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(ewprod_bwd_kernel, dim3(grid_for((long long)rows * out_cols, 256, ctx->num_sms)), dim3(256), 0, ctx->stream, 1, 
      in, in_stride, out_deriv, od_stride, in_deriv, id_stride, rows, out_cols));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

static int mask_blocks(const int32_t* widths, int nb, int cols, MaskBlocks* b) {
  TDNNF_REQUIRE(widths && nb >= 1 && nb <= 8, "1 to 8 candidate blocks");
  int end = 0;
  b->nb = nb;
  for (int j = 0; j < 8; ++j) {
    if (j < nb) {
      TDNNF_REQUIRE(widths[j] > 0, "empty candidate block");
      end += widths[j];
    }
    b->end[j] = end;
  }
  TDNNF_REQUIRE(end == cols, "block widths must add up to the bottleneck dimension");
  return TDNNF_OK;
}

extern "C" int tdnnf_shared_mask_fwd(tdnnf_ctx* ctx, const float* p, int rows, int nb, int p_stride, const float* lin, int cols,
                                     int lin_stride, float* out, int out_stride, const int32_t* widths, float scale) {
  TDNNF_REQUIRE(ctx && p && lin && out, "null argument");
  TDNNF_REQUIRE(p_stride >= nb && lin_stride >= cols && out_stride >= cols, "stride < cols");
  MaskBlocks b;
  int rc = mask_blocks(widths, nb, cols, &b);
  if (rc) return rc;
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const bool vec = cols % 4 == 0 && lin_stride % 4 == 0 && out_stride % 4 == 0 && ((uintptr_t)lin & 15) == 0 && ((uintptr_t)out & 15) == 0;
  const int blocks = (int)std::min<long long>(((long long)rows + 7) / 8, (long long)ctx->num_sms * 8);
  TDNNF_CUDA_OK(launch_pdl(shared_mask_fwd_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, p, p_stride, lin, lin_stride, out, out_stride, rows, cols, b, scale, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

extern "C" int tdnnf_shared_mask_bwd(tdnnf_ctx* ctx, const float* p, int p_stride, const float* lin, int lin_stride,
                                     const float* d_out, int do_stride, float* d_lin, int dl_stride, float* d_p, int dp_stride,
                                     int rows, int cols, int nb, const int32_t* widths, float scale) {
  TDNNF_REQUIRE(ctx && p && lin && d_out && d_p, "null argument");
  TDNNF_REQUIRE(p_stride >= nb && dp_stride >= nb && lin_stride >= cols && do_stride >= cols && (!d_lin || dl_stride >= cols),
                "stride < cols");
  MaskBlocks b;
  int rc = mask_blocks(widths, nb, cols, &b);
  if (rc) return rc;
  if (rows == 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const bool vec = cols % 4 == 0 && lin_stride % 4 == 0 && do_stride % 4 == 0 && (!d_lin || dl_stride % 4 == 0) &&
                   ((uintptr_t)lin & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && ((uintptr_t)d_lin & 15) == 0;
  const int blocks = (int)std::min<long long>(((long long)rows + 7) / 8, (long long)ctx->num_sms * 8);
  TDNNF_CUDA_OK(launch_pdl(shared_mask_bwd_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, p, p_stride, lin, lin_stride, d_out, do_stride, d_lin, dl_stride, d_p,
                                                         dp_stride, rows, cols, b, scale, vec));
  LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
