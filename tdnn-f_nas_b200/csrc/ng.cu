// Device-side helpers of the natural-gradient update (OnlineNaturalGradient::PreconditionDirections at
// ref tdnn.cc:598-599, simple.cc:9542) that are not GEMMs: the trace of X~ X~^T over the spliced input
// WITHOUT materialising X~ (the reference builds the R x (n*D_in+1) matrix in_value_temp, tdnn.cc:476-514),
// the "scale" factor kept on the device (no host sync), and an axpy whose coefficient lives on the device.
#include "context.h"

using namespace tdnnf;

namespace {

// sumsq[i] += sum over the rows of view i (rows row_offsets[i] + k*row_stride, k < out_rows) of ||row||^2.
// One warp per row; each input row is read exactly once whatever the number of overlapping views.
__global__ void __launch_bounds__(256)
view_sumsq_kernel(const float* __restrict__ in, int in_rows, int in_dim, long long ld, int out_rows, int n,
                  TdnnfOffsets offs, int row_stride, double* __restrict__ sumsq) {
  __shared__ double acc[TDNNF_MAX_OFFSETS];
  if (threadIdx.x < TDNNF_MAX_OFFSETS) acc[threadIdx.x] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const bool vec = ((ld & 3) == 0) && ((in_dim & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  double local[TDNNF_MAX_OFFSETS];
#pragma unroll
  for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) local[i] = 0.0;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < in_rows; r += (long long)gridDim.x * wpb) {
    const float* row = in + r * ld;
    float s = 0.f;
    if (vec) {
      for (int c = lane * 4; c < in_dim; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + c);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    } else {
      for (int c = lane; c < in_dim; c += 32) s += row[c] * row[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) {
        if (i < n) {
          const long long d = r - offs.v[i];
          if (d >= 0 && d % row_stride == 0 && d / row_stride < out_rows) local[i] += (double)s;
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i)
      if (i < n && local[i] != 0.0) atomicAdd(&acc[i], local[i]);
  }
  __syncthreads();
  if (threadIdx.x < n && acc[threadIdx.x] != 0.0) atomicAdd(&sumsq[threadIdx.x], acc[threadIdx.x]);
}

// out[0] = tr(X X^T) = sum_i w_i^2 sumsq_i + ones_rows
// out[1] = tr(X^ X^^T) with X^ = X - (X W^T) W:  tr(XX^T) - 2 tr(L) + <L, W W^T>,  L = H^T H, H = X W^T
// out[2] = sqrt(out[0] / out[1])   (1 when tr(X X^T) <= 0), the "scale" of PreconditionDirections
__global__ void ng_scale_kernel(const double* __restrict__ sumsq, const float* __restrict__ weff, int n, float ones_rows,
                                const float* __restrict__ L, int l_ld, const float* __restrict__ WWt, int w_ld, int r,
                                float* __restrict__ out) {
  // 1024 threads, row-strided: every thread has ~r*r/1024 independent loads in flight (the single-block version with a
  // dependent idx/r loop took 17 us at r = 80, all of it load latency)
  __shared__ double red_tr[32], red_dot[32];
  double tr = 0.0, dot = 0.0;
  for (int i = threadIdx.x / 32; i < r; i += blockDim.x / 32) {
    for (int j = threadIdx.x % 32; j < r; j += 32) {
      const double l = L[(long long)i * l_ld + j];
      if (i == j) tr += l;
      dot += l * (double)WWt[(long long)i * w_ld + j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red_tr[threadIdx.x >> 5] = tr;
    red_dot[threadIdx.x >> 5] = dot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      red_tr[0] += red_tr[w];
      red_dot[0] += red_dot[w];
    }
  }
  if (threadIdx.x == 0) {
    double initial = (double)ones_rows;
    for (int i = 0; i < n; ++i) {
      const double w = weff ? (double)weff[i] : 1.0;
      initial += w * w * sumsq[i];
    }
    const double fin = initial - 2.0 * red_tr[0] + red_dot[0];
    out[0] = (float)initial;
    out[1] = (float)fin;
    out[2] = (initial <= 0.0 || !(fin > 0.0)) ? 1.0f : (float)sqrt(initial / fin);
  }
}

__global__ void mat_axpy_dev_kernel(float alpha, const float* __restrict__ f1, const float* __restrict__ f2,
                                    const float* __restrict__ src, long long ss, float* __restrict__ dst, long long ds,
                                    int rows, int cols) {
  float a = alpha;
  if (f1) a *= *f1;
  if (f2) a *= *f2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)rows * cols;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    dst[r * ds + c] += a * src[r * ss + c];
  }
}

}  // namespace

extern "C" int tdnnf_darts_view_sumsq(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride,
                                      int out_rows, int n, const int32_t* row_offsets, int row_stride, double* sumsq) {
  TDNNF_REQUIRE(ctx && in && row_offsets && sumsq, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS && row_stride >= 1 && in_rows > 0 && in_dim > 0 && in_stride >= in_dim,
                "bad view");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TdnnfOffsets offs;
  for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) offs.v[i] = i < n ? row_offsets[i] : 0;
  TDNNF_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(double) * n, ctx->stream));
  const int blocks = std::min((in_rows + 7) / 8, ctx->num_sms * 8);
  view_sumsq_kernel<<<blocks, 256, 0, ctx->stream>>>(in, in_rows, in_dim, in_stride, out_rows, n, offs, row_stride, sumsq);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ng_scale(tdnnf_ctx* ctx, const double* sumsq, const float* weff, int n, float ones_rows,
                              const float* L, int l_stride, const float* WWt, int w_stride, int rank, float* out3) {
  TDNNF_REQUIRE(ctx && sumsq && L && WWt && out3, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS && rank >= 1 && l_stride >= rank && w_stride >= rank, "bad argument");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  ng_scale_kernel<<<1, 1024, 0, ctx->stream>>>(sumsq, weff, n, ones_rows, L, l_stride, WWt, w_stride, rank, out3);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_mat_axpy_dev(tdnnf_ctx* ctx, float alpha, const float* factor1_dev, const float* factor2_dev,
                                  const float* src, int src_stride, float* dst, int dst_stride, int rows, int cols) {
  TDNNF_REQUIRE(ctx && src && dst, "null argument");
  if (rows <= 0 || cols <= 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  long long b = ((long long)rows * cols + 255) / 256;
  b = std::max(1LL, std::min(b, (long long)ctx->num_sms * 16));
  mat_axpy_dev_kernel<<<(int)b, 256, 0, ctx->stream>>>(alpha, factor1_dev, factor2_dev, src, src_stride, dst, dst_stride,
                                                       rows, cols);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_get_stream(tdnnf_ctx* ctx, void** stream) {
  TDNNF_REQUIRE(ctx && stream, "null argument");
  *stream = ctx->stream;
  return TDNNF_OK;
}
