// ConstrainOrthonormalInternal (ref: nnet-utils.cc:914-1035): the semi-orthogonal constraint of the TDNN-F
// bottleneck (the `linear` half of every tdnnf-layer of the manual / derived systems, BASELINE configs[1]).
//   P = M M^T,  scale^2 = tr(P P^T) / tr(P) when floating,  M <- M - 4 (nu / scale^2) (P - scale^2 I) M
// The reference runs SymAddMat2 + CopyLowerToUpper + Trace + TraceMatMat (two host syncs) + AddToDiag + AddMatMat +
// AddMat, and a transposed copy of M in and out when rows > cols (utils.cc:1067-1074).  Here: three launches, no host
// sync, no copy of M: (1) split-K Gram partials, (2) one CTA per row sums them and leaves that row's share of the two
// traces, (3) M is updated IN PLACE panel by panel, every CTA re-deriving the floating scale and the ratio-driven
// update speed from the row traces -- column j of the update
// depends only on column j of M (row j when M is used transposed), so a CTA that holds a panel of 32 such vectors in
// shared memory can overwrite it.  fp32 FMAs throughout (P is at most a few hundred squared; K = a few thousand:
// 0.16 GFLOP per call, every fourth minibatch on average -- not tensor-core work).
#include <algorithm>

#include "context.h"

using namespace tdnnf;

namespace {

constexpr int kTile = 32;      // Gram tile edge
constexpr int kPanel = 32;     // vectors per CTA in the update
constexpr int kMaxDim = 512;   // largest P dimension (shared-memory panel of kMaxDim x 33 floats)

// element (i, k) of the logical n x K matrix A: A = M (trans == 0) or A = M^T (trans == 1)
__device__ __forceinline__ float load_a(const float* __restrict__ M, long long ld, int trans, int i, int k) {
  return trans ? M[(long long)k * ld + i] : M[(long long)i * ld + k];
}

// partial[z][i][j] = sum over the z-th K range of A[i][k] A[j][k]
__global__ void __launch_bounds__(256) ortho_gram_kernel(const float* __restrict__ M, long long ld, int trans, int n, int K,
                                                         int k_per_split, float* __restrict__ partial) {
  __shared__ float As[kTile][kTile + 1], Bs[kTile][kTile + 1];
  const int i0 = blockIdx.y * kTile, j0 = blockIdx.x * kTile;
  if (j0 > i0) return;  // lower triangle only; the finish kernel mirrors it
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 x 2 outputs each
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = k_begin; k0 < k_end; k0 += kTile) {
    // stage 32 x 32 of each operand; the fastest-varying thread index follows the contiguous direction of M
    for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
      const int a = e & 31, b = e >> 5;
      const int i = trans ? a : b, k = trans ? b : a;
      const bool kin = k0 + k < k_end;
      As[i][k] = (kin && i0 + i < n) ? load_a(M, ld, trans, i0 + i, k0 + k) : 0.f;
      Bs[i][k] = (kin && j0 + i < n) ? load_a(M, ld, trans, j0 + i, k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kTile; ++k) {
      const float a0 = As[ty][k], a1 = As[ty + 16][k], b0 = Bs[tx][k], b1 = Bs[tx + 16][k];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.z * n * n;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < n && j < n) out[(size_t)i * n + j] = acc[a][b];
    }
}

// CTA i: row i of P = sum_z partial[z] (lower triangle mirrored), written with pitch np (a multiple of 4, pad = 0), and
// this row's share of tr(P) and tr(P P^T) as doubles (summed in a fixed order by the update kernel: deterministic, so
// data-parallel replicas that apply the constraint to identical models stay bit-identical).
__global__ void __launch_bounds__(128) ortho_reduce_kernel(const float* __restrict__ partial, int splits, int n, int np,
                                                           float* __restrict__ P, double* __restrict__ row_tr) {
  __shared__ double red[4];
  const int i = blockIdx.x;
  const size_t plane = (size_t)n * n;
  double tr2 = 0.0;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    float v = 0.f;
    if (j < n) {
      const size_t src = (j <= i) ? (size_t)i * n + j : (size_t)j * n + i;
      for (int z = 0; z < splits; ++z) v += partial[(size_t)z * plane + src];
      tr2 += (double)v * v;
      if (j == i) row_tr[i] = (double)v;
    }
    P[(size_t)i * np + j] = v;
  }
  for (int o = 16; o > 0; o >>= 1) tr2 += __shfl_xor_sync(0xffffffffu, tr2, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr2;
  __syncthreads();
  if (threadIdx.x == 0) row_tr[n + i] = (red[0] + red[1]) + (red[2] + red[3]);
}

// vectors v_j (length n): column j of M (trans == 0) or row j of M (trans == 1);
//   v_j <- v_j + coef * (P - scale^2 I) v_j,  coef = -4 update_speed / scale^2   (utils.cc:941-1033)
// Every CTA re-derives the scalars from the per-row traces (n <= 512 doubles: cheaper than another launch).
__global__ void __launch_bounds__(256) ortho_update_kernel(float* __restrict__ M, long long ld, int trans, int n, int np,
                                                           int num_vec, const float* __restrict__ P,
                                                           const double* __restrict__ row_tr, float scale_in,
                                                           float* __restrict__ info) {
  extern __shared__ float V[];  // n x (kPanel + 1)
  __shared__ double red[2][8];
  __shared__ float s_coef, s_scale2;
  constexpr int P1 = kPanel + 1;
  const int j0 = blockIdx.x * kPanel;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    double t = 0.0, t2 = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { t += row_tr[i]; t2 += row_tr[n + i]; }
    for (int o = 16; o > 0; o >>= 1) {
      t += __shfl_xor_sync(0xffffffffu, t, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) { red[0][warp] = t; red[1][warp] = t2; }
  }
  if (!trans) {
    for (int k = warp; k < n; k += 8) V[k * P1 + lane] = (j0 + lane < num_vec) ? M[(long long)k * ld + j0 + lane] : 0.f;
  } else {
    for (int j = warp; j < kPanel; j += 8)
      for (int k = lane; k < n; k += 32) V[k * P1 + j] = (j0 + j < num_vec) ? M[(long long)(j0 + j) * ld + k] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) { t += red[0][w]; t2 += red[1][w]; }
    const float trace_P = (float)t, trace_P_P = (float)t2;
    float scale = scale_in, speed = 0.125f, ratio = 0.f;
    bool ok = true;
    if (scale_in < 0.f) {  // floating scale (utils.cc:941-983)
      scale = sqrtf(trace_P_P / trace_P);
      ratio = trace_P_P * (float)n / (trace_P * trace_P);
      ok = ratio > 0.999f;  // KALDI_ASSERT in the reference; here the update is skipped and info[1] tells
      if (ratio > 1.02f) {
        speed *= 0.5f;
        if (ratio > 1.1f) speed *= 0.5f;
      }
    }
    const float s2 = scale * scale;
    s_scale2 = s2;
    s_coef = ok ? -4.0f * (speed / s2) : 0.f;
    if (info != nullptr && blockIdx.x == 0) {
      info[0] = scale;
      info[1] = ratio;
      info[2] = speed;
      // ||P - s^2 I||_F^2 = tr(P P^T) - 2 s^2 tr(P) + n s^4
      const double e2 = t2 - 2.0 * (double)s2 * t + (double)n * (double)s2 * (double)s2;
      info[3] = (float)sqrt(e2 > 0.0 ? e2 : 0.0);
    }
  }
  __syncthreads();
  const float coef = s_coef, s2 = s_scale2;
  // warp w owns 8 consecutive rows per pass; lane = vector.  P is symmetric: rows i..i+7 of column k are 8 consecutive
  // floats of row k (two 16-byte broadcast loads).
  for (int ib = warp * 8; ib < n; ib += 64) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool second = ib + 4 < np;
#pragma unroll 2
    for (int k = 0; k < n; ++k) {
      const float v = V[k * P1 + lane];
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(P + (size_t)k * np + ib));
      const float4 p1 = second ? __ldg(reinterpret_cast<const float4*>(P + (size_t)k * np + ib + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[0] = fmaf(p0.x, v, acc[0]);
      acc[1] = fmaf(p0.y, v, acc[1]);
      acc[2] = fmaf(p0.z, v, acc[2]);
      acc[3] = fmaf(p0.w, v, acc[3]);
      acc[4] = fmaf(p1.x, v, acc[4]);
      acc[5] = fmaf(p1.y, v, acc[5]);
      acc[6] = fmaf(p1.z, v, acc[6]);
      acc[7] = fmaf(p1.w, v, acc[7]);
    }
    if (j0 + lane < num_vec) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int i = ib + a;
        if (i < n) {
          const float old = V[i * P1 + lane];
          const float nv = old + coef * (acc[a] - s2 * old);  // (P - s^2 I) v
          if (!trans) M[(long long)i * ld + j0 + lane] = nv;
          else M[(long long)(j0 + lane) * ld + i] = nv;  // rows > cols is the rare case: plain 4-byte scatter
        }
      }
    }
  }
}

}  // namespace

extern "C" int tdnnf_constrain_orthonormal(tdnnf_ctx* ctx, float* M, int rows, int cols, int stride, float scale,
                                           float* info_dev) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  TDNNF_REQUIRE(M && rows > 0 && cols > 0 && stride >= cols, "bad matrix");
  TDNNF_REQUIRE(scale != 0.0f, "ConstrainOrthonormalInternal: scale must not be 0 (KALDI_ASSERT, nnet-utils.cc:915)");
  const int trans = rows > cols ? 1 : 0;  // utils.cc:1067-1074: the constraint acts on the smaller dimension
  const int n = trans ? cols : rows, K = trans ? rows : cols;
  TDNNF_REQUIRE(n <= kMaxDim, "tdnnf_constrain_orthonormal: min(rows, cols) > 512 is not supported");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  int splits = std::max(1, std::min(16, K / 256));
  const int k_per_split = ((K + splits - 1) / splits + kTile - 1) / kTile * kTile;
  splits = (K + k_per_split - 1) / k_per_split;
  const size_t nn = (size_t)n * n;
  const int np = (n + 3) & ~3;
  {
    const int rc = ctx->ws_reserve(sizeof(float) * (nn * splits + (size_t)n * np) + sizeof(double) * 2 * n + 3 * 1024);
    if (rc != TDNNF_OK) return rc;
  }
  ctx->ws_reset();
  float* partial = static_cast<float*>(ctx->ws_alloc(sizeof(float) * nn * splits));
  float* P = static_cast<float*>(ctx->ws_alloc(sizeof(float) * (size_t)n * np));
  double* row_tr = static_cast<double*>(ctx->ws_alloc(sizeof(double) * 2 * n));
  if (!partial || !P || !row_tr) return TDNNF_ERR_CUDA;
  const int tiles = (n + kTile - 1) / kTile;
  // upper-triangle tiles exit at once and leave their partial entries unwritten: the reduce kernel never reads them
  ortho_gram_kernel<<<dim3(tiles, tiles, splits), 256, 0, ctx->stream>>>(M, stride, trans, n, K, k_per_split, partial);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  ortho_reduce_kernel<<<n, 128, 0, ctx->stream>>>(partial, splits, n, np, P, row_tr);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  const size_t smem = sizeof(float) * (size_t)n * (kPanel + 1);
  if (smem > 48 * 1024)  // per device and per call: the attribute is cheap to set and contexts may live on several devices
    TDNNF_CUDA_OK(cudaFuncSetAttribute(ortho_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(float) * kMaxDim * (kPanel + 1))));
  ortho_update_kernel<<<(K + kPanel - 1) / kPanel, 256, smem, ctx->stream>>>(M, stride, trans, n, np, K, P, row_tr, scale,
                                                                             info_dev);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}
