// Host side of the TdnnDARTSV3 GEMMs: operand-plane pre-pass kernels, TMA tensor maps, the
// work decomposition and the C-ABI entry points tdnnf_darts_{propagate,backprop_data,backprop_params}.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "context.h"
#include "splice_gemm.cuh"

namespace tdnnf {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int x, int m) { return (x + m - 1) / m; }

// ------------------------------------------------------------------------------------------
// Pre-pass 1: fp32 rows -> bf16 hi/lo planes with K (the column index) contiguous.
//   dst[plane][c][q][k]  (pitch Kpad, zero padded) =
//        scale[c] * src[(q*r + c*c_row_mul) * ld + c*c_col_mul + k]      k < D, source row < R
// X planes of Propagate (c = row % r), out_deriv planes of Backprop (r = 1) and the per-offset
// weight planes (c = offset, c_col_mul = D_in, scale = weff) all go through here.
// ------------------------------------------------------------------------------------------
// fp16 != 0: ONE fp16 plane, x * pow2_scale(*absmax_in) (absmax_in null: unscaled), written to `hi`.
// By-products of reading every element once (either may be null):
//   absmax_out  max |x| (ordered as an int: values are >= 0)
//   rowsq       rowsq[source row] += sum of squares of that row (segmented warp reduction, one red.add per run of
//               lanes on the same row) -- the natural-gradient tr(X X^T) without another pass over X
__global__ void split_rows_kernel(const float* __restrict__ src, int R, int D, long long ld, int r, int groups,
                                  int c_row_mul, int c_col_mul, const float* __restrict__ scale, int Q, int Kpad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  __nv_bfloat16* __restrict__ lo2, int fp16, const float* __restrict__ absmax_in,
                                  float* __restrict__ absmax_out, float* __restrict__ rowsq) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int kvec = Kpad >> 3;
  const long long total = (long long)groups * Q * kvec;
  const long long total_up = (total + 31) & ~31LL;  // warp-uniform trip count (the by-products use shuffles)
  const bool aligned = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((c_col_mul & 3) == 0);
  const float s16 = (fp16 && absmax_in) ? pow2_scale(*absmax_in) : 1.0f;
  const int lane = threadIdx.x & 31;
  float amax = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_up;
       idx += (long long)gridDim.x * blockDim.x) {
    const bool valid = idx < total;
    const int kv = (int)(idx % kvec);
    const long long rowidx = idx / kvec;
    const int q = (int)(rowidx % Q);
    const int c = (int)(rowidx / Q);
    const long long srow = (long long)q * r + (long long)c * c_row_mul;
    const int k0 = kv * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    const bool live = valid && srow < R && k0 < D;
    if (live) {
      const float* s = src + srow * ld + (long long)c * c_col_mul + k0;
      if (aligned && k0 + 8 <= D) {
        const float4 a = *reinterpret_cast<const float4*>(s);
        const float4 b = *reinterpret_cast<const float4*>(s + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k0 + j < D) v[j] = s[j];
      }
      if (scale != nullptr) {
        const float sc = scale[c];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sc;
      }
    }
    if (absmax_out != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[j]));
    }
    if (rowsq != nullptr && (kvec & 31) == 0) {
      // every aligned group of 32 consecutive idx lies inside one row: plain warp sum, one red.add per warp
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sq = fmaf(v[j], v[j], sq);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0 && valid && srow < R) atomicAdd(rowsq + srow, sq);
    } else if (rowsq != nullptr) {
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sq = fmaf(v[j], v[j], sq);
      const long long key = live ? srow : -1 - (long long)lane;  // dead lanes: singleton runs, nothing to add
      const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float other = __shfl_down_sync(0xffffffffu, sq, o);
        const long long okey = __shfl_down_sync(0xffffffffu, key, o);
        if (lane + o < 32 && okey == key) sq += other;
      }
      if (live && (lane == 0 || prev != key)) atomicAdd(rowsq + srow, sq);
    }
    if (!valid) continue;
    const long long o = rowidx * Kpad + k0;
    if (fp16) {
      __align__(16) __half h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(v[j] * s16);
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
      continue;
    }
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
    __align__(16) __nv_bfloat16 l2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __float2bfloat16_rn(v[j]);
      const float r1 = v[j] - __bfloat162float(h[j]);
      l[j] = __float2bfloat16_rn(r1);
      l2[j] = __float2bfloat16_rn(r1 - __bfloat162float(l[j]));
    }
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
    if (lo2 != nullptr) *reinterpret_cast<uint4*>(lo2 + o) = *reinterpret_cast<const uint4*>(l2);
  }
  if (absmax_out != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.f) atomicMax(reinterpret_cast<int*>(absmax_out), __float_as_int(amax));
  }
}

// max |x| over a matrix (for operands that were not row-split earlier in the cache scope).
__global__ void absmax_kernel(const float* __restrict__ src, int R, int D, long long ld, float* __restrict__ absmax_out) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  float amax = 0.f;
  const long long total = (long long)R * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    amax = fmaxf(amax, fabsf(src[(idx / D) * ld + idx % D]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(reinterpret_cast<int*>(absmax_out), __float_as_int(amax));
}

// ------------------------------------------------------------------------------------------
// Pre-pass 2: fp32 rows -> TRANSPOSED bf16 hi/lo planes (the row index becomes K, contiguous).
//   dst[plane][c][j][q]  (pitch Qp, zero padded) =
//        scale[c] * src[(q*r + c*c_row_mul + c_row_off[c]) * ld + c*c_col_mul + j]   j < J, q < Q, source row < R
// X^T / out_deriv^T planes of the parameter gradient and the W_i^T planes of the data gradient.
// ------------------------------------------------------------------------------------------
struct GroupRowOffsets {
  int v[kMaxSeg];
};

__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ src, int R, int J, long long ld, int r, int c_row_mul, int c_col_mul,
                       GroupRowOffsets c_row_off, const float* __restrict__ scale, int Q, int Qp,
                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ lo2,
                       float* __restrict__ colsum, float colsum_scale, int fp16, const float* __restrict__ absmax_in) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  // 64 (q) x 64 (j) tile: 256-byte coalesced float4 reads along j, 128-byte (8 x bf16 per lane) writes along q.
  __shared__ float tile[64][65];
  const int c = blockIdx.z;
  const int q0 = blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  const float sc = scale ? scale[c] : 1.0f;
  const long long col0 = (long long)c * c_col_mul + j0;
  const bool vec = ((ld & 3) == 0) && ((col0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (j0 + 64 <= J);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qq = (t >> 4) + 16 * i, j4 = (t & 15) * 4;
    const int q = q0 + qq;
    const long long srow = (long long)q * r + (long long)c * c_row_mul + c_row_off.v[c];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < Q && srow < R) {
      const float* sp = src + srow * ld + col0 + j4;
      if (vec) {
        v = *reinterpret_cast<const float4*>(sp);
      } else {
        if (j0 + j4 + 0 < J) v.x = sp[0];
        if (j0 + j4 + 1 < J) v.y = sp[1];
        if (j0 + j4 + 2 < J) v.z = sp[2];
        if (j0 + j4 + 3 < J) v.w = sp[3];
      }
    }
    tile[qq][j4 + 0] = sc * v.x;
    tile[qq][j4 + 1] = sc * v.y;
    tile[qq][j4 + 2] = sc * v.z;
    tile[qq][j4 + 3] = sc * v.w;
  }
  __syncthreads();
  if (colsum != nullptr && t < 64 && j0 + t < J) {
    // fused AddRowSumMat (bias gradient, ref: tdnn.cc:607-617): column sums of this tile, one atomic per column
    float sum = 0.f;
#pragma unroll 8
    for (int qq = 0; qq < 64; ++qq) sum += tile[qq][t];
    atomicAdd(colsum + j0 + t, colsum_scale * sum);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int jj = (t >> 3) + 32 * i, qc = (t & 7) * 8;
    const int j = j0 + jj, q = q0 + qc;
    if (j >= J || q >= Qp) continue;  // Qp is a multiple of 8
    const long long o = ((long long)c * J + j) * Qp + q;
    if (fp16) {  // one fp16 plane, scaled into the fp16 range by a power of two
      const float s16 = absmax_in ? pow2_scale(*absmax_in) : 1.0f;
      __align__(16) __half hh[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hh[k] = __float2half_rn(tile[qc + k][jj] * s16);
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hh);
      continue;
    }
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
    __align__(16) __nv_bfloat16 l2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float a = tile[qc + k][jj];
      h[k] = __float2bfloat16_rn(a);
      const float r1 = a - __bfloat162float(h[k]);
      l[k] = __float2bfloat16_rn(r1);
      l2[k] = __float2bfloat16_rn(r1 - __bfloat162float(l[k]));
    }
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
    if (lo2 != nullptr) *reinterpret_cast<uint4*>(lo2 + o) = *reinterpret_cast<const uint4*>(l2);
  }
}

// out[r, :] = bias (or 0) for all rows: the starting value when Propagate is split along K.
__global__ void init_rows_kernel(float* __restrict__ out, int rows, int cols, long long ld,
                                 const float* __restrict__ bias) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    out[r * ld + c] = bias ? bias[c] : 0.f;
  }
}

// out[q, j] = bias[j] + sum_i Y[(offs[i] + q*row_stride), i*rank + j]     (tdnnf_darts_project, second stage)
__global__ void project_gather_kernel(const float* __restrict__ Y, long long y_ld, int out_rows, int rank, int n,
                                      GroupRowOffsets offs, int row_stride, const float* __restrict__ bias,
                                      float* __restrict__ out, long long out_ld) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)out_rows * rank;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % rank);
    const long long q = idx / rank;
    float acc = bias ? bias[j] : 0.f;
    for (int i = 0; i < n; ++i) acc += Y[(offs.v[i] + q * row_stride) * y_ld + i * rank + j];
    out[q * out_ld + j] = acc;
  }
}

// Hs[r', i*od + j] = weff[i] * od_mat[q, j] where r' = offs[i] + q*row_stride, zero where no such q exists: the n shifted
// copies of a NARROW out_deriv side by side (tdnnf_darts_backprop_params, stacked form)
__global__ void stack_shift_kernel(const float* __restrict__ od_mat, long long od_ld, int out_rows, int od, int n,
                                   GroupRowOffsets offs, int row_stride, const float* __restrict__ weff, int in_rows,
                                   float* __restrict__ Hs, int ncols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)in_rows * ncols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % ncols);
    const long long rp = idx / ncols;
    const int i = c / od, j = c - i * od;
    const long long t = rp - offs.v[i];
    float v = 0.f;
    if (t >= 0 && t % row_stride == 0) {
      const long long q = t / row_stride;
      if (q < out_rows) v = weff[i] * od_mat[q * od_ld + j];
    }
    Hs[idx] = v;
  }
}

// dW[j, i*in_dim + d] += alpha * T[(i*od + j), d]
__global__ void stack_scatter_kernel(const float* __restrict__ T, int in_dim, int od, int n, float alpha, float* __restrict__ dW,
                                     long long dw_ld) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)n * od * in_dim;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % in_dim);
    const int c = (int)(idx / in_dim);
    const int i = c / od, j = c - i * od;
    dW[(long long)j * dw_ld + (long long)i * in_dim + d] += alpha * T[idx];
  }
}

// ------------------------------------------------------------------------------------------
struct Planes {
  __nv_bfloat16* base = nullptr;  // hi plane; lo plane follows at base + plane_elems
  long long plane_elems = 0;
  int K = 0;       // valid extent of the contiguous dimension (TMA zero-fills beyond it)
  int Kpitch = 0;  // pitch in elements (multiple of 8)
  int rows = 0;    // rows per group
  int groups = 0;
  int np = 2;      // planes (2: hi, lo; 3: hi, mid, lo)
};

static size_t planes_bytes(int np, int groups, int rows, int Kpitch) {
  size_t b = (size_t)np * groups * rows * Kpitch * sizeof(__nv_bfloat16);
  return (b + 1023) & ~size_t(1023);
}

static int make_map(tdnnf_ctx* ctx, const Planes& pl, int box_rows, CUtensorMap* out) {
  cuuint64_t dims[4] = {(cuuint64_t)pl.K, (cuuint64_t)pl.rows, (cuuint64_t)pl.groups, (cuuint64_t)pl.np};
  cuuint64_t strides[3] = {(cuuint64_t)pl.Kpitch * 2, (cuuint64_t)pl.rows * pl.Kpitch * 2,
                           (cuuint64_t)pl.plane_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1, (cuuint32_t)pl.np};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, pl.base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TDNNF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return TDNNF_OK;
}

// The same row planes seen MN-major: (64 columns, row, 64-column chunk, group, plane); a box of `chunks` chunks x
// `box_k` rows lands in shared memory as [plane][chunk][k][64 columns], 128-byte swizzled (see umma_desc_mn_sw128).
static int make_map_mn(tdnnf_ctx* ctx, const Planes& pl, int chunks, int box_k, CUtensorMap* out) {
  if (pl.Kpitch % 64 != 0) return fail(TDNNF_ERR_INVALID, "MN-major planes need a pitch that is a multiple of 64");
  cuuint64_t dims[5] = {64, (cuuint64_t)pl.rows, (cuuint64_t)(pl.Kpitch / 64), (cuuint64_t)pl.groups, (cuuint64_t)pl.np};
  cuuint64_t strides[4] = {(cuuint64_t)pl.Kpitch * 2, 128, (cuuint64_t)pl.rows * pl.Kpitch * 2, (cuuint64_t)pl.plane_elems * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)box_k, (cuuint32_t)chunks, 1, (cuuint32_t)pl.np};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = ctx->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, pl.base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TDNNF_ERR_CUDA, "cuTensorMapEncodeTiled (MN-major) failed with CUresult " + std::to_string((int)r));
  return TDNNF_OK;
}

// Cache lookup / insertion for planes of a registered source (see tdnnf_ctx_operand_cache_begin).
static tdnnf_ctx::PlaneCacheEntry make_key(int kind, const float* src, int R, int D, long long ld, int r, int groups,
                                           int c_row_mul, int c_col_mul, const float* scale, int Q, int pitch,
                                           const int32_t* offs, int fp16) {
  tdnnf_ctx::PlaneCacheEntry k;
  memset(&k, 0, sizeof(k));
  k.kind = kind;
  k.fp16 = fp16;
  k.src = src;
  k.R = R;
  k.D = D;
  k.ld = ld;
  k.r = r;
  k.groups = groups;
  k.c_row_mul = groups > 1 ? c_row_mul : 0;  // group index is always 0: the multipliers do not matter
  k.c_col_mul = groups > 1 ? c_col_mul : 0;
  k.scale = scale;
  k.Q = Q;
  k.pitch = pitch;
  for (int i = 0; i < kMaxSeg; ++i) k.offs[i] = (offs && i < groups) ? offs[i] : 0;
  return k;
}

static tdnnf_ctx::PlaneCacheEntry* cache_find(tdnnf_ctx* ctx, const tdnnf_ctx::PlaneCacheEntry& k) {
  for (auto& e : ctx->cache) {
    if (e.kind == k.kind && e.fp16 == k.fp16 && e.src == k.src && e.R == k.R && e.D == k.D && e.ld == k.ld && e.r == k.r &&
        e.groups == k.groups && e.c_row_mul == k.c_row_mul && e.c_col_mul == k.c_col_mul && e.scale == k.scale &&
        e.Q == k.Q && e.pitch == k.pitch && memcmp(e.offs, k.offs, sizeof(k.offs)) == 0)
      return &e;
  }
  return nullptr;
}

static void cache_store(tdnnf_ctx* ctx, tdnnf_ctx::PlaneCacheEntry k, const Planes& pl) {
  k.np = pl.np;
  k.base = pl.base;
  k.plane_elems = pl.plane_elems;
  if (tdnnf_ctx::PlaneCacheEntry* e = cache_find(ctx, k)) *e = k;
  else ctx->cache.push_back(k);
}

// Device-resident max |x| of a source matrix, for the power-of-two scaling of single-plane fp16 operands.
// Registered sources of the open cache scope have a slot that a row split fills as a by-product; anything else
// gets a scratch slot and one absmax_kernel pass.
static int get_absmax(tdnnf_ctx* ctx, const float* src, int R, int D, long long ld, const float** out) {
  const int idx = ctx->cache_source_index(src);
  float* slot = nullptr;
  if (idx >= 0) {
    if (ctx->absmax_valid[idx]) {
      *out = ctx->absmax_dev + idx;
      return TDNNF_OK;
    }
    slot = ctx->absmax_dev + idx;  // zeroed by tdnnf_ctx_operand_cache_begin
    ctx->absmax_valid[idx] = true;
  } else {
    slot = static_cast<float*>(ctx->ws_alloc(1024));
    if (!slot) return TDNNF_ERR_NOMEM;
    TDNNF_CUDA_OK(cudaMemsetAsync(slot, 0, sizeof(float), ctx->stream));
  }
  const long long total = (long long)R * D;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 1023) / 1024, (long long)ctx->num_sms * 8));
  TDNNF_CUDA_OK(launch_pdl(absmax_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, src, R, D, ld, slot));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  *out = slot;
  return TDNNF_OK;
}

// np: 0 = the context's bf16 plane count (2 or 3); 1 = one fp16 plane scaled by pow2_scale(*absmax_in) (null: unscaled)
static int launch_split_rows(tdnnf_ctx* ctx, const float* src, int R, int D, long long ld, int r, int groups,
                             int c_row_mul, int c_col_mul, const float* scale, int Q, int Kpad, Planes* pl, int np = 0,
                             const float* absmax_in = nullptr) {
  pl->plane_elems = (long long)groups * Q * Kpad;
  pl->K = Kpad;
  pl->Kpitch = Kpad;
  pl->rows = Q;
  pl->groups = groups;
  pl->np = np ? np : ctx->gemm_planes;
  const int fp16 = pl->np == 1;
  // planes that outlive this call (kept from Propagate, or written by the producer of the matrix): no split at all
  if (scale == nullptr && !fp16 && !ctx->grad_fast && groups == r && (groups == 1 || (c_row_mul == 1 && c_col_mul == 0))) {
    const tdnnf_planes* a = ctx->find_attached(src, R, D, ld, r);
    if (a && a->np >= pl->np && a->Q == Q && a->Kpad == Kpad) {
      pl->base = static_cast<__nv_bfloat16*>(a->base);
      pl->plane_elems = a->plane_elems;
      ctx->cache_hits++;
      return TDNNF_OK;
    }
  }
  const bool cacheable = ctx->cache_registered(src);
  tdnnf_ctx::PlaneCacheEntry key;
  if (cacheable) {
    key = make_key(0, src, R, D, ld, r, groups, c_row_mul, c_col_mul, scale, Q, Kpad, nullptr, fp16);
    const tdnnf_ctx::PlaneCacheEntry* e = cache_find(ctx, key);
    if (e && e->np >= pl->np) {  // the first np planes of a 3-plane split ARE the 2-plane split
      pl->base = static_cast<__nv_bfloat16*>(e->base);
      ctx->cache_hits++;
      return TDNNF_OK;
    }
    ctx->cache_misses++;
  }
  // The first row split of a registered source also produces its per-row sums of squares (for the natural-gradient
  // trace) and, in fast-gradient mode, its absmax: every element is read here anyway.
  float* absmax_out = nullptr;
  float* rowsq = nullptr;
  const int sidx = ctx->cache_source_index(src);
  if (sidx >= 0 && scale == nullptr && !fp16) {
    if (ctx->grad_fast && !ctx->absmax_valid[sidx]) {
      absmax_out = ctx->absmax_dev + sidx;
      ctx->absmax_valid[sidx] = true;
    }
    if (!ctx->rowsq_valid[sidx]) {
      if (ctx->rowsq_cap[sidx] < (size_t)R) {
        // grow-only; cudaFree waits for earlier kernels that may still read the old array
        if (ctx->rowsq_dev[sidx]) TDNNF_CUDA_OK(cudaFree(ctx->rowsq_dev[sidx]));
        ctx->rowsq_dev[sidx] = nullptr;
        ctx->rowsq_cap[sidx] = 0;
        TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ctx->rowsq_dev[sidx]), sizeof(float) * (size_t)R * 2));
        ctx->rowsq_cap[sidx] = (size_t)R * 2;
      }
      rowsq = ctx->rowsq_dev[sidx];
      TDNNF_CUDA_OK(cudaMemsetAsync(rowsq, 0, sizeof(float) * (size_t)R, ctx->stream));
      ctx->rowsq_valid[sidx] = true;
      ctx->rowsq_rows[sidx] = R;
    }
  }
  const size_t bytes = planes_bytes(pl->np, groups, Q, Kpad);
  pl->base = static_cast<__nv_bfloat16*>(cacheable ? ctx->cws_alloc(bytes) : ctx->ws_alloc(bytes));
  if (!pl->base) return TDNNF_ERR_NOMEM;
  const long long total = (long long)groups * Q * (Kpad >> 3);
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)ctx->num_sms * 16);
  TDNNF_CUDA_OK(launch_pdl(split_rows_kernel, dim3(std::max(blocks, 1)), dim3(threads), 0, ctx->stream, 1, src, R, D, ld, r, groups,
                           c_row_mul, c_col_mul, scale, Q, Kpad, pl->base, pl->base + pl->plane_elems,
                           pl->np == 3 ? pl->base + 2 * pl->plane_elems : (__nv_bfloat16*)nullptr, fp16, absmax_in, absmax_out,
                           rowsq));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  if (cacheable) cache_store(ctx, key, *pl);
  return TDNNF_OK;
}

static int launch_split_transpose(tdnnf_ctx* ctx, const float* src, int R, int J, long long ld, int r, int groups,
                                  int c_row_mul, int c_col_mul, const float* scale, int Q, int Qp, int Kvalid,
                                  Planes* pl, const int32_t* group_row_offsets = nullptr, float* colsum = nullptr,
                                  float colsum_scale = 0.f, int np = 0, const float* absmax_in = nullptr) {
  pl->plane_elems = (long long)groups * J * Qp;
  pl->K = Kvalid;
  pl->Kpitch = Qp;
  pl->rows = J;
  pl->groups = groups;
  pl->np = np ? np : ctx->gemm_planes;
  const int fp16 = pl->np == 1;
  const bool cacheable = ctx->cache_registered(src);
  tdnnf_ctx::PlaneCacheEntry key;
  if (cacheable) {
    key = make_key(1, src, R, J, ld, r, groups, c_row_mul, c_col_mul, scale, Q, Qp, group_row_offsets, fp16);
    const tdnnf_ctx::PlaneCacheEntry* e = cache_find(ctx, key);
    // a request that also wants the fused column sums must run the kernel
    if (colsum == nullptr && e && e->np >= pl->np) {
      pl->base = static_cast<__nv_bfloat16*>(e->base);
      ctx->cache_hits++;
      return TDNNF_OK;
    }
    ctx->cache_misses++;
  }
  const size_t bytes = planes_bytes(pl->np, groups, J, Qp);
  pl->base = static_cast<__nv_bfloat16*>(cacheable ? ctx->cws_alloc(bytes) : ctx->ws_alloc(bytes));
  if (!pl->base) return TDNNF_ERR_NOMEM;
  dim3 grid(ceil_div(Qp, 64), ceil_div(J, 64), groups), block(256);
  GroupRowOffsets gro;
  for (int i = 0; i < kMaxSeg; ++i) gro.v[i] = (group_row_offsets && i < groups) ? group_row_offsets[i] : 0;
  TDNNF_CUDA_OK(launch_pdl(split_transpose_kernel, grid, block, 0, ctx->stream, 1, src, R, J, ld, r, c_row_mul, c_col_mul, gro, scale,
                           Q, Qp, pl->base, pl->base + pl->plane_elems,
                           pl->np == 3 ? pl->base + 2 * pl->plane_elems : (__nv_bfloat16*)nullptr, colsum, colsum_scale, fp16,
                           absmax_in));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  if (cacheable) cache_store(ctx, key, *pl);
  return TDNNF_OK;
}

// Split-K factor: fill the SMs in whole waves without starving a unit of K iterations.
static int choose_splits(int tiles, int iters_per_tile, int num_sms, int min_iters = 8) {
  // Skinny outputs (the natural-gradient Gram matrices H^T H, J J^T: one or two tiles, K = all rows): even 8 splits
  // leave most SMs idle, and the epilogue of an r x r tile is nothing, so spread K over every SM.  (Measured
  // before: 56 single-CTA launches of 130-175 us each per step.)
  if (tiles * 8 < num_sms) return std::max(1, std::min(num_sms / tiles, iters_per_tile / 2));
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && iters_per_tile / s < min_iters) break;
    const double units = (double)tiles * s;
    const double waves = std::ceil(units / num_sms);
    const double eff = units / (waves * num_sms);
    const double score = eff - 0.012 * (s - 1);
    if (score > best_score + 1e-9) {
      best_score = score;
      best = s;
    }
  }
  return best;
}

// CTA-pair kernels (splice_gemm.cuh, PAIR): K-major two-plane GEMMs with BN = 160 / 256 and enough m-tiles.
// TDNNF_GEMM_PAIR=0 turns them off (A/B experiments).
static bool pair_eligible(const tdnnf_ctx* ctx, int bn, int m_tiles, int np, bool mn) {
  static const int on = [] {
    const char* e = getenv("TDNNF_GEMM_PAIR");
    return e ? atoi(e) : 1;
  }();
  return on && !mn && np == 2 && (bn == 160 || bn == 256) && m_tiles >= 4 && ctx->num_sms >= 2;
}

// Sets p->pair and p->splits: `other_tiles` = n-tiles x groups, `iters` = K iterations of one tile.
static void plan_units(const tdnnf_ctx* ctx, GemmParams* p, int bn, int np, bool mn, int other_tiles, int iters) {
  p->pair = pair_eligible(ctx, bn, p->m_tiles, np, mn) ? 1 : 0;
  // A 128 x 256 tile takes ~13 us to leave through red.global.add (the L2 atomic units, measured): a split that leaves a
  // unit fewer than 16 K blocks (~13 us of MMAs) makes the kernel epilogue-bound -- 92 against ~80 us for the data gradient
  // of the linear layers at 2 splits -- so wide tiles split only when K is long.
  const int min_iters = bn >= 256 ? 16 : 8;
  if (p->pair) p->splits = choose_splits((p->m_tiles + 1) / 2 * other_tiles, iters, ctx->num_sms / 2, min_iters);
  else p->splits = choose_splits(p->m_tiles * other_tiles, iters, ctx->num_sms, min_iters);
}

template <int BN, int NPA, int NPB, bool MN = false, bool PAIR = false>
static int launch_gemm_bn(tdnnf_ctx* ctx, const Planes& A, const Planes& B, const GemmParams& p, double algorithmic_flops) {
  using Cfg = GemmCfg<BN, NPA, NPB, MN, PAIR>;
  CUtensorMap tmA, tmB;
  int rc = MN ? make_map_mn(ctx, A, kBM / 64, Cfg::kBKk, &tmA) : make_map(ctx, A, kBM, &tmA);
  if (rc) return rc;
  rc = MN ? make_map_mn(ctx, B, Cfg::kBNs / 64, Cfg::kBKk, &tmB) : make_map(ctx, B, Cfg::kBNs, &tmB);
  if (rc) return rc;
  auto kern = splice_gemm_kernel<BN, NPA, NPB, MN, PAIR>;
  static bool attr_set = false;  // per template instance
  if (!attr_set) {
    TDNNF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int units = p.c_tiles * (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.splits;
  if (units <= 0) return TDNNF_OK;
  const int grid = PAIR ? 2 * std::min(units, ctx->num_sms / 2) : std::min(units, ctx->num_sms);
  tdnnf_ctx::GemmTiming tm;
  if (ctx->gemm_timing) {
    TDNNF_CUDA_OK(cudaEventCreate(&tm.start));
    TDNNF_CUDA_OK(cudaEventCreate(&tm.stop));
    tm.flops = algorithmic_flops;
    tm.products = (NPA == 3) ? 6 : (NPA == 2 && NPB == 2 ? 3 : NPA * NPB);
    TDNNF_CUDA_OK(cudaEventRecord(tm.start, ctx->stream));
  }
  TDNNF_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, ctx->stream, PAIR ? 2 : 1, tmA, tmB, p));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  if (ctx->gemm_timing) {
    TDNNF_CUDA_OK(cudaEventRecord(tm.stop, ctx->stream));
    ctx->gemm_events.push_back(tm);
  }
  return TDNNF_OK;
}

static int pick_bn(int n, int np = 2) {
  if (np == 3) {  // six products: keep two smem stages (BN <= 160)
    if (n <= 32) return 32;
    if (n <= 64) return 64;
    if (n % 160 == 0 || (n > 128 && n <= 160)) return 160;
    return 128;
  }
  static const int wide = [] {  // experiment knob: TDNNF_BN_WIDE=128|256 for outputs that are multiples of 256
    const char* e = getenv("TDNNF_BN_WIDE");
    return e ? atoi(e) : 256;
  }();
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  if (n % 160 == 0) return 160;
  if (n % 256 == 0 && wide == 256) return 256;
  if (n % 128 == 0 || n > 160) return 128;
  if (n <= 128) return 128;
  return 160;
}

template <int NPA, int NPB>
static int launch_gemm_np(tdnnf_ctx* ctx, int bn, const Planes& A, const Planes& B, const GemmParams& p, double fl) {
  switch (bn) {
    case 32: return launch_gemm_bn<32, NPA, NPB>(ctx, A, B, p, fl);
    case 64: return launch_gemm_bn<64, NPA, NPB>(ctx, A, B, p, fl);
    case 128: return launch_gemm_bn<128, NPA, NPB>(ctx, A, B, p, fl);
    case 160:
      if constexpr (NPA == 2)
        if (p.pair) return launch_gemm_bn<160, NPA, NPB, false, true>(ctx, A, B, p, fl);
      return launch_gemm_bn<160, NPA, NPB>(ctx, A, B, p, fl);
    case 256:
      if constexpr (NPA == 3) return fail(TDNNF_ERR_INVALID, "unsupported BN for 3-plane operands");
      else {
        if constexpr (NPA == 2)
          if (p.pair) return launch_gemm_bn<256, NPA, NPB, false, true>(ctx, A, B, p, fl);
        return launch_gemm_bn<256, NPA, NPB>(ctx, A, B, p, fl);
      }
    default: return fail(TDNNF_ERR_INVALID, "unsupported BN");
  }
}

static int launch_gemm(tdnnf_ctx* ctx, int bn, const Planes& A, const Planes& B, const GemmParams& p, double algorithmic_flops) {
  if (A.np == 3 && B.np == 3) return launch_gemm_np<3, 3>(ctx, bn, A, B, p, algorithmic_flops);
  if (A.np == 2 && B.np == 2) return launch_gemm_np<2, 2>(ctx, bn, A, B, p, algorithmic_flops);
  if (A.np == 1 && B.np == 1) return launch_gemm_np<1, 1>(ctx, bn, A, B, p, algorithmic_flops);
  return fail(TDNNF_ERR_INVALID, "unsupported operand plane combination");
}

// Both operands MN-major (row planes, contraction over rows): the parameter gradient.
static int launch_gemm_mn(tdnnf_ctx* ctx, int bn, const Planes& A, const Planes& B, const GemmParams& p, double fl) {
  if (A.np == 2 && B.np == 2) {
    switch (bn) {
      case 32: return launch_gemm_bn<32, 2, 2, true>(ctx, A, B, p, fl);
      case 64: return launch_gemm_bn<64, 2, 2, true>(ctx, A, B, p, fl);
      case 128: return launch_gemm_bn<128, 2, 2, true>(ctx, A, B, p, fl);
      case 160: return launch_gemm_bn<160, 2, 2, true>(ctx, A, B, p, fl);
      case 256: return launch_gemm_bn<256, 2, 2, true>(ctx, A, B, p, fl);
      default: break;
    }
  } else if (A.np == 3 && B.np == 3) {
    switch (bn) {
      case 32: return launch_gemm_bn<32, 3, 3, true>(ctx, A, B, p, fl);
      case 64: return launch_gemm_bn<64, 3, 3, true>(ctx, A, B, p, fl);
      case 128: return launch_gemm_bn<128, 3, 3, true>(ctx, A, B, p, fl);
      case 160: return launch_gemm_bn<160, 3, 3, true>(ctx, A, B, p, fl);
      default: break;
    }
  }
  return fail(TDNNF_ERR_INVALID, "unsupported MN-major GEMM configuration");
}

static int check_offsets(int n, const int32_t* row_offsets, int row_stride, int out_rows, int in_rows) {
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS, "number of time offsets must be in [1,16]");
  TDNNF_REQUIRE(row_stride >= 1 && row_stride <= kMaxSeg, "row_stride must be in [1,16]");
  for (int i = 0; i < n; ++i) {
    // same condition as the KALDI_ASSERT in GetInputPart (ref: tdnn.cc:811-813)
    TDNNF_REQUIRE(row_offsets[i] >= 0 &&
                      (long long)in_rows >= (long long)row_offsets[i] + (long long)row_stride * out_rows - (row_stride - 1),
                  "row offset / stride view does not fit in the input matrix");
  }
  return TDNNF_OK;
}

}  // namespace tdnnf

using namespace tdnnf;

extern "C" int tdnnf_darts_propagate(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride,
                                     float* out, int out_rows, int out_dim, int out_stride, const float* W,
                                     int w_stride, const float* bias, int bias_mode, const float* weff, int n,
                                     const int32_t* row_offsets, int row_stride) {
  TDNNF_REQUIRE(ctx && in && out && W && weff && row_offsets, "null argument");
  TDNNF_REQUIRE(in_rows > 0 && in_dim > 0 && out_rows > 0 && out_dim > 0, "empty matrix");
  TDNNF_REQUIRE(in_stride >= in_dim && out_stride >= out_dim && w_stride >= n * in_dim, "stride < cols");
  TDNNF_REQUIRE(bias_mode >= 0 && bias_mode <= 2 && (bias_mode != 2 || bias), "bad bias_mode");
  int rc = check_offsets(n, row_offsets, row_stride, out_rows, in_rows);
  if (rc) return rc;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));

  const int r = row_stride;
  const int Q = ceil_div(in_rows, r);
  const int Kpad = round_up(in_dim, kBK);
  ctx->ws_reset();
  rc = ctx->ws_reserve(planes_bytes(ctx->gemm_planes, r, Q, Kpad) + planes_bytes(ctx->gemm_planes, n, out_dim, Kpad));
  if (rc) return rc;
  rc = ctx->cws_reserve(planes_bytes(ctx->gemm_planes, r, Q, Kpad) + planes_bytes(ctx->gemm_planes, n, out_dim, Kpad));
  if (rc) return rc;
  Planes A, B;
  rc = launch_split_rows(ctx, in, in_rows, in_dim, in_stride, r, r, 1, 0, nullptr, Q, Kpad, &A);
  if (rc) return rc;
  rc = launch_split_rows(ctx, W, out_dim, in_dim, w_stride, 1, n, 0, in_dim, weff, out_dim, Kpad, &B);
  if (rc) return rc;

  const int bn = pick_bn(out_dim, ctx->gemm_planes);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = ceil_div(out_rows, kBM);
  p.n_tiles = ceil_div(out_dim, bn);
  p.c_tiles = 1;
  p.kb_per_seg = Kpad / kBK;
  p.kb_last_steps = ceil_div(in_dim - (p.kb_per_seg - 1) * kBK, 16);
  p.nseg = n;
  for (int i = 0; i < n; ++i) {
    p.seg_a_m[i] = row_offsets[i] / r;
    p.seg_a_c[i] = row_offsets[i] % r;
    p.seg_b_c[i] = i;
    p.seg_cmatch[i] = -1;
  }
  p.m_valid[0] = out_rows;
  p.n_valid = out_dim;
  p.seg_weight = weff;
  plan_units(ctx, &p, bn, A.np == B.np ? A.np : 0, false, p.n_tiles, n * p.kb_per_seg);
  p.out = out;
  p.out_ld = out_stride;
  p.row_mul = 1;
  p.alpha = 1.0f;
  if (p.splits > 1) {
    if (bias_mode != 0) {
      const long long total = (long long)out_rows * out_dim;
      const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16);
      TDNNF_CUDA_OK(launch_pdl(init_rows_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, out, out_rows, out_dim, out_stride,
                               bias_mode == 2 ? bias : (const float*)nullptr));
      ctx->launches++;
      TDNNF_CUDA_OK(cudaGetLastError());
    }
    p.accumulate = 1;
    p.atomic = 1;
  } else {
    p.accumulate = (bias_mode == 0);
    p.bias = (bias_mode == 2) ? bias : nullptr;
  }
  // all n offsets counted: in uniform-sample mode the skipped offsets make this an upper bound
  return launch_gemm(ctx, bn, A, B, p, 2.0 * out_rows * (double)out_dim * in_dim * n);
}

// out = [w_1 X_1 | ... | w_n X_n | 1] W^T for a SKINNY W (rank rows): the H = X W_t^T product of OnlineNaturalGradient
// on the spliced input.  tdnnf_darts_propagate would stream the activation planes once per offset (n x the operand
// through L2 into a 32-column tile: bandwidth-bound); here ONE un-spliced GEMM Y = X [W_1^T | ... | W_n^T]
// (in_rows x n*rank, every activation tile fetched once) is followed by a gather-sum over the offsets.
extern "C" int tdnnf_darts_project(tdnnf_ctx* ctx, const float* in, int in_rows, int in_dim, int in_stride, float* out,
                                   int out_rows, int rank, int out_stride, const float* W, int w_stride, const float* bias,
                                   const float* weff, int n, const int32_t* row_offsets, int row_stride) {
  TDNNF_REQUIRE(ctx && in && out && W && weff && row_offsets, "null argument");
  TDNNF_REQUIRE(in_rows > 0 && in_dim > 0 && out_rows > 0 && rank > 0, "empty matrix");
  TDNNF_REQUIRE(in_stride >= in_dim && out_stride >= rank && w_stride >= n * in_dim, "stride < cols");
  int rc = check_offsets(n, row_offsets, row_stride, out_rows, in_rows);
  if (rc) return rc;
  const int ncols = n * rank;
  if (ncols > 256)  // wider than one N tile: the spliced GEMM reads the operand as often
    return tdnnf_darts_propagate(ctx, in, in_rows, in_dim, in_stride, out, out_rows, rank, out_stride, W, w_stride, bias,
                                 bias ? 2 : 1, weff, n, row_offsets, row_stride);
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const int r = row_stride;
  const int Q = ceil_div(in_rows, r);
  const int Kpad = round_up(in_dim, kBK);
  const size_t y_bytes = ((size_t)in_rows * ncols * sizeof(float) + 1023) & ~size_t(1023);
  const size_t need = planes_bytes(ctx->gemm_planes, r, Q, Kpad) + planes_bytes(ctx->gemm_planes, n, rank, Kpad);
  ctx->ws_reset();
  rc = ctx->ws_reserve(need + y_bytes);
  if (rc) return rc;
  rc = ctx->cws_reserve(need);
  if (rc) return rc;
  Planes A, B;
  rc = launch_split_rows(ctx, in, in_rows, in_dim, in_stride, r, r, 1, 0, nullptr, Q, Kpad, &A);
  if (rc) return rc;
  rc = launch_split_rows(ctx, W, rank, in_dim, w_stride, 1, n, 0, in_dim, weff, rank, Kpad, &B);
  if (rc) return rc;
  B.rows = ncols;  // [offset][rank rows][K] is contiguous: one group of n*rank rows
  B.groups = 1;
  float* Y = static_cast<float*>(ctx->ws_alloc(y_bytes));
  if (!Y) return TDNNF_ERR_NOMEM;

  const int bn = pick_bn(ncols, ctx->gemm_planes);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = ceil_div(Q, kBM);
  p.n_tiles = ceil_div(ncols, bn);
  p.c_tiles = r;
  p.kb_per_seg = Kpad / kBK;
  p.kb_last_steps = ceil_div(in_dim - (p.kb_per_seg - 1) * kBK, 16);
  p.nseg = r;
  for (int c = 0; c < r; ++c) {
    p.seg_a_c[c] = c;
    p.seg_cmatch[c] = c;
    p.m_valid[c] = (in_rows - c + r - 1) / r;
  }
  p.n_valid = ncols;
  plan_units(ctx, &p, bn, A.np == B.np ? A.np : 0, false, p.n_tiles * r, p.kb_per_seg);
  p.out = Y;
  p.out_ld = ncols;
  p.row_mul = r;
  p.row_cadd = 1;
  p.alpha = 1.0f;
  if (p.splits > 1) {
    { int zrc = zero_async(ctx, Y, (size_t)in_rows * ncols * sizeof(float)); if (zrc) return zrc; }
    p.accumulate = 1;
    p.atomic = 1;
  }
  rc = launch_gemm(ctx, bn, A, B, p, 2.0 * in_rows * (double)ncols * in_dim);
  if (rc) return rc;
  GroupRowOffsets offs;
  for (int i = 0; i < kMaxSeg; ++i) offs.v[i] = i < n ? row_offsets[i] : 0;
  const long long total = (long long)out_rows * rank;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16));
  TDNNF_CUDA_OK(launch_pdl(project_gather_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, Y, ncols, out_rows, rank, n, offs, r, bias, out, out_stride));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_darts_backprop_data(tdnnf_ctx* ctx, const float* out_deriv, int out_rows, int out_dim,
                                         int od_stride, float* in_deriv, int in_rows, int in_dim, int id_stride,
                                         const float* W, int w_stride, const float* weff, int n,
                                         const int32_t* row_offsets, int row_stride) {
  TDNNF_REQUIRE(ctx && out_deriv && in_deriv && W && weff && row_offsets, "null argument");
  TDNNF_REQUIRE(in_rows > 0 && in_dim > 0 && out_rows > 0 && out_dim > 0, "empty matrix");
  TDNNF_REQUIRE(id_stride >= in_dim && od_stride >= out_dim && w_stride >= n * in_dim, "stride < cols");
  int rc = check_offsets(n, row_offsets, row_stride, out_rows, in_rows);
  if (rc) return rc;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));

  const int r = row_stride;
  const int Kpad = round_up(out_dim, kBK);
  ctx->ws_reset();
  rc = ctx->ws_reserve(planes_bytes(ctx->gemm_planes, 1, out_rows, Kpad) + planes_bytes(ctx->gemm_planes, n, in_dim, Kpad));
  if (rc) return rc;
  rc = ctx->cws_reserve(planes_bytes(ctx->gemm_planes, 1, out_rows, Kpad) + planes_bytes(ctx->gemm_planes, n, in_dim, Kpad));
  if (rc) return rc;
  Planes A, B;
  rc = launch_split_rows(ctx, out_deriv, out_rows, out_dim, od_stride, 1, 1, 0, 0, nullptr, out_rows, Kpad, &A);
  if (rc) return rc;
  // B planes: [i][d][o] = weff[i] * W[o][i*in_dim + d]
  rc = launch_split_transpose(ctx, W, out_dim, in_dim, w_stride, 1, n, 0, in_dim, weff, out_dim, Kpad, Kpad, &B);
  if (rc) return rc;

  const int bn = pick_bn(in_dim, ctx->gemm_planes);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int Qin = ceil_div(in_rows, r);
  p.m_tiles = ceil_div(Qin, kBM);
  p.n_tiles = ceil_div(in_dim, bn);
  p.c_tiles = r;
  p.kb_per_seg = Kpad / kBK;
  p.kb_last_steps = ceil_div(out_dim - (p.kb_per_seg - 1) * kBK, 16);
  p.nseg = n;
  for (int i = 0; i < n; ++i) {
    p.seg_a_m[i] = -(row_offsets[i] / r);
    p.seg_b_c[i] = i;
    p.seg_cmatch[i] = row_offsets[i] % r;
  }
  for (int c = 0; c < r; ++c) p.m_valid[c] = (in_rows - c + r - 1) / r;
  p.n_valid = in_dim;
  p.seg_weight = weff;
  plan_units(ctx, &p, bn, A.np == B.np ? A.np : 0, false, p.n_tiles * r, std::max(1, n / r) * p.kb_per_seg);
  p.out = in_deriv;
  p.out_ld = id_stride;
  p.row_mul = r;
  p.row_cadd = 1;
  p.accumulate = 1;
  p.atomic = p.splits > 1;
  p.alpha = 1.0f;
  return launch_gemm(ctx, bn, A, B, p, 2.0 * out_rows * (double)out_dim * in_dim * n);
}

extern "C" int tdnnf_darts_backprop_params(tdnnf_ctx* ctx, const float* in_value, int in_rows, int in_dim,
                                           int in_stride, const float* out_deriv, int out_rows, int out_dim,
                                           int od_stride, const float* W_model, int w_stride, float* dW,
                                           int dw_stride, float* dbias, const float* weff, int n,
                                           const int32_t* row_offsets, int row_stride, float lr, float* s) {
  TDNNF_REQUIRE(ctx && in_value && out_deriv && dW && weff && row_offsets, "null argument");
  TDNNF_REQUIRE(in_rows > 0 && in_dim > 0 && out_rows > 0 && out_dim > 0, "empty matrix");
  TDNNF_REQUIRE(in_stride >= in_dim && od_stride >= out_dim && dw_stride >= n * in_dim, "stride < cols");
  TDNNF_REQUIRE(s == nullptr || (W_model != nullptr && w_stride >= n * in_dim), "s requires W_model");
  int rc = check_offsets(n, row_offsets, row_stride, out_rows, in_rows);
  if (rc) return rc;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));

  const int r = row_stride;
  const int Q = ceil_div(in_rows, r);
  // ---- large operands: contract over the rows of the ROW planes (MN-major operands), which the data gradient and the
  // natural-gradient projections of the same Backprop have already built -- no transposed pre-pass, and the per-offset
  // shift is a row coordinate (no alignment condition).  TDNNF_WGRAD_MN=0 selects the transposed-plane path below.
  {
    static const int env_min_rows = [] {  // process-wide defaults from the environment; the context setting wins
      const char* off = getenv("TDNNF_WGRAD_MN");
      if (off && atoi(off) == 0) return -1;
      const char* e = getenv("TDNNF_WGRAD_MN_MIN_ROWS");
      return e ? atoi(e) : 0;
    }();
    const int mn_min_rows = env_min_rows != 0 ? env_min_rows : ctx->wgrad_mn_min_rows;
    const bool mn_enabled = mn_min_rows >= 0;
    auto waste0 = [](int x) { return (double)round_up(x, kBM) / x; };
    const bool m_is_in0 = waste0(in_dim) <= waste0(out_dim);
    const int n_dim = m_is_in0 ? out_dim : in_dim;
    const int bn0 = pick_bn(n_dim, ctx->gemm_planes);
    const bool fast0 = ctx->grad_fast && ctx->gemm_planes == 2;
    // n-tiles must start on a 64-column chunk; a narrow operand (the rank-20 / rank-80 H of the natural-gradient
    // corrections) is one tile whose chunk is zero-padded by the row split
    const bool tiles_ok = (bn0 % 64 == 0 || ceil_div(n_dim, bn0) == 1) && !(ctx->gemm_planes == 3 && bn0 == 256);
    if (mn_enabled && !fast0 && tiles_ok && out_rows >= mn_min_rows) {
      if (dbias) {  // bias gradient: lr * colsum(out_deriv) (the transposed pre-pass of the other path does it on the fly)
        const tdnnf_planes* pa = ctx->find_attached(out_deriv, out_rows, out_dim, od_stride, 1);
        if (pa && pa->has_colsum && !ctx->grad_fast)  // the producer of out_deriv summed its columns while writing it
          rc = tdnnf_mat_axpy(ctx, lr, pa->colsum, out_dim, dbias, out_dim, 1, out_dim);
        else
          rc = tdnnf_add_row_sum(ctx, out_deriv, out_rows, out_dim, od_stride, lr, dbias);
        if (rc) return rc;
      }
      // ---- stacked form for a NARROW out_deriv (the J = H^T X product of OnlineNaturalGradient on the spliced input:
      // rank 20 against n * 1536 columns).  The per-offset form reads the activation planes once per offset into a
      // 32-column tile (bandwidth-bound: 113 us per call).  Here the n shifted copies of out_deriv stand side by side,
      // Hs = [w_1 S_1 od | ... | w_n S_n od] (in_rows x n*out_dim), and ONE GEMM T = Hs^T X reads every activation tile
      // once; dW[:, i-th block] += lr * T[i-th block of rows].
      static const int stacked_on = [] {
        const char* e = getenv("TDNNF_WGRAD_STACKED");
        return e ? atoi(e) : 1;
      }();
      const int ncols = n * out_dim;
      const int bnS = pick_bn(ncols, ctx->gemm_planes);
      if (stacked_on && s == nullptr && n > 1 && ncols <= 256 && ceil_div(ncols, bnS) == 1 && !(ctx->gemm_planes == 3 && bnS == 256)) {
        const int KpX = round_up(in_dim, kBK), KpS = round_up(ncols, kBK);
        const size_t hs_bytes = ((size_t)in_rows * ncols * sizeof(float) + 1023) & ~size_t(1023);
        const size_t t_bytes = ((size_t)ncols * in_dim * sizeof(float) + 1023) & ~size_t(1023);
        const size_t need = planes_bytes(ctx->gemm_planes, r, Q, KpX) + planes_bytes(ctx->gemm_planes, r, Q, KpS);
        ctx->ws_reset();
        rc = ctx->ws_reserve(need + hs_bytes + t_bytes + 4096);
        if (rc) return rc;
        rc = ctx->cws_reserve(need);
        if (rc) return rc;
        float* Hs = static_cast<float*>(ctx->ws_alloc(hs_bytes));
        float* T = static_cast<float*>(ctx->ws_alloc(t_bytes));
        if (!Hs || !T) return TDNNF_ERR_NOMEM;
        GroupRowOffsets offs;
        for (int i = 0; i < kMaxSeg; ++i) offs.v[i] = i < n ? row_offsets[i] : 0;
        {
          const long long total = (long long)in_rows * ncols;
          const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16));
          TDNNF_CUDA_OK(launch_pdl(stack_shift_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, out_deriv, od_stride, out_rows, out_dim, n, offs, r, weff, in_rows, Hs,
                                                              ncols));
          ctx->launches++;
          TDNNF_CUDA_OK(cudaGetLastError());
        }
        Planes XR, HS;
        rc = launch_split_rows(ctx, in_value, in_rows, in_dim, in_stride, r, r, 1, 0, nullptr, Q, KpX, &XR);
        if (rc) return rc;
        rc = launch_split_rows(ctx, Hs, in_rows, ncols, ncols, r, r, 1, 0, nullptr, Q, KpS, &HS);
        if (rc) return rc;
        { int zrc = zero_async(ctx, T, (size_t)ncols * in_dim * sizeof(float)); if (zrc) return zrc; }
        constexpr int kBKmn = 32;
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.c_tiles = 1;
        p.kb_per_seg = ceil_div(Q, kBKmn);
        p.kb_last_steps = ceil_div(Q - (p.kb_per_seg - 1) * kBKmn, 16);
        p.nseg = r;  // one segment per row group of the planes (row_stride > 1: rows are de-interleaved)
        for (int c = 0; c < r; ++c) {
          p.seg_a_c[c] = c;
          p.seg_b_c[c] = c;
          p.seg_cmatch[c] = -1;
        }
        p.m_valid[0] = in_dim;
        p.m_tiles = ceil_div(in_dim, kBM);
        p.n_tiles = 1;
        p.n_valid = ncols;
        p.out = T;  // acc[d, c] -> T[c, d]
        p.out_ld = in_dim;
        p.transposed = 1;
        p.row_mul = 1;
        p.accumulate = 1;
        p.atomic = 1;
        p.alpha = 1.0f;
        p.splits = choose_splits(p.m_tiles, r * p.kb_per_seg, ctx->num_sms);
        rc = launch_gemm_mn(ctx, bnS, XR, HS, p, 2.0 * out_rows * (double)out_dim * in_dim * n);
        if (rc) return rc;
        const long long total = (long long)ncols * in_dim;
        const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16));
        TDNNF_CUDA_OK(launch_pdl(stack_scatter_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, T, in_dim, out_dim, n, lr, dW, dw_stride));
        ctx->launches++;
        TDNNF_CUDA_OK(cudaGetLastError());
        return TDNNF_OK;
      }
      const int KpX = round_up(in_dim, kBK), KpO = round_up(out_dim, kBK);
      ctx->ws_reset();
      const size_t need = planes_bytes(ctx->gemm_planes, r, Q, KpX) + planes_bytes(ctx->gemm_planes, 1, out_rows, KpO);
      rc = ctx->ws_reserve(need + 4096);
      if (rc) return rc;
      rc = ctx->cws_reserve(need);
      if (rc) return rc;
      Planes XR, ODR;  // the same splits as tdnnf_darts_propagate / _project (X) and tdnnf_darts_backprop_data (out_deriv)
      rc = launch_split_rows(ctx, in_value, in_rows, in_dim, in_stride, r, r, 1, 0, nullptr, Q, KpX, &XR);
      if (rc) return rc;
      rc = launch_split_rows(ctx, out_deriv, out_rows, out_dim, od_stride, 1, 1, 0, 0, nullptr, out_rows, KpO, &ODR);
      if (rc) return rc;
      if (s) { int zrc = zero_async(ctx, s, sizeof(float) * n); if (zrc) return zrc; }
      constexpr int kBKmn = 32;
      GemmParams p;
      memset(&p, 0, sizeof(p));
      p.c_tiles = n;
      p.kb_per_seg = ceil_div(out_rows, kBKmn);
      p.kb_last_steps = ceil_div(out_rows - (p.kb_per_seg - 1) * kBKmn, 16);
      p.nseg = n;
      p.seg_weight = weff;
      p.out = dW;
      p.out_ld = dw_stride;
      p.accumulate = 1;
      p.alpha = lr;
      p.c_scale = weff;
      p.dot_ref = s ? W_model : nullptr;
      p.dot_ld = w_stride;
      p.dot_out = s;
      for (int i = 0; i < n; ++i) {
        (m_is_in0 ? p.seg_a_k : p.seg_b_k)[i] = row_offsets[i] / r;
        (m_is_in0 ? p.seg_a_c : p.seg_b_c)[i] = row_offsets[i] % r;
        p.seg_cmatch[i] = i;
        p.m_valid[i] = m_is_in0 ? in_dim : out_dim;
      }
      p.row_mul = 1;
      if (m_is_in0) {  // acc[d, o] -> dW[o, i*in_dim + d] (transposed store)
        p.m_tiles = ceil_div(in_dim, kBM);
        p.n_tiles = ceil_div(out_dim, bn0);
        p.n_valid = out_dim;
        p.transposed = 1;
        p.row_cadd = in_dim;
      } else {         // acc[o, d] -> dW[o, i*in_dim + d]
        p.m_tiles = ceil_div(out_dim, kBM);
        p.n_tiles = ceil_div(in_dim, bn0);
        p.n_valid = in_dim;
        p.col_cadd = in_dim;
      }
      p.splits = choose_splits(p.m_tiles * p.n_tiles * n, p.kb_per_seg, ctx->num_sms);
      p.atomic = p.splits > 1;
      const double wflops0 = 2.0 * out_rows * (double)out_dim * in_dim * n;
      return m_is_in0 ? launch_gemm_mn(ctx, bn0, XR, ODR, p, wflops0) : launch_gemm_mn(ctx, bn0, ODR, XR, p, wflops0);
    }
  }
  const int Qp = round_up(Q, 8);
  const int Rp = round_up(out_rows, 8);
  // TMA needs the start of every box row 16-byte aligned: with K = the (de-interleaved) row index,
  // the per-offset K shift row_offsets[i]/r must be a multiple of 8 bf16.  True whenever the number
  // of sequences is a multiple of 8 (the recipes use 64/128); otherwise each offset gets its own
  // pre-shifted X^T plane group (n x the pre-pass traffic, same GEMM).
  bool shifts_aligned = true;
  for (int i = 0; i < n; ++i) shifts_aligned = shifts_aligned && ((row_offsets[i] / r) % 8 == 0);
  ctx->ws_reset();
  const size_t need = (shifts_aligned ? planes_bytes(ctx->gemm_planes, r, in_dim, Qp) : planes_bytes(ctx->gemm_planes, n, in_dim, Rp)) +
                      planes_bytes(ctx->gemm_planes, 1, out_dim, Rp);
  rc = ctx->ws_reserve(need + 4096);
  if (rc) return rc;
  rc = ctx->cws_reserve(need);
  if (rc) return rc;
  // Gradient mode "fast": both operands as ONE fp16 plane each, scaled into the fp16 range by a power of two
  // taken from their device-resident absmax (exact to undo in the epilogue): one product instead of three.
  const bool fast = ctx->grad_fast && ctx->gemm_planes == 2;
  const float *amax_x = nullptr, *amax_od = nullptr;
  if (fast) {
    rc = get_absmax(ctx, in_value, in_rows, in_dim, in_stride, &amax_x);
    if (rc) return rc;
    rc = get_absmax(ctx, out_deriv, out_rows, out_dim, od_stride, &amax_od);
    if (rc) return rc;
  }
  const int gnp = fast ? 1 : 0;
  Planes XT, ODT;
  if (shifts_aligned)
    rc = launch_split_transpose(ctx, in_value, in_rows, in_dim, in_stride, r, r, 1, 0, nullptr, Q, Qp, Q, &XT, nullptr,
                                nullptr, 0.f, gnp, amax_x);
  else
    rc = launch_split_transpose(ctx, in_value, in_rows, in_dim, in_stride, r, n, 0, 0, nullptr, out_rows, Rp,
                                out_rows, &XT, row_offsets, nullptr, 0.f, gnp, amax_x);
  if (rc) return rc;
  // the out_deriv^T pre-pass also accumulates dbias += lr * colsum(out_deriv) (each element is read exactly once)
  rc = launch_split_transpose(ctx, out_deriv, out_rows, out_dim, od_stride, 1, 1, 0, 0, nullptr, out_rows, Rp,
                              out_rows, &ODT, nullptr, dbias, lr, gnp, amax_od);
  if (rc) return rc;
  if (s) { int zrc = zero_async(ctx, s, sizeof(float) * n); if (zrc) return zrc; }

  auto waste = [](int x) { return (double)round_up(x, kBM) / x; };
  const bool m_is_in = waste(in_dim) <= waste(out_dim);  // which dimension rides the 128-row MMA M
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.c_tiles = n;
  p.kb_per_seg = ceil_div(out_rows, kBK);
  p.kb_last_steps = ceil_div(out_rows - (p.kb_per_seg - 1) * kBK, 16);
  p.nseg = n;
  p.seg_weight = weff;  // weff[i] == 0 (unsampled offset): X~ block is zero, nothing to add (ref: tdnn.cc:509-513)
  p.out = dW;
  p.out_ld = dw_stride;
  p.accumulate = 1;
  p.alpha = lr;
  p.c_scale = weff;
  p.dot_ref = s ? W_model : nullptr;
  p.dot_ld = w_stride;
  p.dot_out = s;
  p.absmax_a = m_is_in ? amax_x : amax_od;
  p.absmax_b = m_is_in ? amax_od : amax_x;
  int bn;
  if (m_is_in) {
    // acc[d, o] = sum_k X_i^T[d, k] * OD^T[o, k]  ->  dW[o, i*in_dim + d]   (transposed store)
    bn = pick_bn(out_dim, ctx->gemm_planes);
    p.m_tiles = ceil_div(in_dim, kBM);
    p.n_tiles = ceil_div(out_dim, bn);
    for (int i = 0; i < n; ++i) {
      p.seg_a_k[i] = shifts_aligned ? row_offsets[i] / r : 0;
      p.seg_a_c[i] = shifts_aligned ? row_offsets[i] % r : i;
      p.seg_cmatch[i] = i;
      p.m_valid[i] = in_dim;
    }
    p.n_valid = out_dim;
    p.transposed = 1;
    p.row_mul = 1;
    p.row_cadd = in_dim;
  } else {
    // acc[o, d] = sum_k OD^T[o, k] * X_i^T[d, k]  ->  dW[o, i*in_dim + d]   (row-major store)
    bn = pick_bn(in_dim, ctx->gemm_planes);
    p.m_tiles = ceil_div(out_dim, kBM);
    p.n_tiles = ceil_div(in_dim, bn);
    for (int i = 0; i < n; ++i) {
      p.seg_b_k[i] = shifts_aligned ? row_offsets[i] / r : 0;
      p.seg_b_c[i] = shifts_aligned ? row_offsets[i] % r : i;
      p.seg_cmatch[i] = i;
      p.m_valid[i] = out_dim;
    }
    p.n_valid = in_dim;
    p.row_mul = 1;
    p.col_cadd = in_dim;
  }
  p.splits = choose_splits(p.m_tiles * p.n_tiles * n, p.kb_per_seg, ctx->num_sms);
  p.atomic = p.splits > 1;
  const double wflops = 2.0 * out_rows * (double)out_dim * in_dim * n;
  rc = m_is_in ? launch_gemm(ctx, bn, XT, ODT, p, wflops) : launch_gemm(ctx, bn, ODT, XT, p, wflops);
  if (rc) return rc;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_set_gemm_planes(tdnnf_ctx* ctx, int planes) {
  TDNNF_REQUIRE(ctx && (planes == 2 || planes == 3), "planes must be 2 or 3");
  ctx->gemm_planes = planes;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_set_wgrad_mn_min_rows(tdnnf_ctx* ctx, int min_rows) {
  TDNNF_REQUIRE(ctx, "null context");
  ctx->wgrad_mn_min_rows = min_rows;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_set_gradient_mode(tdnnf_ctx* ctx, int fast) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  ctx->grad_fast = fast != 0;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_operand_rowsq(tdnnf_ctx* ctx, const float* source, int rows, const float** rowsq) {
  TDNNF_REQUIRE(ctx && source && rowsq, "null argument");
  *rowsq = nullptr;
  const int idx = ctx->cache_source_index(source);
  if (idx >= 0 && ctx->rowsq_valid[idx] && ctx->rowsq_rows[idx] == rows) *rowsq = ctx->rowsq_dev[idx];
  if (*rowsq == nullptr && !ctx->grad_fast)
    for (const tdnnf_planes* p : ctx->attached_planes)
      if (p->src == source && p->R == rows && p->rowsq) *rowsq = p->rowsq;
  return TDNNF_OK;
}

// ------------------------------------------------------------------------------------------ tdnnf_planes_*
static int planes_new(tdnnf_ctx* ctx, const float* src, int rows, int cols, int stride, int row_stride, tdnnf_planes** out) {
  const int r = row_stride, Q = ceil_div(rows, r), Kpad = round_up(cols, kBK);
  const size_t pb = planes_bytes(2, r, Q, Kpad);
  const size_t rb = ((size_t)rows * sizeof(float) + 255) & ~size_t(255), cb = ((size_t)cols * sizeof(float) + 255) & ~size_t(255);
  const size_t bytes = pb + rb + cb;
  void* block = nullptr;
  for (size_t i = 0; i < ctx->planes_pool.size(); ++i) {
    if (ctx->planes_pool[i].first == bytes) {
      block = ctx->planes_pool[i].second;
      ctx->planes_pool.erase(ctx->planes_pool.begin() + i);
      break;
    }
  }
  if (!block) {
    cudaError_t e = cudaMalloc(&block, bytes);
    if (e != cudaSuccess) return fail(TDNNF_ERR_NOMEM, std::string("cudaMalloc of operand planes failed: ") + cudaGetErrorString(e));
  }
  tdnnf_planes* p = new tdnnf_planes();
  p->ctx = ctx;
  p->src = src;
  p->R = rows;
  p->D = cols;
  p->ld = stride;
  p->r = r;
  p->Q = Q;
  p->Kpad = Kpad;
  p->np = 2;
  p->block = block;
  p->bytes = bytes;
  p->base = block;
  p->plane_elems = (long long)r * Q * Kpad;
  p->rowsq = reinterpret_cast<float*>(static_cast<char*>(block) + pb);
  p->colsum = reinterpret_cast<float*>(static_cast<char*>(block) + pb + rb);
  *out = p;
  return TDNNF_OK;
}

extern "C" int tdnnf_planes_acquire(tdnnf_ctx* ctx, const float* src, int rows, int cols, int stride, int row_stride,
                                    tdnnf_planes** out) {
  TDNNF_REQUIRE(ctx && src && out, "null argument");
  TDNNF_REQUIRE(rows > 0 && cols > 0 && stride >= cols && row_stride >= 1 && row_stride <= kMaxSeg, "bad matrix");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  if (const tdnnf_planes* a = ctx->find_attached(src, rows, cols, stride, row_stride)) {
    tdnnf_planes* p = const_cast<tdnnf_planes*>(a);
    p->refs++;
    *out = p;
    return TDNNF_OK;
  }
  tdnnf_planes* p = nullptr;
  int rc = planes_new(ctx, src, rows, cols, stride, row_stride, &p);
  if (rc) return rc;
  TDNNF_CUDA_OK(cudaMemsetAsync(p->rowsq, 0, sizeof(float) * (size_t)rows, ctx->stream));
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(p->base);
  const long long total = (long long)p->r * p->Q * (p->Kpad >> 3);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16));
  split_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(src, rows, cols, stride, p->r, p->r, 1, 0, nullptr, p->Q, p->Kpad, hi,
                                                     hi + p->plane_elems, nullptr, 0, nullptr, nullptr, p->rowsq);
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  *out = p;
  return TDNNF_OK;
}

// Planes for a matrix whose PRODUCER fills them (the fused tail kernels): nothing is launched here.
int tdnnf::planes_alloc_for_producer(tdnnf_ctx* ctx, const float* src, int rows, int cols, int stride, tdnnf_planes** out) {
  return planes_new(ctx, src, rows, cols, stride, 1, out);
}

extern "C" int tdnnf_planes_release(tdnnf_planes* p) {
  if (!p) return TDNNF_OK;
  if (--p->refs > 0) return TDNNF_OK;
  tdnnf_ctx* ctx = p->ctx;
  for (size_t i = 0; i < ctx->attached_planes.size(); ++i)
    if (ctx->attached_planes[i] == p) {
      ctx->attached_planes.erase(ctx->attached_planes.begin() + i);
      break;
    }
  ctx->planes_pool.push_back(std::make_pair(p->bytes, p->block));  // stream-ordered reuse: everything runs on the context's stream
  delete p;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_planes_attach(tdnnf_ctx* ctx, tdnnf_planes* p) {
  TDNNF_REQUIRE(ctx && p && p->ctx == ctx, "bad argument");
  if (p->attached++ == 0) ctx->attached_planes.push_back(p);
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_planes_detach(tdnnf_ctx* ctx, tdnnf_planes* p) {
  TDNNF_REQUIRE(ctx && p && p->ctx == ctx, "bad argument");
  if (p->attached > 0 && --p->attached == 0)
    for (size_t i = 0; i < ctx->attached_planes.size(); ++i)
      if (ctx->attached_planes[i] == p) {
        ctx->attached_planes.erase(ctx->attached_planes.begin() + i);
        break;
      }
  return TDNNF_OK;
}

extern "C" int tdnnf_planes_matches(const tdnnf_planes* p, const float* src, int rows, int cols, int stride) {
  return p && p->src == src && p->R == rows && p->D == cols && p->ld == stride;
}
