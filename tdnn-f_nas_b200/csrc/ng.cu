// Device-side helpers of the natural-gradient update (OnlineNaturalGradient::PreconditionDirections at
// ref tdnn.cc:598-599, simple.cc:9542) that are not GEMMs: the trace of X~ X~^T over the spliced input
// WITHOUT materialising X~ (the reference builds the R x (n*D_in+1) matrix in_value_temp, tdnn.cc:476-514),
// the "scale" factor kept on the device (no host sync), and an axpy whose coefficient lives on the device.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "context.h"
#include "ptx.cuh"
#include <cstring>

using namespace tdnnf;

namespace {

// sumsq[i] += sum over the rows of view i (rows row_offsets[i] + k*row_stride, k < out_rows) of ||row||^2.
// One warp per row; each input row is read exactly once whatever the number of overlapping views.
__global__ void __launch_bounds__(256)
view_sumsq_kernel(const float* __restrict__ in, int in_rows, int in_dim, long long ld, int out_rows, int n,
                  TdnnfOffsets offs, int row_stride, double* __restrict__ sumsq) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
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
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
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

// L (r x r) += H^T H for a tall skinny H (N x r, r <= 16*TB): the Gram matrix of the natural-gradient projection
// (kaldi: L_t = H_t^T H_t), in plain fp32 FMAs.  Thread (ti, tj) of a 16 x 16 block owns the TB x TB outputs
// (ti*TB.., tj*TB..).  A CTA owns ONE contiguous slab of rows and stages it in shared memory in a single pass (every load
// in flight at once: the first version staged 32 rows at a time, four dependent load -> barrier -> FMA rounds per CTA,
// and measured 30 us at rank 80 for 0.2 GFLOP), then adds its partial into one of kGramBufs accumulation buffers with
// red.add: chains of ~18 same-address reductions instead of 148 (all CTAs into one buffer measured 37 us), and the
// finish kernel sums 8 buffers instead of 148 per-CTA partials.
constexpr int kGramBufs = 8;
constexpr int kGramSlab = 256;  // rows staged per pass (dynamic shared memory: kGramSlab x (16*TB + 1) floats)
constexpr int kGramCluster = 8;
// CL = true (launched in clusters of kGramCluster CTAs): the CTAs of a cluster leave their r x r partial sums in their own
// shared memory, and CTA c of the cluster adds up slice c of all eight through distributed shared memory and stores it:
// one partial per CLUSTER in global memory, plain stores, no atomics (the 6400 red.adds per CTA of the buffer scheme were
// ~20 of the 28 us at rank 80).
template <int TB, bool CL>
__global__ void __launch_bounds__(256) ng_gram_kernel(const float* __restrict__ H, int N, int r, long long ld,
                                                      float* __restrict__ partials,
                                                      const float* __restrict__ rowsq, int in_rows, int n, TdnnfOffsets offs,
                                                      int row_stride, double* __restrict__ sumsq) {
  constexpr int W = 16 * TB;
  extern __shared__ float gram_tile[];  // [slab rows][W + 1]
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  // optional second job: this CTA's share of sum over the rows of view i of rowsq[row] (view i = rows offs[i] +
  // k*row_stride, k < N) -> sumsq[blockIdx.x][i]: tr(X X^T) of the spliced operand from the per-row sums of squares
  // the operand split left behind
  if (rowsq != nullptr) {
    double local[TDNNF_MAX_OFFSETS];
#pragma unroll
    for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) local[i] = 0.0;
    for (int row = blockIdx.x * 256 + threadIdx.x; row < in_rows; row += gridDim.x * 256) {
      const float sq = rowsq[row];
#pragma unroll
      for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) {
        if (i < n) {
          const int d = row - offs.v[i];
          const bool in_view = row_stride == 1 ? (d >= 0 && d < N) : (d >= 0 && d % row_stride == 0 && d / row_stride < N);
          if (in_view) local[i] += (double)sq;
        }
      }
    }
    // block reduction (warp shuffles, then 8 warps through shared memory); the per-CTA result goes to
    // view_partials[blockIdx.x][i] -- same-address double atomics from every CTA serialise in L2
    __shared__ double view_red[8][TDNNF_MAX_OFFSETS];
#pragma unroll
    for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) {
      double v = i < n ? local[i] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) view_red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < TDNNF_MAX_OFFSETS) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += view_red[w][threadIdx.x];
      sumsq[(size_t)blockIdx.x * TDNNF_MAX_OFFSETS + threadIdx.x] = v;
    }
  }
  const int ti = threadIdx.x / 16, tj = threadIdx.x % 16;
  float acc[TB][TB];
#pragma unroll
  for (int a = 0; a < TB; ++a)
#pragma unroll
    for (int b = 0; b < TB; ++b) acc[a][b] = 0.f;
  const bool vec = (r % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(H) & 15) == 0);
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const long long first = (long long)blockIdx.x * per;
  const long long last = first + per < N ? first + per : N;
  for (long long row0 = first; row0 < last; row0 += kGramSlab) {
    const int rows_here = (int)(last - row0 < kGramSlab ? last - row0 : kGramSlab);
    __syncthreads();
    if (vec) {
      // float4 loads, four per thread in flight before the first store (a plain element loop left ~32 dependent
      // load -> store rounds per thread: most of the 28 us this kernel took at rank 80)
      constexpr int W4 = W / 4;
      const int total4 = rows_here * W4;
      for (int base = threadIdx.x; base < total4; base += 256 * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * 256;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < total4) {
            const int rr = idx / W4, c = (idx % W4) * 4;
            if (c < r) v[u] = *reinterpret_cast<const float4*>(H + (row0 + rr) * ld + c);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * 256;
          if (idx < total4) {
            float* d = gram_tile + (idx / W4) * (W + 1) + (idx % W4) * 4;
            d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
          }
        }
      }
    } else {
      for (int idx = threadIdx.x; idx < rows_here * W; idx += 256) {
        const int rr = idx / W, c = idx % W;
        gram_tile[rr * (W + 1) + c] = c < r ? H[(row0 + rr) * ld + c] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < rows_here; ++rr) {
      float a[TB], b[TB];
      const float* row = gram_tile + rr * (W + 1);
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        a[k] = row[ti * TB + k];
        b[k] = row[tj * TB + k];
      }
#pragma unroll
      for (int x = 0; x < TB; ++x)
#pragma unroll
        for (int y = 0; y < TB; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
  }
  if constexpr (CL) {
    __syncthreads();  // the staging tile is no longer read
    float* part = gram_tile;  // [r * r]
#pragma unroll
    for (int x = 0; x < TB; ++x)
#pragma unroll
      for (int y = 0; y < TB; ++y) {
        const int i = ti * TB + x, j = tj * TB + y;
        if (i < r && j < r) part[i * r + j] = acc[x][y];
      }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int rr2 = r * r, slice = (rr2 + kGramCluster - 1) / kGramCluster;
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(part));
    float* mine = partials + (size_t)(blockIdx.x / kGramCluster) * rr2;
    for (int e = (int)crank * slice + threadIdx.x; e < min(((int)crank + 1) * slice, rr2); e += 256) {
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < kGramCluster; ++c) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + 4u * (uint32_t)e), "r"(c));
        float t;
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t) : "r"(remote));
        v += t;
      }
      mine[e] = v;
    }
    __syncthreads();
    // nobody leaves while its shared memory is still being read
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    float* mine = partials + (size_t)(blockIdx.x % kGramBufs) * r * r;
#pragma unroll
    for (int x = 0; x < TB; ++x)
#pragma unroll
      for (int y = 0; y < TB; ++y) {
        const int i = ti * TB + x, j = tj * TB + y;
        if (i < r && j < r && first < last) atomicAdd(mine + i * r + j, acc[x][y]);
      }
  }
}

// L[i][j] = sum of the accumulation buffers (which are zeroed again as they are read); tr(L) and <L, W W^T> reduced across CTAs in double; the last CTA to finish
// turns them into out[0..2] = {tr(X X^T), tr(X^ X^^T), scale} exactly as ng_scale_kernel, and re-arms the scratch.
// 1024 threads = 128 consecutive elements x 8 slices of the partials: coalesced, and only nblk/8 loads per thread.
__global__ void __launch_bounds__(1024) ng_gram_finish_kernel(float* __restrict__ partials, int nblk, int r,
                                                              float* __restrict__ L, long long l_ld,
                                                              const float* __restrict__ WWt, long long w_ld,
                                                              const double* __restrict__ sumsq,
                                                              const double* __restrict__ view_partials /* [nview][16] or null */,
                                                              int nview,
                                                              const float* __restrict__ weff, int n, float ones_rows,
                                                              double* __restrict__ acc /* [2] */,
                                                              unsigned int* __restrict__ counter, float* __restrict__ out) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float red[8][128];
  __shared__ double red_tr[4], red_dot[4];
  __shared__ double view_sum[TDNNF_MAX_OFFSETS];
  __shared__ bool last;
  const int col = threadIdx.x & 127, slice = threadIdx.x >> 7;
  const int e = blockIdx.x * 128 + col;
  float sum = 0.f;
  if (e < r * r) {
#pragma unroll 4
    for (int b = slice; b < nblk; b += 8) {
      sum += partials[(size_t)b * r * r + e];
      partials[(size_t)b * r * r + e] = 0.f;  // re-armed for the next call on this stream
    }
  }
  red[slice][col] = sum;
  __syncthreads();
  if (slice == 0) {
    double tr = 0.0, dot = 0.0;
    if (e < r * r) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += red[k][col];
      const int i = e / r, j = e % r;
      L[i * l_ld + j] = tot;
      if (i == j) tr = (double)tot;
      dot = (double)tot * (double)WWt[i * w_ld + j];
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
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&acc[0], red_tr[0] + red_tr[1] + red_tr[2] + red_tr[3]);
    atomicAdd(&acc[1], red_dot[0] + red_dot[1] + red_dot[2] + red_dot[3]);
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // the last CTA: view sums (from the per-CTA partials if given), then the three scalars
  if (threadIdx.x < TDNNF_MAX_OFFSETS) view_sum[threadIdx.x] = 0.0;
  __syncthreads();
  if (view_partials != nullptr) {
    const int i = threadIdx.x & 15;
    double v = 0.0;
    for (int b = threadIdx.x >> 4; b < nview; b += 64) v += view_partials[(size_t)b * TDNNF_MAX_OFFSETS + i];
    if (i < n && v != 0.0) atomicAdd(&view_sum[i], v);
  } else if (threadIdx.x < n) {
    view_sum[threadIdx.x] = sumsq[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const double t = atomicAdd(&acc[0], 0.0), d = atomicAdd(&acc[1], 0.0);  // read through L2
    double initial = (double)ones_rows;
    for (int i = 0; i < n; ++i) {
      const double w = weff ? (double)weff[i] : 1.0;
      initial += w * w * view_sum[i];
    }
    const double fin = initial - 2.0 * t + d;
    out[0] = (float)initial;
    out[1] = (float)fin;
    out[2] = (initial <= 0.0 || !(fin > 0.0)) ? 1.0f : (float)sqrt(initial / fin);
    acc[0] = 0.0;
    acc[1] = 0.0;
    *counter = 0u;
  }
}

// W_next[i][d] = sum_{k < r} A[i][k] J[k][d] + sum_{k < r} AC[i][k] W[k][d]     (i < r <= 128, d < D)
// The W_{t+1} = A_t (J_t + diag(c) W_t) step of OnlineNaturalGradient in exact fp32 FMAs: K = 2r is far too short for
// the tensor-core path (it took a zero-fill, four operand splits and two 6-product GEMMs per preconditioner).
// A block owns 16 rows x 64 columns and splits K over 4 thread groups (256 threads): the 16 x 2r coefficient strip goes
// to shared memory once, every thread streams its quarter of its column of [J; W] in batches of 8 independent coalesced
// loads, and the four partial sums meet in shared memory.  (History: 32-row chunks of [J; W] through shared memory with
// two barriers each measured 57 us at rank 80; one thread per column over the whole K, 20 dependent load batches with 2
// warps per SM, 26 us.)
constexpr int kWuRows = 16, kWuCols = 64, kWuGroups = 4;
__global__ void __launch_bounds__(kWuCols* kWuGroups)
ng_w_update_kernel(const float* __restrict__ A, int a_ld, const float* __restrict__ AC, int ac_ld, const float* __restrict__ J,
                   long long j_ld, const float* __restrict__ W, long long w_ld, int r, int D, float* __restrict__ out,
                   long long out_ld) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  extern __shared__ __align__(16) float sM[];  // [kWuRows][Kp] coefficients, then [kWuGroups][kWuRows][kWuCols] partial sums
  const int K = 2 * r, Kp = (K + 7) & ~7;
  float* sP = sM + kWuRows * Kp;
  const int d0 = blockIdx.x * kWuCols, i0 = blockIdx.y * kWuRows;
  const int tx = threadIdx.x % kWuCols, grp = threadIdx.x / kWuCols;
  for (int idx = threadIdx.x; idx < kWuRows * Kp; idx += kWuCols * kWuGroups) {
    const int u = idx / Kp, k = idx % Kp, i = i0 + u;
    float v = 0.f;
    if (i < r && k < K) v = k < r ? A[i * a_ld + k] : AC[i * ac_ld + (k - r)];
    sM[idx] = v;
  }
  __syncthreads();
  const int d = d0 + tx;
  float acc[kWuRows];
#pragma unroll
  for (int u = 0; u < kWuRows; ++u) acc[u] = 0.f;
  const int kq = ((Kp / 8 + kWuGroups - 1) / kWuGroups) * 8;  // K share of one group, a multiple of 8
  const int k_begin = grp * kq, k_end = min(k_begin + kq, Kp);
  if (d < D) {
    for (int k0 = k_begin; k0 < k_end; k0 += 8) {
      float b[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = k0 + kk;
        b[kk] = k < r ? J[k * j_ld + d] : (k < K ? W[(k - r) * w_ld + d] : 0.f);
      }
#pragma unroll
      for (int u = 0; u < kWuRows; ++u) {
        const float4 m0 = *reinterpret_cast<const float4*>(&sM[u * Kp + k0]);
        const float4 m1 = *reinterpret_cast<const float4*>(&sM[u * Kp + k0 + 4]);
        acc[u] = fmaf(m0.x, b[0], fmaf(m0.y, b[1], fmaf(m0.z, b[2], fmaf(m0.w, b[3], acc[u]))));
        acc[u] = fmaf(m1.x, b[4], fmaf(m1.y, b[5], fmaf(m1.z, b[6], fmaf(m1.w, b[7], acc[u]))));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kWuRows; ++u) sP[(grp * kWuRows + u) * kWuCols + tx] = acc[u];
  __syncthreads();
  // group g finishes rows 4g .. 4g+3
  if (d < D) {
#pragma unroll
    for (int uu = 0; uu < kWuRows / kWuGroups; ++uu) {
      const int u = grp * (kWuRows / kWuGroups) + uu;
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < kWuGroups; ++g) v += sP[(g * kWuRows + u) * kWuCols + tx];
      if (i0 + u < r) out[(long long)(i0 + u) * out_ld + d] = v;
    }
  }
}

template <bool ZERO>
__global__ void mat_axpy_dev_kernel(float alpha, const float* __restrict__ f1, const float* __restrict__ f2,
                                    float* __restrict__ src, long long ss, float* __restrict__ dst, long long ds,
                                    int rows, int cols) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  float a = alpha;
  if (f1) a *= *f1;
  if (f2) a *= *f2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)rows * cols;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = (int)(idx % cols);
    dst[r * ds + c] += a * src[r * ss + c];
    if (ZERO) src[r * ss + c] = 0.f;
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
  TDNNF_CUDA_OK(launch_pdl(view_sumsq_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, in, in_rows, in_dim, in_stride, out_rows, n, offs, row_stride, sumsq));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ng_scale(tdnnf_ctx* ctx, const double* sumsq, const float* weff, int n, float ones_rows,
                              const float* L, int l_stride, const float* WWt, int w_stride, int rank, float* out3) {
  TDNNF_REQUIRE(ctx && sumsq && L && WWt && out3, "null argument");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS && rank >= 1 && l_stride >= rank && w_stride >= rank, "bad argument");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(launch_pdl(ng_scale_kernel, dim3(1), dim3(1024), 0, ctx->stream, 1, sumsq, weff, n, ones_rows, L, l_stride, WWt, w_stride, rank, out3));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ng_gram_scale(tdnnf_ctx* ctx, const float* H, int rows, int rank, int h_stride, float* L,
                                   int l_stride, const float* WWt, int w_stride, const float* rowsq, double* sumsq,
                                   int in_rows, int n, const int32_t* row_offsets, int row_stride, const float* weff,
                                   float ones_rows, float* out3) {
  TDNNF_REQUIRE(ctx && H && L && WWt && out3 && (rowsq || sumsq), "null argument");
  TDNNF_REQUIRE(rows > 0 && rank >= 1 && rank <= 128 && h_stride >= rank && l_stride >= rank && w_stride >= rank,
                "bad argument (rank <= 128)");
  TDNNF_REQUIRE(n >= 1 && n <= TDNNF_MAX_OFFSETS, "bad number of views");
  TDNNF_REQUIRE(rowsq == nullptr || (row_offsets && row_stride >= 1 && in_rows > 0), "bad view description");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  int blocks = std::max(1, std::min((rows + 63) / 64, ctx->num_sms));
  const int tb = (rank + 15) / 16;
  // wide Gram matrices from many CTAs: clusters reduce through distributed shared memory (one partial per cluster)
  const bool clustered = tb >= 3 && blocks >= 2 * kGramCluster;
  if (clustered) blocks = blocks / kGramCluster * kGramCluster;
  const int nparts = clustered ? blocks / kGramCluster : std::min(blocks, kGramBufs);
  // scratch: [0,16) acc[2] doubles, [16,20) counter, [64, 64 + 128*num_sms) per-CTA view sums, then the partial Gram matrices
  const size_t view_bytes = (size_t)ctx->num_sms * TDNNF_MAX_OFFSETS * sizeof(double);
  const size_t max_parts = std::max(kGramBufs, ctx->num_sms / kGramCluster + 1);
  const size_t need = 64 + view_bytes + max_parts * rank * rank * sizeof(float);
  if (ctx->ng_scratch_bytes < need) {
    if (ctx->ng_scratch) TDNNF_CUDA_OK(cudaFree(ctx->ng_scratch));  // waits for kernels still using it
    ctx->ng_scratch = nullptr;
    ctx->ng_scratch_bytes = 0;
    const size_t want = std::max(need, 64 + view_bytes + max_parts * 128 * 128 * sizeof(float));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ctx->ng_scratch), want));
    // {acc[2], counter} and the accumulation buffers start at zero and re-arm themselves in the finish kernel
    TDNNF_CUDA_OK(cudaMemsetAsync(ctx->ng_scratch, 0, want, ctx->stream));
    ctx->ng_scratch_bytes = want;
  }
  double* acc = reinterpret_cast<double*>(ctx->ng_scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(ctx->ng_scratch + 16);
  double* view_partials = reinterpret_cast<double*>(ctx->ng_scratch + 64);
  float* partials = reinterpret_cast<float*>(ctx->ng_scratch + 64 + view_bytes);
  TdnnfOffsets offs;
  for (int i = 0; i < TDNNF_MAX_OFFSETS; ++i) offs.v[i] = (rowsq && i < n) ? row_offsets[i] : 0;
  const int per = (rows + blocks - 1) / blocks;
  const size_t tile_bytes = std::max((size_t)std::min(per, kGramSlab) * (16 * tb + 1), (size_t)rank * rank) * sizeof(float);
  auto launch = [&](auto kern, bool cl) -> cudaError_t {
    // every instance has the same pointer TYPE (one instantiation of this lambda): remember the functions, not a flag
    static std::vector<const void*> attr_set;
    if (std::find(attr_set.begin(), attr_set.end(), (const void*)kern) == attr_set.end()) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSlab * (16 * 8 + 1) * (int)sizeof(float));
      if (e != cudaSuccess) return e;
      attr_set.push_back((const void*)kern);
    }
    return launch_pdl(kern, dim3(blocks), dim3(256), tile_bytes, ctx->stream, cl ? kGramCluster : 1, H, rows, rank, h_stride, partials,
                      rowsq, in_rows, n, offs, row_stride, view_partials);
  };
  switch (tb) {
    case 1: TDNNF_CUDA_OK(launch(ng_gram_kernel<1, false>, false)); break;
    case 2: TDNNF_CUDA_OK(launch(ng_gram_kernel<2, false>, false)); break;
    case 3: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<3, true>, true) : launch(ng_gram_kernel<3, false>, false)); break;
    case 4: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<4, true>, true) : launch(ng_gram_kernel<4, false>, false)); break;
    case 5: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<5, true>, true) : launch(ng_gram_kernel<5, false>, false)); break;
    case 6: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<6, true>, true) : launch(ng_gram_kernel<6, false>, false)); break;
    case 7: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<7, true>, true) : launch(ng_gram_kernel<7, false>, false)); break;
    default: TDNNF_CUDA_OK(clustered ? launch(ng_gram_kernel<8, true>, true) : launch(ng_gram_kernel<8, false>, false)); break;
  }
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  TDNNF_CUDA_OK(launch_pdl(ng_gram_finish_kernel, dim3((rank * rank + 127) / 128), dim3(1024), 0, ctx->stream, 1, partials, nparts, rank, L, l_stride, WWt, w_stride, sumsq,
                                                                            rowsq ? view_partials : nullptr, blocks, weff, n, ones_rows, acc,
                                                                            counter, out3));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ng_w_update(tdnnf_ctx* ctx, const float* A, int a_stride, const float* AC, int ac_stride, const float* J,
                                 int j_stride, const float* W, int w_stride, int rank, int dim, float* W_next, int out_stride) {
  TDNNF_REQUIRE(ctx && A && AC && J && W && W_next, "null argument");
  TDNNF_REQUIRE(rank >= 1 && rank <= 128 && dim >= 1 && a_stride >= rank && ac_stride >= rank && j_stride >= dim &&
                    w_stride >= dim && out_stride >= dim,
                "bad argument (rank <= 128)");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const int Kp = (2 * rank + 7) & ~7;
  TDNNF_CUDA_OK(launch_pdl(ng_w_update_kernel, dim3((dim + kWuCols - 1) / kWuCols, (rank + kWuRows - 1) / kWuRows), dim3(kWuCols * kWuGroups), sizeof(float) * (kWuRows * Kp + kWuGroups * kWuRows * kWuCols), ctx->stream, 1, 
      A, a_stride, AC, ac_stride, J, j_stride, W, w_stride, rank, dim, W_next, out_stride));
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
  TDNNF_CUDA_OK(launch_pdl(mat_axpy_dev_kernel<false>, dim3((int)b), dim3(256), 0, ctx->stream, 1, alpha, factor1_dev, factor2_dev, const_cast<float*>(src), src_stride, dst,
                                                              dst_stride, rows, cols));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_mat_axpy_dev_zero(tdnnf_ctx* ctx, float alpha, const float* factor1_dev, const float* factor2_dev, float* src,
                                       int src_stride, float* dst, int dst_stride, int rows, int cols) {
  TDNNF_REQUIRE(ctx && src && dst, "null argument");
  if (rows <= 0 || cols <= 0) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  long long b = ((long long)rows * cols + 255) / 256;
  b = std::max(1LL, std::min(b, (long long)ctx->num_sms * 16));
  TDNNF_CUDA_OK(launch_pdl(mat_axpy_dev_kernel<true>, dim3((int)b), dim3(256), 0, ctx->stream, 1, alpha, factor1_dev, factor2_dev, src, src_stride, dst, dst_stride, rows,
                                                             cols));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

// ---- G <- (I - Wo^T Wo) G (I - Wi^T Wi): the two rank-r projections of the preconditioned gradient, in fp32 FMAs.
// As four calls of the tensor-core GEMM (G Wi^T, G -= . Wi, Wo G, G -= Wo^T .) they were 14 launches per component (eight
// operand splits, two zero-fills) for 0.7 GFLOP: ~2.5 ms per step of fixed costs.  Three small kernels instead:
//   ng_rt_kernel  Ht[k, o] += sum_c Wi[k, c] G[o, c]      (contraction over the columns, split over column ranges)
//   ng_lt_kernel  T[k, c]  += sum_o Wo[k, o] G[o, c]      (contraction over the rows, split over row ranges)
//   ng_lu_kernel  G[o, c]  -= sum_k P[k, o] Q[k, c]       (P = Ht, Q = Wi  or  P = Wo, Q = T)
constexpr int kNgMaxRank = 128;
constexpr int kNgBufs = 8;  // accumulation buffers of the split reductions (chains of same-address red.adds serialise in L2)

// grid (o-tiles of 32 rows, column splits); 256 threads: thread (to = tid % 32, tk = tid / 32) owns Ht[tk + 8 i][o0 + to],
// i < NK (NK = ceil(rank / 8)).  Split s adds into buffer s % kNgBufs.
template <int NK>
__global__ void __launch_bounds__(256) ng_rt_kernel(const float* __restrict__ G, int rows, int cols, long long g_ld,
                                                    const float* __restrict__ Wi, int ri, long long wi_ld, int cols_per_split,
                                                    float* __restrict__ Ht /* [kNgBufs][ri][rows] */) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float Gs[2][32][33];
  __shared__ float Ws[2][NK * 8][33];
  const int to = threadIdx.x & 31, tk = threadIdx.x >> 5;
  const int o0 = blockIdx.x * 32;
  const int c_begin = blockIdx.y * cols_per_split, c_end = min(c_begin + cols_per_split, cols);
  float acc[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) acc[i] = 0.f;
  // register-staged double buffering: the loads of chunk j + 1 are in flight while chunk j is multiplied
  float gl[4], wl[NK];
  auto load = [&](int c0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tk + 8 * j, c = c0 + to;
      gl[j] = (o < rows && c < c_end) ? G[(long long)o * g_ld + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const int k = tk + 8 * i, c = c0 + to;
      wl[i] = (k < ri && c < c_end) ? Wi[(long long)k * wi_ld + c] : 0.f;
    }
  };
  auto stash = [&](int b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) Gs[b][tk + 8 * j][to] = gl[j];
#pragma unroll
    for (int i = 0; i < NK; ++i) Ws[b][tk + 8 * i][to] = wl[i];
  };
  load(c_begin);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    const bool more = c0 + 32 < c_end;
    if (more) load(c0 + 32);
#pragma unroll 8
    for (int cc = 0; cc < 32; ++cc) {
      const float g = Gs[buf][to][cc];
#pragma unroll
      for (int i = 0; i < NK; ++i) acc[i] = fmaf(Ws[buf][tk + 8 * i][cc], g, acc[i]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const int o = o0 + to;
  float* mine = Ht + (size_t)(blockIdx.y % kNgBufs) * ri * rows;
  if (o < rows) {
#pragma unroll
    for (int i = 0; i < NK; ++i)
      if (tk + 8 * i < ri) atomicAdd(mine + (long long)(tk + 8 * i) * rows + o, acc[i]);
  }
}

// grid (c-tiles of 32 columns, row splits); thread (tc = tid % 32, tk = tid / 32) owns T[tk + 8 i][c0 + tc], i < NK.
template <int NK>
__global__ void __launch_bounds__(256) ng_lt_kernel(const float* __restrict__ G, int rows, int cols, long long g_ld,
                                                    const float* __restrict__ Wo, int ro, long long wo_ld, int rows_per_split,
                                                    float* __restrict__ T /* [kNgBufs][ro][cols] */) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float Gs[2][32][32];
  __shared__ __align__(16) float Ps[2][NK * 8][36];
  const int tc = threadIdx.x & 31, tk = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, c = c0 + tc;
  const int o_begin = blockIdx.y * rows_per_split, o_end = min(o_begin + rows_per_split, rows);
  float acc[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) acc[i] = 0.f;
  float gl[4], pl[NK];
  auto load = [&](int o0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tk + 8 * j;
      gl[j] = (o < o_end && c < cols) ? G[(long long)o * g_ld + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const int k = tk + 8 * i, o = o0 + tc;
      pl[i] = (k < ro && o < o_end) ? Wo[(long long)k * wo_ld + o] : 0.f;
    }
  };
  auto stash = [&](int b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) Gs[b][tk + 8 * j][tc] = gl[j];
#pragma unroll
    for (int i = 0; i < NK; ++i) Ps[b][tk + 8 * i][tc] = pl[i];
  };
  load(o_begin);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int o0 = o_begin; o0 < o_end; o0 += 32) {
    const bool more = o0 + 32 < o_end;
    if (more) load(o0 + 32);
#pragma unroll 2
    for (int oo = 0; oo < 32; oo += 4) {
      const float g0 = Gs[buf][oo][tc], g1 = Gs[buf][oo + 1][tc], g2 = Gs[buf][oo + 2][tc], g3 = Gs[buf][oo + 3][tc];
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        const float4 p = *reinterpret_cast<const float4*>(&Ps[buf][tk + 8 * i][oo]);
        acc[i] = fmaf(p.x, g0, fmaf(p.y, g1, fmaf(p.z, g2, fmaf(p.w, g3, acc[i]))));
      }
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  float* mine = T + (size_t)(blockIdx.y % kNgBufs) * ro * cols;
  if (c < cols) {
#pragma unroll
    for (int i = 0; i < NK; ++i)
      if (tk + 8 * i < ro) atomicAdd(mine + (long long)(tk + 8 * i) * cols + c, acc[i]);
  }
}

// buf[0][e] += buf[1][e] + ... + buf[n-1][e]
__global__ void ng_sum_bufs_kernel(float* __restrict__ buf, long long elems, int n) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < elems; e += (long long)gridDim.x * blockDim.x) {
    float v = buf[e];
    for (int b = 1; b < n; ++b) v += buf[b * elems + e];
    buf[e] = v;
  }
}

// G[o, c] -= sum_k P[k, o] Q[k, c]; grid (c-tiles of 64, o-tiles of 64); thread (tc = tid % 32, to = tid / 32) owns rows
// o0 + 8 to .. + 7 and columns c0 + tc, c0 + tc + 32 (r x 128 floats staged per 4096 outputs)
__global__ void __launch_bounds__(256) ng_lu_kernel(float* __restrict__ G, int rows, int cols, long long g_ld,
                                                    const float* __restrict__ P, long long p_ld, const float* __restrict__ Q,
                                                    long long q_ld, int r) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  extern __shared__ __align__(16) float lu_smem[];  // Ps[r][64], Qs[r][64]
  float* Ps = lu_smem;
  float* Qs = lu_smem + r * 64;
  const int tc = threadIdx.x & 31, to = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64, o0 = blockIdx.y * 64;
#pragma unroll 2
  for (int k = to; k < r; k += 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = o0 + tc + 32 * h, c = c0 + tc + 32 * h;
      Ps[k * 64 + tc + 32 * h] = o < rows ? P[(long long)k * p_ld + o] : 0.f;
      Qs[k * 64 + tc + 32 * h] = c < cols ? Q[(long long)k * q_ld + c] : 0.f;
    }
  }
  __syncthreads();
  float acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.f;
#pragma unroll 2
  for (int k = 0; k < r; ++k) {
    const float4 pa = *reinterpret_cast<const float4*>(&Ps[k * 64 + to * 8]);
    const float4 pb = *reinterpret_cast<const float4*>(&Ps[k * 64 + to * 8 + 4]);
    const float q0 = Qs[k * 64 + tc], q1 = Qs[k * 64 + tc + 32];
    const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j][0] = fmaf(pv[j], q0, acc[j][0]);
      acc[j][1] = fmaf(pv[j], q1, acc[j][1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = o0 + to * 8 + j;
    if (o >= rows) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + tc + 32 * h;
      if (c < cols) G[(long long)o * g_ld + c] -= acc[j][h];
    }
  }
}

template <typename F>
static bool ng_dispatch_nk(int rank, F&& f) {  // NK = ceil(rank / 8) rounded up to an instantiated value
  const int nk = (rank + 7) / 8;
  if (nk <= 1) f(std::integral_constant<int, 1>());
  else if (nk <= 3) f(std::integral_constant<int, 3>());
  else if (nk <= 5) f(std::integral_constant<int, 5>());
  else if (nk <= 8) f(std::integral_constant<int, 8>());
  else if (nk <= 10) f(std::integral_constant<int, 10>());
  else if (nk <= 16) f(std::integral_constant<int, 16>());
  else return false;
  return true;
}

extern "C" int tdnnf_ng_project_gradient(tdnnf_ctx* ctx, float* G, int rows, int cols, int g_stride, const float* Wi, int ri,
                                         int wi_stride, const float* Wo, int ro, int wo_stride) {
  TDNNF_REQUIRE(ctx && G && rows > 0 && cols > 0 && g_stride >= cols, "bad gradient matrix");
  TDNNF_REQUIRE(!Wi || (ri >= 1 && ri <= kNgMaxRank && wi_stride >= cols), "bad in-side projection (rank <= 128)");
  TDNNF_REQUIRE(!Wo || (ro >= 1 && ro <= kNgMaxRank && wo_stride >= rows), "bad out-side projection (rank <= 128)");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t ht_floats = Wi ? (size_t)kNgBufs * ri * rows : 0, t_floats = Wo ? (size_t)kNgBufs * ro * cols : 0;
  const size_t ht_bytes = (ht_floats * sizeof(float) + 255) & ~size_t(255);
  const size_t t_bytes = (t_floats * sizeof(float) + 255) & ~size_t(255);
  ctx->ws_reset();
  int rc = ctx->ws_reserve(ht_bytes + t_bytes + 1024);
  if (rc) return rc;
  float* scratch = static_cast<float*>(ctx->ws_alloc(ht_bytes + t_bytes));
  if (!scratch) return TDNNF_ERR_NOMEM;
  float* Ht = scratch;
  float* T = scratch + ht_bytes / sizeof(float);
  const int want_blocks = 4 * ctx->num_sms;
  {
    static bool lu_attr = false;
    if (!lu_attr) {
      TDNNF_CUDA_OK(cudaFuncSetAttribute(ng_lu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(float) * kNgMaxRank * 128));
      lu_attr = true;
    }
  }
  if (Wi) {
    const int o_tiles = (rows + 31) / 32;
    int splits = std::max(1, std::min(want_blocks / o_tiles, (cols + 63) / 64));
    const int per = (((cols + splits - 1) / splits) + 31) / 32 * 32;
    splits = (cols + per - 1) / per;
    const int bufs = std::min(splits, kNgBufs);
    { int zrc = zero_async(ctx, Ht, (size_t)bufs * ri * rows * sizeof(float)); if (zrc) return zrc; }
    {
      cudaError_t err = cudaSuccess;
      if (!ng_dispatch_nk(ri, [&](auto nk) { err = launch_pdl(ng_rt_kernel<decltype(nk)::value>, dim3(o_tiles, splits), dim3(256), 0, ctx->stream, 1, G, rows, cols, g_stride, Wi, ri, wi_stride,
                                                                                          per, Ht); }))
        return fail(TDNNF_ERR_INVALID, "unsupported rank");
      TDNNF_CUDA_OK(err);
    }
    if (bufs > 1) {
      TDNNF_CUDA_OK(launch_pdl(ng_sum_bufs_kernel, dim3(std::max(1, std::min((ri * rows + 255) / 256, ctx->num_sms * 4))), dim3(256), 0, ctx->stream, 1, 
          Ht, (long long)ri * rows, bufs));
      ctx->launches++;
    }
    TDNNF_CUDA_OK(launch_pdl(ng_lu_kernel, dim3((cols + 63) / 64, (rows + 63) / 64), dim3(256), sizeof(float) * ri * 128, ctx->stream, 1, G, rows, cols, g_stride, Ht,
                                                                                                    rows, Wi, wi_stride, ri));
    ctx->launches += 2;
    TDNNF_CUDA_OK(cudaGetLastError());
  }
  if (Wo) {
    const int c_tiles = (cols + 31) / 32;
    int splits = std::max(1, std::min(want_blocks / c_tiles, (rows + 63) / 64));
    const int per = (((rows + splits - 1) / splits) + 31) / 32 * 32;
    splits = (rows + per - 1) / per;
    const int bufs = std::min(splits, kNgBufs);
    { int zrc = zero_async(ctx, T, (size_t)bufs * ro * cols * sizeof(float)); if (zrc) return zrc; }
    {
      cudaError_t err = cudaSuccess;
      if (!ng_dispatch_nk(ro, [&](auto nk) { err = launch_pdl(ng_lt_kernel<decltype(nk)::value>, dim3(c_tiles, splits), dim3(256), 0, ctx->stream, 1, G, rows, cols, g_stride, Wo, ro, wo_stride,
                                                                                          per, T); }))
        return fail(TDNNF_ERR_INVALID, "unsupported rank");
      TDNNF_CUDA_OK(err);
    }
    if (bufs > 1) {
      TDNNF_CUDA_OK(launch_pdl(ng_sum_bufs_kernel, dim3(std::max(1, std::min((ro * cols + 255) / 256, ctx->num_sms * 4))), dim3(256), 0, ctx->stream, 1, 
          T, (long long)ro * cols, bufs));
      ctx->launches++;
    }
    TDNNF_CUDA_OK(launch_pdl(ng_lu_kernel, dim3((cols + 63) / 64, (rows + 63) / 64), dim3(256), sizeof(float) * ro * 128, ctx->stream, 1, G, rows, cols, g_stride, Wo,
                                                                                                    wo_stride, T, cols, ro));
    ctx->launches += 2;
    TDNNF_CUDA_OK(cudaGetLastError());
  }
  return TDNNF_OK;
}

// Up to 4 strided 2-D copies in ONE launch, either side of which may be page-locked HOST memory (device-accessible
// under unified addressing): the R x R matrices of the natural-gradient eigen-update travel device -> host -> device
// without the copy engine.  (As cudaMemcpy2DAsync calls they were ~280 engine switches per refresh period in the middle of
// the backward pass: the finish step ran 2 ms longer than the sum of its kernels.)
struct CopyBlocks {
  const float* src[4];
  float* dst[4];
  int src_stride[4], dst_stride[4], rows[4], cols[4];
  int n;
};
__global__ void copy_blocks_kernel(const CopyBlocks b) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  for (int k = 0; k < b.n; ++k) {
    const int total = b.rows[k] * b.cols[k];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
      const int r = idx / b.cols[k], c = idx % b.cols[k];
      b.dst[k][(long long)r * b.dst_stride[k] + c] = b.src[k][(long long)r * b.src_stride[k] + c];
    }
  }
}

extern "C" int tdnnf_copy_blocks(tdnnf_ctx* ctx, int n, const float* const* src, const int32_t* src_strides, float* const* dst,
                                 const int32_t* dst_strides, const int32_t* rows, const int32_t* cols) {
  TDNNF_REQUIRE(ctx && src && dst && src_strides && dst_strides && rows && cols && n >= 1 && n <= 4, "bad argument (1 <= n <= 4)");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  CopyBlocks b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  long long most = 0;
  for (int k = 0; k < n; ++k) {
    TDNNF_REQUIRE(src[k] && dst[k] && rows[k] >= 0 && cols[k] >= 0 && src_strides[k] >= cols[k] && dst_strides[k] >= cols[k],
                  "bad block");
    b.src[k] = src[k]; b.dst[k] = dst[k];
    b.src_stride[k] = src_strides[k]; b.dst_stride[k] = dst_strides[k];
    b.rows[k] = rows[k]; b.cols[k] = cols[k];
    most = std::max(most, (long long)rows[k] * cols[k]);
  }
  if (most == 0) return TDNNF_OK;
  const int blocks = (int)std::max(1LL, std::min((most + 255) / 256, 64LL));
  TDNNF_CUDA_OK(launch_pdl(copy_blocks_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, b));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_get_stream(tdnnf_ctx* ctx, void** stream) {
  TDNNF_REQUIRE(ctx && stream, "null argument");
  *stream = ctx->stream;
  return TDNNF_OK;
}
