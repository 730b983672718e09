// LF-MMI denominator forward-backward (kaldi: chain/chain-denominator.{h,cc}, chain-den-graph.{h,cc},
// chain-kernels.cu -- upstream Kaldi, not shipped with the reference; restated from SURVEY.md App. B).
//
// Layout in HBM (all fp32, the sequence index s is always the contiguous one):
//   E     [T][P][S]      exp(clamp(nnet_output, -30, 30)), transposed once per call
//   alpha [T+1][N][S]    un-"dashed" alpha; tot[T+1][S] holds sum_h alpha(t,h,s)  (Kaldi keeps
//                        alpha-dash plus the sums in S trailing columns; alpha-dash is re-formed on
//                        read as alpha + leaky*init[h]*tot, which saves a full pass per frame)
//   betad [2][N][S]      beta-dash ping-pong; bsum[2][S] = sum_g init[g]*betad(g,s)
//   gamma [T][P][S]      occupation probabilities, transposed-added into nnet_output_deriv at the end
// Arithmetic is Kaldi's: probability domain, every frame divided by the previous frame's total.
#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "context.h"
#include "ptx.cuh"
#include "den_slices.h"

using namespace tdnnf;

struct tdnnf_den_graph {
  tdnnf_ctx* ctx = nullptr;
  int num_states = 0, num_pdfs = 0, num_transitions = 0;
  int2* fwd_ranges = nullptr;   // device [N]
  int2* bwd_ranges = nullptr;   // device [N]
  float4* trans = nullptr;      // device [A]: {prob, pdf (as int bits), state (as int bits), init[state]}
  float* init = nullptr;        // device [N]
  // states sorted by decreasing arc count (in-arcs for the forward recursion, out-arcs for the backward one): a
  // block's 32 states then have lists of (nearly) the same length, instead of finishing with its longest one
  int* order_in = nullptr;      // device [N]
  int* order_out = nullptr;     // device [N]
  float init_sum = 0.f;         // sum_h init[h] (host copy, fp32 sequential sum)
  // host copies for the per-computation arc plans of the resident kernels (den2_*)
  std::vector<int> h_fwd_ranges, h_bwd_ranges, h_pdf, h_state;
  std::vector<float> h_prob, h_init;
};

// One direction of the resident kernels' work list.  States are sorted by arc count and cut into warp-tasks of 32
// (lane = state), whose arcs are stored interleaved (arc k of lane l at base + 32 k + l, zero-probability padding up
// to the task's longest list): a warp reads 256 contiguous bytes per step and its lanes finish together.
struct DenPlan {
  int num_tasks = 0;
  int* task_base = nullptr;   // device [num_tasks]
  int* task_len = nullptr;    // device [num_tasks]
  int* task_state = nullptr;  // device [32 * num_tasks], -1 = no state
  uint2* arcs = nullptr;      // device: {prob bits, (pdf << state_bits) | other state}
  void destroy() {
    cudaFree(task_base);
    cudaFree(task_len);
    cudaFree(task_state);
    cudaFree(arcs);
  }
};

struct tdnnf_den_comp {
  tdnnf_ctx* ctx = nullptr;
  const tdnnf_den_graph* g = nullptr;
  int S = 0, T = 0;
  float leaky = 0.f;
  float* E = nullptr;
  float* alpha = nullptr;
  float* tot = nullptr;     // [(T+1)][S]
  float* betad = nullptr;   // [2][N][S]
  float* bsum = nullptr;    // [2][S]
  float* gamma = nullptr;   // [T][P][S]
  float* tot_prob = nullptr;  // [S]
  double* scalars = nullptr;  // [2]: logprob, alpha.beta check
  bool forward_done = false;
  // resident path (den2_*): V sequences per cluster of C CTAs; layouts E/gamma [T][S/V][P][V], alpha [T+1][S/V][N][V],
  // betad [2][S/V][N][V], bsum [T+1][S]
  bool resident = false;
  int V = 0, C = 0, state_bits = 0;
  DenPlan plan_fwd, plan_bwd;
  size_t smem_fwd = 0, smem_bwd = 0;
  // sequence-slice path (den_slices.cu): the default wherever it applies
  tdnnf_den_slices* slices = nullptr;
};

namespace {

template <int V>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ static Vec load(const float* p) { Vec r; r.v[0] = *p; return r; }
  __device__ void store(float* p) const { *p = v[0]; }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ static Vec load(const float* p) { const float2 t = *reinterpret_cast<const float2*>(p); Vec r; r.v[0] = t.x; r.v[1] = t.y; return r; }
  __device__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ static Vec load(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
  __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

template <int V>
__device__ __forceinline__ void red_add_vec(float* p, const float (&x)[V]) {
  if constexpr (V == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
  } else if constexpr (V == 2) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x[0]), "f"(x[1]) : "memory");
  } else {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(x[0]) : "memory");
  }
}

constexpr int kDenThreads = 256;
constexpr int kStatesPerBlock = 32;

// E[t][p][s] = exp(clamp(x[t*S+s][p])) : 32x32 tiled transpose per frame.
__global__ void den_exp_transpose_kernel(const float* __restrict__ x, long long ld, int S, int P, float* __restrict__ E) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int s = s0 + i, p = p0 + threadIdx.x;
    float v = 0.f;
    if (s < S && p < P) v = expf(fminf(fmaxf(x[((long long)t * S + s) * ld + p], -30.f), 30.f));
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    if (s < S && p < P) E[((long long)t * P + p) * S + s] = tile[threadIdx.x][i];
  }
}

// alpha(0,h,s) = init[h]; tot(0,s) = sum_h init[h]
__global__ void den_alpha_first_kernel(const float* __restrict__ init, int N, int S, float init_sum,
                                       float* __restrict__ alpha0, float* __restrict__ tot0) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    alpha0[i] = init[i / S];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
    tot0[i] = init_sum;
}

// One forward frame.  Thread = (state h, V consecutive sequences); lanes run along s, so every
// arc is two coalesced row reads (alpha(t-1,g,:) and E(t-1,pdf,:)) plus broadcast scalars.
template <int V, int U>
__global__ void __launch_bounds__(kDenThreads)
den_alpha_frame_kernel(const int2* __restrict__ bwd_ranges, const float4* __restrict__ trans, const int* __restrict__ order,
                       const float* __restrict__ init, int N, int spb, int S, float leaky, const float* __restrict__ alpha_prev,
                       const float* __restrict__ tot_prev, const float* __restrict__ E_prev,
                       float* __restrict__ alpha_cur, float* __restrict__ tot_cur) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int SV = S / V;
  const int tps = kDenThreads / SV;  // states processed concurrently by the block (SV <= 256 divides 256 or not: see host)
  const int sv = threadIdx.x % SV;
  const int hl = threadIdx.x / SV;
  const int s = sv * V;
  const bool active = hl < tps;
  float lt[V], inv[V], part[V];
  {
    const Vec<V> tp = Vec<V>::load(tot_prev + s);
#pragma unroll
    for (int j = 0; j < V; ++j) { lt[j] = leaky * tp.v[j]; inv[j] = 1.0f / tp.v[j]; part[j] = 0.f; }
  }
  // Block b takes 32 consecutive positions of the sorted order: lists of (nearly) equal length, so the block does not
  // wait for one long list; blocks are launched longest first, so each SM gets a mix (measured -14% at N = 16384;
  // striding the sorted order across blocks instead was slower: a third, mostly idle pass per thread).
  if (active) {
    for (int pos = blockIdx.x * spb + hl; pos < min((blockIdx.x + 1) * spb, N); pos += tps) {
      const int h = order[pos];
      const int2 rg = bwd_ranges[h];
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
      // arcs in groups of U: all transition records first, then all row gathers, then the FMAs, so that 2U independent
      // L2 reads are in flight per thread instead of a dependent chain
      for (int a = rg.x; a < rg.y; a += U) {
        float4 tr[U];
        Vec<V> al[U], e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) tr[u] = (a + u < rg.y) ? trans[a + u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pdf = __float_as_int(tr[u].y), g = __float_as_int(tr[u].z);  // (0,0) for the padding arcs: valid rows, weight 0
          al[u] = Vec<V>::load(alpha_prev + (long long)g * S + s);
          e[u] = Vec<V>::load(E_prev + (long long)pdf * S + s);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += ((al[u].v[j] + tr[u].w * lt[j]) * tr[u].x) * e[u].v[j];
      }
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) { o.v[j] = acc[j] * inv[j]; part[j] += o.v[j]; }
      o.store(alpha_cur + (long long)h * S + s);
    }
  }
  // block reduction of the per-sequence totals over the states of this block
  __shared__ float red[kDenThreads * V];
#pragma unroll
  for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = active ? part[j] : 0.f;
  __syncthreads();
  if (threadIdx.x < SV) {
    float sum[V];
#pragma unroll
    for (int j = 0; j < V; ++j) sum[j] = 0.f;
    for (int k = 0; k < tps; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) sum[j] += red[(k * SV + threadIdx.x) * V + j];
    red_add_vec<V>(tot_cur + s, sum);
  }
}

// tot_prob[s] = sum_h alpha'(T,h,s) = tot(T,s) * (1 + leaky*sum(init));  logprob = sum_s log tot_prob + sum_{t<T,s} log tot(t,s)
// Also seeds the backward pass: betad(T,h,s) = 1/tot_prob[s]  =>  bsum(T,s) = sum(init)/tot_prob[s].
__global__ void den_loglike_kernel(const float* __restrict__ tot, int T, int S, float leaky, float init_sum,
                                   float* __restrict__ tot_prob, double* __restrict__ scalars) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ double red[256];
  double acc = 0.0;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float tp = tot[(long long)T * S + s] + leaky * init_sum * tot[(long long)T * S + s];
    tot_prob[s] = tp;
    acc += (double)logf(tp);
  }
  for (long long i = threadIdx.x; i < (long long)T * S; i += blockDim.x) acc += (double)logf(tot[i]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scalars[0] = red[0];
}

__global__ void den_beta_last_kernel(const float* __restrict__ tot_prob, int N, int S, float init_sum,
                                     float* __restrict__ betad, float* __restrict__ bsum) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    betad[i] = 1.0f / tot_prob[i % S];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
    bsum[i] = init_sum / tot_prob[i];
}

// One backward frame (kaldi: BetaDashGeneralFrame + Beta): thread = (source state h, V sequences).
//   vf = p * beta(t+1,g,s) * E(t,pdf,s);  gamma(t,pdf,s) += vf * alpha'(t,h,s)/tot(t,s);  betad(t,h,s) = sum vf / tot(t,s)
template <int V, int U>
__global__ void __launch_bounds__(kDenThreads)
den_beta_frame_kernel(const int2* __restrict__ fwd_ranges, const float4* __restrict__ trans, const int* __restrict__ order,
                      const float* __restrict__ init, int N, int spb, int S, float leaky, const float* __restrict__ alpha_t,
                      const float* __restrict__ tot_t, const float* __restrict__ E_t,
                      const float* __restrict__ betad_next, const float* __restrict__ bsum_next,
                      float* __restrict__ betad_cur, float* __restrict__ bsum_cur, float* __restrict__ gamma_t,
                      double* __restrict__ check /* null unless t == 0 */) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const int SV = S / V;
  const int tps = kDenThreads / SV;
  const int sv = threadIdx.x % SV;
  const int hl = threadIdx.x / SV;
  const int s = sv * V;
  const bool active = hl < tps;
  float lt[V], inv[V], lb[V], part[V], chk[V];
  {
    const Vec<V> tp = Vec<V>::load(tot_t + s);
    const Vec<V> bs = Vec<V>::load(bsum_next + s);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      lt[j] = leaky * tp.v[j];
      inv[j] = 1.0f / tp.v[j];
      lb[j] = leaky * bs.v[j];
      part[j] = 0.f;
      chk[j] = 0.f;
    }
  }
  if (active) {
    for (int pos = blockIdx.x * spb + hl; pos < min((blockIdx.x + 1) * spb, N); pos += tps) {
      const int h = order[pos];
      const int2 rg = fwd_ranges[h];
      const float ih = init[h];
      const Vec<V> al = Vec<V>::load(alpha_t + (long long)h * S + s);
      float occ[V], totv[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { occ[j] = (al.v[j] + ih * lt[j]) * inv[j]; totv[j] = 0.f; }
      for (int a = rg.x; a < rg.y; a += U) {
        float4 tr[U];
        Vec<V> b[U], e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) tr[u] = (a + u < rg.y) ? trans[a + u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pdf = __float_as_int(tr[u].y), g = __float_as_int(tr[u].z);
          b[u] = Vec<V>::load(betad_next + (long long)g * S + s);
          e[u] = Vec<V>::load(E_t + (long long)pdf * S + s);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (a + u < rg.y) {
            const int pdf = __float_as_int(tr[u].y);
            float op[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const float vf = (tr[u].x * (b[u].v[j] + lb[j])) * e[u].v[j];
              totv[j] += vf;
              op[j] = vf * occ[j];
            }
            red_add_vec<V>(gamma_t + (long long)pdf * S + s, op);
          }
        }
      }
      Vec<V> o;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        o.v[j] = totv[j] * inv[j];
        part[j] += ih * o.v[j];
        chk[j] += (al.v[j] + ih * lt[j]) * o.v[j];  // alpha'(t,h,s) * betad(t,h,s)
      }
      o.store(betad_cur + (long long)h * S + s);
    }
  }
  __shared__ float red[kDenThreads * V];
#pragma unroll
  for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = active ? part[j] : 0.f;
  __syncthreads();
  if (threadIdx.x < SV) {
    float sum[V];
#pragma unroll
    for (int j = 0; j < V; ++j) sum[j] = 0.f;
    for (int k = 0; k < tps; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) sum[j] += red[(k * SV + threadIdx.x) * V + j];
    red_add_vec<V>(bsum_cur + s, sum);
  }
  if (check != nullptr) {
    __syncthreads();
    float c = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) c += active ? chk[j] : 0.f;
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = kDenThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(check, (double)red[0]);
  }
}

// nnet_output_deriv[t*S+s][p] += w * gamma[t][p][s]   (32x32 tiled transpose-add)
__global__ void den_deriv_transpose_add_kernel(const float* __restrict__ gamma, int S, int P, float w,
                                               float* __restrict__ deriv, long long ld) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && p < P) ? gamma[((long long)t * P + p) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int s = s0 + i, p = p0 + threadIdx.x;
    if (s < S && p < P) deriv[((long long)t * S + s) * ld + p] += w * tile[threadIdx.x][i];
  }
}

// =====================================================================================================
// Resident path -- an EXPERIMENT, off by default (TDNNF_DEN_RESIDENT=1 enables it; parity-tested, measured slower).
// The per-frame kernels above gather alpha(t-1, src, :) and E(t-1, pdf, :) rows out of L2 for every arc (8 B per arc,
// sequence and frame: ~25x the algorithmic HBM bytes, measured 4.8-7.4 TB/s of L2 traffic).  Here a cluster of C CTAs
// owns V sequences for the WHOLE recursion and keeps alpha'(t-1, :, v) and E(t-1, :, v) of those sequences in shared
// memory ((N + P) V floats; (N + 2P) V in the backward pass, which also accumulates the frame's posteriors there):
// no per-frame launches, no global posterior atomics per arc.  The C CTAs split the warp-tasks of the plan, publish
// their slice of alpha(t) / beta'(t) to global memory and meet at a cluster barrier per frame.
// Measured on B200 (tools/den_sweep.py, N = 16384, S = 128, T = 100): 23.0 ms against 8.8 ms for the per-frame kernels.
// Shared memory caps V at 2 for Switchboard-sized graphs, so (i) every CTA re-streams its share of the arc list from
// L2 each frame (8 B per arc for 2 sequences: no better than the 8 B per arc and sequence it replaces once V = 2),
// (ii) the random 8-byte gathers cost ~6 shared-memory wavefronts per warp instruction (bank conflicts), a floor of
// ~13 us per frame even with perfect latency hiding, and (iii) fp32 shared-memory atomics are CAS loops
// (ATOMS.CAST.SPIN).  Kept for graphs small enough for V = 4 and as the record of why the L2-gather design stays.
// =====================================================================================================
namespace cg = cooperative_groups;

constexpr int kDen2Threads = 512;

// E[t][sb][p][v] = exp(clamp(x[t*S + sb*V + v][p]))
__global__ void den2_exp_kernel(const float* __restrict__ x, long long ld, int S, int P, int V, float* __restrict__ E) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int s = s0 + i, p = p0 + threadIdx.x;
    float v = 0.f;
    if (s < S && p < P) v = expf(fminf(fmaxf(x[((long long)t * S + s) * ld + p], -30.f), 30.f));
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    if (s < S && p < P) E[(((long long)t * (S / V) + s / V) * P + p) * V + s % V] = tile[threadIdx.x][i];
  }
}

// deriv[t*S + s][p] += w * gamma[t][sb][p][v]
__global__ void den2_deriv_kernel(const float* __restrict__ gamma, int S, int P, int V, float w, float* __restrict__ deriv,
                                  long long ld) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && p < P) ? gamma[(((long long)t * (S / V) + s / V) * P + p) * V + s % V] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int s = s0 + i, p = p0 + threadIdx.x;
    if (s < S && p < P) deriv[((long long)t * S + s) * ld + p] += w * tile[threadIdx.x][i];
  }
}

// alpha(0, h, :) = init[h] in the [sb][N][V] layout; tot(0, s) = sum(init)
__global__ void den2_alpha_first_kernel(const float* __restrict__ init, int N, int S, int V, float init_sum,
                                        float* __restrict__ alpha0, float* __restrict__ tot0) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    alpha0[i] = init[(i / V) % N];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
    tot0[i] = init_sum;
}

template <int V>
__device__ __forceinline__ void block_sum_to_global(float (&part)[V], float* red /* smem [V * warps] */, float* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part[v] += __shfl_xor_sync(0xffffffffu, part[v], o);
    if (lane == 0) red[v * nw + warp] = part[v];
  }
  __syncthreads();
  if (threadIdx.x < V) {
    float sum = 0.f;
    for (int w = 0; w < nw; ++w) sum += red[threadIdx.x * nw + w];
    atomicAdd(dst + threadIdx.x, sum);
  }
}

template <int V>
__global__ void __launch_bounds__(kDen2Threads, 1)
den2_forward_kernel(int num_tasks, const int* __restrict__ task_base, const int* __restrict__ task_len,
                    const int* __restrict__ task_state, const uint2* __restrict__ arcs, const float* __restrict__ init,
                    int N, int P, int S, int T, float leaky, int state_bits, const float* __restrict__ E,
                    float* __restrict__ alpha, float* __restrict__ tot) {
  extern __shared__ float den2_smem[];
  float* a_s = den2_smem;                   // alpha'(t-1, h, v)   [N][V]
  float* e_s = a_s + (size_t)N * V;         // E(t-1, p, v)        [P][V]
  float* red = e_s + (size_t)P * V;         // [V][warps]
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int sb = blockIdx.x / C;
  const int SB = S / V;
  const int s0 = sb * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t smask = (1u << state_bits) - 1u;
  for (int t = 1; t <= T; ++t) {
    const float* a_prev = alpha + ((size_t)(t - 1) * SB + sb) * N * V;
    float* a_cur = alpha + ((size_t)t * SB + sb) * N * V;
    const float* e_prev = E + ((size_t)(t - 1) * SB + sb) * P * V;
    float lt[V], inv[V], part[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float tp = __ldcg(tot + (size_t)(t - 1) * S + s0 + v);
      lt[v] = leaky * tp;
      inv[v] = 1.0f / tp;
      part[v] = 0.f;
    }
    // stage alpha-dash and E of the previous frame (alpha was written by all CTAs of the cluster: L2 loads)
    for (int h = threadIdx.x; h < N; h += blockDim.x) {
      const float ih = init[h];
#pragma unroll
      for (int v = 0; v < V; ++v) a_s[h * V + v] = __ldcg(a_prev + (size_t)h * V + v) + ih * lt[v];
    }
    for (int i = threadIdx.x; i < P * V; i += blockDim.x) e_s[i] = e_prev[i];
    __syncthreads();
    for (int task = rank + C * warp; task < num_tasks; task += C * nw) {
      const int h = task_state[task * 32 + lane];
      const int len = task_len[task];
      const uint2* ap = arcs + task_base[task] + lane;
      float acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll 4
      for (int k = 0; k < len; ++k) {
        const uint2 arc = ap[k * 32];
        const float p = __uint_as_float(arc.x);
        const uint32_t g = arc.y & smask, pdf = arc.y >> state_bits;
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += (a_s[g * V + v] * p) * e_s[pdf * V + v];
      }
      if (h >= 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float o = acc[v] * inv[v];
          a_cur[(size_t)h * V + v] = o;
          part[v] += o;
        }
      }
    }
    block_sum_to_global<V>(part, red, tot + (size_t)t * S + s0);
    cluster.sync();  // alpha(t) and tot(t) of every CTA visible to the cluster; shared memory free for the next frame
  }
}

template <int V>
__global__ void __launch_bounds__(kDen2Threads, 1)
den2_backward_kernel(int num_tasks, const int* __restrict__ task_base, const int* __restrict__ task_len,
                     const int* __restrict__ task_state, const uint2* __restrict__ arcs, const float* __restrict__ init,
                     int N, int P, int S, int T, float leaky, float init_sum, int state_bits, const float* __restrict__ E,
                     const float* __restrict__ alpha, const float* __restrict__ tot, const float* __restrict__ tot_prob,
                     float* __restrict__ betad /* [2][SB][N][V] */, float* __restrict__ bsum /* [T+1][S], zeroed */,
                     float* __restrict__ gamma /* [T][SB][P][V], zeroed */, double* __restrict__ check) {
  extern __shared__ float den2_smem[];
  float* b_s = den2_smem;                   // beta(t+1, g, v) = beta'(t+1, g, v) + leaky * bsum(t+1, v)   [N][V]
  float* e_s = b_s + (size_t)N * V;         // E(t, p, v)
  float* g_s = e_s + (size_t)P * V;         // posteriors of this frame, this CTA's arcs
  float* red = g_s + (size_t)P * V;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int sb = blockIdx.x / C;
  const int SB = S / V;
  const int s0 = sb * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t smask = (1u << state_bits) - 1u;
  float chk = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const float* bn = betad + ((size_t)((t + 1) & 1) * SB + sb) * N * V;
    float* bc = betad + ((size_t)(t & 1) * SB + sb) * N * V;
    const float* a_t = alpha + ((size_t)t * SB + sb) * N * V;
    const float* e_t = E + ((size_t)t * SB + sb) * P * V;
    float* g_t = gamma + ((size_t)t * SB + sb) * P * V;
    float lt[V], inv[V], part[V], lb[V], last[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float tp = __ldcg(tot + (size_t)t * S + s0 + v);
      lt[v] = leaky * tp;
      inv[v] = 1.0f / tp;
      part[v] = 0.f;
      // beta'(T, :, s) = 1 / tot_prob[s], bsum(T, s) = sum(init) / tot_prob[s]
      last[v] = 1.0f / tot_prob[s0 + v];
      lb[v] = leaky * ((t == T - 1) ? init_sum * last[v] : __ldcg(bsum + (size_t)(t + 1) * S + s0 + v));
    }
    for (int h = threadIdx.x; h < N; h += blockDim.x) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        b_s[h * V + v] = ((t == T - 1) ? last[v] : __ldcg(bn + (size_t)h * V + v)) + lb[v];
    }
    for (int i = threadIdx.x; i < P * V; i += blockDim.x) {
      e_s[i] = e_t[i];
      g_s[i] = 0.f;
    }
    __syncthreads();
    for (int task = rank + C * warp; task < num_tasks; task += C * nw) {
      const int h = task_state[task * 32 + lane];
      const int len = task_len[task];
      const uint2* ap = arcs + task_base[task] + lane;
      float occ[V], totv[V], ad[V];
      const float ih = h >= 0 ? init[h] : 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        ad[v] = h >= 0 ? (a_t[(size_t)h * V + v] + ih * lt[v]) : 0.f;  // alpha'(t, h, v)
        occ[v] = ad[v] * inv[v];
        totv[v] = 0.f;
      }
#pragma unroll 2
      for (int k = 0; k < len; ++k) {
        const uint2 arc = ap[k * 32];
        const float p = __uint_as_float(arc.x);
        if (p != 0.f) {  // padding arcs carry probability 0
          const uint32_t g = arc.y & smask, pdf = arc.y >> state_bits;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float vf = (p * b_s[g * V + v]) * e_s[pdf * V + v];
            totv[v] += vf;
            atomicAdd(&g_s[pdf * V + v], vf * occ[v]);
          }
        }
      }
      if (h >= 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float o = totv[v] * inv[v];
          bc[(size_t)h * V + v] = o;
          part[v] += ih * o;
          chk += ad[v] * o;  // alpha'(t, h, s) * beta'(t, h, s): summed at t == 0 only (below)
        }
      }
    }
    if (t != 0) chk = 0.f;
    block_sum_to_global<V>(part, red, bsum + (size_t)t * S + s0);  // has a __syncthreads: g_s is complete after it
    for (int i = threadIdx.x; i < P * V; i += blockDim.x) {
      const float gv = g_s[i];
      if (gv != 0.f) atomicAdd(g_t + i, gv);
    }
    cluster.sync();
  }
  if (check != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(0xffffffffu, chk, o);
    if (lane == 0 && chk != 0.f) atomicAdd(check, (double)chk);
  }
}

// Builds one direction of the plan on the host.  ranges: [N][2] into (prob, pdf, state).
int build_den_plan(const std::vector<int>& ranges, const std::vector<float>& prob, const std::vector<int>& pdf,
                   const std::vector<int>& state, int N, int state_bits, DenPlan* plan) {
  std::vector<int> order(N);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return ranges[2 * a + 1] - ranges[2 * a] > ranges[2 * b + 1] - ranges[2 * b];
  });
  const int tasks = (N + 31) / 32;
  std::vector<int> base(tasks), len(tasks), st((size_t)tasks * 32, -1);
  size_t total = 0;
  for (int t = 0; t < tasks; ++t) {
    const int h0 = order[t * 32];
    len[t] = ranges[2 * h0 + 1] - ranges[2 * h0];  // sorted: the first lane has the longest list
    base[t] = (int)total;
    total += (size_t)len[t] * 32;
  }
  if (total > (size_t)INT32_MAX) return fail(TDNNF_ERR_UNSUPPORTED, "denominator graph too large for the resident plan");
  std::vector<uint2> arcs(std::max<size_t>(total, 1), make_uint2(0u, 0u));
  for (int t = 0; t < tasks; ++t) {
    for (int l = 0; l < 32 && t * 32 + l < N; ++l) {
      const int h = order[t * 32 + l];
      st[(size_t)t * 32 + l] = h;
      int k = 0;
      for (int a = ranges[2 * h]; a < ranges[2 * h + 1]; ++a, ++k) {
        uint32_t bits;
        memcpy(&bits, &prob[a], 4);
        arcs[(size_t)base[t] + (size_t)k * 32 + l] = make_uint2(bits, ((uint32_t)pdf[a] << state_bits) | (uint32_t)state[a]);
      }
    }
  }
  plan->num_tasks = tasks;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up(reinterpret_cast<void**>(&plan->task_base), base.data(), sizeof(int) * tasks);
  up(reinterpret_cast<void**>(&plan->task_len), len.data(), sizeof(int) * tasks);
  up(reinterpret_cast<void**>(&plan->task_state), st.data(), sizeof(int) * st.size());
  up(reinterpret_cast<void**>(&plan->arcs), arcs.data(), sizeof(uint2) * arcs.size());
  if (e != cudaSuccess) return fail(TDNNF_ERR_CUDA, std::string("den plan upload failed: ") + cudaGetErrorString(e));
  return TDNNF_OK;
}

template <typename Kern, typename... Args>
cudaError_t launch_cluster(Kern kern, int grid, int cluster, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kDen2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

void frame_grid(int N, int num_sms, int* blocks, int* spb) {
  (void)num_sms;
  *spb = kStatesPerBlock;
  *blocks = (N + kStatesPerBlock - 1) / kStatesPerBlock;
}

int pick_vec(int S) {
  // V consecutive sequences per thread, lanes along s: the widest loads that still leave >= 16 lanes per state.
  // Measured at N = 16384, S = 64, T = 50: V = 4 (4 arcs per group) 2.43 ms, V = 2 with 8 arcs per group (same bytes in
  // flight per thread, twice the threads, 98 registers) 3.12 ms, V = 1 4.71 ms.  TDNNF_DEN_VEC overrides.
  static const int forced = [] {
    const char* e = getenv("TDNNF_DEN_VEC");
    return e ? atoi(e) : 0;
  }();
  if ((forced == 4 || forced == 2 || forced == 1) && S % forced == 0) return forced;
  if (S % 4 == 0 && S >= 64) return 4;
  if (S % 2 == 0 && S >= 32) return 2;
  return 1;
}

}  // namespace

extern "C" int tdnnf_den_graph_create(tdnnf_ctx* ctx, int num_states, int num_pdfs, int num_transitions,
                                      const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* trans_prob,
                                      const int32_t* trans_pdf, const int32_t* trans_state, const float* initial_probs,
                                      tdnnf_den_graph** out) {
  TDNNF_REQUIRE(ctx && fwd_ranges && bwd_ranges && trans_prob && trans_pdf && trans_state && initial_probs && out,
                "null argument");
  TDNNF_REQUIRE(num_states > 0 && num_pdfs > 0 && num_transitions > 0, "empty graph");
  for (int h = 0; h < num_states; ++h) {
    TDNNF_REQUIRE(fwd_ranges[2 * h] >= 0 && fwd_ranges[2 * h] <= fwd_ranges[2 * h + 1] &&
                      fwd_ranges[2 * h + 1] <= num_transitions && bwd_ranges[2 * h] >= 0 &&
                      bwd_ranges[2 * h] <= bwd_ranges[2 * h + 1] && bwd_ranges[2 * h + 1] <= num_transitions,
                  "transition range out of bounds");
  }
  for (int a = 0; a < num_transitions; ++a)
    TDNNF_REQUIRE(trans_pdf[a] >= 0 && trans_pdf[a] < num_pdfs && trans_state[a] >= 0 && trans_state[a] < num_states,
                  "transition pdf-id / hmm-state out of range");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  tdnnf_den_graph* g = new tdnnf_den_graph();
  g->ctx = ctx;
  g->num_states = num_states;
  g->num_pdfs = num_pdfs;
  g->num_transitions = num_transitions;
  std::vector<float4> tr(num_transitions);
  for (int a = 0; a < num_transitions; ++a) {
    tr[a].x = trans_prob[a];
    memcpy(&tr[a].y, &trans_pdf[a], 4);
    memcpy(&tr[a].z, &trans_state[a], 4);
    tr[a].w = initial_probs[trans_state[a]];  // saves a dependent load per arc in the forward recursion
  }
  float isum = 0.f;
  for (int h = 0; h < num_states; ++h) isum += initial_probs[h];
  g->init_sum = isum;
  g->h_fwd_ranges.assign(fwd_ranges, fwd_ranges + 2 * (size_t)num_states);
  g->h_bwd_ranges.assign(bwd_ranges, bwd_ranges + 2 * (size_t)num_states);
  g->h_prob.assign(trans_prob, trans_prob + num_transitions);
  g->h_pdf.assign(trans_pdf, trans_pdf + num_transitions);
  g->h_state.assign(trans_state, trans_state + num_transitions);
  g->h_init.assign(initial_probs, initial_probs + num_states);
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up(reinterpret_cast<void**>(&g->fwd_ranges), fwd_ranges, sizeof(int2) * num_states);
  up(reinterpret_cast<void**>(&g->bwd_ranges), bwd_ranges, sizeof(int2) * num_states);
  up(reinterpret_cast<void**>(&g->trans), tr.data(), sizeof(float4) * num_transitions);
  up(reinterpret_cast<void**>(&g->init), initial_probs, sizeof(float) * num_states);
  {
    std::vector<int> order(num_states);
    auto by_len = [&](const int32_t* ranges) {
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return ranges[2 * a + 1] - ranges[2 * a] > ranges[2 * b + 1] - ranges[2 * b];
      });
    };
    by_len(bwd_ranges);
    up(reinterpret_cast<void**>(&g->order_in), order.data(), sizeof(int) * num_states);
    by_len(fwd_ranges);
    up(reinterpret_cast<void**>(&g->order_out), order.data(), sizeof(int) * num_states);
  }
  if (e != cudaSuccess) {
    tdnnf_den_graph_destroy(g);
    return fail(TDNNF_ERR_CUDA, std::string("den graph upload failed: ") + cudaGetErrorString(e));
  }
  *out = g;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_graph_destroy(tdnnf_den_graph* g) {
  if (!g) return TDNNF_OK;
  cudaFree(g->fwd_ranges);
  cudaFree(g->bwd_ranges);
  cudaFree(g->trans);
  cudaFree(g->init);
  cudaFree(g->order_in);
  cudaFree(g->order_out);
  delete g;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_create(tdnnf_ctx* ctx, const tdnnf_den_graph* g, int num_seqs, int frames_per_seq,
                                float leaky_hmm_coefficient, tdnnf_den_comp** out) {
  TDNNF_REQUIRE(ctx && g && out, "null argument");
  TDNNF_REQUIRE(num_seqs > 0 && frames_per_seq > 0, "empty minibatch");
  TDNNF_REQUIRE(num_seqs <= 1024, "num_seqs > 1024 is not supported");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  tdnnf_den_comp* c = new tdnnf_den_comp();
  c->ctx = ctx;
  c->g = g;
  c->S = num_seqs;
  c->T = frames_per_seq;
  c->leaky = leaky_hmm_coefficient;
  const size_t N = g->num_states, P = g->num_pdfs, S = num_seqs, T = frames_per_seq;
  // Resident path (opt-in experiment, TDNNF_DEN_RESIDENT=1): V sequences per cluster such that (N + 2P) V floats
  // (backward pass) fit in shared memory, and the packed arc word holds a state and a pdf-id.
  {
    int sbits = 1, pbits = 1;
    while ((1ull << sbits) < N) ++sbits;
    while ((1ull << pbits) < P) ++pbits;
    const char* env = getenv("TDNNF_DEN_RESIDENT");
    const bool allowed = (env && env[0] == '1') && sbits + pbits <= 32;
    const size_t smem_max = 232448 - 1024;
    int V = 0;
    for (int v : {4, 2, 1}) {
      if (S % v == 0 && (N + 2 * P) * v * sizeof(float) + 4 * (kDen2Threads / 32) * sizeof(float) <= smem_max) {
        V = v;
        break;
      }
    }
    if (allowed && V > 0) {
      c->resident = true;
      c->V = V;
      c->state_bits = sbits;
      const int clusters = (int)(S / V);
      int C = 1;
      while (C < 8 && clusters * C * 2 <= ctx->num_sms && (size_t)C * 2 * 32 <= N) C *= 2;
      c->C = C;
      const size_t red = (size_t)V * (kDen2Threads / 32) * sizeof(float);
      c->smem_fwd = (N + P) * V * sizeof(float) + red;
      c->smem_bwd = (N + 2 * P) * V * sizeof(float) + red;
      int rc = build_den_plan(g->h_bwd_ranges, g->h_prob, g->h_pdf, g->h_state, (int)N, sbits, &c->plan_fwd);
      if (rc == TDNNF_OK) rc = build_den_plan(g->h_fwd_ranges, g->h_prob, g->h_pdf, g->h_state, (int)N, sbits, &c->plan_bwd);
      if (rc != TDNNF_OK) {
        tdnnf_den_destroy(c);
        return rc;
      }
      cudaError_t ea = cudaFuncSetAttribute(V == 4 ? den2_forward_kernel<4> : (V == 2 ? den2_forward_kernel<2> : den2_forward_kernel<1>),
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_fwd);
      if (ea == cudaSuccess)
        ea = cudaFuncSetAttribute(V == 4 ? den2_backward_kernel<4> : (V == 2 ? den2_backward_kernel<2> : den2_backward_kernel<1>),
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bwd);
      if (ea != cudaSuccess) {
        tdnnf_den_destroy(c);
        return fail(TDNNF_ERR_CUDA, std::string("cudaFuncSetAttribute (den2) failed: ") + cudaGetErrorString(ea));
      }
    }
  }
  if (!c->resident) {
    const int rc = den_slices_create(ctx, (int)N, (int)P, (int)S, (int)T, g->h_fwd_ranges, g->h_bwd_ranges, g->h_prob, g->h_pdf,
                                     g->h_state, g->h_init, &c->slices);
    if (rc != TDNNF_OK) {
      tdnnf_den_destroy(c);
      return rc;
    }
  }
  cudaError_t e = cudaSuccess;
  auto al = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  if (!c->slices) {  // the slice path keeps E / alpha / beta / gamma in its own layouts
    al(reinterpret_cast<void**>(&c->E), sizeof(float) * T * P * S);
    al(reinterpret_cast<void**>(&c->alpha), sizeof(float) * (T + 1) * N * S);
    al(reinterpret_cast<void**>(&c->betad), sizeof(float) * 2 * N * S);
    al(reinterpret_cast<void**>(&c->gamma), sizeof(float) * T * P * S);
  }
  al(reinterpret_cast<void**>(&c->tot), sizeof(float) * (T + 1) * S);
  al(reinterpret_cast<void**>(&c->bsum), sizeof(float) * (c->resident ? (T + 1) : 2) * S);
  al(reinterpret_cast<void**>(&c->tot_prob), sizeof(float) * S);
  al(reinterpret_cast<void**>(&c->scalars), sizeof(double) * 2);
  if (e != cudaSuccess) {
    tdnnf_den_destroy(c);
    return fail(TDNNF_ERR_NOMEM, std::string("denominator workspace allocation failed: ") + cudaGetErrorString(e));
  }
  *out = c;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_destroy(tdnnf_den_comp* c) {
  if (!c) return TDNNF_OK;
  cudaFree(c->E);
  cudaFree(c->alpha);
  cudaFree(c->tot);
  cudaFree(c->betad);
  cudaFree(c->bsum);
  cudaFree(c->gamma);
  cudaFree(c->tot_prob);
  cudaFree(c->scalars);
  c->plan_fwd.destroy();
  c->plan_bwd.destroy();
  den_slices_destroy(c->slices);
  delete c;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_describe(const tdnnf_den_comp* c, int* path, int* cluster, int* parts, int* ctas) {
  TDNNF_REQUIRE(c && path && cluster && parts && ctas, "null argument");
  *path = 0;
  *cluster = *parts = *ctas = 0;
  if (c->slices) {
    *path = 2;
    den_slices_describe(c->slices, cluster, parts, ctas);
  } else if (c->resident) {
    *path = 1;
    *cluster = c->C;
    *ctas = (c->S / c->V) * c->C;
  }
  return TDNNF_OK;
}

#define DEN_LAUNCH_CHECK(ctx)            \
  do {                                   \
    (ctx)->launches++;                   \
    TDNNF_CUDA_OK(cudaGetLastError());   \
  } while (0)

extern "C" int tdnnf_den_forward(tdnnf_den_comp* c, const float* nnet_output, int stride, float* logprob) {
  TDNNF_REQUIRE(c && nnet_output && logprob, "null argument");
  const tdnnf_den_graph* g = c->g;
  tdnnf_ctx* ctx = c->ctx;
  const int N = g->num_states, P = g->num_pdfs, S = c->S, T = c->T;
  TDNNF_REQUIRE(stride >= P, "stride < num_pdfs");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (c->slices) {
    int rc = den_slices_forward(ctx, c->slices, nnet_output, stride, g->init, g->init_sum, N, P, S, T, c->leaky, c->tot);
    if (rc) return rc;
    TDNNF_CUDA_OK(launch_pdl(den_loglike_kernel, dim3(1), dim3(256), 0, st, 1, c->tot, T, S, c->leaky, g->init_sum, c->tot_prob, c->scalars));
    DEN_LAUNCH_CHECK(ctx);
    double lp = 0.0;
    TDNNF_CUDA_OK(cudaMemcpyAsync(&lp, c->scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
    TDNNF_CUDA_OK(cudaStreamSynchronize(st));
    *logprob = (float)lp;
    c->forward_done = true;
    return TDNNF_OK;
  }
  if (c->resident) {
    const int V = c->V, C = c->C;
    TDNNF_CUDA_OK(launch_pdl(den2_exp_kernel, dim3((P + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st, 1, nnet_output, stride, S, P, V, c->E));
    DEN_LAUNCH_CHECK(ctx);
    TDNNF_CUDA_OK(cudaMemsetAsync(c->tot, 0, sizeof(float) * (size_t)(T + 1) * S, st));
    TDNNF_CUDA_OK(launch_pdl(den2_alpha_first_kernel, dim3(ctx->num_sms * 4), dim3(256), 0, st, 1, g->init, N, S, V, g->init_sum, c->alpha, c->tot));
    DEN_LAUNCH_CHECK(ctx);
    const DenPlan& pl = c->plan_fwd;
    const int grid = (S / V) * C;
    cudaError_t le;
    if (V == 4)
      le = launch_cluster(den2_forward_kernel<4>, grid, C, c->smem_fwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, c->state_bits,
                          (const float*)c->E, c->alpha, c->tot);
    else if (V == 2)
      le = launch_cluster(den2_forward_kernel<2>, grid, C, c->smem_fwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, c->state_bits,
                          (const float*)c->E, c->alpha, c->tot);
    else
      le = launch_cluster(den2_forward_kernel<1>, grid, C, c->smem_fwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, c->state_bits,
                          (const float*)c->E, c->alpha, c->tot);
    TDNNF_CUDA_OK(le);
    DEN_LAUNCH_CHECK(ctx);
    TDNNF_CUDA_OK(launch_pdl(den_loglike_kernel, dim3(1), dim3(256), 0, st, 1, c->tot, T, S, c->leaky, g->init_sum, c->tot_prob, c->scalars));
    DEN_LAUNCH_CHECK(ctx);
    double lp = 0.0;
    TDNNF_CUDA_OK(cudaMemcpyAsync(&lp, c->scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
    TDNNF_CUDA_OK(cudaStreamSynchronize(st));
    *logprob = (float)lp;
    c->forward_done = true;
    return TDNNF_OK;
  }
  TDNNF_CUDA_OK(launch_pdl(den_exp_transpose_kernel, dim3((P + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st, 1, nnet_output, stride, S, P, c->E));
  DEN_LAUNCH_CHECK(ctx);
  TDNNF_CUDA_OK(cudaMemsetAsync(c->tot, 0, sizeof(float) * (size_t)(T + 1) * S, st));
  TDNNF_CUDA_OK(launch_pdl(den_alpha_first_kernel, dim3(ctx->num_sms * 4), dim3(256), 0, st, 1, g->init, N, S, g->init_sum, c->alpha, c->tot));
  DEN_LAUNCH_CHECK(ctx);
  const int V = pick_vec(S);
  int blocks, spb;
  frame_grid(N, ctx->num_sms, &blocks, &spb);
  TDNNF_REQUIRE(S / V <= kDenThreads, "num_seqs too large for the frame kernels");
  for (int t = 1; t <= T; ++t) {
    const float* ap = c->alpha + (size_t)(t - 1) * N * S;
    float* ac = c->alpha + (size_t)t * N * S;
    const float* tp = c->tot + (size_t)(t - 1) * S;
    float* tc = c->tot + (size_t)t * S;
    const float* Ep = c->E + (size_t)(t - 1) * P * S;
    if (V == 4)
      TDNNF_CUDA_OK(launch_pdl(den_alpha_frame_kernel<4, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->bwd_ranges, g->trans, g->order_in, g->init, N, spb, S, c->leaky, ap, tp, Ep, ac, tc));
    else if (V == 2)
      TDNNF_CUDA_OK(launch_pdl(den_alpha_frame_kernel<2, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->bwd_ranges, g->trans, g->order_in, g->init, N, spb, S, c->leaky, ap, tp, Ep, ac, tc));
    else
      TDNNF_CUDA_OK(launch_pdl(den_alpha_frame_kernel<1, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->bwd_ranges, g->trans, g->order_in, g->init, N, spb, S, c->leaky, ap, tp, Ep, ac, tc));
    DEN_LAUNCH_CHECK(ctx);
  }
  TDNNF_CUDA_OK(launch_pdl(den_loglike_kernel, dim3(1), dim3(256), 0, st, 1, c->tot, T, S, c->leaky, g->init_sum, c->tot_prob, c->scalars));
  DEN_LAUNCH_CHECK(ctx);
  double lp = 0.0;
  TDNNF_CUDA_OK(cudaMemcpyAsync(&lp, c->scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
  TDNNF_CUDA_OK(cudaStreamSynchronize(st));
  *logprob = (float)lp;
  c->forward_done = true;
  return TDNNF_OK;
}

extern "C" int tdnnf_den_backward(tdnnf_den_comp* c, float deriv_weight, float* nnet_output_deriv, int stride, int* ok) {
  TDNNF_REQUIRE(c && nnet_output_deriv && ok, "null argument");
  TDNNF_REQUIRE(c->forward_done, "Backward() called before Forward()");
  const tdnnf_den_graph* g = c->g;
  tdnnf_ctx* ctx = c->ctx;
  const int N = g->num_states, P = g->num_pdfs, S = c->S, T = c->T;
  TDNNF_REQUIRE(stride >= P, "stride < num_pdfs");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  TDNNF_CUDA_OK(cudaMemsetAsync(c->scalars + 1, 0, sizeof(double), st));
  if (c->slices) {
    int rc = den_slices_backward(ctx, c->slices, g->init, g->init_sum, N, P, S, T, c->leaky, c->tot, c->tot_prob, deriv_weight,
                                 nnet_output_deriv, stride, c->scalars + 1);
    if (rc) return rc;
    double chk = 0.0;
    TDNNF_CUDA_OK(cudaMemcpyAsync(&chk, c->scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    TDNNF_CUDA_OK(cudaStreamSynchronize(st));
    *ok = (chk == chk && fabs(chk - (double)S) <= 2.0) ? 1 : 0;
    return TDNNF_OK;
  }
  TDNNF_CUDA_OK(cudaMemsetAsync(c->gamma, 0, sizeof(float) * (size_t)T * P * S, st));
  if (c->resident) {
    const int V = c->V, C = c->C;
    TDNNF_CUDA_OK(cudaMemsetAsync(c->bsum, 0, sizeof(float) * (size_t)(T + 1) * S, st));
    const DenPlan& pl = c->plan_bwd;
    const int grid = (S / V) * C;
    cudaError_t le;
    if (V == 4)
      le = launch_cluster(den2_backward_kernel<4>, grid, C, c->smem_bwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, g->init_sum,
                          c->state_bits, (const float*)c->E, (const float*)c->alpha, (const float*)c->tot, (const float*)c->tot_prob,
                          c->betad, c->bsum, c->gamma, c->scalars + 1);
    else if (V == 2)
      le = launch_cluster(den2_backward_kernel<2>, grid, C, c->smem_bwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, g->init_sum,
                          c->state_bits, (const float*)c->E, (const float*)c->alpha, (const float*)c->tot, (const float*)c->tot_prob,
                          c->betad, c->bsum, c->gamma, c->scalars + 1);
    else
      le = launch_cluster(den2_backward_kernel<1>, grid, C, c->smem_bwd, st, pl.num_tasks, (const int*)pl.task_base, (const int*)pl.task_len,
                          (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)g->init, N, P, S, T, c->leaky, g->init_sum,
                          c->state_bits, (const float*)c->E, (const float*)c->alpha, (const float*)c->tot, (const float*)c->tot_prob,
                          c->betad, c->bsum, c->gamma, c->scalars + 1);
    TDNNF_CUDA_OK(le);
    DEN_LAUNCH_CHECK(ctx);
    TDNNF_CUDA_OK(launch_pdl(den2_deriv_kernel, dim3((P + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st, 1, c->gamma, S, P, V, deriv_weight,
                                                                                    nnet_output_deriv, stride));
    DEN_LAUNCH_CHECK(ctx);
    double chk = 0.0;
    TDNNF_CUDA_OK(cudaMemcpyAsync(&chk, c->scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    TDNNF_CUDA_OK(cudaStreamSynchronize(st));
    *ok = (chk == chk && fabs(chk - (double)S) <= 2.0) ? 1 : 0;
    return TDNNF_OK;
  }
  TDNNF_CUDA_OK(launch_pdl(den_beta_last_kernel, dim3(ctx->num_sms * 4), dim3(256), 0, st, 1, c->tot_prob, N, S, g->init_sum, c->betad + (size_t)(T & 1) * N * S,
                                                         c->bsum + (size_t)(T & 1) * S));
  DEN_LAUNCH_CHECK(ctx);
  const int V = pick_vec(S);
  int blocks, spb;
  frame_grid(N, ctx->num_sms, &blocks, &spb);
  for (int t = T - 1; t >= 0; --t) {
    const float* bn = c->betad + (size_t)((t + 1) & 1) * N * S;
    const float* sn = c->bsum + (size_t)((t + 1) & 1) * S;
    float* bc = c->betad + (size_t)(t & 1) * N * S;
    float* sc = c->bsum + (size_t)(t & 1) * S;
    TDNNF_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(float) * S, st));
    const float* at = c->alpha + (size_t)t * N * S;
    const float* tt = c->tot + (size_t)t * S;
    const float* Et = c->E + (size_t)t * P * S;
    float* gt = c->gamma + (size_t)t * P * S;
    double* chk = (t == 0) ? c->scalars + 1 : nullptr;
    if (V == 4)
      TDNNF_CUDA_OK(launch_pdl(den_beta_frame_kernel<4, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->fwd_ranges, g->trans, g->order_out, g->init, N, spb, S, c->leaky, at, tt, Et, bn, sn, bc, sc, gt, chk));
    else if (V == 2)
      TDNNF_CUDA_OK(launch_pdl(den_beta_frame_kernel<2, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->fwd_ranges, g->trans, g->order_out, g->init, N, spb, S, c->leaky, at, tt, Et, bn, sn, bc, sc, gt, chk));
    else
      TDNNF_CUDA_OK(launch_pdl(den_beta_frame_kernel<1, 4>, dim3(blocks), dim3(kDenThreads), 0, st, 1, g->fwd_ranges, g->trans, g->order_out, g->init, N, spb, S, c->leaky, at, tt, Et, bn, sn, bc, sc, gt, chk));
    DEN_LAUNCH_CHECK(ctx);
  }
  TDNNF_CUDA_OK(launch_pdl(den_deriv_transpose_add_kernel, dim3((P + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st, 1, c->gamma, S, P, deriv_weight,
                                                                                               nnet_output_deriv, stride));
  DEN_LAUNCH_CHECK(ctx);
  double chk = 0.0;
  TDNNF_CUDA_OK(cudaMemcpyAsync(&chk, c->scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
  TDNNF_CUDA_OK(cudaStreamSynchronize(st));
  // kaldi: BetaGeneralFrameDebug at t == 0: |sum alpha'.betad - num_sequences| > 2  =>  ok_ = false
  *ok = (chk == chk && fabs(chk - (double)S) <= 2.0) ? 1 : 0;
  return TDNNF_OK;
}
