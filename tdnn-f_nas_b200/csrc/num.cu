// Chain numerator, generic (per-sequence FST) form: log-domain forward-backward
// (kaldi: chain/chain-generic-numerator.cc -- upstream Kaldi, not shipped with the reference; SURVEY "next" row N3).
// One CTA per sequence: the FSTs are tiny (tens to hundreds of states), so alpha for all frames lives in a
// per-sequence global scratch strip and each frame is one pass of "thread = state, loop over its arcs" with a
// block barrier between frames; log-sum-exp in fp32 with a running max.  Posteriors go to nnet_output_deriv
// with atomicAdd (a handful per frame).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.h"
#include "ptx.cuh"

using namespace tdnnf;

struct tdnnf_num_graph {
  tdnnf_ctx* ctx = nullptr;
  int num_seqs = 0, num_states = 0, num_arcs = 0, max_states = 0;
  int* state_offsets = nullptr;  // device [S+1]
  int2* fwd_ranges = nullptr;    // device [num_states]
  int2* bwd_ranges = nullptr;
  float* arc_logprob = nullptr;  // device [2A]
  int* arc_pdf = nullptr;
  int* arc_state = nullptr;
  float* final_logprob = nullptr;  // device [num_states]
  float* alpha = nullptr;          // device scratch, grown on demand: [num_states][T+1] laid out per sequence
  size_t alpha_elems = 0;
  double* scalars = nullptr;  // [2]: total logprob, number of failed sequences
  // tdnnf_num_graph_update: capacities of the device arrays, a pinned staging buffer and the event of its last copy
  int cap_seqs = 0, cap_states = 0, cap_arcs = 0;
  char* staging = nullptr;
  size_t staging_bytes = 0;
  cudaEvent_t staged = nullptr;
};

namespace {

constexpr float kLogZero = -1.0e30f;
constexpr int kNumThreads = 128;

__device__ __forceinline__ float log_add(float a, float b) {
  if (a < b) { const float t = a; a = b; b = t; }
  if (b <= kLogZero) return a;
  return a + log1pf(expf(b - a));
}

// alpha strip of sequence s: [(T+1)][ns] with ns = number of states of that sequence.
__global__ void __launch_bounds__(kNumThreads)
num_fwd_bwd_kernel(const int* __restrict__ state_offsets, const int2* __restrict__ fwd_ranges,
                   const int2* __restrict__ bwd_ranges, const float* __restrict__ arc_logprob,
                   const int* __restrict__ arc_pdf, const int* __restrict__ arc_state,
                   const float* __restrict__ final_logprob, const float* __restrict__ x, long long x_stride, int S, int T,
                   float* __restrict__ alpha_all, float deriv_weight, float* deriv, long long d_stride,
                   double* __restrict__ scalars) {
  ptx::grid_dep_launch_dependents();
  ptx::grid_dep_wait();
  extern __shared__ float beta_sm[];  // [2][ns]
  __shared__ float red[kNumThreads];
  const int s = blockIdx.x;
  const int s0 = state_offsets[s], ns = state_offsets[s + 1] - s0;
  float* alpha = alpha_all + (size_t)s0 * (T + 1);
  // ---- forward
  for (int h = threadIdx.x; h < ns; h += blockDim.x) alpha[h] = (h == 0) ? 0.f : kLogZero;
  __syncthreads();
  for (int t = 1; t <= T; ++t) {
    const float* prev = alpha + (size_t)(t - 1) * ns;
    float* cur = alpha + (size_t)t * ns;
    const float* xrow = x + ((long long)(t - 1) * S + s) * x_stride;
    for (int h = threadIdx.x; h < ns; h += blockDim.x) {
      const int2 rg = bwd_ranges[s0 + h];
      float acc = kLogZero;
      for (int a = rg.x; a < rg.y; ++a) {
        const float p = prev[arc_state[a] - s0];
        if (p > kLogZero) acc = log_add(acc, p + arc_logprob[a] + xrow[arc_pdf[a]]);
      }
      cur[h] = acc;
    }
    __syncthreads();
  }
  // ---- total
  float tot = kLogZero;
  for (int h = threadIdx.x; h < ns; h += blockDim.x) {
    const float f = final_logprob[s0 + h], a = alpha[(size_t)T * ns + h];
    if (f > kLogZero && a > kLogZero) tot = log_add(tot, a + f);
  }
  red[threadIdx.x] = tot;
  __syncthreads();
  for (int o = kNumThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = log_add(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  tot = red[0];
  if (threadIdx.x == 0) {
    if (tot > kLogZero) atomicAdd(scalars, (double)tot);
    else atomicAdd(scalars + 1, 1.0);
  }
  if (deriv == nullptr || !(tot > kLogZero)) return;
  // ---- backward: beta(T,h) = final(h); posterior(arc,t) = exp(alpha(t,src) + w + x(t,pdf) + beta(t+1,dst) - tot)
  float* bnext = beta_sm;
  float* bcur = beta_sm + ns;
  for (int h = threadIdx.x; h < ns; h += blockDim.x) bnext[h] = final_logprob[s0 + h];
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* al = alpha + (size_t)t * ns;
    const float* xrow = x + ((long long)t * S + s) * x_stride;
    float* drow = deriv + ((long long)t * S + s) * d_stride;
    for (int h = threadIdx.x; h < ns; h += blockDim.x) {
      const int2 rg = fwd_ranges[s0 + h];
      const float a_h = al[h];
      float acc = kLogZero;
      for (int a = rg.x; a < rg.y; ++a) {
        const float b = bnext[arc_state[a] - s0];
        if (b <= kLogZero) continue;
        const int pdf = arc_pdf[a];
        const float v = arc_logprob[a] + xrow[pdf] + b;
        acc = log_add(acc, v);
        if (a_h > kLogZero) atomicAdd(drow + pdf, deriv_weight * expf(a_h + v - tot));
      }
      bcur[h] = acc;
    }
    __syncthreads();
    float* tmp = bnext; bnext = bcur; bcur = tmp;
  }
}

}  // namespace

extern "C" int tdnnf_num_graph_create(tdnnf_ctx* ctx, int num_seqs, const int32_t* state_offsets, int num_arcs,
                                      const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* arc_logprob,
                                      const int32_t* arc_pdf, const int32_t* arc_state, const float* final_logprob,
                                      tdnnf_num_graph** out) {
  TDNNF_REQUIRE(ctx && state_offsets && fwd_ranges && bwd_ranges && arc_logprob && arc_pdf && arc_state && final_logprob && out,
                "null argument");
  TDNNF_REQUIRE(num_seqs > 0 && num_arcs > 0 && state_offsets[0] == 0, "empty numerator graph");
  const int num_states = state_offsets[num_seqs];
  int max_states = 0;
  for (int s = 0; s < num_seqs; ++s) {
    TDNNF_REQUIRE(state_offsets[s + 1] > state_offsets[s], "a sequence has no states");
    max_states = std::max(max_states, state_offsets[s + 1] - state_offsets[s]);
  }
  TDNNF_REQUIRE(max_states <= 12000, "numerator FST too large for the per-sequence kernel");
  for (int h = 0; h < num_states; ++h)
    TDNNF_REQUIRE(fwd_ranges[2 * h] <= fwd_ranges[2 * h + 1] && fwd_ranges[2 * h + 1] <= 2 * num_arcs &&
                      bwd_ranges[2 * h] <= bwd_ranges[2 * h + 1] && bwd_ranges[2 * h + 1] <= 2 * num_arcs,
                  "arc range out of bounds");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  tdnnf_num_graph* g = new tdnnf_num_graph();
  g->ctx = ctx;
  g->num_seqs = num_seqs;
  g->num_states = num_states;
  g->num_arcs = num_arcs;
  g->max_states = max_states;
  g->cap_seqs = num_seqs;
  g->cap_states = num_states;
  g->cap_arcs = num_arcs;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up(reinterpret_cast<void**>(&g->state_offsets), state_offsets, sizeof(int) * (num_seqs + 1));
  up(reinterpret_cast<void**>(&g->fwd_ranges), fwd_ranges, sizeof(int2) * num_states);
  up(reinterpret_cast<void**>(&g->bwd_ranges), bwd_ranges, sizeof(int2) * num_states);
  up(reinterpret_cast<void**>(&g->arc_logprob), arc_logprob, sizeof(float) * 2 * num_arcs);
  up(reinterpret_cast<void**>(&g->arc_pdf), arc_pdf, sizeof(int) * 2 * num_arcs);
  up(reinterpret_cast<void**>(&g->arc_state), arc_state, sizeof(int) * 2 * num_arcs);
  up(reinterpret_cast<void**>(&g->final_logprob), final_logprob, sizeof(float) * num_states);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->scalars), sizeof(double) * 2);
  if (e != cudaSuccess) {
    tdnnf_num_graph_destroy(g);
    return fail(TDNNF_ERR_CUDA, std::string("numerator graph upload failed: ") + cudaGetErrorString(e));
  }
  *out = g;
  return TDNNF_OK;
}

extern "C" int tdnnf_num_graph_destroy(tdnnf_num_graph* g) {
  if (!g) return TDNNF_OK;
  cudaFree(g->state_offsets);
  cudaFree(g->fwd_ranges);
  cudaFree(g->bwd_ranges);
  cudaFree(g->arc_logprob);
  cudaFree(g->arc_pdf);
  cudaFree(g->arc_state);
  cudaFree(g->final_logprob);
  cudaFree(g->alpha);
  cudaFree(g->scalars);
  if (g->staging) cudaFreeHost(g->staging);
  if (g->staged) cudaEventDestroy(g->staged);
  delete g;
  return TDNNF_OK;
}

// A new minibatch's supervision into the SAME device arrays (every minibatch brings its own numerator FSTs: kaldi
// NnetChainExample -> Supervision::e2e_fsts): one pinned staging buffer, asynchronous copies on the context's stream,
// no allocation and no device synchronisation as long as the new graphs fit the capacities of the first ones.
extern "C" int tdnnf_num_graph_update(tdnnf_num_graph* g, int num_seqs, const int32_t* state_offsets, int num_arcs,
                                      const int32_t* fwd_ranges, const int32_t* bwd_ranges, const float* arc_logprob,
                                      const int32_t* arc_pdf, const int32_t* arc_state, const float* final_logprob) {
  TDNNF_REQUIRE(g && state_offsets && fwd_ranges && bwd_ranges && arc_logprob && arc_pdf && arc_state && final_logprob, "null argument");
  TDNNF_REQUIRE(num_seqs > 0 && num_arcs > 0 && state_offsets[0] == 0, "empty numerator graph");
  const int num_states = state_offsets[num_seqs];
  int max_states = 0;
  for (int s = 0; s < num_seqs; ++s) {
    TDNNF_REQUIRE(state_offsets[s + 1] > state_offsets[s], "a sequence has no states");
    max_states = std::max(max_states, state_offsets[s + 1] - state_offsets[s]);
  }
  TDNNF_REQUIRE(max_states <= 12000, "numerator FST too large for the per-sequence kernel");
  for (int h = 0; h < num_states; ++h)
    TDNNF_REQUIRE(fwd_ranges[2 * h] <= fwd_ranges[2 * h + 1] && fwd_ranges[2 * h + 1] <= 2 * num_arcs &&
                      bwd_ranges[2 * h] <= bwd_ranges[2 * h + 1] && bwd_ranges[2 * h + 1] <= 2 * num_arcs,
                  "arc range out of bounds");
  tdnnf_ctx* ctx = g->ctx;
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (num_seqs > g->cap_seqs || num_states > g->cap_states || num_arcs > g->cap_arcs) {
    // grow (with head-room): the one path that synchronises
    TDNNF_CUDA_OK(cudaStreamSynchronize(st));
    const int cs = std::max(num_seqs, g->cap_seqs), cn = std::max(num_states + num_states / 4, g->cap_states),
              ca = std::max(num_arcs + num_arcs / 4, g->cap_arcs);
    cudaFree(g->state_offsets); cudaFree(g->fwd_ranges); cudaFree(g->bwd_ranges); cudaFree(g->arc_logprob);
    cudaFree(g->arc_pdf); cudaFree(g->arc_state); cudaFree(g->final_logprob);
    g->state_offsets = nullptr; g->fwd_ranges = g->bwd_ranges = nullptr; g->arc_logprob = nullptr;
    g->arc_pdf = g->arc_state = nullptr; g->final_logprob = nullptr;
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->state_offsets), sizeof(int) * (cs + 1)));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->fwd_ranges), sizeof(int2) * cn));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->bwd_ranges), sizeof(int2) * cn));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->arc_logprob), sizeof(float) * 2 * ca));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->arc_pdf), sizeof(int) * 2 * ca));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->arc_state), sizeof(int) * 2 * ca));
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->final_logprob), sizeof(float) * cn));
    g->cap_seqs = cs; g->cap_states = cn; g->cap_arcs = ca;
  }
  const size_t sizes[7] = {sizeof(int) * (size_t)(num_seqs + 1), sizeof(int2) * (size_t)num_states, sizeof(int2) * (size_t)num_states,
                           sizeof(float) * 2 * (size_t)num_arcs, sizeof(int) * 2 * (size_t)num_arcs, sizeof(int) * 2 * (size_t)num_arcs,
                           sizeof(float) * (size_t)num_states};
  const void* srcs[7] = {state_offsets, fwd_ranges, bwd_ranges, arc_logprob, arc_pdf, arc_state, final_logprob};
  void* dsts[7] = {g->state_offsets, g->fwd_ranges, g->bwd_ranges, g->arc_logprob, g->arc_pdf, g->arc_state, g->final_logprob};
  size_t total = 0;
  for (size_t b : sizes) total += (b + 255) & ~size_t(255);
  if (!g->staged) TDNNF_CUDA_OK(cudaEventCreateWithFlags(&g->staged, cudaEventDisableTiming));
  else TDNNF_CUDA_OK(cudaEventSynchronize(g->staged));  // the previous minibatch's copies have left the staging buffer
  if (g->staging_bytes < total) {
    if (g->staging) cudaFreeHost(g->staging);
    g->staging = nullptr;
    TDNNF_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&g->staging), total * 2));
    g->staging_bytes = total * 2;
  }
  size_t off = 0;
  for (int i = 0; i < 7; ++i) {
    memcpy(g->staging + off, srcs[i], sizes[i]);
    TDNNF_CUDA_OK(cudaMemcpyAsync(dsts[i], g->staging + off, sizes[i], cudaMemcpyHostToDevice, st));
    off += (sizes[i] + 255) & ~size_t(255);
  }
  TDNNF_CUDA_OK(cudaEventRecord(g->staged, st));
  g->num_seqs = num_seqs;
  g->num_states = num_states;
  g->num_arcs = num_arcs;
  g->max_states = max_states;
  return TDNNF_OK;
}

extern "C" int tdnnf_num_forward_backward(tdnnf_ctx* ctx, const tdnnf_num_graph* g_in, const float* nnet_output, int stride,
                                          int frames_per_seq, float deriv_weight, float* nnet_output_deriv,
                                          int deriv_stride, float* logprob, int* ok) {
  TDNNF_REQUIRE(ctx && g_in && nnet_output && logprob && ok, "null argument");
  TDNNF_REQUIRE(frames_per_seq > 0, "empty minibatch");
  tdnnf_num_graph* g = const_cast<tdnnf_num_graph*>(g_in);
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  const size_t need = (size_t)g->num_states * (frames_per_seq + 1);
  if (need > g->alpha_elems) {
    TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (g->alpha) cudaFree(g->alpha);
    g->alpha = nullptr;
    TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g->alpha), sizeof(float) * need));
    g->alpha_elems = need;
  }
  TDNNF_CUDA_OK(cudaMemsetAsync(g->scalars, 0, sizeof(double) * 2, ctx->stream));
  const size_t smem = sizeof(float) * 2 * g->max_states;
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) {
      TDNNF_CUDA_OK(cudaFuncSetAttribute(num_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      set = true;
    }
  }
  TDNNF_CUDA_OK(launch_pdl(num_fwd_bwd_kernel, dim3(g->num_seqs), dim3(kNumThreads), smem, ctx->stream, 1, 
      g->state_offsets, g->fwd_ranges, g->bwd_ranges, g->arc_logprob, g->arc_pdf, g->arc_state, g->final_logprob, nnet_output,
      stride, g->num_seqs, frames_per_seq, g->alpha, deriv_weight, nnet_output_deriv, deriv_stride, g->scalars));
  ctx->launches++;
  TDNNF_CUDA_OK(cudaGetLastError());
  double h[2] = {0, 0};
  TDNNF_CUDA_OK(cudaMemcpyAsync(h, g->scalars, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  *logprob = (float)h[0];
  *ok = (h[1] == 0.0) ? 1 : 0;
  return TDNNF_OK;
}
