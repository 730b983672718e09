// LF-MMI denominator forward-backward, "sequence-slice" kernels (kaldi: chain/chain-denominator.{h,cc}).
//
// The per-frame kernels of den.cu gather, for every arc, sequence and frame, a row of alpha(t-1) AND a row of
// E(t-1) = exp(nnet_output) out of L2, plus a row of posterior atomics on the way back: 5 row visits per arc and
// frame-pair, one launch per frame and direction with a partial tail wave each.  Here:
//   * a thread-block CLUSTER of G CTAs owns one slice of 8 sequences for the whole recursion (sequences are
//     independent): one launch per direction, no grid-wide barrier, a hardware cluster barrier per frame;
//   * E(t, :, slice) (P x 8 floats = 188 KB at P = 6008) is staged in the shared memory of every CTA of the cluster
//     by ONE multicast bulk copy (cp.async.bulk ... .multicast::cluster: each CTA issues 1/G of it), double-buffered
//     by pdf halves so that the copy for the next frame runs under the arcs of this one: the E gather leaves L2;
//   * alpha / beta rows of a slice are 32 bytes = one L2 sector per arc; transition records are streamed coalesced
//     (8 B per arc and slice, states grouped 16 to a warp with interleaved lists of equal length);
//   * per-sequence totals: warp shuffles, then a fixed-order sum of the G CTAs' partials exchanged through
//     distributed shared memory (bit-identical in every CTA, no atomics).
// MEASURED (B200, N = 16384 states, 262 k arcs, P = 6008, T = 50; tools/den_sweep.py --variants, profiles/r02_den.md):
// S = 64: 4.32 ms with E in one part, 6.04 ms in two parts, against 2.43 ms for the per-frame kernels; S = 128: 6.73 / 9.39
// against 4.10 ms.  Parity-green (tests/test_gpu_den.py), but slower, so it is OPT-IN (TDNNF_DEN_PATH=slices).  Why: (i) with
// 188 KB of E per CTA the occupancy API grants clusters of at most 10 (S = 64) / 6 (S = 128) CTAs, i.e. 80 / 96 of 148
// SMs; (ii) one CTA of 512 threads per SM with 4 x 16-byte gathers in flight per thread keeps 32 KB in flight per SM
// (the per-frame kernels: 3 CTAs x 256 threads x 128 B = 96 KB), and at ~1 us of loaded L2 latency that is 2.6 TB/s of row
// visits over 80 SMs -- the 2.3 TB/s measured: latency-bound on memory-level parallelism, not on L2 bandwidth;
// (iii) the second E part costs a second cluster barrier per frame and a partial-sum round trip.  The traffic saved
// (3 row visits per arc instead of 5) is partly paid back by re-streaming the transition records per slice (8-12 B per
// arc and slice) -- the net L2 bytes per frame pair are 249 MB against 351 MB, too small a gain to carry (i) and (ii).
// Layouts (fp32): E3 / gamma3 [T][S/8][Ppad][8], alpha3 [T+1][S/8][N][8], betad3 [2][S/8][N][8], tot / bsum [T+1][S].
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "context.h"
#include "den_slices.h"
#include "ptx.cuh"

using namespace tdnnf;

namespace {

constexpr int kV = 8;            // sequences per slice
constexpr int kThreads = 512;    // 16 warps; a warp works on 16 states x 8 sequences (thread = state x 4 sequences)
constexpr int kWarps = kThreads / 32;
constexpr int kMaxG = 16;
constexpr int kU = 4;            // arcs in flight per thread

struct SlicePlan {               // one direction
  int num_tasks = 0;             // warp-tasks of 16 states, all CTAs
  int* cta_begin = nullptr;      // device [G + 1]: tasks [cta_begin[c], cta_begin[c + 1]) belong to CTA c
  int4* task_info = nullptr;     // device [num_tasks]: {base of part 0, length of part 0, base of part 1, length of part 1}
  int* task_state = nullptr;     // device [16 * num_tasks], -1 = no state
  uint2* arcs = nullptr;         // device: {prob bits, (pdf - part * Ph) << state_bits | other state}; arc k of slot j at base + 16 k + j
  float* arcs_q = nullptr;       // device (forward plan only): prob * init[source state]
  void destroy() {
    cudaFree(cta_begin);
    cudaFree(task_info);
    cudaFree(task_state);
    cudaFree(arcs);
    cudaFree(arcs_q);
  }
};

}  // namespace

struct tdnnf_den_slices {
  int G = 0, NP = 0, Ph = 0, Ppad = 0, state_bits = 0, SB = 0;
  size_t smem = 0;
  SlicePlan fwd, bwd;
  float *E = nullptr, *alpha = nullptr, *betad = nullptr, *gamma = nullptr, *bsum = nullptr;
};

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
__device__ __forceinline__ void st_cluster_f32(float* local, uint32_t cta, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ptx::smem_u32(local)), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
// 1-D bulk copy global -> the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier (same offset) gets
// complete_tx for the bytes it received.
__device__ __forceinline__ void bulk_load_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(ptx::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory carve-up (dynamic): [E buffers: NP x Ph x 8 floats][mbar: 2 x u64][red: kWarps x 8][part: 2 x kMaxG x 8][cur: 8]
struct Smem {
  float* e;
  uint64_t* mbar;
  float* red;
  float* part;
  float* cur;
};
__device__ __forceinline__ Smem carve(float* base, int NP, int Ph) {
  Smem s;
  s.e = base;
  s.mbar = reinterpret_cast<uint64_t*>(base + (size_t)NP * Ph * kV);
  s.red = reinterpret_cast<float*>(s.mbar + 2);
  s.part = s.red + kWarps * kV;
  s.cur = s.part + 2 * kMaxG * kV;
  return s;
}
size_t smem_bytes(int NP, int Ph) { return (size_t)NP * Ph * kV * 4 + 16 + (kWarps * kV + 2 * kMaxG * kV + kV) * 4; }

// One elected thread: arm this CTA's barrier for part p and copy this CTA's share of E(t, part p, slice) to every CTA.
__device__ __forceinline__ void issue_e_load(const Smem& sm, const float* E_slice_t /* [Ppad][8] */, int p, int Ph, int G, int rank) {
  const uint32_t part_bytes = (uint32_t)Ph * kV * 4;
  ptx::mbar_expect_tx(&sm.mbar[p], part_bytes);
  const int chunk = (Ph + G - 1) / G;
  const int r0 = rank * chunk, r1 = min(Ph, r0 + chunk);
  if (r1 > r0) {
    float* dst = sm.e + ((size_t)p * Ph + r0) * kV;
    const float* src = E_slice_t + ((size_t)p * Ph + r0) * kV;
    const uint32_t bytes = (uint32_t)(r1 - r0) * kV * 4;
    if (G > 1) bulk_load_multicast(dst, src, bytes, &sm.mbar[p], (uint16_t)((1u << G) - 1u));
    else bulk_load(dst, src, bytes, &sm.mbar[p]);
  }
}

// Sum over the warp's 16 state slots (lane = 2 * slot + half), leaves the result in lanes 0 / 1 (half 0 / 1).
__device__ __forceinline__ void reduce_slots(float (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) x[i] += __shfl_xor_sync(0xffffffffu, x[i], o);
  }
}

// Every CTA contributes 8 per-sequence partial sums; after the cluster barrier that follows, all_reduce_read returns the same
// fixed-order total in every CTA.  `buf` alternates per frame (a fast CTA may write frame t+1 while a slow one still reads t).
__device__ __forceinline__ void all_reduce_write(const Smem& sm, float (&mine)[4], int buf, int G, int rank) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  reduce_slots(mine);
  if (lane < 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.red[warp * kV + lane * 4 + i] = mine[i];
  }
  __syncthreads();
  if (threadIdx.x < kV) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += sm.red[w * kV + threadIdx.x];
    float* slot = sm.part + ((size_t)buf * kMaxG + rank) * kV + threadIdx.x;
    for (int c = 0; c < G; ++c) st_cluster_f32(slot, (uint32_t)c, s);
  }
}
__device__ __forceinline__ void all_reduce_read(const Smem& sm, int buf, int G, int half, float (&out)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
    for (int c = 0; c < G; ++c) s += sm.part[((size_t)buf * kMaxG + c) * kV + half * 4 + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// alpha(t, h, s) = sum_{arcs (g -> h, pdf, p)} (alpha(t-1, g, s) + leaky init[g] tot(t-1, s)) p E(t-1, pdf, s) / tot(t-1, s)
template <int NP>
__global__ void __launch_bounds__(kThreads, 1)
den3_forward_kernel(const int* __restrict__ cta_begin, const int4* __restrict__ task_info, const int* __restrict__ task_state,
                    const uint2* __restrict__ arcs, const float* __restrict__ arcs_q, int N, int Ph, int S, int T, float leaky,
                    float init_sum, int state_bits, int G, const float* __restrict__ E, float* __restrict__ alpha,
                    float* __restrict__ tot) {
  extern __shared__ __align__(128) float den3_smem[];
  const Smem sm = carve(den3_smem, NP, Ph);
  const int rank = (int)cluster_ctarank();
  const int sb = blockIdx.x / G, SB = S / kV, Ppad = NP * Ph;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = lane >> 1, half = lane & 1;
  const uint32_t smask = (1u << state_bits) - 1u;
  const int t0 = cta_begin[rank], t1 = cta_begin[rank + 1];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) ptx::mbar_init(&sm.mbar[p], 1);
    ptx::fence_mbar_init();
  }
  cluster_sync_all();  // every CTA's barriers exist before any multicast copy can signal them
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) issue_e_load(sm, E + ((size_t)0 * SB + sb) * Ppad * kV, p, Ph, G, rank);
  }
  float totp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) totp[i] = init_sum;
  uint32_t phase = 0;
  for (int t = 1; t <= T; ++t) {
    const float* a_prev = alpha + ((size_t)(t - 1) * SB + sb) * N * kV + half * 4;
    float* a_cur = alpha + ((size_t)t * SB + sb) * N * kV + half * 4;
    float lt[4], inv[4], part[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lt[i] = leaky * totp[i];
      inv[i] = 1.0f / totp[i];
      part[i] = 0.f;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ptx::mbar_wait(&sm.mbar[p], phase);
      const float* e_s = sm.e + (size_t)p * Ph * kV + half * 4;
      for (int task = t0 + warp; task < t1; task += kWarps) {
        const int4 info = task_info[task];
        const int base = p == 0 ? info.x : info.z, len = p == 0 ? info.y : info.w;
        const int h = task_state[task * 16 + slot];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (p > 0 && h >= 0) {  // partial sum of the earlier parts (written by this very thread)
          const float4 v = __ldcg(reinterpret_cast<const float4*>(a_cur + (size_t)h * kV));
          acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
        }
        const uint2* ap = arcs + base + slot;
        const float* aq = arcs_q + base + slot;
        uint2 rec[kU];
        float q[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          rec[u] = u < len ? ap[u * 16] : make_uint2(0u, 0u);
          q[u] = u < len ? aq[u * 16] : 0.f;
        }
        for (int k0 = 0; k0 < len; k0 += kU) {
          float4 a[kU], e[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const uint32_t g = rec[u].y & smask, pl = rec[u].y >> state_bits;
            a[u] = rec[u].x != 0u ? __ldcg(reinterpret_cast<const float4*>(a_prev + (size_t)g * kV)) : make_float4(0.f, 0.f, 0.f, 0.f);
            e[u] = *reinterpret_cast<const float4*>(e_s + (size_t)pl * kV);
          }
          uint2 recn[kU];
          float qn[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int k = k0 + kU + u;
            recn[u] = k < len ? ap[k * 16] : make_uint2(0u, 0u);
            qn[u] = k < len ? aq[k * 16] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const float pr = __uint_as_float(rec[u].x);
            acc[0] += (a[u].x * pr + lt[0] * q[u]) * e[u].x;
            acc[1] += (a[u].y * pr + lt[1] * q[u]) * e[u].y;
            acc[2] += (a[u].z * pr + lt[2] * q[u]) * e[u].z;
            acc[3] += (a[u].w * pr + lt[3] * q[u]) * e[u].w;
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            rec[u] = recn[u];
            q[u] = qn[u];
          }
        }
        if (h >= 0) {
          if (p == NP - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i] *= inv[i];
              part[i] += acc[i];
            }
          }
          __stcg(reinterpret_cast<float4*>(a_cur + (size_t)h * kV), make_float4(acc[0], acc[1], acc[2], acc[3]));
        }
      }
      if (p == NP - 1) all_reduce_write(sm, part, t & 1, G, rank);
      // everybody has finished reading this part of E(t-1) (and, after the last part, has published alpha(t) and its totals)
      cluster_sync_all();
      if (threadIdx.x == 0 && t < T) issue_e_load(sm, E + ((size_t)t * SB + sb) * Ppad * kV, p, Ph, G, rank);
    }
    phase ^= 1u;
    all_reduce_read(sm, t & 1, G, half, totp);
    if (rank == 0 && warp == 0 && lane < 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tot[(size_t)t * S + sb * kV + lane * 4 + i] = totp[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// vf = p beta(t+1, g, s) E(t, pdf, s);  gamma(t, pdf, s) += vf alpha'(t, h, s) / tot(t, s);  betad(t, h, s) = sum vf / tot(t, s)
// beta(t+1, g, s) = betad(t+1, g, s) + leaky bsum(t+1, s),  bsum(t, s) = sum_g init[g] betad(t, g, s)
template <int NP>
__global__ void __launch_bounds__(kThreads, 1)
den3_backward_kernel(const int* __restrict__ cta_begin, const int4* __restrict__ task_info, const int* __restrict__ task_state,
                     const uint2* __restrict__ arcs, const float* __restrict__ init, int N, int Ph, int S, int T, float leaky,
                     float init_sum, int state_bits, int G, const float* __restrict__ E, const float* __restrict__ alpha,
                     const float* __restrict__ tot, const float* __restrict__ tot_prob, float* __restrict__ betad,
                     float* __restrict__ gamma, double* __restrict__ check) {
  extern __shared__ __align__(128) float den3_smem[];
  const Smem sm = carve(den3_smem, NP, Ph);
  const int rank = (int)cluster_ctarank();
  const int sb = blockIdx.x / G, SB = S / kV, Ppad = NP * Ph;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = lane >> 1, half = lane & 1;
  const uint32_t smask = (1u << state_bits) - 1u;
  const int t0 = cta_begin[rank], t1 = cta_begin[rank + 1];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) ptx::mbar_init(&sm.mbar[p], 1);
    ptx::fence_mbar_init();
  }
  cluster_sync_all();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NP; ++p) issue_e_load(sm, E + ((size_t)(T - 1) * SB + sb) * Ppad * kV, p, Ph, G, rank);
  }
  // betad(T, h, s) = 1 / tot_prob[s]  =>  bsum(T, s) = sum(init) / tot_prob[s]
  float last[4], bs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    last[i] = 1.0f / tot_prob[sb * kV + half * 4 + i];
    bs[i] = init_sum * last[i];
  }
  float chk = 0.f;
  uint32_t phase = 0;
  for (int t = T - 1; t >= 0; --t) {
    const float* b_next = betad + ((size_t)((t + 1) & 1) * SB + sb) * N * kV + half * 4;
    float* b_cur = betad + ((size_t)(t & 1) * SB + sb) * N * kV + half * 4;
    const float* a_t = alpha + ((size_t)t * SB + sb) * N * kV + half * 4;
    float* g_t = gamma + ((size_t)t * SB + sb) * Ppad * kV + half * 4;
    float lt[4], inv[4], lb[4], part[4];
    {
      const float4 tp = *reinterpret_cast<const float4*>(tot + (size_t)t * S + sb * kV + half * 4);
      const float tv[4] = {tp.x, tp.y, tp.z, tp.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lt[i] = leaky * tv[i];
        inv[i] = 1.0f / tv[i];
        lb[i] = leaky * bs[i];
        part[i] = 0.f;
      }
    }
    const bool first = t == T - 1;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ptx::mbar_wait(&sm.mbar[p], phase);
      const float* e_s = sm.e + (size_t)p * Ph * kV + half * 4;
      float* g_p = g_t + (size_t)p * Ph * kV;
      for (int task = t0 + warp; task < t1; task += kWarps) {
        const int4 info = task_info[task];
        const int base = p == 0 ? info.x : info.z, len = p == 0 ? info.y : info.w;
        const int h = task_state[task * 16 + slot];
        float ih = 0.f, ad[4] = {0.f, 0.f, 0.f, 0.f}, occ[4], totv[4] = {0.f, 0.f, 0.f, 0.f};
        if (h >= 0) {
          ih = init[h];
          const float4 v = __ldcg(reinterpret_cast<const float4*>(a_t + (size_t)h * kV));
          ad[0] = v.x + ih * lt[0]; ad[1] = v.y + ih * lt[1]; ad[2] = v.z + ih * lt[2]; ad[3] = v.w + ih * lt[3];
          if (p > 0) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(b_cur + (size_t)h * kV));
            totv[0] = w.x; totv[1] = w.y; totv[2] = w.z; totv[3] = w.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) occ[i] = ad[i] * inv[i];
        const uint2* ap = arcs + base + slot;
        uint2 rec[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) rec[u] = u < len ? ap[u * 16] : make_uint2(0u, 0u);
        for (int k0 = 0; k0 < len; k0 += kU) {
          float4 b[kU], e[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const uint32_t g = rec[u].y & smask, pl = rec[u].y >> state_bits;
            if (first) b[u] = make_float4(last[0], last[1], last[2], last[3]);
            else b[u] = rec[u].x != 0u ? __ldcg(reinterpret_cast<const float4*>(b_next + (size_t)g * kV)) : make_float4(0.f, 0.f, 0.f, 0.f);
            e[u] = *reinterpret_cast<const float4*>(e_s + (size_t)pl * kV);
          }
          uint2 recn[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int k = k0 + kU + u;
            recn[u] = k < len ? ap[k * 16] : make_uint2(0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            if (rec[u].x != 0u) {  // padding arcs carry probability 0
              const float pr = __uint_as_float(rec[u].x);
              const uint32_t pl = rec[u].y >> state_bits;
              const float v0 = (pr * (b[u].x + lb[0])) * e[u].x, v1 = (pr * (b[u].y + lb[1])) * e[u].y;
              const float v2 = (pr * (b[u].z + lb[2])) * e[u].z, v3 = (pr * (b[u].w + lb[3])) * e[u].w;
              totv[0] += v0; totv[1] += v1; totv[2] += v2; totv[3] += v3;
              red_add_v4(g_p + (size_t)pl * kV, v0 * occ[0], v1 * occ[1], v2 * occ[2], v3 * occ[3]);
            }
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) rec[u] = recn[u];
        }
        if (h >= 0) {
          if (p == NP - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              totv[i] *= inv[i];
              part[i] += ih * totv[i];
              if (t == 0) chk += ad[i] * totv[i];  // alpha'(0, h, s) betad(0, h, s)
            }
          }
          __stcg(reinterpret_cast<float4*>(b_cur + (size_t)h * kV), make_float4(totv[0], totv[1], totv[2], totv[3]));
        }
      }
      if (p == NP - 1) all_reduce_write(sm, part, t & 1, G, rank);
      cluster_sync_all();
      if (threadIdx.x == 0 && t > 0) issue_e_load(sm, E + ((size_t)(t - 1) * SB + sb) * Ppad * kV, p, Ph, G, rank);
    }
    phase ^= 1u;
    all_reduce_read(sm, t & 1, G, half, bs);
  }
  if (check != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chk += __shfl_xor_sync(0xffffffffu, chk, o);
    if (lane == 0 && chk != 0.f) atomicAdd(check, (double)chk);
  }
}

// E3[t][sb][p][v] = exp(clamp(x[t*S + sb*8 + v][p], -30, 30)), zero for the padding pdfs
__global__ void den3_exp_kernel(const float* __restrict__ x, long long ld, int S, int P, int Ppad, float* __restrict__ E) {
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
  // thread (x, y): sequence s0 + x of pdf p0 + i: 8 consecutive x are one 32-byte row of the slice layout
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    if (s < S && p < Ppad) E[(((long long)t * (S / kV) + s / kV) * Ppad + p) * kV + s % kV] = tile[threadIdx.x][i];
  }
}

// deriv[t*S + s][p] += w * gamma3[t][sb][p][v]
__global__ void den3_deriv_kernel(const float* __restrict__ gamma, int S, int P, int Ppad, float w, float* __restrict__ deriv,
                                  long long ld) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int p0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && p < P) ? gamma[(((long long)t * (S / kV) + s / kV) * Ppad + p) * kV + s % kV] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int s = s0 + i, p = p0 + threadIdx.x;
    if (s < S && p < P) deriv[((long long)t * S + s) * ld + p] += w * tile[threadIdx.x][i];
  }
}

// alpha3[0][sb][h][v] = init[h]; tot[0][s] = sum(init)
__global__ void den3_alpha_first_kernel(const float* __restrict__ init, int N, int S, float init_sum, float* __restrict__ alpha0,
                                        float* __restrict__ tot0) {
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    alpha0[i] = init[(i / kV) % N];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
    tot0[i] = init_sum;
}

// ------------------------------------------------------------------------------------------------ host: the plan
// ranges: [N][2] into (prob, pdf, state); dir_q: init[] when the records also carry prob * init[other state] (forward).
int build_plan(const std::vector<int>& ranges, const std::vector<float>& prob, const std::vector<int>& pdf,
               const std::vector<int>& state, const float* init_for_q, int N, int G, int NP, int Ph, int state_bits,
               SlicePlan* plan) {
  auto len_of = [&](int h) { return ranges[2 * h + 1] - ranges[2 * h]; };
  auto len0_of = [&](int h) {
    int c = 0;
    for (int a = ranges[2 * h]; a < ranges[2 * h + 1]; ++a) c += pdf[a] < Ph ? 1 : 0;
    return c;
  };
  std::vector<int> order(N);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len_of(a) > len_of(b); });
  if (NP > 1) {
    // inside windows of 16 tasks (nearly equal totals) sort by the length of part 0: the 16 states of a task then agree
    // on both part lengths, and the interleaved lists need next to no padding
    std::vector<int> l0(N);
    for (int h = 0; h < N; ++h) l0[h] = len0_of(h);
    for (int w = 0; w < N; w += 256)
      std::stable_sort(order.begin() + w, order.begin() + std::min(N, w + 256), [&](int a, int b) { return l0[a] > l0[b]; });
  }
  const int tasks = (N + 15) / 16;
  // task i (in order of decreasing length) -> CTA i mod G: equal shares of every length class
  std::vector<std::vector<int>> per_cta(G);
  for (int i = 0; i < tasks; ++i) per_cta[i % G].push_back(i);
  std::vector<int> cta_begin(G + 1, 0);
  std::vector<int4> info(tasks);
  std::vector<int> st((size_t)tasks * 16, -1);
  size_t total = 0;
  std::vector<uint2> arcs;
  std::vector<float> arcs_q;
  int slot_index = 0;
  for (int c = 0; c < G; ++c) {
    cta_begin[c] = slot_index;
    for (int ti : per_cta[c]) {
      int lens[2] = {0, 0};
      for (int j = 0; j < 16 && ti * 16 + j < N; ++j) {
        const int h = order[ti * 16 + j];
        const int a0 = len0_of(h), all = len_of(h);
        const int l0 = NP > 1 ? a0 : all, l1 = NP > 1 ? all - a0 : 0;
        lens[0] = std::max(lens[0], l0);
        lens[1] = std::max(lens[1], l1);
      }
      int4 inf;
      inf.x = (int)total;
      inf.y = lens[0];
      inf.z = (int)(total + (size_t)lens[0] * 16);
      inf.w = lens[1];
      const size_t need = (size_t)(lens[0] + lens[1]) * 16;
      if (total + need > (size_t)INT32_MAX) return fail(TDNNF_ERR_UNSUPPORTED, "denominator graph too large for the slice plan");
      arcs.resize(total + need, make_uint2(0u, 0u));
      if (init_for_q) arcs_q.resize(total + need, 0.f);
      for (int j = 0; j < 16 && ti * 16 + j < N; ++j) {
        const int h = order[ti * 16 + j];
        st[(size_t)slot_index * 16 + j] = h;
        int k[2] = {0, 0};
        for (int a = ranges[2 * h]; a < ranges[2 * h + 1]; ++a) {
          const int part = NP > 1 ? (pdf[a] < Ph ? 0 : 1) : 0;
          const size_t at = (size_t)(part == 0 ? inf.x : inf.z) + (size_t)k[part] * 16 + j;
          ++k[part];
          uint32_t bits;
          memcpy(&bits, &prob[a], 4);
          arcs[at] = make_uint2(bits, ((uint32_t)(pdf[a] - part * Ph) << state_bits) | (uint32_t)state[a]);
          if (init_for_q) arcs_q[at] = prob[a] * init_for_q[state[a]];
        }
      }
      info[slot_index] = inf;
      total += need;
      ++slot_index;
    }
  }
  cta_begin[G] = slot_index;
  if (arcs.empty()) arcs.push_back(make_uint2(0u, 0u));
  if (init_for_q && arcs_q.empty()) arcs_q.push_back(0.f);
  plan->num_tasks = tasks;
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up(reinterpret_cast<void**>(&plan->cta_begin), cta_begin.data(), sizeof(int) * (G + 1));
  up(reinterpret_cast<void**>(&plan->task_info), info.data(), sizeof(int4) * tasks);
  up(reinterpret_cast<void**>(&plan->task_state), st.data(), sizeof(int) * st.size());
  up(reinterpret_cast<void**>(&plan->arcs), arcs.data(), sizeof(uint2) * arcs.size());
  if (init_for_q) up(reinterpret_cast<void**>(&plan->arcs_q), arcs_q.data(), sizeof(float) * arcs_q.size());
  if (e != cudaSuccess) return fail(TDNNF_ERR_CUDA, std::string("den slice plan upload failed: ") + cudaGetErrorString(e));
  return TDNNF_OK;
}

template <typename Kern, typename... Args>
cudaError_t launch_cluster(Kern kern, int grid, int cluster, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
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

template <typename Kern>
int max_active_clusters(Kern kern, int grid, int cluster, size_t smem) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <typename Kern>
cudaError_t prepare_kernel(Kern kern, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}

}  // namespace

#define DEN3_LAUNCH_CHECK(ctx)          \
  do {                                  \
    (ctx)->launches++;                  \
    TDNNF_CUDA_OK(cudaGetLastError());  \
  } while (0)

// Returns TDNNF_OK with *out == nullptr when the slice path does not apply to this shape (the caller keeps the frame kernels).
int tdnnf::den_slices_create(tdnnf_ctx* ctx, int N, int P, int S, int T, const std::vector<int>& fwd_ranges,
                             const std::vector<int>& bwd_ranges, const std::vector<float>& prob, const std::vector<int>& pdf,
                             const std::vector<int>& state, const std::vector<float>& init, tdnnf_den_slices** out) {
  *out = nullptr;
  // Opt-in (TDNNF_DEN_PATH=slices): measured on B200 at N = 16384, P = 6008, T = 50 (profiles/r02_den.md) this path runs
  // 4.3 ms (S = 64, E in one part) against 2.4 ms for the per-frame kernels: see the note at the top of the file.
  const char* env = getenv("TDNNF_DEN_PATH");
  if (!(env && strcmp(env, "slices") == 0)) return TDNNF_OK;
  if (S % kV != 0) return TDNNF_OK;
  const int SB = S / kV;
  int NP = 2;
  if (const char* e = getenv("TDNNF_DEN_PARTS")) NP = atoi(e) == 1 ? 1 : 2;
  const int Ph = ((P + NP - 1) / NP + 3) / 4 * 4;
  int sbits = 1, pbits = 1;
  while ((1ll << sbits) < N) ++sbits;
  while ((1ll << pbits) < Ph) ++pbits;
  if (sbits + pbits > 32) return TDNNF_OK;
  const size_t smem = smem_bytes(NP, Ph);
  if (smem > 232448) return TDNNF_OK;  // E(t, :, slice) must fit the 227 KB of shared memory
  cudaError_t pe = NP == 1 ? prepare_kernel(den3_forward_kernel<1>, smem) : prepare_kernel(den3_forward_kernel<2>, smem);
  if (pe == cudaSuccess) pe = NP == 1 ? prepare_kernel(den3_backward_kernel<1>, smem) : prepare_kernel(den3_backward_kernel<2>, smem);
  if (pe != cudaSuccess) {
    cudaGetLastError();
    return TDNNF_OK;
  }
  // the largest cluster such that all slices are resident at once
  int G = std::min(kMaxG, std::max(1, ctx->num_sms / SB));
  if (const char* e = getenv("TDNNF_DEN_CLUSTER")) G = std::max(1, std::min(kMaxG, atoi(e)));
  for (; G >= 1; --G) {
    const int n = NP == 1 ? max_active_clusters(den3_backward_kernel<1>, SB * G, G, smem)
                          : max_active_clusters(den3_backward_kernel<2>, SB * G, G, smem);
    if (n >= SB || (n >= 1 && G == 1)) break;
  }
  if (G < 1) return TDNNF_OK;
  tdnnf_den_slices* s = new tdnnf_den_slices();
  s->G = G;
  s->NP = NP;
  s->Ph = Ph;
  s->Ppad = NP * Ph;
  s->state_bits = sbits;
  s->SB = SB;
  s->smem = smem;
  int rc = build_plan(bwd_ranges, prob, pdf, state, init.data(), N, G, NP, Ph, sbits, &s->fwd);
  if (rc == TDNNF_OK) rc = build_plan(fwd_ranges, prob, pdf, state, nullptr, N, G, NP, Ph, sbits, &s->bwd);
  cudaError_t e = cudaSuccess;
  auto al = [&](float** p, size_t floats) {
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(float) * floats);
  };
  if (rc == TDNNF_OK) {
    al(&s->E, (size_t)T * SB * s->Ppad * kV);
    al(&s->gamma, (size_t)T * SB * s->Ppad * kV);
    al(&s->alpha, (size_t)(T + 1) * N * S);
    al(&s->betad, (size_t)2 * N * S);
    if (e != cudaSuccess) rc = fail(TDNNF_ERR_NOMEM, std::string("denominator workspace allocation failed: ") + cudaGetErrorString(e));
  }
  if (rc != TDNNF_OK) {
    den_slices_destroy(s);
    return rc;
  }
  *out = s;
  return TDNNF_OK;
}

void tdnnf::den_slices_destroy(tdnnf_den_slices* s) {
  if (!s) return;
  s->fwd.destroy();
  s->bwd.destroy();
  cudaFree(s->E);
  cudaFree(s->gamma);
  cudaFree(s->alpha);
  cudaFree(s->betad);
  delete s;
}

void tdnnf::den_slices_describe(const tdnnf_den_slices* s, int* cluster, int* parts, int* ctas) {
  *cluster = s->G;
  *parts = s->NP;
  *ctas = s->G * s->SB;
}

int tdnnf::den_slices_forward(tdnnf_ctx* ctx, tdnnf_den_slices* s, const float* nnet_output, int stride, const float* init,
                              float init_sum, int N, int P, int S, int T, float leaky, float* tot) {
  cudaStream_t st = ctx->stream;
  den3_exp_kernel<<<dim3((s->Ppad + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st>>>(nnet_output, stride, S, P, s->Ppad, s->E);
  DEN3_LAUNCH_CHECK(ctx);
  den3_alpha_first_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(init, N, S, init_sum, s->alpha, tot);
  DEN3_LAUNCH_CHECK(ctx);
  const SlicePlan& pl = s->fwd;
  cudaError_t le;
  if (s->NP == 1)
    le = launch_cluster(den3_forward_kernel<1>, s->SB * s->G, s->G, s->smem, st, (const int*)pl.cta_begin, (const int4*)pl.task_info,
                        (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)pl.arcs_q, N, s->Ph, S, T, leaky, init_sum,
                        s->state_bits, s->G, (const float*)s->E, s->alpha, tot);
  else
    le = launch_cluster(den3_forward_kernel<2>, s->SB * s->G, s->G, s->smem, st, (const int*)pl.cta_begin, (const int4*)pl.task_info,
                        (const int*)pl.task_state, (const uint2*)pl.arcs, (const float*)pl.arcs_q, N, s->Ph, S, T, leaky, init_sum,
                        s->state_bits, s->G, (const float*)s->E, s->alpha, tot);
  TDNNF_CUDA_OK(le);
  DEN3_LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}

int tdnnf::den_slices_backward(tdnnf_ctx* ctx, tdnnf_den_slices* s, const float* init, float init_sum, int N, int P, int S, int T,
                               float leaky, const float* tot, const float* tot_prob, float deriv_weight, float* nnet_output_deriv,
                               int stride, double* check) {
  cudaStream_t st = ctx->stream;
  TDNNF_CUDA_OK(cudaMemsetAsync(s->gamma, 0, sizeof(float) * (size_t)T * s->SB * s->Ppad * kV, st));
  const SlicePlan& pl = s->bwd;
  cudaError_t le;
  if (s->NP == 1)
    le = launch_cluster(den3_backward_kernel<1>, s->SB * s->G, s->G, s->smem, st, (const int*)pl.cta_begin, (const int4*)pl.task_info,
                        (const int*)pl.task_state, (const uint2*)pl.arcs, init, N, s->Ph, S, T, leaky, init_sum, s->state_bits, s->G,
                        (const float*)s->E, (const float*)s->alpha, tot, tot_prob, s->betad, s->gamma, check);
  else
    le = launch_cluster(den3_backward_kernel<2>, s->SB * s->G, s->G, s->smem, st, (const int*)pl.cta_begin, (const int4*)pl.task_info,
                        (const int*)pl.task_state, (const uint2*)pl.arcs, init, N, s->Ph, S, T, leaky, init_sum, s->state_bits, s->G,
                        (const float*)s->E, (const float*)s->alpha, tot, tot_prob, s->betad, s->gamma, check);
  TDNNF_CUDA_OK(le);
  DEN3_LAUNCH_CHECK(ctx);
  den3_deriv_kernel<<<dim3((P + 31) / 32, (S + 31) / 32, T), dim3(32, 8), 0, st>>>(s->gamma, S, P, s->Ppad, deriv_weight,
                                                                                  nnet_output_deriv, stride);
  DEN3_LAUNCH_CHECK(ctx);
  return TDNNF_OK;
}
