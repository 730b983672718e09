// Segment GEMM on sm_100a: the one tensor-core kernel behind TdnnDARTSV3Component's
// Propagate / Backprop / parameter-gradient GEMMs (ref: tdnn.cc:292-328, 366-416, 482-539, 619-624).
//
//   acc[m, n] = sum over segments g (matching the tile's group c), sum over k:
//                  A[plane][g.a_c][m0 + m + g.a_m][k + g.a_k] * B[plane'][g.b_c][n0 + n + g.b_n][k + g.b_k]
//
// A and B are bf16 "hi/lo" operand planes (x = hi + lo + O(2^-18 x)); each K block issues the
// three products hi*hi, hi*lo, lo*hi into one fp32 TMEM accumulator, which restores ~fp32
// accuracy (Kaldi computes in fp32) at bf16 tensor-core rate.  Operands arrive through two 4-D
// TMA tensor maps (k, row, group, plane) with 128-byte swizzle; out-of-range rows / k are zero
// filled by TMA, which is what implements the splice boundaries of the data-gradient.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over work units):
//   warp 0   TMA producer            (one lane)
//   warp 1   tcgen05.mma issuer      (one lane) + TMEM allocator
//   warp 2-5 epilogue: tcgen05.ld -> registers -> scaled store / red.add / dot-product reduce
// Two TMEM accumulator buffers let the epilogue of unit j overlap the MMAs of unit j+1.
#pragma once
#include "ptx.cuh"

namespace tdnnf {

constexpr int kMaxSeg = 16;
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kAccCols = 256;  // TMEM columns per accumulator buffer

struct GemmParams {
  int m_tiles, n_tiles, c_tiles, splits;
  int pair;        // 1: CTA-pair kernel (cta_group::2): a unit is TWO consecutive m-tiles, one per CTA of a cluster of 2
  int kb_per_seg;  // K blocks (of 64) per segment
  int kb_last_steps;  // 16-wide MMA steps that hold data in the LAST K block of a segment (1..4): zero padding is skipped
  int nseg;
  int seg_a_k[kMaxSeg], seg_a_m[kMaxSeg], seg_a_c[kMaxSeg];
  int seg_b_k[kMaxSeg], seg_b_n[kMaxSeg], seg_b_c[kMaxSeg];
  int seg_cmatch[kMaxSeg];  // segment applies to tiles of group c == cmatch (-1: every group)
  int m_valid[kMaxSeg];     // valid accumulator rows per group c
  int n_valid;
  const float* seg_weight;  // device [nseg] or null: a segment whose weight is 0 is skipped
  // epilogue
  float* out;
  long long out_ld;
  int row_mul, row_cadd;  // R = m*row_mul + c*row_cadd
  int col_cadd;           // C = n + c*col_cadd
  int transposed;         // element index = transposed ? C*ld + R : R*ld + C
  int accumulate;         // 1: out += v ; 0: out = v
  int atomic;             // 1: use red.global.add (split-K or shared outputs)
  const float* bias;      // [n_valid] added by split 0 only, or null
  float alpha;            // v = alpha * c_scale[c] * acc (+ bias)
  const float* c_scale;   // device [c_tiles] or null
  const float* dot_ref;   // optional: dot_out[c] += sum(acc * dot_ref[index'])  (un-scaled acc;
  long long dot_ld;       //           index' uses dot_ld in place of out_ld)
  float* dot_out;
  // single-plane fp16 operands are stored as x * pow2_scale(absmax): the device-resident absmax of each such
  // operand (null: not scaled) lets the epilogue undo it, v *= 1 / (pow2_scale(*absmax_a) * pow2_scale(*absmax_b))
  const float* absmax_a;
  const float* absmax_b;
};

// Power-of-two factor that brings a tensor whose largest magnitude is `absmax` into [2^14, 2^15): the top of the
// fp16 range, so that elements down to 2^-29 * absmax stay normal.  Exact to apply and to undo.
__host__ __device__ __forceinline__ float pow2_scale(float absmax) {
  if (!(absmax > 0.f) || absmax > 3.0e38f) return 1.0f;
  int e;
  frexpf(absmax, &e);  // absmax = m * 2^e, m in [0.5, 1)
  int k = 15 - e;
  k = k > 120 ? 120 : (k < -120 ? -120 : k);
  return ldexpf(1.0f, k);
}

struct GemmSmemMeta {
  int seg_a_k[kMaxSeg], seg_a_m[kMaxSeg], seg_a_c[kMaxSeg];
  int seg_b_k[kMaxSeg], seg_b_n[kMaxSeg], seg_b_c[kMaxSeg];
  int list[kMaxSeg][kMaxSeg];  // per group c: active matching segments
  int cnt[kMaxSeg];
  int m_valid[kMaxSeg];
  float c_scale[kMaxSeg];
};

// NPA / NPB = planes of the A / B operand.  Plane formats and products per K step:
//   (2,2)  bf16 hi/lo each (x to ~2^-17): hi*hi + hi*lo + lo*hi, three products -- Propagate (1e-4 tolerance)
//   (3,3)  bf16 hi/mid/lo (x to 2^-24): six products -- the natural-gradient update, whose eigen-problem amplifies rounding
//   (1,1)  ONE fp16 plane each (x to 2^-11), scaled by a power of two into the fp16 range: one product -- the
//          parameter gradient (1e-3 tolerance; measured 2.9e-4)
// (Mixing a bf16 operand with an fp16 one in one kind::f16 MMA -- which would give a two-product data gradient --
// is rejected by the hardware: "illegal instruction", measured on B200.)
// MN = true: both operands are MN-major (the contraction runs over ROWS of the row planes: the parameter gradient).
// K blocks are then 32 rows (no swizzle-width constraint on K), the B tile is whole 64-column chunks.
// PAIR = true: two CTAs of a cluster (one TPC) issue ONE tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own 128
// rows of A and HALF of the B tile, and reads the other half from its partner's shared memory.  The kernels are bound
// by L2 -> SM bandwidth (hi + lo planes: 72-96 KB per K block of 64 at one CTA per tile, i.e. 62-75 B/clk/SM at full
// tensor rate against ~45-50 B/clk/SM measured): the pair form stages 52-64 KB per SM for the same MMA work, and at
// BN = 256 a third pipeline stage fits.
template <int BN, int NPA, int NPB, bool MN = false, bool PAIR = false>
struct GemmCfg {
  static constexpr int kBKk = MN ? 32 : kBK;                 // K elements per pipeline stage
  static constexpr int kBNs = (MN ? (BN + 63) / 64 * 64 : BN) / (PAIR ? 2 : 1);  // B columns held in this CTA's shared memory
  static constexpr int kAPlane = kBM * kBKk * 2;
  static constexpr int kBPlane = kBNs * kBKk * 2;
  static constexpr int kABytes = NPA * kAPlane;
  static constexpr int kBBytes = NPB * kBPlane;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMetaBytes = 2048 + (int)sizeof(GemmSmemMeta);
  static constexpr int kAvail = 232448 - 1024 - kMetaBytes;
  static constexpr int kStagesMax = kAvail / kStageBytes >= 4 ? 4 : kAvail / kStageBytes;
  // Epilogue staging (one 32 x 32 fp32 block per epilogue warp, rows padded to 36 floats): the accumulator leaves TMEM
  // with one ROW per lane, so a direct st / red.global.v4 of a warp touches 32 rows x 16 B; through the staging block a
  // warp instruction covers 4 rows x 128 B.  Only where it does not cost a pipeline stage.
  static constexpr int kEpiPitch = 36;
  static constexpr int kEpiWant = 4 * 32 * kEpiPitch * 4;
  static constexpr int kStagesEpi = (kAvail - kEpiWant) / kStageBytes >= 4 ? 4 : (kAvail - kEpiWant) / kStageBytes;
  static constexpr int kEpiBytes = (kStagesEpi == kStagesMax) ? kEpiWant : 0;
  static constexpr int kStages = kStagesMax;
  static constexpr int kSmemBytes = kStages * kStageBytes + kMetaBytes + kEpiBytes + 1024;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N must be a multiple of 16 in [16,256]");
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(NPA == NPB && NPA >= 1 && NPA <= 3, "unsupported plane combination");
  static_assert(!PAIR || (!MN && BN % 32 == 0), "pair kernels: K-major operands, N/2 a multiple of 16");
  static constexpr int kClusterBytes = (PAIR ? 2 : 1) * kStageBytes;  // what one full barrier waits for
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- CTA-pair (cta_group::2) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier of any CTA of the cluster (address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0,
                                                 int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D (128 rows in each CTA's TMEM) (+)= [A_cta0; A_cta1] * [B_cta0; B_cta1]^T, issued by one thread of the leader CTA
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// No "memory" clobber on purpose: the reductions are fire-and-forget and nothing in the thread reads the
// locations back, so the compiler may hoist independent loads (dot_ref, bias) above them.
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v));
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}

struct UnitCoord {
  int c, m_t, n_t, split;
  int it0, it1;
};

__device__ __forceinline__ UnitCoord decode_unit(int u, const GemmParams& p, const GemmSmemMeta* meta) {
  // Order: group c fastest, then n-tile, then k-split, then m-tile.  Units that run at the same time on
  // different SMs then share operand tiles through L2: the 7 offsets of a parameter-gradient tile read the
  // same activation rows (shifted by one K block) and the n-tiles of a wide output share the A tile.  (With
  // c slowest every wave re-streamed the ~130 MB of operand planes from HBM in 128-byte pieces.)
  UnitCoord uc;
  uc.c = u % p.c_tiles;
  u /= p.c_tiles;
  uc.n_t = u % p.n_tiles;
  u /= p.n_tiles;
  uc.split = u % p.splits;
  uc.m_t = u / p.splits;  // pair kernels: the pair index; the caller turns it into 2 * pair + cta rank
  const long long total = (long long)meta->cnt[uc.c] * p.kb_per_seg;
  uc.it0 = (int)((total * uc.split) / p.splits);
  uc.it1 = (int)((total * (uc.split + 1)) / p.splits);
  return uc;
}

template <int BN, int NPA, int NPB, bool MN = false, bool PAIR = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
splice_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, NPA, NPB, MN, PAIR>;
  // pair kernels: clusters of 2 consecutive CTAs; CTA `rank` owns m-tile 2 * (pair index) + rank and rows
  // [rank * BN / 2, +BN / 2) of the B tile; rank 0 (the leader) issues the MMAs and owns the full / tmem_empty barriers
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kStages = Cfg::kStages;
  constexpr int kBKk = Cfg::kBKk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* meta_base = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(meta_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  GemmSmemMeta* meta = reinterpret_cast<GemmSmemMeta*>(meta_base + 2048);
  float* epi_stage = reinterpret_cast<float*>(meta_base + Cfg::kMetaBytes);  // kEpiBytes (16-byte aligned)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  ptx::grid_dep_launch_dependents();  // the next kernel may queue its CTAs behind ours
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], PAIR ? 8 : 4);  // the epilogue warps of both CTAs release the leader's accumulator
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_ptr, 2 * kAccCols);
    } else {
      ptx::tmem_alloc(tmem_ptr, 2 * kAccCols);
      ptx::tmem_relinquish();
    }
  }
  // everything above is private to this CTA; what follows reads what earlier kernels wrote (seg_weight, c_scale, operands)
  ptx::grid_dep_wait();
  for (int i = threadIdx.x; i < kMaxSeg; i += blockDim.x) {
    meta->seg_a_k[i] = p.seg_a_k[i];
    meta->seg_a_m[i] = p.seg_a_m[i];
    meta->seg_a_c[i] = p.seg_a_c[i];
    meta->seg_b_k[i] = p.seg_b_k[i];
    meta->seg_b_n[i] = p.seg_b_n[i];
    meta->seg_b_c[i] = p.seg_b_c[i];
    meta->m_valid[i] = p.m_valid[i];
    meta->c_scale[i] = (p.c_scale != nullptr && i < p.c_tiles) ? p.c_scale[i] : 1.0f;
    // active segment list of group c = i
    int n = 0;
    if (i < p.c_tiles) {
      for (int g = 0; g < p.nseg; ++g) {
        const bool active = (p.seg_weight == nullptr) || (p.seg_weight[g] != 0.0f);
        if (active && (p.seg_cmatch[g] < 0 || p.seg_cmatch[g] == i)) meta->list[i][n++] = g;
      }
    }
    meta->cnt[i] = n;
  }
  ptx::tc_fence_before_sync();
  __syncwarp();
  if constexpr (PAIR) cluster_sync_all();  // the partner's barriers exist before anything is signalled across
  else __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int total_units = p.c_tiles * (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.splits;

  if (warp == 0) {
    // ================================================= TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = worker; u < total_units; u += workers) {
        UnitCoord uc = decode_unit(u, p, meta);
        if constexpr (PAIR) uc.m_t = 2 * uc.m_t + (int)rank;
        // Iteration order: K block outer, segment (time offset) inner.  Offset i of m-tile m and offset i-2 of m-tile
        // m+1 read the SAME activation rows; with the segment outer they did so 2*kb_per_seg iterations apart, and
        // with one unit per SM (no split-K) the ~225 MB streamed in between pushed them out of the 126 MB L2:
        // ncu showed 774 MB of DRAM reads for a 144 MB operand.  Now the reuse distance is 2 iterations.
        const int cnt = max(meta->cnt[uc.c], 1);
        int kb = uc.it0 / cnt, j = uc.it0 % cnt;
        for (int it = uc.it0; it < uc.it1; ++it) {
          const int g = meta->list[uc.c][j];
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (!PAIR || rank == 0) ptx::mbar_expect_tx(&full_bar[stage], Cfg::kClusterBytes);
          uint8_t* sa = tiles + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (PAIR) {
            // both CTAs' bytes are counted on the LEADER's full barrier (its expect_tx may come after the partner's
            // first complete_tx: the phase cannot end before the leader's own arrival)
            const uint32_t bar = mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
            tma_load_4d_pair(sa, &tmA, bar, kb * kBK + meta->seg_a_k[g], uc.m_t * kBM + meta->seg_a_m[g], meta->seg_a_c[g], 0);
            tma_load_4d_pair(sb, &tmB, bar, kb * kBK + meta->seg_b_k[g],
                             uc.n_t * BN + (int)rank * (BN / 2) + meta->seg_b_n[g], meta->seg_b_c[g], 0);
          } else if constexpr (MN) {
            // row planes as (64 columns, row = k, 64-column chunk, group, plane): smem gets [plane][chunk][k][64]
            tma_load_5d(sa, &tmA, &full_bar[stage], 0, kb * kBKk + meta->seg_a_k[g], (uc.m_t * kBM + meta->seg_a_m[g]) >> 6,
                        meta->seg_a_c[g], 0);
            tma_load_5d(sb, &tmB, &full_bar[stage], 0, kb * kBKk + meta->seg_b_k[g], (uc.n_t * BN + meta->seg_b_n[g]) >> 6,
                        meta->seg_b_c[g], 0);
          } else {
            tma_load_4d(sa, &tmA, &full_bar[stage], kb * kBK + meta->seg_a_k[g], uc.m_t * kBM + meta->seg_a_m[g],
                        meta->seg_a_c[g], 0);
            tma_load_4d(sb, &tmB, &full_bar[stage], kb * kBK + meta->seg_b_k[g], uc.n_t * BN + meta->seg_b_n[g],
                        meta->seg_b_c[g], 0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++j == cnt) { j = 0; ++kb; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer (pair kernels: the leader CTA only)
    // The WHOLE warp runs the loop and one elected lane issues: with the loop under `if (lane == 0)` the operands of
    // every tcgen05.mma lived in per-thread registers and reached the instruction's uniform registers through an
    // ELECT + 5 x R2UR.BROADCAST + branch sequence (14 SASS instructions per MMA, ~175 per K block of the MN-major
    // kernel: the single issuing thread, not the tensor pipe or the TMA loads, bounded these kernels at 51-65 % tensor
    // activity).  Values that every lane computes identically but that come out of shared memory are passed through
    // __shfl_sync so that the compiler knows they are warp-uniform.
    if (rank == 0) {
      // single-plane operands are fp16 (format 0), multi-plane operands are bf16 (format 1)
      constexpr uint32_t idesc =
          ptx::umma_idesc_f16(PAIR ? 2 * kBM : kBM, BN, NPA == 1 ? 0u : 1u, NPB == 1 ? 0u : 1u, MN ? 1u : 0u);
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if constexpr (PAIR) umma_bf16_pair(d, a, b, id, acc);
        else ptx::umma_bf16(d, a, b, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else ptx::umma_commit(bar);
      };
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tiles_u = __shfl_sync(0xffffffffu, ptx::smem_u32(tiles), 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc_buf = 0;
      uint32_t acc_phase = 0;
      for (int u = worker; u < total_units; u += workers) {
        const UnitCoord uc = decode_unit(u, p, meta);
        const int it0 = __shfl_sync(0xffffffffu, uc.it0, 0), it1 = __shfl_sync(0xffffffffu, uc.it1, 0);
        const int cnt = max(__shfl_sync(0xffffffffu, meta->cnt[uc.c], 0), 1);
        if (it1 <= it0) continue;
        ptx::mbar_wait(&tmem_empty[acc_buf], acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base_u + acc_buf * kAccCols;
        int kb = it0 / cnt, j = it0 % cnt;  // same order as the producer: K block outer, segment inner
        for (int it = it0; it < it1; ++it) {
          const int ksteps = (kb == p.kb_per_seg - 1) ? p.kb_last_steps : kBKk / 16;
          if (++j == cnt) { j = 0; ++kb; }
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = tiles_u + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          // MN-major: chunks of 64 columns are kBKk rows x 128 B apart (LBO), groups of 8 k-rows 1 KB apart (SBO)
          auto desc = [](uint32_t addr) {
            return MN ? ptx::umma_desc_mn_sw128(addr, kBKk * 128, 1024) : ptx::umma_desc_k_sw128(addr);
          };
          const uint64_t a_hi = desc(sa);
          const uint64_t a_lo = desc(sa + (NPA > 1 ? 1 : 0) * Cfg::kAPlane);
          const uint64_t b_hi = desc(sb);
          const uint64_t b_lo = desc(sb + (NPB > 1 ? 1 : 0) * Cfg::kBPlane);
          const uint64_t a_l2 = desc(sa + (NPA - 1) * Cfg::kAPlane);  // third plane (NP == 3)
          const uint64_t b_l2 = desc(sb + (NPB - 1) * Cfg::kBPlane);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBKk / 16; ++k) {
              if (k >= ksteps) break;
              // K-major: 16 elements = 32 B = 2 x 16 B units inside the swizzle atom; MN-major: 16 k-rows = 2 KB = 128 units
              const uint64_t adv = (uint64_t)(MN ? k * 128 : k * 2);
              const uint32_t first = (it > it0 || k > 0) ? 1u : 0u;
              if (NPA == 3) {
                // smallest products first: (lo, hi) and (mid, mid) are ~2^-16 of (hi, hi); (mid, lo), (lo, lo) < 2^-24 dropped
                mma(d_tmem, a_l2 + adv, b_hi + adv, idesc, first);
                mma(d_tmem, a_hi + adv, b_l2 + adv, idesc, 1u);
                mma(d_tmem, a_lo + adv, b_lo + adv, idesc, 1u);
                mma(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                mma(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                mma(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              } else if (NPA == 2 && NPB == 2) {
                mma(d_tmem, a_hi + adv, b_hi + adv, idesc, first);
                mma(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                mma(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              } else {  // (1,1)
                mma(d_tmem, a_hi + adv, b_hi + adv, idesc, first);
              }
            }
            commit(&empty_bar[stage]);  // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) commit(&tmem_full[acc_buf]);
        __syncwarp();
        acc_buf ^= 1;
        if (acc_buf == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================================================= epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (!p.transposed) && ((p.out_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                        ((p.col_cadd & 3) == 0);
    // undo the power-of-two scaling of single-plane fp16 operands (exact)
    const float descale = 1.0f / ((p.absmax_a ? pow2_scale(*p.absmax_a) : 1.0f) * (p.absmax_b ? pow2_scale(*p.absmax_b) : 1.0f));
    const bool dot_vec_ok = p.dot_ref != nullptr && ((reinterpret_cast<uintptr_t>(p.dot_ref) & 15) == 0) &&
                            ((p.dot_ld & 3) == 0) && ((p.col_cadd & 3) == 0);
    const uint32_t tmem_empty_leader[2] = {PAIR ? mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0) : 0u,
                                           PAIR ? mapa_shared(ptx::smem_u32(&tmem_empty[1]), 0) : 0u};
    for (int u = worker; u < total_units; u += workers) {
      UnitCoord uc = decode_unit(u, p, meta);
      if constexpr (PAIR) uc.m_t = 2 * uc.m_t + (int)rank;
      const bool has_acc = uc.it1 > uc.it0;
      if (!has_acc && p.accumulate) continue;  // nothing to add
      if (has_acc) {
        ptx::mbar_wait(&tmem_full[acc_buf], acc_phase);
        ptx::tc_fence_after_sync();
      }
      const int m = uc.m_t * kBM + quarter * 32 + lane;
      const bool row_ok = m < meta->m_valid[uc.c];
      const long long R = (long long)m * p.row_mul + (long long)uc.c * p.row_cadd;
      const float scale = p.alpha * meta->c_scale[uc.c] * descale;
      const bool add_bias = (p.bias != nullptr) && (uc.split == 0);
      if (Cfg::kEpiBytes > 0 && vec_ok && p.dot_ref == nullptr) {
        // ---- staged epilogue: 32 columns at a time through this warp's staging block
        float* stg = epi_stage + (warp - 2) * (32 * Cfg::kEpiPitch);
        const int cq = lane & 7, rsub = lane >> 3;
#pragma unroll 1
        for (int g = 0; g < BN / 32; ++g) {
          const int n = uc.n_t * BN + g * 32 + cq * 4;  // first of this lane's 4 columns in the row-major phase
          if (uc.n_t * BN + g * 32 >= p.n_valid) break;  // warp-uniform
          uint32_t v[32];
          __syncwarp();  // the previous block has been read; tcgen05.ld is .sync.aligned
          if (has_acc) {
            ptx::tmem_ld_32x32(tmem_base + acc_buf * kAccCols + ((uint32_t)(quarter * 32) << 16) + g * 32, v);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = scale * __uint_as_float(v[j]);
            o.y = scale * __uint_as_float(v[j + 1]);
            o.z = scale * __uint_as_float(v[j + 2]);
            o.w = scale * __uint_as_float(v[j + 3]);
            *reinterpret_cast<float4*>(stg + lane * Cfg::kEpiPitch + j) = o;
          }
          __syncwarp();
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add_bias && n < p.n_valid) {  // the bias tail is not 16-byte aligned in general
            b4.x = __ldg(p.bias + n);
            if (n + 1 < p.n_valid) b4.y = __ldg(p.bias + n + 1);
            if (n + 2 < p.n_valid) b4.z = __ldg(p.bias + n + 2);
            if (n + 3 < p.n_valid) b4.w = __ldg(p.bias + n + 3);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            const int mm = uc.m_t * kBM + quarter * 32 + rr;
            if (mm >= meta->m_valid[uc.c] || n >= p.n_valid) continue;
            float4 o = *reinterpret_cast<const float4*>(stg + rr * Cfg::kEpiPitch + cq * 4);
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            float* dst = p.out + ((long long)mm * p.row_mul + (long long)uc.c * p.row_cadd) * p.out_ld + n +
                         (long long)uc.c * p.col_cadd;
            if (n + 4 <= p.n_valid) {
              if (p.atomic || p.accumulate) red_add_v4_f32(dst, o.x, o.y, o.z, o.w);
              else *reinterpret_cast<float4*>(dst) = o;
            } else {
              const float e[4] = {o.x, o.y, o.z, o.w};
              for (int j = 0; j < 4 && n + j < p.n_valid; ++j) {
                if (p.atomic || p.accumulate) red_add_f32(dst + j, e[j]);
                else dst[j] = e[j];
              }
            }
          }
        }
        __syncwarp();
        if (has_acc) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader[acc_buf]);
            else ptx::mbar_arrive(&tmem_empty[acc_buf]);
          }
          acc_buf ^= 1;
          if (acc_buf == 0) acc_phase ^= 1;
        }
        continue;
      }
      float dot = 0.f;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 16; ++chunk) {
        uint32_t v[16];
        float ref[16];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
        const int n0 = uc.n_t * BN + chunk * 16;
        const bool live = row_ok && n0 < p.n_valid;
        const long long C0 = (long long)n0 + (long long)uc.c * p.col_cadd;
        const bool full = (n0 + 16 <= p.n_valid);
        if (has_acc)
          ptx::tmem_ld_32x16(tmem_base + acc_buf * kAccCols + ((uint32_t)(quarter * 32) << 16) + chunk * 16, v);
        // the dot_ref reads are issued before waiting for TMEM so that their L2 latency overlaps the load
        if (p.dot_ref != nullptr && live) {
          if (!p.transposed) {
            const float* rp = p.dot_ref + R * p.dot_ld + C0;
            if (full && dot_vec_ok) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                ref[j] = r4.x; ref[j + 1] = r4.y; ref[j + 2] = r4.z; ref[j + 3] = r4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) ref[j] = (n0 + j < p.n_valid) ? rp[j] : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ref[j] = (n0 + j < p.n_valid) ? p.dot_ref[(C0 + j) * p.dot_ld + R] : 0.f;
          }
        }
        if (has_acc) {
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (live) {
          if (p.dot_ref != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dot += __uint_as_float(v[j]) * ref[j];  // descaled once per unit below
          }
          if (!p.transposed) {
            float* dst = p.out + R * p.out_ld + C0;
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 o;
                o.x = scale * __uint_as_float(v[j]);
                o.y = scale * __uint_as_float(v[j + 1]);
                o.z = scale * __uint_as_float(v[j + 2]);
                o.w = scale * __uint_as_float(v[j + 3]);
                if (add_bias) {  // the bias tail starts n floats into bias_params_: not 16-byte aligned in general
                  const float* bj = p.bias + n0 + j;
                  o.x += __ldg(bj); o.y += __ldg(bj + 1); o.z += __ldg(bj + 2); o.w += __ldg(bj + 3);
                }
                // accumulation is always a fire-and-forget red.add (L2 does the read-modify-write): loading the old
                // value first puts an L2 round trip per 16 columns on the epilogue's critical path
                if (p.atomic || p.accumulate) red_add_v4_f32(dst + j, o.x, o.y, o.z, o.w);
                else *reinterpret_cast<float4*>(dst + j) = o;
              }
            } else {
              for (int j = 0; j < 16; ++j) {
                if (n0 + j >= p.n_valid) break;
                float o = scale * __uint_as_float(v[j]);
                if (add_bias) o += p.bias[n0 + j];
                if (p.atomic || p.accumulate) red_add_f32(dst + j, o);
                else dst[j] = o;
              }
            }
          } else {
            // transposed: consecutive lanes (rows m) are consecutive addresses -> coalesced per register
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n0 + j < p.n_valid) {
                const long long idx = (C0 + j) * p.out_ld + R;
                float o = scale * __uint_as_float(v[j]);
                if (add_bias) o += p.bias[n0 + j];
                // accumulate always goes through red.add: a plain `+=` here is a chain of BN dependent
                // load-add-store round trips to L2 (~0.6 us each) per thread
                if (p.atomic || p.accumulate) red_add_f32(p.out + idx, o);
                else p.out[idx] = o;
              }
            }
          }
        }
      }
      __syncwarp();
      if (p.dot_ref != nullptr) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        if (lane == 0 && dot != 0.f) atomicAdd(p.dot_out + uc.c, dot * descale);
      }
      if (has_acc) {
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader[acc_buf]);
          else ptx::mbar_arrive(&tmem_empty[acc_buf]);
        }
        acc_buf ^= 1;
        if (acc_buf == 0) acc_phase ^= 1;
      }
    }
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncwarp();  // the single-lane role branches above: barrier.cluster is .aligned
  if constexpr (PAIR) cluster_sync_all();  // the partner may still be reading this CTA's B half / signalling its barriers
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after_sync();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 2 * kAccCols);
    else ptx::tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}

}  // namespace tdnnf
