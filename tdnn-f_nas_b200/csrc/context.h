// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tdnnf_nas_b200.h"

namespace tdnnf {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void set_error(const std::string& msg);
}  // namespace tdnnf
struct tdnnf_ctx;
struct tdnnf_planes;
namespace tdnnf {
int planes_alloc_for_producer(tdnnf_ctx* ctx, const float* src, int rows, int cols, int stride, tdnnf_planes** out);
int fail(int code, const std::string& msg);

#define TDNNF_CUDA_OK(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::tdnnf::fail(TDNNF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

// Kernel launch with programmatic dependent launch allowed (the kernel must begin with ptx::grid_dep_wait() before it
// touches anything a predecessor wrote): its CTAs are scheduled while the tail of the previous kernel drains, so launch
// latency and prologue leave the critical path.  TDNNF_PDL=0 turns the attribute off.  cluster_x > 1 adds a cluster.
bool pdl_enabled(const char* file = nullptr);  // TDNNF_PDL_OFF=<substrings of source file names, comma separated> turns it off per file
int zero_async(tdnnf_ctx* ctx, void* p, size_t bytes);  // zero-fill on the context's stream (a kernel)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_at(const char* file, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled(file)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

#define launch_pdl(...) launch_pdl_at(__FILE__, __VA_ARGS__)

#define TDNNF_REQUIRE(cond, msg)                                                   \
  do {                                                                             \
    if (!(cond)) return ::tdnnf::fail(TDNNF_ERR_INVALID, std::string(msg));        \
  } while (0)

}  // namespace tdnnf

// Operand planes with a life of their own (tdnnf_planes_*): the bf16 hi/lo row planes (and the per-row sums of squares /
// column sums that come with building them) of one fp32 matrix, kept from Propagate to Backprop in the component's memo
// or written by the kernel that produces the matrix.  While attached to the context, every split of that matrix is a hit.
struct tdnnf_planes {
  tdnnf_ctx* ctx = nullptr;
  const float* src = nullptr;
  int R = 0, D = 0, r = 1, Q = 0, Kpad = 0, np = 2;
  long long ld = 0;
  void* block = nullptr;      // one pooled allocation: [planes][rowsq: R floats][colsum: D floats]
  size_t bytes = 0;
  void* base = nullptr;       // hi plane; plane p at base + p * plane_elems (bf16)
  long long plane_elems = 0;
  float* rowsq = nullptr;     // [R] sum of squares of every source row
  float* colsum = nullptr;    // [D] column sums (valid only when has_colsum)
  bool has_colsum = false;
  int refs = 1, attached = 0;
};

// per-offset row offsets passed by value to kernels
struct TdnnfOffsets {
  int v[TDNNF_MAX_OFFSETS];
};

// The opaque context of the C ABI.  One per (host thread, device); owns the stream the
// kernels are launched on and a grow-only scratch arena for operand planes.
struct tdnnf_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  tdnnf::EncodeTiledFn encode = nullptr;
  // scratch arena (device).  Carved per public call by ws_reset()/ws_alloc(); grown on demand
  // (growth synchronises the stream, steady state does not).
  char* ws = nullptr;
  size_t ws_bytes = 0;
  size_t ws_off = 0;
  // operand planes of the tensor-core GEMMs: 2 = bf16 hi/lo (three products), 3 = hi/mid/lo (six products)
  int gemm_planes = 2;
  // counters for bench.py's gpu_launches claim
  unsigned long long launches = 0;
  // optional per-GEMM-launch event timing (roofline instrumentation)
  bool gemm_timing = false;
  struct GemmTiming {
    cudaEvent_t start, stop;
    double flops;
    int products;  // tensor-core products per K step of this launch (1, 2, 3 or 6)
  };
  std::vector<GemmTiming> gemm_events;

  void ws_reset() { ws_off = 0; }
  // Returns nullptr on failure (error string set).
  void* ws_alloc(size_t bytes);
  int ws_reserve(size_t bytes);

  // Operand-plane cache (tdnnf_ctx_operand_cache_begin/end): inside a scope the planes built from a REGISTERED
  // source matrix are kept in a second arena and reused by later calls that split the same matrix the same way
  // (Backprop splits in_value / out_deriv for the data gradient, the two natural-gradient projections and the
  // parameter gradient).  The caller promises that registered sources do not change inside the scope.
  struct PlaneCacheEntry {
    int kind;  // 0: rows, 1: transposed
    int fp16;  // 1: one scaled fp16 plane, 0: bf16 hi/lo(/lo2) planes
    const float* src;
    int R, D;
    long long ld;
    int r, groups, c_row_mul, c_col_mul;
    const float* scale;
    int Q, pitch;
    int offs[16];
    int np;
    void* base;
    long long plane_elems;
  };
  bool cache_on = false;
  std::vector<const float*> cache_srcs;
  std::vector<PlaneCacheEntry> cache;
  char* cws = nullptr;
  size_t cws_bytes = 0, cws_off = 0;
  unsigned long long cache_hits = 0, cache_misses = 0;
  bool cache_registered(const float* p) const {
    if (!cache_on) return false;
    for (const float* q : cache_srcs)
      if (q == p) return true;
    return false;
  }
  int cache_source_index(const float* p) const {
    if (!cache_on) return -1;
    for (size_t i = 0; i < cache_srcs.size(); ++i)
      if (cache_srcs[i] == p) return (int)i;
    return -1;
  }
  // max |x| of each registered source (device floats), filled by the first row split or by absmax_kernel
  float* absmax_dev = nullptr;
  bool absmax_valid[8] = {false, false, false, false, false, false, false, false};
  // per-row sums of squares of each registered source (device floats), by-product of its first row split
  float* rowsq_dev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t rowsq_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int rowsq_rows[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool rowsq_valid[8] = {false, false, false, false, false, false, false, false};
  // scratch of tdnnf_ng_gram_scale: {double acc[2], uint counter, pad to 64 B, per-CTA partial Gram matrices}
  char* ng_scratch = nullptr;
  size_t ng_scratch_bytes = 0;
  // parameter-gradient GEMM precision: false = three bf16 products; true = one fp16 x fp16 product (power-of-two scaled)
  bool grad_fast = false;
  // parameter gradient: operands with at least this many rows take the MN-major form (row planes, no transposed
  // pre-pass); < 0 = never.  Default 512 (TDNNF_WGRAD_MN=0 disables, TDNNF_WGRAD_MN_MIN_ROWS overrides).
  int wgrad_mn_min_rows = 512;
  // tdnnf_planes_*: attached plane sets (matched by source pointer, shape, stride and row stride) and the pool their
  // blocks come from / return to (nothing is handed back to the driver)
  std::vector<tdnnf_planes*> attached_planes;
  std::vector<std::pair<size_t, void*>> planes_pool;
  const tdnnf_planes* find_attached(const float* src, int R, int D, long long ld, int r) const {
    for (const tdnnf_planes* p : attached_planes)
      if (p->src == src && p->R == R && p->D == D && p->ld == ld && p->r == r) return p;
    return nullptr;
  }
  // Makes room for `bytes` more cached planes; called at the top of a public call, before any plane of that call
  // exists (growing drops every cached plane).
  int cws_reserve(size_t bytes);
  void* cws_alloc(size_t bytes);
};
