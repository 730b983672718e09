// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tdnnf_nas_b200.h"

namespace tdnnf {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define TDNNF_CUDA_OK(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::tdnnf::fail(TDNNF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define TDNNF_REQUIRE(cond, msg)                                                   \
  do {                                                                             \
    if (!(cond)) return ::tdnnf::fail(TDNNF_ERR_INVALID, std::string(msg));        \
  } while (0)

}  // namespace tdnnf

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
  };
  std::vector<GemmTiming> gemm_events;

  void ws_reset() { ws_off = 0; }
  // Returns nullptr on failure (error string set).
  void* ws_alloc(size_t bytes);
  int ws_reserve(size_t bytes);
};
