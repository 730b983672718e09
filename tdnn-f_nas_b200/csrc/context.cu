// Context, error reporting and scratch arena of the C ABI.
#include <algorithm>
#include <cstdio>
#include <string>

#include "context.h"
#include <cstdlib>

namespace tdnnf {

// zero-fill as a kernel: unlike a memset node it takes part in programmatic dependent launch chains
__global__ void zero_words_kernel(uint32_t* __restrict__ p, size_t n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

int zero_async(tdnnf_ctx* ctx, void* p, size_t bytes) {
  if (bytes == 0) return TDNNF_OK;
  if ((bytes & 3) != 0 || (reinterpret_cast<uintptr_t>(p) & 3) != 0) {
    TDNNF_CUDA_OK(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    return TDNNF_OK;
  }
  const size_t n = bytes / 4;
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 8));
  TDNNF_CUDA_OK(launch_pdl(zero_words_kernel, dim3(blocks), dim3(256), 0, ctx->stream, 1, static_cast<uint32_t*>(p), n));
  ctx->launches++;
  return TDNNF_OK;
}

bool pdl_enabled(const char* file) {
  static const bool on = [] {
    const char* e = getenv("TDNNF_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  static const std::string off = [] {
    const char* e = getenv("TDNNF_PDL_OFF");
    return std::string(e ? e : "");
  }();
  if (!on) return false;
  if (file == nullptr || off.empty()) return true;
  const char* base = strrchr(file, '/');
  base = base ? base + 1 : file;
  size_t pos = 0;
  while (pos <= off.size()) {
    size_t comma = off.find(',', pos);
    if (comma == std::string::npos) comma = off.size();
    const std::string tok = off.substr(pos, comma - pos);
    if (!tok.empty() && strstr(base, tok.c_str()) != nullptr) return false;
    pos = comma + 1;
  }
  return true;
}


static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

}  // namespace tdnnf

using namespace tdnnf;

void* tdnnf_ctx::ws_alloc(size_t bytes) {
  bytes = (bytes + 1023) & ~size_t(1023);
  if (ws_off + bytes > ws_bytes) {
    set_error("internal: scratch arena overflow (ws_reserve was not called with the full size)");
    return nullptr;
  }
  void* p = ws + ws_off;
  ws_off += bytes;
  return p;
}

int tdnnf_ctx::ws_reserve(size_t bytes) {
  bytes = ((bytes + 1023) & ~size_t(1023)) + 4096;
  if (bytes <= ws_bytes) return TDNNF_OK;
  // Growing: earlier kernels on the stream may still read the old arena.
  TDNNF_CUDA_OK(cudaStreamSynchronize(stream));
  if (ws) TDNNF_CUDA_OK(cudaFree(ws));
  ws = nullptr;
  ws_bytes = 0;
  size_t want = bytes + bytes / 4;  // head-room so that slightly larger minibatches do not regrow
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ws), want);
  if (e != cudaSuccess) {
    want = bytes;
    e = cudaMalloc(reinterpret_cast<void**>(&ws), want);
  }
  if (e != cudaSuccess) return fail(TDNNF_ERR_NOMEM, std::string("cudaMalloc of scratch arena failed: ") + cudaGetErrorString(e));
  ws_bytes = want;
  return TDNNF_OK;
}

int tdnnf_ctx::cws_reserve(size_t bytes) {
  if (!cache_on) return TDNNF_OK;
  bytes = ((bytes + 1023) & ~size_t(1023)) + 8192;
  if (cws_off + bytes <= cws_bytes) return TDNNF_OK;
  // Growing: kernels on the stream may still read cached planes, and the cached planes move.
  TDNNF_CUDA_OK(cudaStreamSynchronize(stream));
  cache.clear();
  const size_t want = std::max(cws_off + bytes, cws_bytes) * 2;
  if (cws) TDNNF_CUDA_OK(cudaFree(cws));
  cws = nullptr;
  cws_bytes = cws_off = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&cws), want);
  if (e != cudaSuccess) return fail(TDNNF_ERR_NOMEM, std::string("cudaMalloc of the operand cache failed: ") + cudaGetErrorString(e));
  cws_bytes = want;
  return TDNNF_OK;
}

void* tdnnf_ctx::cws_alloc(size_t bytes) {
  bytes = (bytes + 1023) & ~size_t(1023);
  if (cws_off + bytes > cws_bytes) {
    set_error("internal: operand cache overflow (cws_reserve was not called with the full size)");
    return nullptr;
  }
  void* p = cws + cws_off;
  cws_off += bytes;
  return p;
}

extern "C" int tdnnf_ctx_operand_cache_begin(tdnnf_ctx* ctx, const float* const* sources, int num_sources) {
  TDNNF_REQUIRE(ctx != nullptr && (num_sources == 0 || sources != nullptr), "null argument");
  TDNNF_REQUIRE(num_sources >= 0 && num_sources <= 8, "at most 8 cached sources");
  TDNNF_REQUIRE(!ctx->cache_on, "operand cache scopes do not nest");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  if (!ctx->absmax_dev) TDNNF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ctx->absmax_dev), sizeof(float) * 8));
  TDNNF_CUDA_OK(cudaMemsetAsync(ctx->absmax_dev, 0, sizeof(float) * 8, ctx->stream));
  for (bool& v : ctx->absmax_valid) v = false;
  for (bool& v : ctx->rowsq_valid) v = false;
  ctx->cache_on = true;
  ctx->cache.clear();
  ctx->cws_off = 0;
  ctx->cache_srcs.assign(sources, sources + num_sources);
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_operand_cache_end(tdnnf_ctx* ctx) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  ctx->cache_on = false;
  ctx->cache.clear();
  ctx->cache_srcs.clear();
  for (bool& v : ctx->rowsq_valid) v = false;
  ctx->cws_off = 0;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_operand_cache_stats(const tdnnf_ctx* ctx, uint64_t* hits, uint64_t* misses) {
  TDNNF_REQUIRE(ctx && hits && misses, "null argument");
  *hits = ctx->cache_hits;
  *misses = ctx->cache_misses;
  return TDNNF_OK;
}

extern "C" const char* tdnnf_last_error(void) { return g_last_error.c_str(); }

extern "C" int tdnnf_abi_version(void) { return 1002; }  // 1002: orthonormal constraint, dropout, log-softmax, MN-major parameter-gradient switch

extern "C" int tdnnf_ctx_create(int device, tdnnf_ctx** out) {
  TDNNF_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(TDNNF_ERR_CUDA, std::string("no CUDA device available (there is no CPU path): ") +
                                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  TDNNF_REQUIRE(device >= 0 && device < count, "device index out of range");
  TDNNF_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  TDNNF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(TDNNF_ERR_UNSUPPORTED, "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                           "; this library contains sm_100a code only");
  tdnnf_ctx* ctx = new tdnnf_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete ctx;
    return fail(TDNNF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  *out = ctx;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_destroy(tdnnf_ctx* ctx) {
  if (!ctx) return TDNNF_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->cws) cudaFree(ctx->cws);
  if (ctx->absmax_dev) cudaFree(ctx->absmax_dev);
  if (ctx->ng_scratch) cudaFree(ctx->ng_scratch);
  for (float* p : ctx->rowsq_dev)
    if (p) cudaFree(p);
  delete ctx;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_set_stream(tdnnf_ctx* ctx, void* stream) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  ctx->stream = static_cast<cudaStream_t>(stream);
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_reserve(tdnnf_ctx* ctx, uint64_t bytes) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  return ctx->ws_reserve((size_t)bytes);
}

extern "C" uint64_t tdnnf_ctx_launch_count(const tdnnf_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int tdnnf_ctx_gemm_timing_enable(tdnnf_ctx* ctx, int enable) {
  TDNNF_REQUIRE(ctx != nullptr, "null context");
  ctx->gemm_timing = enable != 0;
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_gemm_timing_read(tdnnf_ctx* ctx, double* total_ms, double* total_flops, uint64_t* launches) {
  TDNNF_REQUIRE(ctx && total_ms && total_flops && launches, "null argument");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  double ms = 0.0, fl = 0.0;
  for (auto& t : ctx->gemm_events) {
    float e = 0.f;
    TDNNF_CUDA_OK(cudaEventElapsedTime(&e, t.start, t.stop));
    ms += e;
    fl += t.flops;
    cudaEventDestroy(t.start);
    cudaEventDestroy(t.stop);
  }
  *total_ms = ms;
  *total_flops = fl;
  *launches = ctx->gemm_events.size();
  ctx->gemm_events.clear();
  return TDNNF_OK;
}

extern "C" int tdnnf_ctx_gemm_timing_read_ex(tdnnf_ctx* ctx, double min_flops, double* total_ms, double* total_flops,
                                             double* tensor_pipe_flops, uint64_t* launches, double* other_ms,
                                             uint64_t* other_launches) {
  TDNNF_REQUIRE(ctx && total_ms && total_flops && tensor_pipe_flops && launches && other_ms && other_launches, "null argument");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  TDNNF_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  double ms = 0.0, fl = 0.0, raw = 0.0, oms = 0.0;
  uint64_t n = 0, on = 0;
  for (auto& t : ctx->gemm_events) {
    float e = 0.f;
    TDNNF_CUDA_OK(cudaEventElapsedTime(&e, t.start, t.stop));
    if (t.flops >= min_flops) {
      ms += e;
      fl += t.flops;
      raw += t.flops * t.products;
      n++;
    } else {
      oms += e;
      on++;
    }
    cudaEventDestroy(t.start);
    cudaEventDestroy(t.stop);
  }
  *total_ms = ms;
  *total_flops = fl;
  *tensor_pipe_flops = raw;
  *launches = n;
  *other_ms = oms;
  *other_launches = on;
  ctx->gemm_events.clear();
  return TDNNF_OK;
}
