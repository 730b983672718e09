// Host glue of one training step above the kernels, for a C++ host that has no Python:
//   * ComputeChainObjfAndDeriv (kaldi: chain/chain-training.cc) on the denominator / numerator entries;
//   * the data-parallel reduction of the deltas over NCCL (replaces the multi-job nnet3-average of
//     steps/libs/nnet3/train/common.py:144-164: summing per-job deltas).
// NCCL is bound at run time (dlopen of the libnccl already in the process, else libnccl.so.2): the library itself has no
// link-time dependency on it, and a single-GPU host never loads it.
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <string>

#include "context.h"

using namespace tdnnf;

extern "C" int tdnnf_chain_objf_and_deriv(tdnnf_ctx* ctx, tdnnf_den_comp* den, const tdnnf_num_graph* num, const float* nnet_output,
                                          int stride, int num_seqs, int frames_per_seq, int num_pdfs, float supervision_weight,
                                          float l2_regularize, float out_of_range_regularize, int oor_row_step, int oor_row_offset,
                                          float* nnet_output_deriv, int deriv_stride, float* xent_output_deriv, int xent_stride,
                                          float* objf, float* l2_term, float* weight) {
  TDNNF_REQUIRE(ctx && den && num && nnet_output && objf && l2_term && weight, "null argument");
  TDNNF_REQUIRE(num_seqs > 0 && frames_per_seq > 0 && num_pdfs > 0 && stride >= num_pdfs, "bad shape");
  TDNNF_REQUIRE(nnet_output_deriv != nullptr || xent_output_deriv == nullptr, "xent_output_deriv needs nnet_output_deriv");
  const int rows = num_seqs * frames_per_seq;
  const float w = supervision_weight;
  *weight = w * num_seqs * frames_per_seq;
  *l2_term = 0.f;
  int rc;
  if (nnet_output_deriv) {
    TDNNF_REQUIRE(deriv_stride >= num_pdfs, "deriv stride < num_pdfs");
    if ((rc = tdnnf_mat_set(ctx, nnet_output_deriv, rows, num_pdfs, deriv_stride, 0.f))) return rc;
  }
  // the denominator first, as upstream (it frees its workspace before the numerator allocates)
  float den_logprob = 0.f;
  int den_ok = 1;
  if ((rc = tdnnf_den_forward(den, nnet_output, stride, &den_logprob))) return rc;
  const float den_logprob_weighted = w * den_logprob;
  if (nnet_output_deriv && (rc = tdnnf_den_backward(den, -w, nnet_output_deriv, deriv_stride, &den_ok))) return rc;
  float num_logprob = 0.f;
  int num_ok = 1;
  if (xent_output_deriv) {
    TDNNF_REQUIRE(xent_stride >= num_pdfs, "xent stride < num_pdfs");
    if ((rc = tdnnf_mat_set(ctx, xent_output_deriv, rows, num_pdfs, xent_stride, 0.f))) return rc;
    if ((rc = tdnnf_num_forward_backward(ctx, num, nnet_output, stride, frames_per_seq, w, xent_output_deriv, xent_stride,
                                         &num_logprob, &num_ok)))
      return rc;
    if (num_ok && (rc = tdnnf_mat_axpy(ctx, 1.f, xent_output_deriv, xent_stride, nnet_output_deriv, deriv_stride, rows, num_pdfs)))
      return rc;
  } else {
    if ((rc = tdnnf_num_forward_backward(ctx, num, nnet_output, stride, frames_per_seq, w, nnet_output_deriv, deriv_stride,
                                         &num_logprob, &num_ok)))
      return rc;
  }
  const float num_logprob_weighted = w * num_logprob;
  *objf = num_logprob_weighted - den_logprob_weighted;
  if (!((*objf) - (*objf) == 0.f) || !den_ok || !num_ok) {
    // inf / NaN, or a failed alpha.beta check: zero derivatives, -10 per frame
    if (nnet_output_deriv && (rc = tdnnf_mat_set(ctx, nnet_output_deriv, rows, num_pdfs, deriv_stride, 0.f))) return rc;
    if (xent_output_deriv && (rc = tdnnf_mat_set(ctx, xent_output_deriv, rows, num_pdfs, xent_stride, 0.f))) return rc;
    *objf = -10.f * *weight;
  }
  if (l2_regularize != 0.f && num_ok) {
    const float scale = w * l2_regularize;
    float tr = 0.f;
    if ((rc = tdnnf_mat_dot(ctx, nnet_output, stride, nnet_output, stride, rows, num_pdfs, &tr))) return rc;
    *l2_term = -0.5f * scale * tr;
    if (nnet_output_deriv && (rc = tdnnf_mat_axpy(ctx, -scale, nnet_output, stride, nnet_output_deriv, deriv_stride, rows, num_pdfs)))
      return rc;
  }
  if (nnet_output_deriv && out_of_range_regularize != 0.f) {
    // PenalizeOutOfRange(nnet_output, 30.0, 2.0 * out_of_range_regularize, deriv) on every row_step-th row, the scale
    // multiplied by row_step to compensate for the sub-sampling.
    const int step = oor_row_step >= 1 ? oor_row_step : 1;
    if ((rc = tdnnf_penalize_out_of_range(ctx, nnet_output, rows, num_pdfs, stride, 30.f, 2.f * out_of_range_regularize * step, step,
                                          oor_row_offset, nnet_output_deriv, deriv_stride)))
      return rc;
  }
  return TDNNF_OK;
}

// ------------------------------------------------------------------------------------------- NCCL (bound at run time)
namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
};
constexpr int kNcclFloat = 7, kNcclSum = 0;  // ncclFloat32, ncclSum (stable enum values of nccl.h)

std::string g_nccl_error = "NCCL is not loadable (libnccl.so.2)";

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  // a host process that already carries an NCCL (torch bundles one) must not get a second copy
  for (const char* nm : names)
    if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  for (const char* nm : names)
    if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) {
    const char* why = dlerror();
    g_nccl_error = api.error = std::string("NCCL is not loadable: ") + (why ? why : "libnccl.so.2 not found");
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char* nm) {
    void* p = dlsym(api.handle, nm);
    if (!p) {
      ok = false;
      api.error = std::string("NCCL symbol missing: ") + nm;
    }
    return p;
  };
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(ncclComm_t*, int, NcclUniqueId, int)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)>(sym("ncclAllReduce"));
  api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  if (!ok) {
    g_nccl_error = api.error;
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

}  // namespace

struct tdnnf_dp_comm {
  tdnnf_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  bool owned = true;
  cudaStream_t side = nullptr;     // bucketed reductions overlapped with the backward pass run here
  cudaEvent_t ready = nullptr, done = nullptr;
};

#define TDNNF_NCCL_OK(api, expr)                                                                                \
  do {                                                                                                          \
    int _r = (expr);                                                                                            \
    if (_r != 0) return fail(TDNNF_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(_r));            \
  } while (0)

extern "C" int tdnnf_dp_unique_id(char* id_out, int id_bytes) {
  TDNNF_REQUIRE(id_out && id_bytes >= 128, "id_out must hold 128 bytes");
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  NcclUniqueId id;
  TDNNF_NCCL_OK(api, api->GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return TDNNF_OK;
}

extern "C" int tdnnf_dp_comm_create(tdnnf_ctx* ctx, int nranks, int rank, const char* unique_id, tdnnf_dp_comm** out) {
  TDNNF_REQUIRE(ctx && unique_id && out, "null argument");
  TDNNF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  NcclUniqueId id;
  memcpy(id.internal, unique_id, 128);
  tdnnf_dp_comm* c = new tdnnf_dp_comm();
  c->ctx = ctx;
  c->nranks = nranks;
  c->rank = rank;
  int r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    delete c;
    return fail(TDNNF_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
  }
  cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
  *out = c;
  return TDNNF_OK;
}

// Adopts an ncclComm_t the host already has (a Kaldi-side trainer that set NCCL up itself); not destroyed here.
extern "C" int tdnnf_dp_comm_adopt(tdnnf_ctx* ctx, void* nccl_comm, int nranks, int rank, tdnnf_dp_comm** out) {
  TDNNF_REQUIRE(ctx && nccl_comm && out, "null argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  TDNNF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
  TDNNF_CUDA_OK(cudaSetDevice(ctx->device));
  tdnnf_dp_comm* c = new tdnnf_dp_comm();
  c->ctx = ctx;
  c->comm = static_cast<ncclComm_t>(nccl_comm);
  c->nranks = nranks;
  c->rank = rank;
  c->owned = false;
  cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
  *out = c;
  return TDNNF_OK;
}

extern "C" int tdnnf_dp_comm_destroy(tdnnf_dp_comm* c) {
  if (!c) return TDNNF_OK;
  NcclApi* api = nccl_api();
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ready) cudaEventDestroy(c->ready);
  if (c->done) cudaEventDestroy(c->done);
  if (api && c->owned && c->comm) api->CommDestroy(c->comm);
  delete c;
  return TDNNF_OK;
}

// Sum over ranks, in place, of n delta buffers (device pointers, `counts` floats each; a strided matrix is passed as
// rows * stride floats: its pitch padding is zero on every rank).  One NCCL group on the context's stream.
extern "C" int tdnnf_dp_allreduce_deltas(tdnnf_dp_comm* c, int n, float* const* bufs, const int64_t* counts) {
  TDNNF_REQUIRE(c && bufs && counts && n >= 0, "bad argument");
  if (c->nranks == 1 || n == 0) return TDNNF_OK;
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  TDNNF_CUDA_OK(cudaSetDevice(c->ctx->device));
  TDNNF_NCCL_OK(api, api->GroupStart());
  for (int i = 0; i < n; ++i) {
    if (counts[i] <= 0) continue;
    int r = api->AllReduce(bufs[i], bufs[i], (size_t)counts[i], kNcclFloat, kNcclSum, c->comm, c->ctx->stream);
    if (r != 0) {
      api->GroupEnd();
      return fail(TDNNF_ERR_CUDA, std::string("ncclAllReduce: ") + api->GetErrorString(r));
    }
  }
  TDNNF_NCCL_OK(api, api->GroupEnd());
  return TDNNF_OK;
}

// The same reduction for ONE bucket, overlapped with whatever the context's stream does next: the side stream waits
// for the work already queued on the context's stream (the backward pass that produced the bucket), reduces, and
// tdnnf_dp_allreduce_wait makes the context's stream wait for every bucket issued so far.
extern "C" int tdnnf_dp_allreduce_bucket_async(tdnnf_dp_comm* c, float* buf, int64_t count) {
  TDNNF_REQUIRE(c && buf && count >= 0, "bad argument");
  if (c->nranks == 1 || count == 0) return TDNNF_OK;
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  TDNNF_CUDA_OK(cudaSetDevice(c->ctx->device));
  TDNNF_CUDA_OK(cudaEventRecord(c->ready, c->ctx->stream));
  TDNNF_CUDA_OK(cudaStreamWaitEvent(c->side, c->ready, 0));
  TDNNF_NCCL_OK(api, api->AllReduce(buf, buf, (size_t)count, kNcclFloat, kNcclSum, c->comm, c->side));
  return TDNNF_OK;
}

extern "C" int tdnnf_dp_allreduce_wait(tdnnf_dp_comm* c) {
  TDNNF_REQUIRE(c, "null argument");
  if (c->nranks == 1) return TDNNF_OK;
  TDNNF_CUDA_OK(cudaSetDevice(c->ctx->device));
  TDNNF_CUDA_OK(cudaEventRecord(c->done, c->side));
  TDNNF_CUDA_OK(cudaStreamWaitEvent(c->ctx->stream, c->done, 0));
  return TDNNF_OK;
}

extern "C" int tdnnf_dp_nccl_version(int* version) {
  TDNNF_REQUIRE(version, "null argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(TDNNF_ERR_UNSUPPORTED, g_nccl_error);
  TDNNF_NCCL_OK(api, api->GetVersion(version));
  return TDNNF_OK;
}
