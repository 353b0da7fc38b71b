// fx_shard.cu -- the row-panel sharded SpMM from HOST buffers on one rank of a G-rank job (the reference is
// single-GPU, so there is no counterpart: this is the "upload B once per job, all-gather it over NVLink" step the survey
// names as the one real exchange of the sharded path).
//
// Each rank holds the tiles of its row-panel shard (fx_build with row_begin/row_end).  fx_spmm_host on every rank would push
// all of B through the host's PCIe root G times per SpMM (8 GPUs, Reddit-shape k=128: 5.96 ms against 3.81 ms on one).  Here
// rank r uploads only rows [r*ceil(n/G), (r+1)*ceil(n/G)) of B, ncclAllGather assembles B on every GPU over NVLink, the rank
// multiplies its shard and copies its own rows of C back -- pipelined over column chunks exactly like fx_spmm_host: chunk c+1
// is uploaded and gathered while chunk c is multiplied and chunk c-1 is copied out (four streams, events between them).
// On the device a chunk is a contiguous [rows x cw] block (row stride cw), which is what NCCL needs and what the kernels
// take as (k = row stride, width).
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the process -- torch's under Python -- or the
// system's), so libflexb200.so has no link-time NCCL dependency and the single-GPU paths never touch it.
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "fx_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7;  // ncclFloat32 (nccl.h)

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOLOAD | RTLD_NOW);  // the copy the host program already uses, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api;
}

#define FX_NCCL(call)                                                                                        \
  do {                                                                                                       \
    ncclResult_t r__ = (call);                                                                               \
    if (r__ != 0) {                                                                                          \
      fx::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r__));              \
      return FX_ERR_CUDA;                                                                                    \
    }                                                                                                        \
  } while (0)

}  // namespace

struct fx_comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // per-communicator device staging, grown on demand: this rank's slice of B, the gathered B, this rank's rows of C
  float *slice_dev = nullptr, *B_dev = nullptr, *C_dev = nullptr;
  size_t slice_cap = 0, B_cap = 0, C_cap = 0;
  cudaStream_t s_in = nullptr, s_comm = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[8] = {}, ev_ag[8] = {}, ev_k[8] = {}, k0[8] = {}, e0 = nullptr, e1 = nullptr;
};

extern "C" int fx_comm_unique_id(char id[128]) {
  FX_REQUIRE(id, FX_ERR_ARG, "fx_comm_unique_id: null");
  FX_REQUIRE(nccl().ok, FX_ERR_UNSUPPORTED, "NCCL not found (libnccl.so.2)");
  ncclUniqueId u;
  FX_NCCL(nccl().GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return FX_OK;
}

extern "C" int fx_comm_init(int nranks, int rank, const char id[128], fx_comm** out) {
  FX_REQUIRE(out && id && nranks >= 1 && rank >= 0 && rank < nranks, FX_ERR_ARG, "fx_comm_init: bad argument");
  FX_REQUIRE(nccl().ok, FX_ERR_UNSUPPORTED, "NCCL not found (libnccl.so.2)");
  auto c = new fx_comm();
  c->nranks = nranks; c->rank = rank;
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, u, rank);
  if (r != 0) { fx::set_error("ncclCommInitRank -> %s", nccl().GetErrorString(r)); delete c; return FX_ERR_CUDA; }
  bool ok = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->e0) == cudaSuccess && cudaEventCreate(&c->e1) == cudaSuccess;
  for (int i = 0; i < 8 && ok; ++i)
    ok = cudaEventCreate(&c->ev_in[i]) == cudaSuccess && cudaEventCreate(&c->ev_ag[i]) == cudaSuccess &&
         cudaEventCreate(&c->ev_k[i]) == cudaSuccess && cudaEventCreate(&c->k0[i]) == cudaSuccess;
  if (!ok) { fx::set_error("fx_comm_init: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())); fx_comm_free(c); return FX_ERR_CUDA; }
  *out = c;
  return FX_OK;
}

extern "C" void fx_comm_free(fx_comm* c) {
  if (!c) return;
  cudaFree(c->slice_dev); cudaFree(c->B_dev); cudaFree(c->C_dev);
  for (cudaStream_t s : {c->s_in, c->s_comm, c->s_k, c->s_out}) if (s) cudaStreamDestroy(s);
  for (int i = 0; i < 8; ++i)
    for (cudaEvent_t e : {c->ev_in[i], c->ev_ag[i], c->ev_k[i], c->k0[i]}) if (e) cudaEventDestroy(e);
  if (c->e0) cudaEventDestroy(c->e0);
  if (c->e1) cudaEventDestroy(c->e1);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

extern "C" int fx_comm_slice(const fx_comm* c, int64_t n, int64_t* row_lo, int64_t* row_hi) {
  FX_REQUIRE(c && n >= 0, FX_ERR_ARG, "fx_comm_slice: bad argument");
  const int64_t per = (n + c->nranks - 1) / c->nranks;
  if (row_lo) *row_lo = std::min<int64_t>(n, c->rank * per);
  if (row_hi) *row_hi = std::min<int64_t>(n, (c->rank + 1) * per);
  return FX_OK;
}

static int grow(float** p, size_t* cap, size_t elems) {
  if (*cap >= elems) return FX_OK;
  cudaFree(*p);
  *p = nullptr; *cap = 0;
  FX_CUDA(cudaMalloc(p, sizeof(float) * std::max<size_t>(elems, 1)));
  FX_CUDA(cudaMemset(*p, 0, sizeof(float) * std::max<size_t>(elems, 1)));  // the padding rows of the last slice stay zero
  *cap = elems;
  return FX_OK;
}

extern "C" int fx_spmm_sharded_host(const fx_tiles* t, fx_comm* c, const float* B_rows_host, float* C_local_host, int k,
                                    float* total_ms, float* tElap_ms) {
  FX_REQUIRE(t && c && C_local_host && k > 0, FX_ERR_ARG, "fx_spmm_sharded_host: bad argument");
  FX_REQUIRE(t->format == FX_FMT_ASPT || t->format == FX_FMT_TCW, FX_ERR_UNSUPPORTED, "fx_spmm_sharded_host: ASpT / tensor-window tiles");
  FX_REQUIRE(k % 4 == 0 && k <= t->k, FX_ERR_UNSUPPORTED, "fx_spmm_sharded_host: k %% 4 == 0 and k <= the build's k");
  const int64_t n = t->mat->n, per = (n + c->nranks - 1) / c->nranks;
  const int64_t lo = std::min<int64_t>(n, c->rank * per), hi = std::min<int64_t>(n, (c->rank + 1) * per);
  const int64_t nloc = t->row_end - t->row_begin;
  FX_REQUIRE(hi == lo || B_rows_host, FX_ERR_ARG, "fx_spmm_sharded_host: this rank's rows of B are missing");
  // equal column chunks, each a multiple of 32 features (as fx_spmm_host)
  static const int want_chunks = getenv("FLEX_HOST_CHUNKS") ? atoi(getenv("FLEX_HOST_CHUNKS")) : 2;
  int nchunk = 1;
  if (k % 32 == 0 && k >= 64 && want_chunks > 1) {
    nchunk = std::min(std::min(want_chunks, k / 32), 8);
    while ((k / 32) % nchunk) --nchunk;
  }
  const int cw = k / nchunk;
  const size_t slice_elems = (size_t)per * cw, full_elems = slice_elems * c->nranks;
  int rc;
  if ((rc = grow(&c->slice_dev, &c->slice_cap, slice_elems * nchunk))) return rc;
  if ((rc = grow(&c->B_dev, &c->B_cap, full_elems * nchunk))) return rc;
  if ((rc = grow(&c->C_dev, &c->C_cap, (size_t)std::max<int64_t>(nloc, 1) * k))) return rc;
  const size_t hpitch = sizeof(float) * (size_t)k, dpitch = sizeof(float) * (size_t)cw;
  FX_CUDA(cudaEventRecord(c->e0, c->s_in));
  for (int i = 0; i < nchunk; ++i) {
    float* slice = c->slice_dev + (size_t)i * slice_elems;
    float* Bfull = c->B_dev + (size_t)i * full_elems;           // [nranks*per x cw], row stride cw
    float* Cchunk = c->C_dev + (size_t)i * (size_t)nloc * cw;   // [nloc x cw], row stride cw
    if (hi > lo)
      FX_CUDA(cudaMemcpy2DAsync(slice, dpitch, B_rows_host + (size_t)i * cw, hpitch, dpitch, (size_t)(hi - lo), cudaMemcpyHostToDevice, c->s_in));
    FX_CUDA(cudaEventRecord(c->ev_in[i], c->s_in));
    FX_CUDA(cudaStreamWaitEvent(c->s_comm, c->ev_in[i], 0));
    FX_NCCL(nccl().AllGather(slice, Bfull, slice_elems, kNcclFloat, c->comm, c->s_comm));
    FX_CUDA(cudaEventRecord(c->ev_ag[i], c->s_comm));
    FX_CUDA(cudaStreamWaitEvent(c->s_k, c->ev_ag[i], 0));
    FX_CUDA(cudaEventRecord(c->k0[i], c->s_k));
    if ((rc = fx::spmm_aspt(t, Bfull, Cchunk, cw, c->s_k, cw)) != FX_OK) return rc;
    FX_CUDA(cudaEventRecord(c->ev_k[i], c->s_k));
    FX_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0));
    if (nloc > 0)
      FX_CUDA(cudaMemcpy2DAsync(C_local_host + (size_t)i * cw, hpitch, Cchunk, dpitch, dpitch, (size_t)nloc, cudaMemcpyDeviceToHost, c->s_out));
  }
  FX_CUDA(cudaEventRecord(c->e1, c->s_out));
  FX_CUDA(cudaEventSynchronize(c->e1));
  if (total_ms) FX_CUDA(cudaEventElapsedTime(total_ms, c->e0, c->e1));
  if (tElap_ms) {
    float sum = 0.f, ms = 0.f;
    for (int i = 0; i < nchunk; ++i) { FX_CUDA(cudaEventElapsedTime(&ms, c->k0[i], c->ev_k[i])); sum += ms; }
    *tElap_ms = sum;
  }
  static const bool prof = getenv("FLEX_SHARD_PROF") != nullptr;  // time line of the call, ms after its start, per column chunk
  if (prof) {
    float tot = 0.f;
    cudaEventElapsedTime(&tot, c->e0, c->e1);
    fprintf(stderr, "[fx_spmm_sharded_host rank %d] total %.3f ms;", c->rank, tot);
    for (int i = 0; i < nchunk; ++i) {
      float a = 0.f, b = 0.f, k0 = 0.f, k1 = 0.f;
      cudaEventElapsedTime(&a, c->e0, c->ev_in[i]); cudaEventElapsedTime(&b, c->e0, c->ev_ag[i]);
      cudaEventElapsedTime(&k0, c->e0, c->k0[i]); cudaEventElapsedTime(&k1, c->e0, c->ev_k[i]);
      fprintf(stderr, " chunk %d: slice in %.3f, all-gather done %.3f, SpMM %.3f-%.3f;", i, a, b, k0, k1);
    }
    fprintf(stderr, "\n");
  }
  return FX_OK;
}
