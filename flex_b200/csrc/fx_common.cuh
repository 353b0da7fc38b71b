// fx_common.cuh -- shared internals of libflexb200.so (error handling, launch counting,
// the device arena, handle layouts).  Not part of the public ABI (include/flexb200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flexb200.h"
#include "fx_flex.cuh"

namespace fx {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// Reference: CUDA_CHECK throws (common.h:53-60).  A C ABI cannot throw, so the failing
// call is recorded and FX_ERR_CUDA propagates to the caller.
#define FX_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      fx::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return FX_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define FX_LAUNCH_CHECK()                                                       \
  do {                                                                          \
    fx::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      fx::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,           \
                    cudaGetErrorString(e__));                                   \
      return FX_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

#define FX_REQUIRE(cond, status, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      fx::set_error(__VA_ARGS__);          \
      return (status);                     \
    }                                      \
  } while (0)

// One cudaMalloc per handle, carved by a bump pointer: nothing is allocated inside the
// timed build (the reference allocates ~20 times inside tPre, aspt/sspmm_128.cu:1268-1326).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t bytes) {
    if (base) cudaFree(base);
    base = nullptr; cap = used = 0;
    if (bytes == 0) bytes = 256;
    cudaError_t e = cudaMalloc(&base, bytes);
    if (e != cudaSuccess) { set_error("arena cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e)); return FX_ERR_NOMEM; }
    cap = bytes;
    return FX_OK;
  }
  template <class T> T* take(size_t count) {
    size_t off = (used + 255) & ~size_t(255);
    size_t bytes = count * sizeof(T);
    if (off + bytes > cap) return nullptr;
    used = off + bytes;
    return reinterpret_cast<T*>(base + off);
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
  void release() { if (base) cudaFree(base); base = nullptr; cap = used = 0; }
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: the "already raised" cache is kept per device
// ordinal (a process-wide flag made the first launch on a second GPU of the same process fail with "invalid argument").
struct SmemAttr {
  size_t set[64] = {};
  template <class K>
  int ensure(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return FX_OK;
    const int dev = current_device();
    if (smem > set[dev]) {
      FX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set[dev] = smem;
    }
    return FX_OK;
  }
};

// Threads per CTA of the persistent builder kernels (k_heavy, k_tcw_select, k_tcw_split); FLEX_BUILD_THREADS overrides.
// One 1024-thread CTA per SM instead of two of 512: the same threads per SM, but a panel -- the unit of work, one CTA each --
// is walked by twice as many, and the build's duration is set by its heaviest panels (hub-first orderings put a million nz in
// one panel).  Measured, tPre: Reddit-shape 1.95 -> 1.90 ms, + DEG 24.5 -> 15.3, yelp-shape + DEG 8.1 -> 5.3, Amazon-shape
// 22.9 -> 20.3; and half the counter memory (one n-sized array per CTA).
inline int build_threads() {
  static const int t = getenv("FLEX_BUILD_THREADS") ? atoi(getenv("FLEX_BUILD_THREADS")) : 1024;
  return (t >= 128 && t <= 1024 && t % 32 == 0) ? t : 1024;
}

inline int sm_count_of_current_device() {
  static int sm[64] = {};
  const int dev = current_device();
  if (!sm[dev]) cudaDeviceGetAttribute(&sm[dev], cudaDevAttrMultiProcessorCount, dev);
  return sm[dev] > 0 ? sm[dev] : 148;
}

}  // namespace fx

// ---------------------------------------------------------------------------------------
// Handle layouts
// ---------------------------------------------------------------------------------------
struct fx_matrix {
  int64_t n = 0, nnz = 0;
  int k = 0;
  fx_matrix_info info{};
  // host CSR (DataLoader::rowPtr/col/vals); empty when created from device arrays
  std::vector<uint32_t> rowptr, col;
  std::vector<float> val;
  std::vector<int32_t> vo_mp;  // vo_mp[new] = old (DataLoader.cuh:38)
  // device CSR (rowPtr_dev/col_dev/vals_dev DataLoader.cuh:53-55)
  uint32_t *rowptr_dev = nullptr, *col_dev = nullptr;
  float* val_dev = nullptr;
  int32_t* vo_mp_dev = nullptr;
  bool census_done = false;
};

struct fx_aspt_dev {  // device-side ASpT format (aspt/sspmm_128.cu:76-90 globals)
  int n = 0, nr = 0, npanel = 0, ne = 0, BW = 128, row0 = 0;
  int *csr_v = nullptr;      // nr+1 padded row pointer (local nz offsets)
  int *mcsr_chk = nullptr, *mcsr_cnt = nullptr, *tcount = nullptr;
  int *mcsr_e = nullptr, *mcsr_list = nullptr, *baddr = nullptr, *saddr = nullptr;
  int *perm = nullptr, *csr_e = nullptr;
  float* csr_ev = nullptr;
  uint16_t* key2 = nullptr;
  int2* heavy = nullptr;  // per panel region at csr_v[p*BH]/16 + p
  int* nheavy = nullptr;
  unsigned* cnt_scratch = nullptr;  // G x ncols saturating per-column counters (kept zeroed)
  int G = 0;
  int *spec_cnt = nullptr, *spec_off = nullptr, *special = nullptr, *special2 = nullptr;
  // execution order of the 512-chunks (L2-residency scheduling): chunk ids sorted by the first column they read
  int *spec_order = nullptr, *spec_iota = nullptr;
  unsigned *spec_keys = nullptr, *spec_keys_out = nullptr;
  void* spec_sort_tmp = nullptr;
  size_t spec_sort_tmp_bytes = 0;
  int *plist_plain = nullptr, *plist_tiled = nullptr;  // panels without / with dense tiles (ascending)
  int n_plain = 0, n_tiled = 0;
  // work lists of the SpMM launches: one entry per CTA = (panel, part | parts << 8); a panel with several times
  // the average work is cut into up to 16 parts at row boundaries (skewed orderings put all hubs in a few panels)
  int2 *wl_all = nullptr, *wl_plain = nullptr, *wl_tiled = nullptr;
  int n_wl_all = 0, n_wl_plain = 0, n_wl_tiled = 0, wl_cap = 0;
  unsigned long long* stats = nullptr;  // [0]=S1 [1]=S2 [2]=special chunks [3]=num_dense [4]=any flag
  float* partial = nullptr;             // special_p_cap x k partial sums of 512-chunks
  size_t partial_cap_floats = 0;
  int mcsr_e_cap = 0, list_cap_tiles = 0, special_cap = 0;
  // host scalars filled by the build
  int num_dense = 0, any_flag = 0, regime = 0, special_p = 0, max_tp = 0;
  long long S1 = 0, S2 = 0;
  double avg = 0, vari = 0;
  bool aliased = false;  // csr_e/csr_ev alias the matrix arrays (no dense tile anywhere)
  // the arrays the SpMM kernels read (aliases of csr_v / the CSR when no panel has a dense tile,
  // as the reference does: _mcsr_e = _csr_v, aspt/sspmm_128.cu:1227-1229)
  const int *mcsr_e_use = nullptr, *csr_e_use = nullptr;
  const float* csr_ev_use = nullptr;
};

struct fx_tcw_dev {  // tensor-window format (fx_tcw_build.cu): per-panel column lists + window nz + remainder CSR
  int T = 4, W = 256, min_gain = 1024, chunk_cost = 224;
  long long min_total = 1000000, net_gain = 0;
  int *tc_cols = nullptr, *tc_ncol = nullptr, *tc_slot = nullptr, *tc_panels = nullptr;
  int *csr_v = nullptr, *win_len = nullptr, *win_rowptr = nullptr, *chunk_len = nullptr, *win_cptr = nullptr;
  uint32_t *rest_rowptr = nullptr, *rest_col = nullptr;
  uint16_t* win_code = nullptr;
  float *win_val = nullptr, *rest_val = nullptr;
  float* tc_out = nullptr;  // [ntc][128][k] products of the window parts
  unsigned long long* stats = nullptr;
  long long win_nnz = 0, ncols_listed = 0;
  int ntc = 0, dropped = 0;
  std::vector<int32_t> h_cols, h_ncol, h_win_cptr;
  std::vector<uint16_t> h_win_code;
  std::vector<uint32_t> h_rest_rowptr, h_rest_col;
  std::vector<float> h_win_val, h_rest_val;
};

struct fx_tiles {
  const fx_matrix* mat = nullptr;
  // CSR the ASpT builder reads; null = the matrix itself (FX_FMT_TCW points it at the remainder)
  const uint32_t *src_rowptr = nullptr, *src_col = nullptr;
  const float* src_val = nullptr;
  int src_row0 = 0;
  fx_tcw_dev tcw;
  fx_build_opts opts{};
  int format = FX_FMT_ASPT;
  int k = 0;
  int row_begin = 0, row_end = 0;  // shard
  int64_t nnz_local = 0;
  fx::Arena arena;
  fx_aspt_dev aspt;
  fx_flex_dev flex;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long* stats_host = nullptr;  // pinned mirror of aspt.stats
  // fx_axw: the intermediate factor (X*W, or A*X)
  float* axw_scratch = nullptr;
  size_t axw_cap = 0;
  // host staging for fx_spmm_host
  float *B_stage_dev = nullptr, *C_stage_dev = nullptr;
  float *B_pinned = nullptr, *C_pinned = nullptr;
  size_t stage_elems = 0;
  cudaStream_t own_stream = nullptr;
  // fx_spmm_host pipeline: copy-in, compute and copy-out streams + per-chunk events (created on first use)
  cudaStream_t pipe_s[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pipe_in[8] = {}, pipe_k0[8] = {}, pipe_k1[8] = {}, pipe_e0 = nullptr, pipe_e1 = nullptr;
  cudaEvent_t pipe_g[64] = {};  // [chunk][row group]: "this range of panels is multiplied"
  // host copies for export
  std::vector<int32_t> h_chk, h_cnt, h_e, h_list, h_baddr, h_saddr, h_perm, h_csr_e, h_special, h_special2;
  std::vector<float> h_csr_ev;
};

namespace fx {
int ensure_device(const fx_matrix* m);
// fx_aspt_build.cu
size_t aspt_arena_bytes(int64_t n_rows, int64_t ncols, int64_t ne, int BW, int k, int G);
int aspt_carve(fx_tiles* t, int64_t ncols, size_t extra_bytes = 0);
int aspt_build(fx_tiles* t, cudaStream_t s);
// fx_tcw_build.cu
size_t tcw_arena_bytes(const fx_tiles* t);
int tcw_carve(fx_tiles* t);
int tcw_build(fx_tiles* t, cudaStream_t s);
// fx_flex_build.cu
int flex_carve(fx_tiles* t);
int flex_build(fx_tiles* t, cudaStream_t s);
int flex_spmm(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s);
void flex_release(fx_tiles* t);
// fx_spmm.cu
int spmm_csr(const uint32_t* rowptr, const uint32_t* col, const float* val, int64_t nrows, const float* B,
             float* C, int k, cudaStream_t s);
// width > 0: compute only `width` feature columns starting at the B / C pointers (row stride stays k)
// width: feature columns computed from the B/C pointers on (0 = k); [p_lo, p_hi): the 128-row panels multiplied (-1 = all)
int spmm_aspt(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s, int width = 0, int p_lo = 0, int p_hi = -1);
int spmm_aspt_times(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s, float ms[4]);
int gemm_xw(const float* X, const float* W, float* out, int64_t rows, int k, int c, cudaStream_t s);
int permute_rows(const int32_t* map, int64_t n, int k, const float* src, float* dst, bool scatter,
                 cudaStream_t s);
}  // namespace fx
