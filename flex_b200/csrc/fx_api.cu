// fx_api.cu -- extern "C" entry points for build / SpMM / validation (include/flexb200.h).
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "fx_common.cuh"

namespace {

int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fx::set_error("no CUDA device: %s (libflexb200 has no CPU path)", cudaGetErrorString(e));
    return FX_ERR_CUDA;
  }
  return FX_OK;
}

int time_region(fx_tiles* t, cudaStream_t s, float* ms, int (*fn)(fx_tiles*, cudaStream_t)) {
  FX_CUDA(cudaEventRecord(t->ev0, s));
  int rc = fn(t, s);
  if (rc != FX_OK) return rc;
  FX_CUDA(cudaEventRecord(t->ev1, s));
  FX_CUDA(cudaEventSynchronize(t->ev1));
  if (ms) FX_CUDA(cudaEventElapsedTime(ms, t->ev0, t->ev1));
  return FX_OK;
}

int build_dispatch(fx_tiles* t, cudaStream_t s) {
  switch (t->format) {
    case FX_FMT_CSR: return FX_OK;
    case FX_FMT_ASPT: return fx::aspt_build(t, s);
    case FX_FMT_TCW: {
      int rc = fx::tcw_build(t, s);
      return rc != FX_OK ? rc : fx::aspt_build(t, s);
    }
    case FX_FMT_TILE: case FX_FMT_SEG: case FX_FMT_PILLAR: return fx::flex_build(t, s);
    default: fx::set_error("format %d not built by this entry point", t->format); return FX_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int fx_build(const fx_matrix* m, const fx_build_opts* opts, fx_tiles** out, float* tPre_ms) {
  FX_REQUIRE(m && out, FX_ERR_ARG, "fx_build: null argument");
  int rc = check_device();
  if (rc != FX_OK) return rc;
  rc = fx::ensure_device(m);
  if (rc != FX_OK) return rc;
  auto t = new fx_tiles();
  t->mat = m;
  if (opts) t->opts = *opts;
  t->format = opts ? opts->format : FX_FMT_ASPT;
  t->k = (int)m->k;
  t->row_begin = t->opts.row_begin;
  t->row_end = t->opts.row_end;
  if (t->row_begin == 0 && t->row_end == 0) t->row_end = (int)m->n;
  if (t->row_begin < 0 || t->row_end > m->n || t->row_begin > t->row_end) {
    delete t;
    fx::set_error("fx_build: bad row range");
    return FX_ERR_ARG;
  }
  t->nnz_local = (int64_t)m->rowptr[t->row_end] - (int64_t)m->rowptr[t->row_begin];
  auto fail = [&](int code) { fx_tiles_free(t); return code; };
  if (cudaEventCreate(&t->ev0) != cudaSuccess || cudaEventCreate(&t->ev1) != cudaSuccess ||
      cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(&t->stats_host, sizeof(unsigned long long) * 16) != cudaSuccess) {
    fx::set_error("fx_build: event/stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(FX_ERR_CUDA);
  }
  if (t->format == FX_FMT_ASPT || t->format == FX_FMT_TCW) {
    fx_aspt_dev& a = t->aspt;
    const int nloc = t->row_end - t->row_begin;
    a.n = nloc;
    a.nr = (nloc + 127) / 128 * 128;
    a.npanel = a.nr / 128;
    a.ne = (int)t->nnz_local;
    a.row0 = t->row_begin;
    a.BW = t->opts.bw ? t->opts.bw : (m->k >= 64 ? 128 : 256);  // sspmm_128.cu:34 vs sspmm_32.cu:33
    if (a.BW != 128 && a.BW != 256) { fx::set_error("bw must be 128 or 256"); return fail(FX_ERR_ARG); }
    size_t extra = 0;
    if (t->format == FX_FMT_TCW) {
      fx_tcw_dev& w = t->tcw;
      if (t->opts.tc_threshold) w.T = t->opts.tc_threshold;
      if (t->opts.tc_width) w.W = t->opts.tc_width;
      if (t->opts.tc_min_gain) w.min_gain = std::max(0, t->opts.tc_min_gain);  // negative = no minimum
      if (t->opts.tc_chunk_cost) w.chunk_cost = std::max(0, t->opts.tc_chunk_cost);
      if (t->opts.tc_min_total) w.min_total = std::max(0, t->opts.tc_min_total);
      // The whole-matrix gate is a statement about the MATRIX (do its windows pay for a second kernel?): a row-panel shard
      // answers for its share of the nz, so that 1/8 shards of a matrix that keeps its windows keep theirs (with an absolute
      // threshold every shard of Reddit-shape at 8 GPUs fell back to the slower ASpT path).
      if (m->nnz > 0 && t->nnz_local < m->nnz) w.min_total = w.min_total * t->nnz_local / m->nnz;
      if (w.T < 2 || w.W < 32 || w.W > 4096 || w.W % 32) { fx::set_error("tc_threshold must be >= 2 and tc_width a multiple of 32 in [32,4096]"); return fail(FX_ERR_ARG); }
      extra = fx::tcw_arena_bytes(t);
    }
    rc = fx::aspt_carve(t, m->n, extra);
    if (rc != FX_OK) return fail(rc);
    if (t->format == FX_FMT_TCW && (rc = fx::tcw_carve(t)) != FX_OK) return fail(rc);
  } else if (t->format == FX_FMT_TILE || t->format == FX_FMT_SEG || t->format == FX_FMT_PILLAR) {
    rc = fx::flex_carve(t);
    if (rc != FX_OK) return fail(rc);
  } else if (t->format != FX_FMT_CSR) {
    fx::set_error("fx_build: unknown format %d", t->format);
    return fail(FX_ERR_ARG);
  }
  rc = time_region(t, t->own_stream, tPre_ms, build_dispatch);
  if (rc != FX_OK) return fail(rc);
  *out = t;
  return FX_OK;
}

extern "C" int fx_rebuild(fx_tiles* t, float* tPre_ms) {
  FX_REQUIRE(t, FX_ERR_ARG, "fx_rebuild: null");
  return time_region(t, t->own_stream, tPre_ms, build_dispatch);
}

extern "C" void fx_tiles_free(fx_tiles* t) {
  if (!t) return;
  fx::flex_release(t);
  t->arena.release();
  if (t->ev0) cudaEventDestroy(t->ev0);
  if (t->ev1) cudaEventDestroy(t->ev1);
  if (t->own_stream) cudaStreamDestroy(t->own_stream);
  for (int i = 0; i < 64; ++i) if (t->pipe_g[i]) cudaEventDestroy(t->pipe_g[i]);
  for (int i = 0; i < 3; ++i) if (t->pipe_s[i]) cudaStreamDestroy(t->pipe_s[i]);
  for (int i = 0; i < 8; ++i) {
    if (t->pipe_in[i]) cudaEventDestroy(t->pipe_in[i]);
    if (t->pipe_k0[i]) cudaEventDestroy(t->pipe_k0[i]);
    if (t->pipe_k1[i]) cudaEventDestroy(t->pipe_k1[i]);
  }
  if (t->pipe_e0) cudaEventDestroy(t->pipe_e0);
  if (t->pipe_e1) cudaEventDestroy(t->pipe_e1);
  if (t->stats_host) cudaFreeHost(t->stats_host);
  cudaFree(t->B_stage_dev); cudaFree(t->C_stage_dev); cudaFree(t->axw_scratch);
  if (t->B_pinned) cudaFreeHost(t->B_pinned);
  if (t->C_pinned) cudaFreeHost(t->C_pinned);
  delete t;
}

template <class T>
static int d2h(std::vector<T>& dst, const void* src, size_t count) {
  dst.resize(count);
  if (count) FX_CUDA(cudaMemcpy(dst.data(), src, sizeof(T) * count, cudaMemcpyDeviceToHost));
  return FX_OK;
}

extern "C" int fx_tiles_export_aspt(fx_tiles* t, fx_aspt_arrays* o) {
  FX_REQUIRE(t && o && t->format == FX_FMT_ASPT, FX_ERR_ARG, "fx_tiles_export_aspt: not an ASpT handle");
  const fx_aspt_dev& a = t->aspt;
  FX_CUDA(cudaDeviceSynchronize());
  int rc;
  const size_t nd = a.num_dense;
  if ((rc = d2h(t->h_chk, a.mcsr_chk, a.npanel))) return rc;
  if ((rc = d2h(t->h_cnt, a.mcsr_cnt, a.npanel + 1))) return rc;
  if ((rc = d2h(t->h_e, a.mcsr_e_use, (size_t)128 * (nd + a.npanel) + 1))) return rc;
  if ((rc = d2h(t->h_list, a.mcsr_list, nd * a.BW))) return rc;
  if ((rc = d2h(t->h_baddr, a.baddr, nd))) return rc;
  if ((rc = d2h(t->h_saddr, a.saddr, nd))) return rc;
  if ((rc = d2h(t->h_csr_e, a.csr_e_use, a.ne))) return rc;
  if ((rc = d2h(t->h_csr_ev, a.csr_ev_use, a.ne))) return rc;
  if (a.aliased) {
    t->h_perm.resize(a.ne);
    for (int i = 0; i < a.ne; ++i) t->h_perm[i] = i;
  } else if ((rc = d2h(t->h_perm, a.perm, a.ne))) return rc;
  // the reference only materialises the special lists when vari >= 200 (:1319-1328)
  const int sp = a.vari >= 200 ? a.special_p : 0;
  if ((rc = d2h(t->h_special, a.special, sp))) return rc;
  if ((rc = d2h(t->h_special2, a.special2, sp))) return rc;
  o->n = a.n; o->nr = a.nr; o->npanel = a.npanel; o->ne = a.ne; o->BH = 128; o->BW = a.BW;
  o->num_dense = a.num_dense; o->any_flag = a.any_flag; o->regime = a.regime; o->special_p = sp;
  o->S1 = a.S1; o->S2 = a.S2; o->avg = a.avg; o->vari = a.vari;
  o->mcsr_chk = t->h_chk.data(); o->mcsr_cnt = t->h_cnt.data(); o->mcsr_e = t->h_e.data();
  o->mcsr_list = t->h_list.data(); o->baddr = t->h_baddr.data(); o->saddr = t->h_saddr.data();
  o->perm = t->h_perm.data(); o->csr_e = t->h_csr_e.data(); o->csr_ev = t->h_csr_ev.data();
  o->special = t->h_special.data(); o->special2 = t->h_special2.data();
  return FX_OK;
}

static int spmm_dispatch(const fx_tiles* t, const float* B, float* C, int k, cudaStream_t s) {
  const fx_matrix* m = t->mat;
  const bool tcw = t->format == FX_FMT_TCW;
  const bool aspt_like = t->format == FX_FMT_ASPT || tcw;
  // wider than the build's k: the per-handle scratch (512-chunk partial sums, window products) is sized for t->k
  if (t->format == FX_FMT_CSR || (aspt_like && (k % 4 != 0 || k > t->k))) {
    // raw CSR over the shard's rows (run_ge_spmm path flex.cu:4285; "ssparse" regime :629)
    return fx::spmm_csr(m->rowptr_dev + t->row_begin, m->col_dev, m->val_dev, t->row_end - t->row_begin, B, C, k, s);
  }
  if (t->format == FX_FMT_ASPT || tcw) return fx::spmm_aspt(t, B, C, k, s);
  if (t->format == FX_FMT_TILE || t->format == FX_FMT_SEG || t->format == FX_FMT_PILLAR) return fx::flex_spmm(t, B, C, k, s);
  fx::set_error("fx_spmm: format %d has its own entry point", t->format);
  return FX_ERR_UNSUPPORTED;
}

extern "C" int fx_spmm(const fx_tiles* t, const float* B_dev, float* C_dev, int k, void* stream, float* tElap_ms) {
  FX_REQUIRE(t && B_dev && C_dev && k > 0, FX_ERR_ARG, "fx_spmm: bad argument");
  FX_REQUIRE((int64_t)t->mat->n * k / 4 < (1ll << 32), FX_ERR_UNSUPPORTED, "n*k too large for 32-bit float4 offsets");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!tElap_ms) return spmm_dispatch(t, B_dev, C_dev, k, s);
  fx_tiles* tt = const_cast<fx_tiles*>(t);
  FX_CUDA(cudaEventRecord(tt->ev0, s));
  int rc = spmm_dispatch(t, B_dev, C_dev, k, s);
  if (rc != FX_OK) return rc;
  FX_CUDA(cudaEventRecord(tt->ev1, s));
  FX_CUDA(cudaEventSynchronize(tt->ev1));
  FX_CUDA(cudaEventElapsedTime(tElap_ms, tt->ev0, tt->ev1));
  return FX_OK;
}

extern "C" int fx_set_device(int ordinal) {
  if (int rc = check_device()) return rc;
  FX_CUDA(cudaSetDevice(ordinal));
  return FX_OK;
}

// Row-panel shards of the multi-GPU path: contiguous ranges of 128-row panels with about nnz/G nonzeros each, cut by
// prefix sums over rowPtr (the rule flex_b200/shard.py:panel_shards states for the CPU tests).
extern "C" int fx_panel_shards(const fx_matrix* m, int nranks, int64_t* cuts) {
  FX_REQUIRE(m && cuts && nranks >= 1, FX_ERR_ARG, "fx_panel_shards: bad argument");
  const int64_t n = m->n, npanel = (n + 127) / 128, total = m->rowptr.empty() ? 0 : (int64_t)m->rowptr[n];
  auto pnnz = [&](int64_t p) { return (int64_t)m->rowptr[std::min<int64_t>(p * 128, n)]; };
  int64_t prev = 0;
  cuts[0] = 0;
  for (int r = 1; r < nranks; ++r) {
    // first panel boundary whose nz prefix reaches total*r/nranks (numpy searchsorted, side="left", on the prefix)
    const double want = (double)total * r / nranks;
    int64_t lo = 0, hi = npanel + 1;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if ((double)pnnz(mid) < want) lo = mid + 1; else hi = mid; }
    prev = std::max(prev, std::min(lo, npanel));
    cuts[r] = std::min<int64_t>(prev * 128, n);
  }
  cuts[nranks] = n;
  return FX_OK;
}

extern "C" int fx_spmm_kernel_times(const fx_tiles* t, const float* B_dev, float* C_dev, int k, void* stream, float ms[4]) {
  FX_REQUIRE(t && B_dev && C_dev && k > 0 && ms, FX_ERR_ARG, "fx_spmm_kernel_times: bad argument");
  FX_REQUIRE((t->format == FX_FMT_ASPT || t->format == FX_FMT_TCW) && k % 4 == 0 && k <= t->k, FX_ERR_UNSUPPORTED,
             "fx_spmm_kernel_times: ASpT / tensor-window handles, k %% 4 == 0 and k <= the build's k");
  return fx::spmm_aspt_times(t, B_dev, C_dev, k, static_cast<cudaStream_t>(stream), ms);
}

// AXW (cusp.cu:3-208, main.cu:22-79): order 0 = run1, C = A*(X*W); order 1 = run2, C = (A*X)*W
extern "C" int fx_axw(const fx_tiles* tc, const float* X_dev, const float* W_dev, float* C_dev, int k, int c, int order,
                      void* stream, float* gemm_ms, float* spmm_ms) {
  FX_REQUIRE(tc && X_dev && W_dev && C_dev && k > 0 && c > 0 && (order == 0 || order == 1), FX_ERR_ARG, "fx_axw: bad argument");
  FX_REQUIRE(k % 4 == 0 && c % 4 == 0, FX_ERR_UNSUPPORTED, "fx_axw: k and c must be multiples of 4");
  fx_tiles* t = const_cast<fx_tiles*>(tc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = t->mat->n, nloc = t->row_end - t->row_begin;
  const size_t need = order == 0 ? (size_t)n * c : (size_t)std::max<int64_t>(nloc, 1) * k;
  if (t->axw_cap < need) {
    cudaFree(t->axw_scratch);
    t->axw_scratch = nullptr; t->axw_cap = 0;
    FX_CUDA(cudaMalloc(&t->axw_scratch, sizeof(float) * need));
    t->axw_cap = need;
  }
  cudaEvent_t e[3] = {};
  const bool timed = gemm_ms || spmm_ms;
  if (timed) for (auto& x : e) FX_CUDA(cudaEventCreate(&x));
  int rc = FX_OK;
  if (timed) FX_CUDA(cudaEventRecord(e[0], s));
  if (order == 0) {
    rc = fx::gemm_xw(X_dev, W_dev, t->axw_scratch, n, k, c, s);            // B = X*W   (cusp.cu:30-32)
    if (rc == FX_OK && timed) FX_CUDA(cudaEventRecord(e[1], s));
    if (rc == FX_OK) rc = spmm_dispatch(t, t->axw_scratch, C_dev, c, s);     // C = A*B   (cusp.cu:62-83)
  } else {
    rc = spmm_dispatch(t, X_dev, t->axw_scratch, k, s);                      // B = A*X
    if (rc == FX_OK && timed) FX_CUDA(cudaEventRecord(e[1], s));
    if (rc == FX_OK) rc = fx::gemm_xw(t->axw_scratch, W_dev, C_dev, nloc, k, c, s);  // C = B*W
  }
  if (rc == FX_OK && timed) {
    FX_CUDA(cudaEventRecord(e[2], s));
    FX_CUDA(cudaEventSynchronize(e[2]));
    float a = 0, b = 0;
    FX_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
    FX_CUDA(cudaEventElapsedTime(&b, e[1], e[2]));
    if (gemm_ms) *gemm_ms = order == 0 ? a : b;
    if (spmm_ms) *spmm_ms = order == 0 ? b : a;
  }
  if (timed) for (auto& x : e) cudaEventDestroy(x);
  return rc;
}

extern "C" int fx_spmm_host(const fx_tiles* tc, const float* B_host, float* C_host, int k, float* total_ms,
                            float* tElap_ms) {
  FX_REQUIRE(tc && B_host && C_host && k > 0, FX_ERR_ARG, "fx_spmm_host: bad argument");
  fx_tiles* t = const_cast<fx_tiles*>(tc);
  const size_t nB = (size_t)t->mat->n * k, nC = (size_t)(t->row_end - t->row_begin) * k;
  if (t->stage_elems < std::max(nB, nC)) {
    cudaFree(t->B_stage_dev); cudaFree(t->C_stage_dev);
    t->B_stage_dev = t->C_stage_dev = nullptr;
    t->stage_elems = 0;
    FX_CUDA(cudaMalloc(&t->B_stage_dev, sizeof(float) * std::max<size_t>(nB, 1)));
    FX_CUDA(cudaMalloc(&t->C_stage_dev, sizeof(float) * std::max<size_t>(nC, 1)));
    t->stage_elems = std::max(nB, nC);
  }
  // Column-chunk pipeline (ASpT / tensor-window formats): the features are cut into (by default two) chunks of a multiple of 32; chunk
  // i+1 is copied in while chunk i is multiplied and chunk i-1 is copied out (PCIe is full duplex and
  // strided 2D copies of >= 128-byte rows run at the rate of one contiguous copy on this platform:
  // scripts/pcie_probe.py).  The kernels take the row stride k and the chunk width separately.
  static const int want_chunks = getenv("FLEX_HOST_CHUNKS") ? atoi(getenv("FLEX_HOST_CHUNKS")) : 2;  // measured: 4.87 ms (1), 3.92 (2), 4.53 (4) on Reddit-shape k=128
  int nchunk = 1;
  if ((t->format == FX_FMT_ASPT || t->format == FX_FMT_TCW) && k % 32 == 0 && k >= 64 && want_chunks > 1 &&
      k <= t->k) {
    nchunk = std::min(std::min(want_chunks, k / 32), 8);
    while ((k / 32) % nchunk) --nchunk;  // equal chunks, each a multiple of 32 features
  }
  if (nchunk > 1) {
    if (!t->pipe_s[0]) {
      for (int i = 0; i < 3; ++i) FX_CUDA(cudaStreamCreateWithFlags(&t->pipe_s[i], cudaStreamNonBlocking));
      for (int i = 0; i < 8; ++i) {
        FX_CUDA(cudaEventCreateWithFlags(&t->pipe_in[i], cudaEventDisableTiming));
        FX_CUDA(cudaEventCreate(&t->pipe_k0[i]));
        FX_CUDA(cudaEventCreate(&t->pipe_k1[i]));
      }
      FX_CUDA(cudaEventCreate(&t->pipe_e0));
      FX_CUDA(cudaEventCreate(&t->pipe_e1));
    }
    cudaStream_t sin = t->pipe_s[0], sk = t->pipe_s[1], sout = t->pipe_s[2];
    const int cw = k / nchunk;
    const size_t pitch = sizeof(float) * (size_t)k, wbytes = sizeof(float) * (size_t)cw;
    const size_t nrowsB = (size_t)t->mat->n, nrowsC = (size_t)(t->row_end - t->row_begin);
    // Row groups: a chunk's multiply is launched over NG ranges of 128-row panels, and the rows of a range are copied out as
    // soon as that range is done -- the copy-out of chunk i starts one range (not one chunk) after its copy-in ended, so the
    // D2H engine is free earlier for the last chunk, whose copy-out nothing overlaps.
    static const int want_groups = getenv("FLEX_HOST_GROUPS") ? atoi(getenv("FLEX_HOST_GROUPS")) : 4;
    const int npanel = t->aspt.npanel;
    const int NG = std::max(1, std::min(std::min(want_groups, 8), npanel / 256));  // a range of fewer panels is under one wave
    if (!t->pipe_g[0]) for (int i = 0; i < 64; ++i) FX_CUDA(cudaEventCreateWithFlags(&t->pipe_g[i], cudaEventDisableTiming));
    FX_CUDA(cudaEventRecord(t->pipe_e0, sin));
    for (int i = 0; i < nchunk; ++i) {
      const size_t c0 = (size_t)i * cw;
      FX_CUDA(cudaMemcpy2DAsync(t->B_stage_dev + c0, pitch, B_host + c0, pitch, wbytes, nrowsB, cudaMemcpyHostToDevice, sin));
      FX_CUDA(cudaEventRecord(t->pipe_in[i], sin));
      FX_CUDA(cudaStreamWaitEvent(sk, t->pipe_in[i], 0));
      FX_CUDA(cudaEventRecord(t->pipe_k0[i], sk));
      for (int g = 0; g < NG; ++g) {
        const int p_lo = (int)((long long)npanel * g / NG), p_hi = (int)((long long)npanel * (g + 1) / NG);
        int rc = fx::spmm_aspt(t, t->B_stage_dev + c0, t->C_stage_dev + c0, k, sk, cw, p_lo, p_hi);
        if (rc != FX_OK) return rc;
        cudaEvent_t done = g + 1 < NG ? t->pipe_g[i * 8 + g] : t->pipe_k1[i];
        FX_CUDA(cudaEventRecord(done, sk));
        FX_CUDA(cudaStreamWaitEvent(sout, done, 0));
        const size_t r_lo = std::min(nrowsC, (size_t)p_lo * 128), r_hi = std::min(nrowsC, (size_t)p_hi * 128);
        if (r_hi > r_lo)
          FX_CUDA(cudaMemcpy2DAsync(C_host + r_lo * k + c0, pitch, t->C_stage_dev + r_lo * k + c0, pitch, wbytes, r_hi - r_lo,
                                    cudaMemcpyDeviceToHost, sout));
      }
    }
    FX_CUDA(cudaEventRecord(t->pipe_e1, sout));
    FX_CUDA(cudaEventSynchronize(t->pipe_e1));
    if (total_ms) FX_CUDA(cudaEventElapsedTime(total_ms, t->pipe_e0, t->pipe_e1));
    if (tElap_ms) {
      float sum = 0.f, ms = 0.f;
      for (int i = 0; i < nchunk; ++i) { FX_CUDA(cudaEventElapsedTime(&ms, t->pipe_k0[i], t->pipe_k1[i])); sum += ms; }
      *tElap_ms = sum;
    }
    return FX_OK;
  }
  cudaStream_t s = t->own_stream;
  cudaEvent_t e0, e1, k0, k1;
  FX_CUDA(cudaEventCreate(&e0)); FX_CUDA(cudaEventCreate(&e1));
  FX_CUDA(cudaEventCreate(&k0)); FX_CUDA(cudaEventCreate(&k1));
  FX_CUDA(cudaEventRecord(e0, s));
  FX_CUDA(cudaMemcpyAsync(t->B_stage_dev, B_host, sizeof(float) * nB, cudaMemcpyHostToDevice, s));
  FX_CUDA(cudaEventRecord(k0, s));
  int rc = spmm_dispatch(t, t->B_stage_dev, t->C_stage_dev, k, s);
  if (rc != FX_OK) return rc;
  FX_CUDA(cudaEventRecord(k1, s));
  FX_CUDA(cudaMemcpyAsync(C_host, t->C_stage_dev, sizeof(float) * nC, cudaMemcpyDeviceToHost, s));
  FX_CUDA(cudaEventRecord(e1, s));
  FX_CUDA(cudaEventSynchronize(e1));
  if (total_ms) FX_CUDA(cudaEventElapsedTime(total_ms, e0, e1));
  if (tElap_ms) FX_CUDA(cudaEventElapsedTime(tElap_ms, k0, k1));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(k0); cudaEventDestroy(k1);
  return FX_OK;
}

extern "C" int fx_permute_rows(const fx_matrix* m, const float* B_dev, float* shadowB_dev, int k, void* stream) {
  FX_REQUIRE(m && B_dev && shadowB_dev, FX_ERR_ARG, "fx_permute_rows: bad argument");
  if (int rc = fx::ensure_device(m)) return rc;
  return fx::permute_rows(m->vo_mp_dev, m->n, k, B_dev, shadowB_dev, false, static_cast<cudaStream_t>(stream));
}

extern "C" int fx_unpermute_rows(const fx_matrix* m, const float* C_dev, float* C_out_dev, int k, void* stream) {
  FX_REQUIRE(m && C_dev && C_out_dev, FX_ERR_ARG, "fx_unpermute_rows: bad argument");
  if (int rc = fx::ensure_device(m)) return rc;
  return fx::permute_rows(m->vo_mp_dev, m->n, k, C_dev, C_out_dev, true, static_cast<cudaStream_t>(stream));
}

// resCheck (flex.cu:4155-4213) + ASpT validator (aspt/sspmm_128.cu:1425-1446) + the 1e-5 contract
extern "C" int fx_check(const float* gold, const float* res, int64_t n, int k, const uint32_t* rowptr, fx_report* rep) {
  FX_REQUIRE(gold && res && rep, FX_ERR_ARG, "fx_check: null");
  int64_t flex = 0, tight = 0, aspt = 0;
  double max_err = 0;
  for (int64_t r = 0; r < n; ++r) {
    const int rnnz = rowptr ? (int)(rowptr[r + 1] - rowptr[r]) : 1;
    const double tol = (double)FLT_EPSILON * rnnz * 4;
    double rowmax = 1.0;  // the 1e-5 contract is row-normwise: |d| <= 1e-5*max(1, ||gold[r,:]||_inf)
    for (int j = 0; j < k; ++j) rowmax = std::max(rowmax, (double)std::fabs(gold[r * k + j]));
    for (int j = 0; j < k; ++j) {
      const float g = gold[r * k + j], x = res[r * k + j];
      const double d = std::fabs((double)g - (double)x);
      const double err = std::fabs(g) < 1 ? d : d / std::fabs(g);
      if (!(err <= tol)) flex++;
      if (err > max_err) max_err = err;
      const float p1 = std::fabs(g), p2 = std::fabs(x);
      if (std::fabs(p1 - p2) / std::max(p1, p2) > 0.01f) aspt++;
      if (!(d <= 1e-5 * rowmax)) tight++;
    }
  }
  rep->errs_flex = flex;
  rep->errs_tight = tight;
  rep->errs_aspt_pct = (n * k) != 0 ? (double)aspt / (double)(n * k) * 100 : 0;
  rep->max_err = max_err;
  return FX_OK;
}
