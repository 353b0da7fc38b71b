// flex_driver.hpp -- C++ facade over the C ABI with the reference's class names, so that code
// written against DataLoader / Mat / run() (main.cu:13,80; flex.cuh:59) reads the same.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "flexb200.h"

namespace flexb200 {

inline void ck(int rc) {
  if (rc != FX_OK) throw std::runtime_error(std::string("libflexb200: ") + fx_last_error());  // CUDA_CHECK common.h:53-60
}

class DataLoader {  // DataLoader.cuh:21-112
 public:
  DataLoader(const std::string& path, int k) : dim(k) { ck(fx_csr_load(path.c_str(), k, &h_)); load_info(); }
  DataLoader(fx_matrix* h, int k) : dim(k), h_(h) { load_info(); }
  ~DataLoader() { fx_matrix_free(h_); }
  DataLoader(const DataLoader&) = delete;
  fx_matrix* handle() const { return h_; }
  size_t m = 0, n = 0, nnz = 0, c = 0;
  int dim = 0;
  std::string graph_name, vertex_order_abbr;
  const uint32_t *rowPtr = nullptr, *col = nullptr;
  const float* vals = nullptr;
  const int32_t* vo_mp = nullptr;
  fx_matrix_info info{};

 protected:
  void load_info() {
    ck(fx_matrix_get_info(h_, &info));
    m = info.m; n = info.n; nnz = info.nnz; c = info.c;
    graph_name = info.graph_name; vertex_order_abbr = info.order_abbr;
    ck(fx_matrix_host_csr(h_, &rowPtr, &col, &vals));
    ck(fx_permutation(h_, &vo_mp));
  }
  fx_matrix* h_ = nullptr;
};

inline DataLoader* reorder(const DataLoader& dl, fx_order o) {  // DataLoaderDeg/Rcm/Gorder(dl)
  fx_matrix* h = nullptr;
  ck(fx_reorder(dl.handle(), o, &h));
  return new DataLoader(h, dl.dim);
}
inline DataLoader* DataLoaderDeg(const DataLoader& dl) { return reorder(dl, FX_ORDER_DEG); }
inline DataLoader* DataLoaderRcm(const DataLoader& dl) { return reorder(dl, FX_ORDER_RCM); }
inline DataLoader* DataLoaderGorder(const DataLoader& dl) { return reorder(dl, FX_ORDER_GOR); }

class Mat {  // mat.cuh:67-229
 public:
  Mat(const DataLoader& dl, int format = FX_FMT_ASPT, int tm = 4, int tn = 4) {
    fx_build_opts o{};
    o.format = format; o.tm = tm; o.tn = tn;
    ck(fx_build(dl.handle(), &o, &h_, &tPre_ms));
  }
  ~Mat() { fx_tiles_free(h_); }
  Mat(const Mat&) = delete;
  fx_tiles* handle() const { return h_; }
  float tPre_ms = 0;

 private:
  fx_tiles* h_ = nullptr;
};

// flex_spmm(A, B, k): B, C host buffers (n*k); returns the tPre/tElap/GFlops/Errs report
inline fx_report flex_spmm(const DataLoader& A, const float* B, float* C, int k, int format = FX_FMT_ASPT,
                           const float* gold = nullptr, int tm = 4, int tn = 4) {
  Mat mat(A, format, tm, tn);
  float total = 0, telap = 0;
  ck(fx_spmm_host(mat.handle(), B, C, k, nullptr, nullptr));  // untimed first run (module load, staging buffers), as the
                                                              // reference warms up before its timed loop (flex.cu:5051)
  ck(fx_spmm_host(mat.handle(), B, C, k, &total, &telap));
  fx_report rep{};
  if (gold) ck(fx_check(gold, C, (int64_t)A.n, k, A.rowPtr, &rep));
  rep.tPre_ms = mat.tPre_ms;
  rep.tElap_ms = telap;
  rep.gflops = telap > 0 ? 2.0 * A.nnz * k / (telap * 1e-3) / 1e9 : 0;
  rep.tpre_over_telap = telap > 0 ? mat.tPre_ms / telap : 0;
  return rep;
}

}  // namespace flexb200
