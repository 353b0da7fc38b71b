// fx_matrix.cu -- L0 of the hot path: 3-line CSV -> CSR -> HBM  (replaces DataLoader.cu:9-218).
//
// Host work is deliberately different from the reference: one pass over an in-memory copy of the
// file with from_chars (the reference tokenises through stringstream/stoi per value), and the
// direction / zero-degree census (DataLoader.cu:86-115) is computed from a transposed CSR built by
// counting sort instead of vector<map<int,float>> (50-100 B/edge there; 12 B/edge here).
#include <algorithm>
#include <charconv>
#include <fstream>
#include <numeric>

#include "fx_common.cuh"

namespace fx {
static thread_local std::string g_err;
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
}  // namespace fx

extern "C" const char* fx_last_error(void) { return fx::g_err.c_str(); }
extern "C" int fx_version(void) { return 100; }
extern "C" int64_t fx_launch_count(void) { return fx::g_launches.load(); }
extern "C" int fx_device_sm_count(int* n_sm) {
  int dev = 0;
  FX_CUDA(cudaGetDevice(&dev));
  FX_CUDA(cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, dev));
  return FX_OK;
}

namespace {

int class_count(const std::string& name) {  // DataLoader.cu:62-84
  static const struct { const char* f; int c; } tab[] = {
      {"polblogs.csv", 2}, {"cora.csv", 7},  {"citeseer.csv", 6}, {"pubmed.csv", 3}, {"ppi.csv", 121},
      {"reddit.csv", 41},  {"flickr.csv", 7}, {"yelp.csv", 100},  {"amazon.csv", 107}};
  for (auto& e : tab)
    if (name == e.f) return e.c;
  return 100;
}

template <class T>
bool parse_line(const char*& p, const char* end, std::vector<T>& out) {
  // comma separated numbers up to '\n' (or end); tolerant of a trailing comma / CR
  while (p < end && *p != '\n') {
    while (p < end && (*p == ' ' || *p == '\r')) ++p;
    if (p >= end || *p == '\n') break;
    T v{};
    if constexpr (std::is_floating_point<T>::value) {
      if (*p == '+') ++p;
    }
    auto r = std::from_chars(p, end, v);
    if (r.ec != std::errc()) return false;
    out.push_back(v);
    p = r.ptr;
    while (p < end && (*p == ' ' || *p == '\r')) ++p;
    if (p < end && *p == ',') ++p;
  }
  if (p < end && *p == '\n') ++p;
  return true;
}

// fx_csr_from_device: the same column checks as validate_csr, on the device (one warp per row).  flag bit 0: a column
// >= n; bit 1: columns of a row not strictly ascending (the GPU builders count with col[e] as an index and assume
// ascending columns: fx_aspt_build.cu k_heavy, fx_tcw_build.cu k_tcw_select / k_tcw_split).
__global__ void k_validate_cols(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, long long n,
                                unsigned* __restrict__ flag) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const uint32_t lo = rowptr[r], hi = rowptr[r + 1];
  unsigned bad = 0;
  for (uint32_t e = lo + lane; e < hi; e += 32) {
    const uint32_t c = col[e];
    if (c >= (uint32_t)n) bad |= 1u;
    if (e > lo && col[e - 1] >= c) bad |= 2u;
  }
  if (bad) atomicOr(flag, bad);
}

__global__ void k_iota(int32_t* __restrict__ p, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

void census(fx_matrix* m) {  // DataLoader.cu:86-115
  const int64_t n = m->n, nnz = m->nnz;
  std::vector<uint32_t> tp(n + 2, 0);
  for (int64_t e = 0; e < nnz; ++e) tp[m->col[e] + 1]++;
  for (int64_t i = 0; i < n; ++i) tp[i + 1] += tp[i];
  std::vector<uint32_t> cur(tp.begin(), tp.begin() + n + 1), tsrc(nnz);
  std::vector<float> tval(nnz);
  for (int64_t r = 0; r < n; ++r)
    for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) {
      uint32_t d = m->col[e];
      tsrc[cur[d]] = (uint32_t)r;
      tval[cur[d]++] = m->val[e];
    }
  int64_t one_way = 0, asym = 0;
  for (int64_t r = 0; r < n; ++r)
    for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) {
      auto b = tsrc.begin() + tp[r], en = tsrc.begin() + tp[r + 1];
      auto it = std::lower_bound(b, en, m->col[e]);
      if (it == en || *it != m->col[e]) one_way++;
      else if (tval[it - tsrc.begin()] != m->val[e]) asym++;
    }
  int zo = 0, zi = 0, zd = 0;
  for (int64_t r = 0; r < n; ++r) {
    bool z_out = m->rowptr[r] == m->rowptr[r + 1], z_in = tp[r] == tp[r + 1];
    zo += z_out; zi += z_in; zd += (z_out && z_in);
  }
  m->info.n_edges_one_way = one_way;
  m->info.n_edges_asymmetric = asym;
  m->info.n_nodes_z_out = zo; m->info.n_nodes_z_in = zi; m->info.n_nodes_z_deg = zd;
  m->info.is_directed = one_way != 0;
  m->census_done = true;
}

int validate_csr(const fx_matrix* m) {
  // preconditions the reference asserts: square (DataLoader.cu:58-59), unique columns per row
  // (:97), columns ascending ("Tiling algorithm needs dests sorted" :272)
  const int64_t n = m->n;
  FX_REQUIRE(m->rowptr.size() == (size_t)n + 1 && m->rowptr[0] == 0 && m->rowptr[n] == (uint32_t)m->nnz,
             FX_ERR_FORMAT, "rowPtr does not span [0,nnz]");
  for (int64_t r = 0; r < n; ++r) {
    FX_REQUIRE(m->rowptr[r] <= m->rowptr[r + 1], FX_ERR_FORMAT, "rowPtr not monotone at row %lld", (long long)r);
    for (uint32_t e = m->rowptr[r]; e < m->rowptr[r + 1]; ++e) {
      FX_REQUIRE(m->col[e] < (uint32_t)n, FX_ERR_FORMAT, "column %u out of range in row %lld", m->col[e], (long long)r);
      FX_REQUIRE(e == m->rowptr[r] || m->col[e - 1] < m->col[e], FX_ERR_FORMAT,
                 "columns of row %lld are not strictly ascending (reference asserts uniqueness, DataLoader.cu:97)",
                 (long long)r);
    }
  }
  return FX_OK;
}

int upload(fx_matrix* m) {  // DataLoader::cuda_alloc_cpy, CSR part (DataLoader.cu:184-191)
  if (m->rowptr_dev) return FX_OK;
  FX_CUDA(cudaMalloc(&m->rowptr_dev, sizeof(uint32_t) * (m->n + 1)));
  FX_CUDA(cudaMalloc(&m->col_dev, sizeof(uint32_t) * std::max<int64_t>(m->nnz, 1)));
  FX_CUDA(cudaMalloc(&m->val_dev, sizeof(float) * std::max<int64_t>(m->nnz, 1)));
  FX_CUDA(cudaMemcpy(m->rowptr_dev, m->rowptr.data(), sizeof(uint32_t) * (m->n + 1), cudaMemcpyHostToDevice));
  if (m->nnz) {
    FX_CUDA(cudaMemcpy(m->col_dev, m->col.data(), sizeof(uint32_t) * m->nnz, cudaMemcpyHostToDevice));
    FX_CUDA(cudaMemcpy(m->val_dev, m->val.data(), sizeof(float) * m->nnz, cudaMemcpyHostToDevice));
  }
  if (!m->vo_mp.empty()) {
    FX_CUDA(cudaMalloc(&m->vo_mp_dev, sizeof(int32_t) * m->n));
    FX_CUDA(cudaMemcpy(m->vo_mp_dev, m->vo_mp.data(), sizeof(int32_t) * m->n, cudaMemcpyHostToDevice));
  }
  return FX_OK;
}

void fill_info(fx_matrix* m, const std::string& file_name, int order) {
  m->info.m = m->info.n = m->n;
  m->info.nnz = m->nnz;
  m->info.dim = m->k;
  m->info.c = class_count(file_name);
  std::string g = file_name.substr(0, file_name.find("."));  // DataLoader.cu:12
  snprintf(m->info.graph_name, sizeof(m->info.graph_name), "%s", g.c_str());
  static const char* abbr[] = {"OVO", "DEG", "RCM", "GOR", "DFS", "RBT", "OVO", "OVO"};
  m->info.order = order;
  snprintf(m->info.order_abbr, sizeof(m->info.order_abbr), "%s", abbr[order & 7]);
  int64_t uni = 0;  // DataLoader.cu:24-27
  for (int64_t i = 1; i <= m->n; ++i) uni += (m->rowptr[i] - m->rowptr[i - 1] == 1);
  m->info.uni_nb = uni;
}

}  // namespace

namespace fx {
// The device copy is made on first use (fx_build, fx_permute_rows, fx_matrix_device_csr), so the
// host-side logic (CSV parse, census, reordering) is usable -- and testable -- without a GPU.
int finish_matrix(fx_matrix* m, const std::string& name, int order, bool do_upload) {
  int rc = validate_csr(m);
  if (rc != FX_OK) return rc;
  fill_info(m, name, order);
  if (do_upload) return upload(m);
  return FX_OK;
}
int ensure_device(const fx_matrix* m) { return upload(const_cast<fx_matrix*>(m)); }
int ensure_census(const fx_matrix* m) {
  if (!m->census_done && !m->col.empty()) census(const_cast<fx_matrix*>(m));
  return FX_OK;
}
}  // namespace fx

extern "C" int fx_csr_load(const char* path, int k, fx_matrix** out) {
  FX_REQUIRE(path && out && k > 0, FX_ERR_ARG, "fx_csr_load: bad argument");
  std::ifstream fin(path, std::ios::binary | std::ios::ate);
  FX_REQUIRE(fin.good(), FX_ERR_IO, "cannot open %s", path);
  std::streamsize sz = fin.tellg();
  fin.seekg(0);
  std::string buf((size_t)sz, '\0');
  fin.read(buf.data(), sz);
  fin.close();
  auto m = new fx_matrix();
  m->k = k;
  const char *p = buf.data(), *end = buf.data() + buf.size();
  std::string sp(path);
  std::string name = sp.substr(sp.find_last_of("/") + 1);
  bool ok = parse_line(p, end, m->rowptr) && parse_line(p, end, m->col);
  if (ok) {
    if (name == "amazon.csv") {  // DataLoader.cu:36-46: no value line
      m->val.resize(m->col.size());
      for (auto& v : m->val) v = 2 * (float)rand() / (float)RAND_MAX - 1.0f;
    } else {
      ok = parse_line(p, end, m->val);
    }
  }
  if (!ok || m->rowptr.size() < 1 || m->col.size() != m->val.size()) {
    delete m;
    fx::set_error("malformed CSV %s (need rowPtr / col / vals lines of matching length)", path);
    return FX_ERR_IO;
  }
  m->n = (int64_t)m->rowptr.size() - 1;
  m->nnz = (int64_t)m->col.size();
  m->vo_mp.resize(m->n);
  std::iota(m->vo_mp.begin(), m->vo_mp.end(), 0);  // DataLoader.cu:117-118
  int rc = fx::finish_matrix(m, name, FX_ORDER_OVO, false);
  if (rc != FX_OK) { fx_matrix_free(m); return rc; }
  *out = m;
  return FX_OK;
}

extern "C" int fx_csr_from_arrays(int64_t n, int64_t nnz, const uint32_t* rowptr, const uint32_t* col,
                                  const float* val, int k, const char* name, fx_matrix** out) {
  FX_REQUIRE(out && rowptr && n >= 0 && nnz >= 0 && k > 0 && (nnz == 0 || (col && val)), FX_ERR_ARG,
             "fx_csr_from_arrays: bad argument");
  FX_REQUIRE(nnz < (1ll << 31) && n < (1ll << 31), FX_ERR_UNSUPPORTED, "indices are 32-bit (reference: int everywhere)");
  auto m = new fx_matrix();
  m->n = n; m->nnz = nnz; m->k = k;
  m->rowptr.assign(rowptr, rowptr + n + 1);
  m->col.assign(col, col + nnz);
  m->val.assign(val, val + nnz);
  m->vo_mp.resize(n);
  std::iota(m->vo_mp.begin(), m->vo_mp.end(), 0);
  int rc = fx::finish_matrix(m, name ? name : "arrays.csv", FX_ORDER_OVO, false);
  if (rc != FX_OK) { fx_matrix_free(m); return rc; }
  *out = m;
  return FX_OK;
}

extern "C" int fx_csr_from_device(int64_t n, int64_t nnz, const uint32_t* rowptr_dev, const uint32_t* col_dev,
                                  const float* val_dev, int k, const char* name, fx_matrix** out) {
  FX_REQUIRE(out && rowptr_dev && n >= 0 && nnz >= 0 && k > 0, FX_ERR_ARG, "fx_csr_from_device: bad argument");
  FX_REQUIRE(nnz < (1ll << 31) && n < (1ll << 31), FX_ERR_UNSUPPORTED, "indices are 32-bit");
  auto m = new fx_matrix();
  m->n = n; m->nnz = nnz; m->k = k;
  // only the row pointer comes back to the host (sharding and arena sizing need it)
  m->rowptr.resize(n + 1);
  cudaError_t e = cudaMemcpy(m->rowptr.data(), rowptr_dev, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { delete m; fx::set_error("rowptr D2H -> %s", cudaGetErrorString(e)); return FX_ERR_CUDA; }
  if (m->rowptr[0] != 0 || m->rowptr[n] != (uint32_t)nnz) { delete m; fx::set_error("rowPtr does not span [0,nnz]"); return FX_ERR_FORMAT; }
  for (int64_t r = 0; r < n; ++r)
    if (m->rowptr[r] > m->rowptr[r + 1]) { delete m; fx::set_error("rowPtr not monotone at row %lld", (long long)r); return FX_ERR_FORMAT; }
  fill_info(m, name ? name : "device.csv", FX_ORDER_OVO);
  int rc = FX_OK;
  unsigned* flag_dev = nullptr;
  unsigned flag = 0;
  do {
    if (cudaMalloc(&m->rowptr_dev, sizeof(uint32_t) * (n + 1)) != cudaSuccess ||
        cudaMalloc(&m->col_dev, sizeof(uint32_t) * std::max<int64_t>(nnz, 1)) != cudaSuccess ||
        cudaMalloc(&m->val_dev, sizeof(float) * std::max<int64_t>(nnz, 1)) != cudaSuccess ||
        cudaMalloc(&m->vo_mp_dev, sizeof(int32_t) * std::max<int64_t>(n, 1)) != cudaSuccess ||
        cudaMalloc(&flag_dev, sizeof(unsigned)) != cudaSuccess) { rc = FX_ERR_NOMEM; break; }
    if (cudaMemcpy(m->rowptr_dev, rowptr_dev, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemcpy(m->col_dev, col_dev, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemcpy(m->val_dev, val_dev, sizeof(float) * nnz, cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemset(flag_dev, 0, sizeof(unsigned)) != cudaSuccess) { rc = FX_ERR_CUDA; break; }
    if (n > 0) {
      // the column checks of validate_csr (the builders index counters with col[e] and assume ascending columns) and the
      // identity vo_mp of an unreordered matrix (fx_permute_rows / fx_unpermute_rows read it)
      k_validate_cols<<<fx::ceil_div(n * 32, 256), 256>>>(m->rowptr_dev, m->col_dev, n, flag_dev);
      k_iota<<<fx::ceil_div(n, 256), 256>>>(m->vo_mp_dev, n);
      if (cudaMemcpy(&flag, flag_dev, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = FX_ERR_CUDA; break; }
    }
  } while (0);
  cudaFree(flag_dev);
  if (rc != FX_OK) { fx::set_error("fx_csr_from_device: device copy failed: %s", cudaGetErrorString(cudaGetLastError())); fx_matrix_free(m); return rc; }
  if (flag) {
    fx::set_error(flag & 1u ? "a column index is >= n (DataLoader.cu:58-59: the matrix must be square)"
                            : "columns of a row are not strictly ascending (DataLoader.cu:97,272: unique, sorted dests)");
    fx_matrix_free(m);
    return FX_ERR_FORMAT;
  }
  m->vo_mp.resize(n);
  std::iota(m->vo_mp.begin(), m->vo_mp.end(), 0);
  *out = m;
  return FX_OK;
}

extern "C" int fx_matrix_get_info(const fx_matrix* m, fx_matrix_info* info) {
  FX_REQUIRE(m && info, FX_ERR_ARG, "fx_matrix_get_info: null");
  if (!m->census_done && !m->col.empty()) census(const_cast<fx_matrix*>(m));
  *info = m->info;
  return FX_OK;
}

extern "C" int fx_matrix_host_csr(const fx_matrix* m, const uint32_t** rowptr, const uint32_t** col, const float** val) {
  FX_REQUIRE(m, FX_ERR_ARG, "null matrix");
  if (rowptr) *rowptr = m->rowptr.data();
  if (col) *col = m->col.empty() ? nullptr : m->col.data();
  if (val) *val = m->val.empty() ? nullptr : m->val.data();
  return FX_OK;
}

extern "C" int fx_matrix_device_csr(const fx_matrix* m, const uint32_t** rowptr_dev, const uint32_t** col_dev,
                                    const float** val_dev) {
  FX_REQUIRE(m, FX_ERR_ARG, "null matrix");
  int rc = fx::ensure_device(m);
  if (rc != FX_OK) return rc;
  if (rowptr_dev) *rowptr_dev = m->rowptr_dev;
  if (col_dev) *col_dev = m->col_dev;
  if (val_dev) *val_dev = m->val_dev;
  return FX_OK;
}

extern "C" void fx_matrix_free(fx_matrix* m) {
  if (!m) return;
  cudaFree(m->rowptr_dev); cudaFree(m->col_dev); cudaFree(m->val_dev); cudaFree(m->vo_mp_dev);
  delete m;
}

extern "C" int fx_rand_B(int64_t n, int k, float* B) {  // DataLoader.cu:198-209 (glibc rand, seed 1)
  FX_REQUIRE(B && n >= 0 && k > 0, FX_ERR_ARG, "fx_rand_B: bad argument");
  srand(1);
  for (int64_t i = 0; i < n * k; ++i) B[i] = 2 * (float)rand() / (float)RAND_MAX - 1.0f;
  return FX_OK;
}

extern "C" int fx_permutation(const fx_matrix* m, const int32_t** vo_mp) {
  FX_REQUIRE(m && vo_mp, FX_ERR_ARG, "fx_permutation: null");
  *vo_mp = m->vo_mp.data();
  return FX_OK;
}
