// fx_io.cu -- N2: the on-disk formats either side of the CSV loader.
//   fx_mtx_load      Matrix Market coordinate file -> CSR, the conversion data/SuiteSparse/mtx2csr.cc
//                    performs (mmio_allinone :57-222: 1-based -> 0-based, symmetric/hermitian files
//                    mirrored, pattern files get 1.0, complex files keep the real part) followed by the
//                    text round trip of writeCSR2csv (:224-246: `ofstream << float`, 6 significant
//                    digits) that the reference's values go through before DataLoader parses them.
//                    Rows are additionally sorted by column (the converter leaves them in file order;
//                    every builder downstream needs ascending columns, DataLoader.cu:272).
//   fx_csr_write_csv the 3-line CSV of writeCSR2csv, same number formatting
//   fx_csr_save_bin / fx_csr_load_bin   a binary CSR cache: parsing 3 GB of ASCII for an Amazon-sized
//                    graph costs far more than everything on the GPU; the cache is one read().
#include <algorithm>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>

#include "fx_common.cuh"

namespace fx {
int finish_matrix(fx_matrix* m, const std::string& name, int order, bool do_upload);
}

namespace {
std::string base_name(const std::string& p) { return p.substr(p.find_last_of("/") + 1); }

float text_round_trip(float v) {  // what `myFile << value[i]` followed by std::stof gives back
  char buf[64];
  snprintf(buf, sizeof(buf), "%g", (double)v);
  return strtof(buf, nullptr);
}
}  // namespace

extern "C" int fx_mtx_load(const char* path, int k, fx_matrix** out) {
  FX_REQUIRE(path && out && k > 0, FX_ERR_ARG, "fx_mtx_load: bad argument");
  std::ifstream f(path);
  FX_REQUIRE(f.good(), FX_ERR_IO, "cannot open %s", path);
  std::string line;
  FX_REQUIRE((bool)std::getline(f, line), FX_ERR_IO, "%s: empty file", path);
  std::string low = line;
  std::transform(low.begin(), low.end(), low.begin(), ::tolower);
  FX_REQUIRE(low.rfind("%%matrixmarket", 0) == 0 && low.find("coordinate") != std::string::npos, FX_ERR_IO,
             "%s: not a Matrix Market coordinate file", path);
  const bool pattern = low.find("pattern") != std::string::npos, cplx = low.find("complex") != std::string::npos;
  const bool sym = low.find("symmetric") != std::string::npos || low.find("hermitian") != std::string::npos;
  do {
    FX_REQUIRE((bool)std::getline(f, line), FX_ERR_IO, "%s: no size line", path);
  } while (!line.empty() && line[0] == '%');
  long long M = 0, N = 0, NZ = 0;
  FX_REQUIRE(sscanf(line.c_str(), "%lld %lld %lld", &M, &N, &NZ) == 3 && M > 0 && NZ >= 0, FX_ERR_IO, "%s: bad size line", path);
  FX_REQUIRE(M == N, FX_ERR_FORMAT, "%s: %lld x %lld is not square (DataLoader.cu:58-59 sets n = m)", path, M, N);
  std::vector<uint32_t> ri(NZ), ci(NZ);
  std::vector<float> vv(NZ);
  std::vector<uint32_t> cnt(M + 1, 0);
  for (long long i = 0; i < NZ; ++i) {
    long long a, b;
    double x = 1.0, y = 0.0;
    f >> a >> b;
    if (!pattern) f >> x;
    if (cplx) f >> y;
    FX_REQUIRE(f.good() || f.eof(), FX_ERR_IO, "%s: entry %lld unreadable", path, i);
    FX_REQUIRE(a >= 1 && a <= M && b >= 1 && b <= N, FX_ERR_IO, "%s: entry %lld out of range", path, i);
    ri[i] = (uint32_t)(a - 1); ci[i] = (uint32_t)(b - 1); vv[i] = (float)x;
    cnt[ri[i]]++;
    if (sym && ri[i] != ci[i]) cnt[ci[i]]++;
  }
  auto m = new fx_matrix();
  m->n = M; m->k = k;
  m->rowptr.assign(M + 1, 0);
  for (long long r = 0; r < M; ++r) m->rowptr[r + 1] = m->rowptr[r] + cnt[r];
  m->nnz = m->rowptr[M];
  m->col.resize(m->nnz); m->val.resize(m->nnz);
  std::vector<uint32_t> fill(m->rowptr.begin(), m->rowptr.end() - 1);
  for (long long i = 0; i < NZ; ++i) {  // placement order of mmio_allinone :173-207
    m->col[fill[ri[i]]] = ci[i]; m->val[fill[ri[i]]++] = vv[i];
    if (sym && ri[i] != ci[i]) { m->col[fill[ci[i]]] = ri[i]; m->val[fill[ci[i]]++] = vv[i]; }
  }
  // ascending columns per row (stable), values through the CSV text round trip
  std::vector<uint32_t> idx;
  std::vector<uint32_t> c2;
  std::vector<float> v2;
  for (long long r = 0; r < M; ++r) {
    const uint32_t lo = m->rowptr[r], hi = m->rowptr[r + 1];
    idx.resize(hi - lo);
    std::iota(idx.begin(), idx.end(), lo);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return m->col[a] < m->col[b]; });
    c2.resize(hi - lo); v2.resize(hi - lo);
    for (size_t q = 0; q < idx.size(); ++q) { c2[q] = m->col[idx[q]]; v2[q] = text_round_trip(m->val[idx[q]]); }
    std::copy(c2.begin(), c2.end(), m->col.begin() + lo);
    std::copy(v2.begin(), v2.end(), m->val.begin() + lo);
  }
  m->vo_mp.resize(M);
  std::iota(m->vo_mp.begin(), m->vo_mp.end(), 0);
  std::string name = base_name(path);
  name = name.substr(0, name.find(".")) + ".csv";
  int rc = fx::finish_matrix(m, name, FX_ORDER_OVO, false);
  if (rc != FX_OK) { fx_matrix_free(m); return rc; }
  *out = m;
  return FX_OK;
}

extern "C" int fx_csr_write_csv(const fx_matrix* m, const char* path) {
  FX_REQUIRE(m && path && (!m->col.empty() || m->nnz == 0), FX_ERR_ARG, "fx_csr_write_csv: needs a host CSR");
  std::ofstream f(path);
  FX_REQUIRE(f.good(), FX_ERR_IO, "cannot create %s", path);
  for (int64_t i = 0; i <= m->n; ++i) { f << m->rowptr[i]; if (i < m->n) f << ","; }
  f << "\n";
  for (int64_t i = 0; i < m->nnz; ++i) { f << m->col[i]; if (i + 1 < m->nnz) f << ","; }
  f << "\n";
  for (int64_t i = 0; i < m->nnz; ++i) { f << m->val[i]; if (i + 1 < m->nnz) f << ","; }  // 6 significant digits, as :240
  f << "\n";
  FX_REQUIRE(f.good(), FX_ERR_IO, "write to %s failed", path);
  return FX_OK;
}

namespace {
struct BinHeader { char magic[8]; int64_t n, nnz; int32_t k, order; };
const char kMagic[8] = {'F', 'X', 'C', 'S', 'R', '0', '1', 0};
}  // namespace

extern "C" int fx_csr_save_bin(const fx_matrix* m, const char* path) {
  FX_REQUIRE(m && path && (!m->col.empty() || m->nnz == 0), FX_ERR_ARG, "fx_csr_save_bin: needs a host CSR");
  std::ofstream f(path, std::ios::binary);
  FX_REQUIRE(f.good(), FX_ERR_IO, "cannot create %s", path);
  BinHeader h{};
  memcpy(h.magic, kMagic, 8);
  h.n = m->n; h.nnz = m->nnz; h.k = m->k; h.order = m->info.order;
  f.write(reinterpret_cast<const char*>(&h), sizeof(h));
  f.write(m->info.graph_name, sizeof(m->info.graph_name));
  f.write(reinterpret_cast<const char*>(m->rowptr.data()), sizeof(uint32_t) * (m->n + 1));
  f.write(reinterpret_cast<const char*>(m->col.data()), sizeof(uint32_t) * m->nnz);
  f.write(reinterpret_cast<const char*>(m->val.data()), sizeof(float) * m->nnz);
  f.write(reinterpret_cast<const char*>(m->vo_mp.data()), sizeof(int32_t) * m->n);
  FX_REQUIRE(f.good(), FX_ERR_IO, "write to %s failed", path);
  return FX_OK;
}

extern "C" int fx_csr_load_bin(const char* path, int k, fx_matrix** out) {
  FX_REQUIRE(path && out, FX_ERR_ARG, "fx_csr_load_bin: bad argument");
  std::ifstream f(path, std::ios::binary);
  FX_REQUIRE(f.good(), FX_ERR_IO, "cannot open %s", path);
  BinHeader h{};
  f.read(reinterpret_cast<char*>(&h), sizeof(h));
  FX_REQUIRE(f.good() && !memcmp(h.magic, kMagic, 8) && h.n >= 0 && h.nnz >= 0, FX_ERR_IO, "%s: not a flex-b200 CSR cache", path);
  char gname[64];
  f.read(gname, sizeof(gname));
  gname[63] = 0;
  auto m = new fx_matrix();
  m->n = h.n; m->nnz = h.nnz; m->k = k > 0 ? k : h.k;
  m->rowptr.resize(h.n + 1); m->col.resize(h.nnz); m->val.resize(h.nnz); m->vo_mp.resize(h.n);
  f.read(reinterpret_cast<char*>(m->rowptr.data()), sizeof(uint32_t) * (h.n + 1));
  f.read(reinterpret_cast<char*>(m->col.data()), sizeof(uint32_t) * h.nnz);
  f.read(reinterpret_cast<char*>(m->val.data()), sizeof(float) * h.nnz);
  f.read(reinterpret_cast<char*>(m->vo_mp.data()), sizeof(int32_t) * h.n);
  if (!f.good()) { delete m; fx::set_error("%s: truncated", path); return FX_ERR_IO; }
  int rc = fx::finish_matrix(m, std::string(gname) + ".csv", h.order, false);
  if (rc != FX_OK) { fx_matrix_free(m); return rc; }
  *out = m;
  return FX_OK;
}
