// flexb200 <path.csv> <k> [--order ovo|deg|rcm|gor|dfs|rbt] [--format tcw|aspt|csr] [--check]
// CLI mirror of `./flex <csv> <k>` (main.cu:7-13) and `./sspmm_128 <csv> <k>`
// (aspt/sspmm_128.cu:1460-1468): loads the CSV, optionally reorders, builds the tile format on the
// GPU, runs C = A*B with the reference's B stream, prints the reference's report lines.  --check adds the
// reference's validation step: the CPU loop of aspt/sspmm_128.cu:1415-1422 as the gold result and the error
// counts of resCheck (flex.cu:4155) and of the ASpT validator (:1425-1446).  The loop is the checker, not a
// fallback: C always comes from the GPU.
#include <cstring>
#include <memory>

#include "flex_driver.hpp"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <path.csv> <k> [--order ovo|deg|rcm|gor|dfs|rbt] [--format tcw|aspt|csr] [--check]\n", argv[0]);
    return 2;
  }
  try {
    const int k = std::atoi(argv[2]);
    fx_order ord = FX_ORDER_OVO;
    int fmt = FX_FMT_TCW;
    bool check = false;
    for (int i = 3; i < argc; ++i) {
      if (!std::strcmp(argv[i], "--check")) {
        check = true;
      } else if (!std::strcmp(argv[i], "--order") && i + 1 < argc) {
        const char* o = argv[++i];
        ord = !std::strcmp(o, "deg") ? FX_ORDER_DEG : !std::strcmp(o, "rcm") ? FX_ORDER_RCM
              : !std::strcmp(o, "gor") ? FX_ORDER_GOR : !std::strcmp(o, "dfs") ? FX_ORDER_DFS
              : !std::strcmp(o, "rbt") ? FX_ORDER_RBT : FX_ORDER_OVO;
      } else if (!std::strcmp(argv[i], "--format") && i + 1 < argc) {
        const char* f = argv[++i];
        fmt = !std::strcmp(f, "csr") ? FX_FMT_CSR : !std::strcmp(f, "aspt") ? FX_FMT_ASPT : FX_FMT_TCW;
      }
    }
    std::printf("-----------  %s  ---------------- start \n", argv[1]);
    flexb200::DataLoader data(argv[1], k);
    std::unique_ptr<flexb200::DataLoader> re;
    const flexb200::DataLoader* A = &data;
    if (ord != FX_ORDER_OVO) { re.reset(flexb200::reorder(data, ord)); A = re.get(); }
    std::printf("graph %s order %s: n = %zu nnz = %zu k = %d directed = %d\n", A->graph_name.c_str(),
                A->vertex_order_abbr.c_str(), A->n, A->nnz, k, A->info.is_directed);
    std::vector<float> B(A->n * (size_t)k), C(A->n * (size_t)k);
    flexb200::ck(fx_rand_B((int64_t)A->n, k, B.data()));
    std::vector<float> gold;
    if (check) {  // aspt/sspmm_128.cu:1415-1422: row-major B and C, fp32 accumulate in CSR order
      gold.assign(A->n * (size_t)k, 0.f);
      for (size_t i = 0; i < A->n; ++i)
        for (uint32_t e = A->rowPtr[i]; e < A->rowPtr[i + 1]; ++e) {
          const float v = A->vals[e];
          const float* b = B.data() + (size_t)A->col[e] * k;
          float* g = gold.data() + i * (size_t)k;
          for (int j = 0; j < k; ++j) g[j] += v * b[j];
        }
    }
    fx_report rep = flexb200::flex_spmm(*A, B.data(), C.data(), k, fmt, check ? gold.data() : nullptr);
    std::printf("tPre: %f ms\ntElap: %f ms\n", rep.tPre_ms, rep.tElap_ms);
    std::printf("GFLOPS: %f\n", rep.gflops);                    // aspt/sspmm_128.cu:1406
    std::printf("t_pre/t_exe: %f\n", rep.tpre_over_telap);      // :1408
    if (check)
      std::printf("errs: %lld (resCheck)  %f %% (ASpT validator)  %lld (1e-5 row-normwise)  max_err %g\n", (long long)rep.errs_flex,
                  rep.errs_aspt_pct, (long long)rep.errs_tight, rep.max_err);
    std::printf("-----------  %s  ----------------- end \n", argv[1]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
