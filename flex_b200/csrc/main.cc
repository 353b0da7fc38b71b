// flexb200 <path.csv> <k> [--order ovo|deg|rcm|gor] [--format tcw|aspt|csr]
// CLI mirror of `./flex <csv> <k>` (main.cu:7-13) and `./sspmm_128 <csv> <k>`
// (aspt/sspmm_128.cu:1460-1468): loads the CSV, optionally reorders, builds the tile format on the
// GPU, runs C = A*B with the reference's B stream, prints the reference's report lines.
#include <cstring>
#include <memory>

#include "flex_driver.hpp"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <path.csv> <k> [--order ovo|deg|rcm|gor] [--format tcw|aspt|csr]\n", argv[0]);
    return 2;
  }
  try {
    const int k = std::atoi(argv[2]);
    fx_order ord = FX_ORDER_OVO;
    int fmt = FX_FMT_TCW;
    for (int i = 3; i + 1 < argc; i += 2) {
      if (!std::strcmp(argv[i], "--order")) {
        const char* o = argv[i + 1];
        ord = !std::strcmp(o, "deg") ? FX_ORDER_DEG : !std::strcmp(o, "rcm") ? FX_ORDER_RCM
              : !std::strcmp(o, "gor") ? FX_ORDER_GOR : FX_ORDER_OVO;
      } else if (!std::strcmp(argv[i], "--format")) {
        fmt = !std::strcmp(argv[i + 1], "csr") ? FX_FMT_CSR : !std::strcmp(argv[i + 1], "aspt") ? FX_FMT_ASPT : FX_FMT_TCW;
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
    fx_report rep = flexb200::flex_spmm(*A, B.data(), C.data(), k, fmt);
    std::printf("tPre: %f ms\ntElap: %f ms\n", rep.tPre_ms, rep.tElap_ms);
    std::printf("GFLOPS: %f\n", rep.gflops);                    // aspt/sspmm_128.cu:1406
    std::printf("t_pre/t_exe: %f\n", rep.tpre_over_telap);      // :1408
    std::printf("-----------  %s  ----------------- end \n", argv[1]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
