// flexb200 <path.csv> <k> [--order ovo|deg|rcm|gor|dfs|rbt] [--format tcw|aspt|csr|tile|seg|pillar] [--tm N] [--tn N]
//          [--gpus N] [--check]
// CLI mirror of `./flex <csv> <k>` (main.cu:7-13) and `./sspmm_128 <csv> <k>` (aspt/sspmm_128.cu:1460-1468): loads the
// CSV, optionally reorders, builds the tile format on the GPU, runs C = A*B with the reference's B stream, prints the
// reference's report lines.  --check adds the reference's validation step: the CPU loop of aspt/sspmm_128.cu:1415-1422 as
// the gold result and the error counts of resCheck (flex.cu:4155) and of the ASpT validator (:1425-1446).  The loop is the
// checker, not a fallback: C always comes from the GPU.
// --gpus N: one process per GPU (forked before anything touches CUDA or NCCL), rank r on device r;
// every rank loads the CSV, builds the tiles of its row-panel shard (fx_panel_shards) and runs fx_spmm_sharded_host:
// its 1/N slice of B goes up, ncclAllGather assembles B, its rows of C come back.  The NCCL unique id and the per-rank
// results travel through pipes.  The reported times are the maximum over the ranks.
#include <sys/wait.h>
#include <unistd.h>

#include <cstring>
#include <memory>

#include "flex_driver.hpp"

namespace {

struct RankResult {
  float tPre_ms, total_ms, tElap_ms;
  long long errs_flex, errs_tight, aspt_bad, rows;
  double max_err;
  int ok;
};

int parse_format(const char* f) {
  return !std::strcmp(f, "csr") ? FX_FMT_CSR : !std::strcmp(f, "aspt") ? FX_FMT_ASPT : !std::strcmp(f, "tile") ? FX_FMT_TILE
         : !std::strcmp(f, "seg") ? FX_FMT_SEG : !std::strcmp(f, "pillar") ? FX_FMT_PILLAR : FX_FMT_TCW;
}

// aspt/sspmm_128.cu:1415-1422 over rows [lo, hi): row-major B and C, fp32 accumulate in CSR order
void cpu_gold(const flexb200::DataLoader& A, const float* B, int k, size_t lo, size_t hi, std::vector<float>& gold) {
  gold.assign((hi - lo) * (size_t)k, 0.f);
  for (size_t i = lo; i < hi; ++i)
    for (uint32_t e = A.rowPtr[i]; e < A.rowPtr[i + 1]; ++e) {
      const float v = A.vals[e];
      const float* b = B + (size_t)A.col[e] * k;
      float* g = gold.data() + (i - lo) * (size_t)k;
      for (int j = 0; j < k; ++j) g[j] += v * b[j];
    }
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr,
                 "usage: %s <path.csv> <k> [--order ovo|deg|rcm|gor|dfs|rbt] [--format tcw|aspt|csr|tile|seg|pillar] [--tm N] [--tn N] "
                 "[--gpus N] [--check]\n", argv[0]);
    return 2;
  }
  const int k = std::atoi(argv[2]);
  fx_order ord = FX_ORDER_OVO;
  int fmt = FX_FMT_TCW, gpus = 1, tm = 4, tn = 4;
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
      fmt = parse_format(argv[++i]);
    } else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) {
      gpus = std::atoi(argv[++i]);
    } else if (!std::strcmp(argv[i], "--tm") && i + 1 < argc) {
      tm = std::atoi(argv[++i]);
    } else if (!std::strcmp(argv[i], "--tn") && i + 1 < argc) {
      tn = std::atoi(argv[++i]);
    }
  }
  const bool flex_fmt = fmt == FX_FMT_TILE || fmt == FX_FMT_SEG || fmt == FX_FMT_PILLAR;
  if (gpus < 1 || gpus > 64 || (gpus > 1 && fmt != FX_FMT_TCW && fmt != FX_FMT_ASPT)) {
    std::fprintf(stderr, "--gpus N needs 1 <= N <= 64 and --format tcw|aspt (row-panel shards)\n");
    return 2;
  }

  // ---- one process per GPU: fork first, so that no CUDA / NCCL state is inherited ----
  int rank = 0;
  std::vector<int> uid_w(gpus, -1), res_r(gpus, -1);  // parent's ends: unique id to rank r, result from rank r
  int uid_in = -1, res_out = -1;                      // a child's ends
  std::vector<pid_t> kids;
  for (int r = 1; r < gpus; ++r) {
    int pu[2], pr[2];
    if (pipe(pu) != 0 || pipe(pr) != 0) { std::perror("pipe"); return 1; }
    const pid_t pid = fork();
    if (pid < 0) { std::perror("fork"); return 1; }
    if (pid == 0) {
      rank = r;
      close(pu[1]); close(pr[0]);
      uid_in = pu[0]; res_out = pr[1];
      for (int q = 1; q < r; ++q) { close(uid_w[q]); close(res_r[q]); }
      break;
    }
    close(pu[0]); close(pr[1]);
    uid_w[r] = pu[1]; res_r[r] = pr[0];
    kids.push_back(pid);
  }

  RankResult mine{};
  try {
    if (gpus > 1) flexb200::ck(fx_set_device(rank));
    if (rank == 0) std::printf("-----------  %s  ---------------- start \n", argv[1]);
    flexb200::DataLoader data(argv[1], k);
    std::unique_ptr<flexb200::DataLoader> re;
    const flexb200::DataLoader* A = &data;
    if (ord != FX_ORDER_OVO) { re.reset(flexb200::reorder(data, ord)); A = re.get(); }
    if (rank == 0)
      std::printf("graph %s order %s: n = %zu nnz = %zu k = %d directed = %d gpus = %d\n", A->graph_name.c_str(),
                  A->vertex_order_abbr.c_str(), A->n, A->nnz, k, A->info.is_directed, gpus);
    const size_t n = A->n;
    std::vector<float> B(n * (size_t)k);
    flexb200::ck(fx_rand_B((int64_t)n, k, B.data()));
    std::vector<float> gold;
    fx_report rep{};
    if (gpus == 1) {
      std::vector<float> C(n * (size_t)k);
      if (flex_fmt && ord != FX_ORDER_OVO) {
        // Flex kernels of a reordered loader read shadow_b = B gathered by vo_mp (flexspmm_v9_permuteX, flex.cu:276) and write
        // C[voMp[row]] (flex.cu:994): C comes out in the ORIGINAL order, so the gold result is the original matrix times B
        std::vector<float> shadow(n * (size_t)k);
        for (size_t r = 0; r < n; ++r) std::memcpy(&shadow[r * k], &B[(size_t)A->vo_mp[r] * k], sizeof(float) * k);
        if (check) cpu_gold(data, B.data(), k, 0, n, gold);
        rep = flexb200::flex_spmm(*A, shadow.data(), C.data(), k, fmt, nullptr, tm, tn);
        if (check) {
          fx_report e{};
          flexb200::ck(fx_check(gold.data(), C.data(), (int64_t)n, k, data.rowPtr, &e));
          rep.errs_flex = e.errs_flex; rep.errs_tight = e.errs_tight; rep.errs_aspt_pct = e.errs_aspt_pct; rep.max_err = e.max_err;
        }
      } else {
        if (check) cpu_gold(*A, B.data(), k, 0, n, gold);
        rep = flexb200::flex_spmm(*A, B.data(), C.data(), k, fmt, check ? gold.data() : nullptr, tm, tn);
      }
    } else {
      // ---- row-panel sharded run ----
      std::vector<int64_t> cuts(gpus + 1);
      flexb200::ck(fx_panel_shards(A->handle(), gpus, cuts.data()));
      const int64_t lo = cuts[rank], hi = cuts[rank + 1];
      char uid[128] = {};
      if (rank == 0) {
        flexb200::ck(fx_comm_unique_id(uid));
        for (int r = 1; r < gpus; ++r)
          if (write(uid_w[r], uid, 128) != 128) throw std::runtime_error("pipe write (unique id)");
      } else if (read(uid_in, uid, 128) != 128) {
        throw std::runtime_error("pipe read (unique id)");
      }
      fx_comm* comm = nullptr;
      flexb200::ck(fx_comm_init(gpus, rank, uid, &comm));
      fx_build_opts o{};
      o.format = fmt; o.row_begin = (int32_t)lo; o.row_end = (int32_t)hi;
      fx_tiles* tiles = nullptr;
      flexb200::ck(fx_build(A->handle(), &o, &tiles, &mine.tPre_ms));
      int64_t slo = 0, shi = 0;
      flexb200::ck(fx_comm_slice(comm, (int64_t)n, &slo, &shi));
      std::vector<float> C((size_t)(hi - lo) * k);
      flexb200::ck(fx_spmm_sharded_host(tiles, comm, B.data() + (size_t)slo * k, C.data(), k, nullptr, nullptr));  // warm-up
      flexb200::ck(fx_spmm_sharded_host(tiles, comm, B.data() + (size_t)slo * k, C.data(), k, &mine.total_ms, &mine.tElap_ms));
      mine.rows = hi - lo;
      if (check && hi > lo) {
        cpu_gold(*A, B.data(), k, (size_t)lo, (size_t)hi, gold);
        fx_report e{};
        flexb200::ck(fx_check(gold.data(), C.data(), hi - lo, k, A->rowPtr + lo, &e));
        mine.errs_flex = e.errs_flex; mine.errs_tight = e.errs_tight; mine.max_err = e.max_err;
        mine.aspt_bad = (long long)(e.errs_aspt_pct / 100.0 * (double)(hi - lo) * k + 0.5);
      }
      fx_tiles_free(tiles);
      fx_comm_free(comm);
      mine.ok = 1;
      if (rank != 0) {
        if (write(res_out, &mine, sizeof(mine)) != (ssize_t)sizeof(mine)) return 1;
        return 0;
      }
      RankResult all = mine;
      for (int r = 1; r < gpus; ++r) {
        RankResult x{};
        if (read(res_r[r], &x, sizeof(x)) != (ssize_t)sizeof(x) || !x.ok) throw std::runtime_error("a rank failed");
        all.tPre_ms = std::max(all.tPre_ms, x.tPre_ms); all.total_ms = std::max(all.total_ms, x.total_ms);
        all.tElap_ms = std::max(all.tElap_ms, x.tElap_ms);
        all.errs_flex += x.errs_flex; all.errs_tight += x.errs_tight; all.aspt_bad += x.aspt_bad; all.rows += x.rows;
        all.max_err = std::max(all.max_err, x.max_err);
      }
      rep.tPre_ms = all.tPre_ms; rep.tElap_ms = all.tElap_ms;
      rep.gflops = all.tElap_ms > 0 ? 2.0 * A->nnz * k / (all.tElap_ms * 1e-3) / 1e9 : 0;
      rep.tpre_over_telap = all.tElap_ms > 0 ? all.tPre_ms / all.tElap_ms : 0;
      rep.errs_flex = all.errs_flex; rep.errs_tight = all.errs_tight; rep.max_err = all.max_err;
      rep.errs_aspt_pct = all.rows > 0 ? 100.0 * (double)all.aspt_bad / ((double)all.rows * k) : 0;
      std::printf("end to end (slice H2D + all-gather + SpMM + D2H, max over ranks): %f ms\n", all.total_ms);
    }
    std::printf("tPre: %f ms\ntElap: %f ms\n", rep.tPre_ms, rep.tElap_ms);
    std::printf("GFLOPS: %f\n", rep.gflops);                    // aspt/sspmm_128.cu:1406
    std::printf("t_pre/t_exe: %f\n", rep.tpre_over_telap);      // :1408
    if (check)
      std::printf("errs: %lld (resCheck)  %f %% (ASpT validator)  %lld (1e-5 row-normwise)  max_err %g\n", (long long)rep.errs_flex,
                  rep.errs_aspt_pct, (long long)rep.errs_tight, rep.max_err);
    std::printf("-----------  %s  ----------------- end \n", argv[1]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "rank %d: %s\n", rank, e.what());
    if (rank != 0 && res_out >= 0) { mine.ok = 0; if (write(res_out, &mine, sizeof(mine)) < 0) {} }
    return 1;
  }
  int rc = 0;
  for (pid_t p : kids) { int st = 0; waitpid(p, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1; }
  return rc;
}
