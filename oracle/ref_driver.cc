// ref_driver.cc -- TEST INFRASTRUCTURE ONLY.
// Thin command-line driver around the UNMODIFIED reference host sources (DataLoader.cu, mat.cu,
// order_*.cu, edgelist.cu, adjlist.cu, algo_bfs.cu, unitheap.cu, tools.cu), compiled by
// oracle/ref_build.sh with g++ against the stand-in headers in oracle/ref_stubs/ (host-only
// cuda_runtime.h: "device" memory is host memory).  It runs the reference's own CSV loader,
// reordering constructors and tile builders on the CPU and dumps their outputs so that tests can
// pin the oracle's C restatement (and through it the CUDA product) against the reference itself.
// A separate process per call: the reference asserts (abort) on inputs it does not support.
//
//   flexref load  <csv> <out>
//   flexref order <csv> <out> deg|rcm|gor|dfs|rbt  (DataLoaderDeg/Rcm/Gorder/DFS/Rabbit DataLoader.cu:324-857)
//   flexref rank  <csv> <out> deg|rcm|gor          (order_deg/order_rcm/complete_gorder)
//   flexref seg   <csv> <out> <tm>                 (Mat::csr2seg_Cmajor per panel, mat.cu:1192)
//   flexref diag  <csv> <out> <tm> <n_sm>          (Mat::csr2_DiagTiling mat.cu:680)
//   flexref tile  <csv> <out> <tm> <tn> R|C        (Mat::csr2flex_Rmajor/Cmajor mat.cu:1345,1438)
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "DataLoader.cuh"
#include "edgelist.cuh"
#include "mat.cuh"
#include "order_deg.cuh"
#include "order_gorder.cuh"
#include "order_rcm.cuh"

int orc_ref_n_sm = 148;
void cuSpmm(DataLoader&, Perfs&) {}  // flex.cu is not linked: the cuSPARSE baseline is not on this path

static FILE* g_out;
template <class T>
static void dump(const char* name, char dtype, const T* p, size_t n) {
  uint32_t l = (uint32_t)strlen(name);
  uint64_t cnt = n;
  fwrite(&l, 4, 1, g_out); fwrite(name, 1, l, g_out); fwrite(&dtype, 1, 1, g_out); fwrite(&cnt, 8, 1, g_out);
  if (n) fwrite(p, sizeof(T), n, g_out);
}
template <class T> static void dumpv(const char* name, char dtype, const std::vector<T>& v) { dump(name, dtype, v.data(), v.size()); }
static void dumps(const char* name, long long v) { dump(name, 'q', &v, 1); }
static void dumpd(const char* name, double v) { dump(name, 'd', &v, 1); }

static void dump_loader(const DataLoader& d) {
  dumpv("rowPtr", 'I', d.rowPtr); dumpv("col", 'I', d.col); dumpv("vals", 'f', d.vals); dumpv("vo_mp", 'i', d.vo_mp);
  dumps("m", d.m); dumps("nnz", d.nnz); dumps("c", d.c); dumps("uni_nb", d.uni_nb);
  dumps("is_directed", d.is_directed); dumps("n_edges_one_way", d.n_edges_one_way);
  dumps("n_edges_asymmetric", d.n_edges_asymmetric); dumps("n_nodes_z_out", d.n_nodes_z_out);
  dumps("n_nodes_z_in", d.n_nodes_z_in); dumps("n_nodes_z_deg", d.n_nodes_z_deg);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: flexref <cmd> <csv> <out> ...\n"); return 2; }
  std::string cmd = argv[1];
  g_out = fopen(argv[3], "wb");
  if (!g_out) return 3;
  DataLoader dl(argv[2], 4);
  if (cmd == "load") {
    dump_loader(dl);
    dumpv("cpuX", 'f', dl.cpuX);
  } else if (cmd == "order") {
    std::string o = argv[4];
    if (o == "deg") { DataLoaderDeg r(dl); dump_loader(r); }
    else if (o == "rcm") { DataLoaderRcm r(dl); dump_loader(r); }
    else if (o == "dfs") { DataLoaderDFS r(dl); dump_loader(r); }
    else if (o == "rbt") { DataLoaderRabbit r(dl); dump_loader(r); }
    else { DataLoaderGorder r(dl); dump_loader(r); }
  } else if (cmd == "rank") {
    std::string o = argv[4];
    Edgelist h(dl);
    if (o == "deg") { auto r = order_deg(h, true); dump("rank", 'Q', r.data(), r.size()); }
    else if (o == "rcm") { auto r = order_rcm(h); dump("rank", 'Q', r.data(), r.size()); }
    else { auto r = complete_gorder(h, 3); dump("rank", 'I', r.data(), r.size()); }
  } else if (cmd == "seg") {
    int tm = atoi(argv[4]);
    Mat mat(dl, tm, 4);
    std::unordered_map<int, std::unordered_set<int>> dup;
    int nnz_rowPtr = 0;
    mat.alpha_rowPtr.push_back(0);
    mat.alpha_pillar_rowPtr.push_back(0);
    std::vector<int> per_panel;
    int tileRows = (mat.m + tm - 1) / tm;
    for (int i = 0; i < tileRows; ++i) per_panel.push_back(mat.csr2seg_Cmajor(i, dup, nnz_rowPtr));
    dumpv("alpha_rowPtr", 'I', mat.alpha_rowPtr); dumpv("alpha_colIdx", 'I', mat.alpha_colIdx);
    dumpv("alpha_vals", 'f', mat.alpha_vals); dumpv("alpha_pillar_rowPtr", 'I', mat.alpha_pillar_rowPtr);
    dumpv("segVoMap", 'I', mat.segVoMap); dumpv("segs_per_panel", 'i', per_panel); dumps("nnz_rowPtr", nnz_rowPtr);
  } else if (cmd == "diag") {
    int tm = atoi(argv[4]);
    orc_ref_n_sm = atoi(argv[5]);
    Mat mat(dl, tm, 4);
    mat.csr2_DiagTiling();
    dumpv("alpha_rowPtr", 'I', mat.alpha_rowPtr); dumpv("alpha_colIdx", 'I', mat.alpha_colIdx);
    dumpv("alpha_vals", 'f', mat.alpha_vals); dumpv("alpha_pillar_rowPtr", 'I', mat.alpha_pillar_rowPtr);
    dumpv("alpha_pillarIdx", 'I', mat.alpha_pillarIdx); dumpv("segVoMap", 'I', mat.segVoMap);
    dumps("n_segs", mat.n_segs); dumps("sms", mat.sms); dumpd("empty_wp_p", mat.empty_wp_p); dumpd("band_nz_p", mat.band_nz_p);
  } else if (cmd == "tile") {
    int tm = atoi(argv[4]), tn = atoi(argv[5]);
    bool cmaj = argv[6][0] == 'C';
    Mat mat(dl, tm, tn);
    int tileRows = (mat.m + tm - 1) / tm;
    for (int i = 0; i < tileRows; ++i) { if (cmaj) mat.csr2flex_Cmajor(i); else mat.csr2flex_Rmajor(i); }
    dumpv("tileRowPtr", 'I', mat.tileRowPtr); dumpv("tileNnz", 'I', mat.tileNnz); dumpv("nnzTile", 'i', mat.nnzTile);
    dumpv("bitMap", 'i', mat.bitMap); dumpv("tileColIdx", 'I', mat.tileColIdx); dumpv("rcOffset", 'i', mat.rcOffset);
    dumpv("newVals", 'f', mat.newVals);
  } else {
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
  }
  fclose(g_out);
  return 0;
}
