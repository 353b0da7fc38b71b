// ref_aspt_dump.cu -- TEST INFRASTRUCTURE: pins the ASpT tile format to a RUN OF THE REFERENCE ITSELF.
//
// This translation unit #includes the reference's own, unmodified aspt/sspmm_128.cu (or sspmm_32.cu) where it lies under
// /root/reference -- nothing is copied -- with its main() renamed, and traces its cudaMalloc calls by the NAME of the
// pointer argument.  After the reference has run (ready2 + process: its pre-processing, kernels and validator) the device
// arrays of its tile format are still allocated (it never frees them), so they are read back and written to
// <out>.bin / <out>.json: mcsr_chk, mcsr_cnt, mcsr_e, mcsr_list, baddr, saddr, csr_e, csr_ev, special, special2 and the host
// scalars nr, nr0, ne, npanel, num_dense, avg, vari, special_p.
// Built by oracle/ref_build.sh into oracle/_ref/sspmm_{128,32}_dump (git-ignored; travels to the GPU box); run there by
// tests/tools/pin_aspt.sh; tests/golden/make_aspt_golden.py turns the dumps into the committed fixtures that
// tests/test_ref_pin.py::test_aspt_pinned compares with the oracle (order-independent projections: the reference's slot
// depths come from atomicAdd arrival order and its nz order from an unstable sort, aspt/sspmm_128.cu:915,957, bb_exch.h:24).
//
//   nvcc ... -DREF_SRC='"/root/reference/aspt/sspmm_128.cu"' oracle/ref_aspt_dump.cu -o oracle/_ref/sspmm_128_dump
//   ./sspmm_128_dump <in.csv> <k> <out-prefix>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>

static std::map<std::string, std::pair<void*, size_t>>& fx_trace() {
  static std::map<std::string, std::pair<void*, size_t>> t;
  return t;
}
static cudaError_t fx_traced_malloc(const char* expr, void** p, size_t bytes) {
  const cudaError_t e = cudaMalloc(p, bytes);
  const char* amp = std::strrchr(expr, '&');  // "(void **) &_mcsr_cnt" -> "_mcsr_cnt"
  std::string name = amp ? amp + 1 : expr;
  while (!name.empty() && name.back() == ' ') name.pop_back();
  fx_trace()[name] = {*p, bytes};
  return e;
}
#define cudaMalloc(p, bytes) fx_traced_malloc(#p, (void**)(p), (bytes))
// The reference's `int main(int argc, char **argv)` has no return statement: renamed as it stands it would be a non-void
// function flowing off its end (undefined behaviour; g++ -O3 drops the epilogue).  The macro turns the definition into
//   int ref_main(int argc, char **argv) { ...; ref_main_body(argc, argv); return 0; }  void ref_main_body(int argc, char **argv) {<the reference's body>}
#define main(...) ref_main(__VA_ARGS__) { void ref_main_body(int, char**); ref_main_body(argc, argv); return 0; } void ref_main_body(__VA_ARGS__)
#include REF_SRC
#undef main
#undef cudaMalloc

static void dump(FILE* bin, FILE* js, const char* name, size_t count, size_t elem, bool last = false) {
  auto it = fx_trace().find(name);
  std::vector<char> h(count * elem);
  if (it == fx_trace().end() || count == 0) count = 0;
  else if (count * elem > it->second.second) { std::fprintf(stderr, "dump: %s smaller than expected\n", name); count = 0; }
  else cudaMemcpy(h.data(), it->second.first, count * elem, cudaMemcpyDeviceToHost);
  const long off = std::ftell(bin);
  if (count) std::fwrite(h.data(), elem, count, bin);
  std::fprintf(js, "  \"%s\": {\"offset\": %ld, \"count\": %zu, \"elem\": %zu}%s\n", name + 1, off, count, elem, last ? "" : ",");
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s <in.csv> <k> <out-prefix>\n", argv[0]); return 2; }
  ref_main(argc, argv);
  cudaDeviceSynchronize();
  const std::string pre = argv[3];
  FILE* bin = std::fopen((pre + ".bin").c_str(), "wb");
  FILE* js = std::fopen((pre + ".json").c_str(), "w");
  if (!bin || !js) return 1;
  std::fprintf(js, "{\n  \"nr\": %d, \"nr0\": %d, \"nc\": %d, \"ne\": %d, \"npanel\": %d, \"num_dense\": %d, \"special_p\": %d,\n", nr, nr0, nc, ne,
               npanel, num_dense, special_p);
  std::fprintf(js, "  \"avg\": %.17g, \"vari\": %.17g, \"BH\": %d, \"BW\": %d, \"k\": %d,\n", avg, vari, (int)BH, (int)BW, sc);
  const size_t nd = (size_t)num_dense, np = (size_t)npanel;
  dump(bin, js, "_mcsr_chk", np, 4);
  dump(bin, js, "_mcsr_cnt", np + 1, 4);
  // with no dense tile anywhere the reference aliases _mcsr_e = _csr_v and _csr_e = _csr_e0 (:1227-1229)
  dump(bin, js, nd ? "_mcsr_e" : "_csr_v", nd ? (size_t)BH * (nd + np) + 1 : (size_t)nr + 1, 4);
  dump(bin, js, "_mcsr_list", nd * BW, 4);
  dump(bin, js, "_baddr", nd ? nd + np : 0, 4);
  dump(bin, js, "_saddr", nd ? nd + np : 0, 4);
  dump(bin, js, nd ? "_csr_e" : "_csr_e0", (size_t)ne, 4);
  dump(bin, js, nd ? "_csr_ev" : "_csr_ev0", (size_t)ne, 4);
  dump(bin, js, "_special", vari >= 200 ? (size_t)special_p : 0, 4);
  dump(bin, js, "_special2", vari >= 200 ? (size_t)special_p : 0, 4, true);
  std::fprintf(js, "}\n");
  std::fclose(bin); std::fclose(js);
  return 0;
}
