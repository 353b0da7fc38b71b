// Standalone check of the tcgen05 window kernel: nvcc -gencode arch=compute_100a,code=sm_100a scripts/tc_test.cu -o scripts/tc_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define FX_TC_DEBUG
#include "../flex_b200/csrc/fx_tc_kernel.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N>
int run(int k) {
  const int n = 1024, npanel = 2, win0 = 300, win1 = 768;
  std::vector<unsigned> col; std::vector<float> val; std::vector<int> lo(npanel * 128), hi(npanel * 128);
  srand(7);
  std::vector<int> wins = {win0, win1};
  for (int p = 0; p < npanel; ++p)
    for (int r = 0; r < 128; ++r) {
      lo[p * 128 + r] = (int)col.size();
      for (int c = 0; c < 256; ++c)
        if ((getenv("TC_ONES") || rand() % 100 < 15) && wins[p] + c < n) { col.push_back(wins[p] + c); val.push_back(getenv("TC_ONES") ? 1.0f : (float)rand() / RAND_MAX * 2 - 1); }
      hi[p * 128 + r] = (int)col.size();
    }
  std::vector<float> B((size_t)n * k);
  for (auto& x : B) x = getenv("TC_ONES") ? 1.0f : (float)rand() / RAND_MAX * 2 - 1;
  std::vector<int> panels = {0, 1};
  unsigned* dcol; float *dval, *dB, *dout; int *dlo, *dhi, *dp, *dw;
  CK(cudaMalloc(&dcol, col.size() * 4)); CK(cudaMalloc(&dval, val.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dout, (size_t)npanel * 128 * k * 4)); CK(cudaMalloc(&dlo, lo.size() * 4)); CK(cudaMalloc(&dhi, hi.size() * 4));
  CK(cudaMalloc(&dp, 8)); CK(cudaMalloc(&dw, 8));
  CK(cudaMemcpy(dcol, col.data(), col.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dval, val.data(), val.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dlo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dhi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, panels.data(), 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, wins.data(), 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, (size_t)npanel * 128 * k * 4));
  fxtc::TcArgs a{dcol, dval, dp, dw, dlo, dhi, dB, dout, k, getenv("TC_NOMMA") ? -n : n};
  { int sl = getenv("TC_MODE") ? atoi(getenv("TC_MODE")) : 0; CK(cudaMemcpyToSymbol(fxtc::g_tc_sleep, &sl, sizeof(int))); }
  const size_t smem = fxtc::tc_smem_bytes<N>();
  CK(cudaFuncSetAttribute(fxtc::k_spmm_tc<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fxtc::k_spmm_tc<N><<<dim3(npanel, k / N), 256, smem>>>(a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)npanel * 128 * k);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int p = 0; p < npanel; ++p)
    for (int r = 0; r < 128; ++r)
      for (int j = 0; j < k; ++j) {
        double s = 0;
        for (int e = lo[p * 128 + r]; e < hi[p * 128 + r]; ++e) s += (double)val[e] * B[(size_t)col[e] * k + j];
        double d = fabs(s - out[((size_t)p * 128 + r) * k + j]);
        if (!(d <= maxerr)) maxerr = d;
        if (fabs(s) > maxref) maxref = fabs(s);
      }
  if (getenv("TC_DEBUG")) {
    for (int r : {0, 1, 8, 33, 127}) {
      printf("row %d out:", r);
      for (int j = 0; j < 6; ++j) printf(" %11.4g", out[(size_t)r * k + j]);
      printf("  ref:");
      for (int j = 0; j < 6; ++j) { double s = 0; for (int e = lo[r]; e < hi[r]; ++e) s += (double)val[e] * B[(size_t)col[e] * k + j]; printf(" %9.4f", s); }
      printf("\n");
    }
  }
  printf("N=%d k=%d nnz=%zu max|err|=%.3e max|ref|=%.3f %s\n", N, k, col.size(), maxerr, maxref, maxerr < 1e-4 ? "OK" : "FAIL");
  return maxerr < 1e-4 ? 0 : 2;
}

int main() {
  int rc = run<128>(128);
  if (getenv("TC_DEBUG")) return rc;
  rc |= run<128>(256);
  rc |= run<64>(64);
  rc |= run<32>(32);
  return rc;
}
