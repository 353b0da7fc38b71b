#pragma once
/* TEST INFRASTRUCTURE -- stand-in for gp/cuda-gpuinfo.h of the LSU "gp" library, which the reference includes (common.h:4) but
 * does not ship (Makefile:1 GP_ROOT ?= ../../gp; no submodule, no version).  Written from the reference's 60-odd use sites
 * (flex.cu:4127-4145,4334-4366,4679-4697,4935-4940,5044,5143): just enough for the reference's own sources to build
 * unmodified for sm_100 with nvcc (oracle/ref_flex_build.sh) so that its Flex v36 kernel can be timed on the B200 box. */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <set>
#include <string>
#include <vector>

#define CE(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t ce_e_ = (call);                                                                           \
    if (ce_e_ != cudaSuccess) {                                                                           \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(ce_e_), __FILE__, __LINE__);   \
      std::exit(1);                                                                                       \
    }                                                                                                     \
  } while (0)

typedef void (*GPU_Info_Func)();

struct Kernel_Info {
  GPU_Info_Func func_ptr = nullptr;
  const char* name = "";
  cudaFuncAttributes cfa{};
  int get_max_active_blocks_per_mp(int threads, size_t dyn_smem = 0) const {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, (const void*)func_ptr, threads, dyn_smem);
    return n;
  }
};

struct GPU_Info {
  cudaDeviceProp cuda_prop{};
  double clock_freq_hz = 0;
  std::deque<Kernel_Info> kernels;  // stable references
  void get_gpu_info(int dev) {
    CE(cudaGetDeviceProperties(&cuda_prop, dev));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    clock_freq_hz = khz * 1e3;
  }
  int get_fp32_per_sm() const { return 128; }
  Kernel_Info& get_info(GPU_Info_Func f) {
    for (auto& k : kernels) if (k.func_ptr == f) return k;
    kernels.emplace_back();
    kernels.back().func_ptr = f;
    cudaFuncGetAttributes(&kernels.back().cfa, (const void*)f);
    return kernels.back();
  }
  template <class K> Kernel_Info& get_info_named(K k, const char* name) {
    Kernel_Info& ki = get_info((GPU_Info_Func)k);
    ki.name = name;
    return ki;
  }
};
#define GET_INFO(k) get_info_named(k, #k)

inline void gpu_info_print() {
  int n = 0;
  cudaGetDeviceCount(&n);
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p{};
    cudaGetDeviceProperties(&p, d);
    std::printf("GPU %d: %s, %d SMs, cc %d.%d\n", d, p.name, p.multiProcessorCount, p.major, p.minor);
  }
}
inline int gpu_choose_index() { return 0; }
