#pragma once
/* TEST INFRASTRUCTURE -- stand-in for nperf.h (common.h:6), the CUPTI counter sampler of the LSU "gp" library.  No counters:
 * every metric reads 0.  The run protocol `for ( NPerf_data_reset(); NPerf_need_run_get(); ) kernel<<<>>>();` executes the
 * kernel once, bracketed by CUDA events, and NPerf_kernel_et_get() returns that time in seconds.  (The reference's own
 * cudaEvent time around its untimed-by-NPerf launch, flex.cu:5051-5068, does not depend on this header.) */
#include <cuda_runtime.h>
struct NPerf_Stub_State { int runs = 0; cudaEvent_t e0 = nullptr, e1 = nullptr; double et = 0; };
inline NPerf_Stub_State& nperf_stub() { static NPerf_Stub_State s; return s; }
inline void NPerf_init(bool = false) {}
inline void NPerf_metric_collect(const char*) {}
inline double NPerf_metric_value_get(const char*) { return 0.0; }
inline void NPerf_metrics_off() {}
inline void NPerf_metrics_on() {}
inline void NPerf_data_reset() { nperf_stub().runs = 0; }
inline bool NPerf_need_run_get() {
  NPerf_Stub_State& s = nperf_stub();
  if (!s.e0) { cudaEventCreate(&s.e0); cudaEventCreate(&s.e1); }
  if (s.runs == 0) { s.runs = 1; cudaEventRecord(s.e0, 0); return true; }
  cudaEventRecord(s.e1, 0);
  cudaEventSynchronize(s.e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, s.e0, s.e1);
  s.et = ms * 1e-3;
  return false;
}
inline double NPerf_kernel_et_get() { return nperf_stub().et; }
