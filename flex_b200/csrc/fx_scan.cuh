// fx_scan.cuh -- exclusive prefix sum of an int array by ONE CTA of 1024 threads (the builders' scans are short
// and sit between dependent kernels, so a single launch without inter-CTA traffic is the cheapest form).
// A thread owns IPT consecutive elements per pass, fetched as 16-byte words when the pointers allow, so a pass
// covers 1024*IPT elements for three block barriers (the first version took one element per thread per pass:
// 228 passes and 0.21 ms for the 233 k rows of Reddit-shape).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

namespace fx {

// The two row-length scans of a build (one entry per row: 233 k on Reddit-shape, 61 us each through one CTA) go through
// cub's single-pass device scan instead (~10 us): out[0] = 0, out[i + 1] = in[0] + ... + in[i], out has n + 1 elements.
inline size_t device_scan_tmp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n);
  return bytes;
}
inline cudaError_t device_exclusive_scan(void* tmp, size_t tmp_bytes, const int* in, int n, int* out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int), s);
  if (e != cudaSuccess || n <= 0) return e;
  return cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out + 1, n, s);
}

// out[i] = in[0] + ... + in[i-1] for i in [0, n]; out has n + 1 elements.  blockDim.x must be 1024.
__device__ __forceinline__ void cta_exclusive_scan(const int* __restrict__ in, int n, int* __restrict__ out) {
  constexpr int IPT = 8;
  __shared__ int scan_warp_sum[33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  int carry = 0;
  for (int base = 0; base < n; base += 1024 * IPT) {
    const int i0 = base + tid * IPT;
    int v[IPT];
    if (vec_ok && i0 + IPT <= n) {
      const int4 a = *reinterpret_cast<const int4*>(in + i0), b = *reinterpret_cast<const int4*>(in + i0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < IPT; ++j) v[j] = i0 + j < n ? in[i0 + j] : 0;
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) s += v[j];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) scan_warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int ws = scan_warp_sum[lane];
      int wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      scan_warp_sum[lane] = wi - ws;
      if (lane == 31) scan_warp_sum[32] = wi;
    }
    __syncthreads();
    int run = carry + scan_warp_sum[warp] + inc - s;  // exclusive prefix of this thread's first element
    carry += scan_warp_sum[32];
    if (vec_ok && i0 + IPT <= n) {
      int4 a, b;
      a.x = run; run += v[0]; a.y = run; run += v[1]; a.z = run; run += v[2]; a.w = run; run += v[3];
      b.x = run; run += v[4]; b.y = run; run += v[5]; b.z = run; run += v[6]; b.w = run;
      *reinterpret_cast<int4*>(out + i0) = a;
      *reinterpret_cast<int4*>(out + i0 + 4) = b;
    } else {
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        if (i0 + j < n) out[i0 + j] = run;
        run += v[j];
      }
    }
    __syncthreads();  // scan_warp_sum is rewritten by the next pass
  }
  if (tid == 0) out[n] = carry;
}

}  // namespace fx
