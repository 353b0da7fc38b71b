/* Host-only stand-in for <cuda_runtime.h>, used ONLY to compile the reference's host sources
 * (mat.cu, DataLoader.cu, order_*.cu ...) with g++ into oracle/_ref/libflexref.so so that the
 * reference's own tile builders / reordering code can be run on the CPU as the oracle's pin.
 * "Device" memory is host memory.  Test infrastructure; never part of the product. */
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <stdexcept>
#include <string>
#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__
#define __forceinline__ inline
#define __launch_bounds__(...)
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct cudaDeviceProp { int multiProcessorCount; char name[256]; };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
extern "C" int orc_ref_n_sm;  /* set by the driver: the SM count the builders see */
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n, size_t = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) { std::memcpy(&sym, s, n); return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = orc_ref_n_sm; p->name[0] = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "stub"; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; }; struct float4 { float x, y, z, w; }; struct int2 { int x, y; }; struct int4 { int x, y, z, w; };
