"""PCIe probe for the end-to-end path: cudaMemcpy2D column chunks vs a zero-copy kernel (the SMs read pinned host memory / write
it directly) for 128- and 256-byte row segments, alone and with both directions at once.  torch only as an array library."""
import time
import torch
from cuda.bindings import runtime as rt

n, k = 232965, 128
h = torch.empty((n, k), dtype=torch.float32).pin_memory(); h.fill_(1.0)
h2 = torch.empty((n, k), dtype=torch.float32).pin_memory()
d = torch.empty((n, k), dtype=torch.float32, device="cuda")
d2 = torch.randn((n, k), dtype=torch.float32, device="cuda")
s, s2 = torch.cuda.Stream(), torch.cuda.Stream()
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


mb = n * k * 4 / 1e6
t = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * k * 4, H2D, s.cuda_stream)); print("1D H2D %.3f ms  %.1f GB/s" % (t, mb / t))
t = timed(lambda: rt.cudaMemcpyAsync(h2.data_ptr(), d2.data_ptr(), n * k * 4, D2H, s.cuda_stream)); print("1D D2H %.3f ms  %.1f GB/s" % (t, mb / t))
def duplex():
    rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * k * 4, H2D, s.cuda_stream)
    rt.cudaMemcpyAsync(h2.data_ptr(), d2.data_ptr(), n * k * 4, D2H, s2.cuda_stream)
t = timed(duplex); print("1D duplex %.3f ms  %.1f GB/s each way" % (t, mb / t))
for w in (32, 64):
    def h2d():
        for c0 in range(0, k, w):
            rt.cudaMemcpy2DAsync(d.data_ptr() + c0 * 4, k * 4, h.data_ptr() + c0 * 4, k * 4, w * 4, n, H2D, s.cuda_stream)
    def d2h():
        for c0 in range(0, k, w):
            rt.cudaMemcpy2DAsync(h2.data_ptr() + c0 * 4, k * 4, d2.data_ptr() + c0 * 4, k * 4, w * 4, n, D2H, s2.cuda_stream)
    print("memcpy2D width %3d B: H2D %.3f ms  D2H %.3f ms  both at once %.3f ms" % (w * 4, timed(h2d), timed(d2h), timed(lambda: (h2d(), d2h()))))
# zero-copy through torch indexing kernels: device tensor <- pinned host tensor slice (torch issues a copy kernel? no: it uses memcpy2D)
# so use a tiny CUDA kernel compiled with NVRTC
from cuda.bindings import nvrtc, driver
src = r'''
extern "C" __global__ void cp_cols(const float4* __restrict__ src, float4* __restrict__ dst, long long n, int k4, int c0, int w4) {
  const long long total = n * w4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / w4; const int c = (int)(i % w4);
    dst[r * k4 + c0 + c] = src[r * k4 + c0 + c];
  }
}'''
err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"cp.cu", 0, [], [])
opts = [b"--gpu-architecture=sm_100"]
nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
err, sz = nvrtc.nvrtcGetCUBINSize(prog)
cubin = b" " * sz
nvrtc.nvrtcGetCUBIN(prog, cubin)
err, mod = driver.cuModuleLoadData(cubin)
err, fn = driver.cuModuleGetFunction(mod, b"cp_cols")
import numpy as np
def launch(srcp, dstp, c0, w, stream, blocks):
    args = [np.array([srcp], np.uint64), np.array([dstp], np.uint64), np.array([n], np.int64), np.array([k // 4], np.int32),
            np.array([c0 // 4], np.int32), np.array([w // 4], np.int32)]
    ptrs = np.array([a.ctypes.data for a in args], np.uint64)
    driver.cuLaunchKernel(fn, blocks, 1, 1, 256, 1, 1, 0, stream.cuda_stream, ptrs.ctypes.data, 0)
for blocks in (148, 592, 2368):
    for w in (32, 64, 128):
        def h2d():
            for c0 in range(0, k, w):
                launch(h.data_ptr(), d.data_ptr(), c0, w, s, blocks)
        def d2h():
            for c0 in range(0, k, w):
                launch(d2.data_ptr(), h2.data_ptr(), c0, w, s2, blocks)
        print("zero-copy kernel %4d CTAs width %3d B: H2D %.3f ms  D2H %.3f ms  both at once %.3f ms" % (blocks, w * 4, timed(h2d), timed(d2h), timed(lambda: (h2d(), d2h()))))
