"""H2D/D2H rate of strided (2D) copies vs one contiguous copy, pinned host memory (for the column-chunk pipeline of fx_spmm_host)."""
import torch, time
from cuda.bindings import runtime as rt
n, k = 232965, 128
h = torch.empty((n, k), dtype=torch.float32).pin_memory(); h.fill_(1.0)
d = torch.empty((n, k), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost
print("1D H2D %.3f ms" % timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * k * 4, H2D, s.cuda_stream)))
print("1D D2H %.3f ms" % timed(lambda: rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), n * k * 4, D2H, s.cuda_stream)))
for w in (32, 64):
    def h2d():
        for c0 in range(0, k, w):
            rt.cudaMemcpy2DAsync(d.data_ptr() + c0 * 4, k * 4, h.data_ptr() + c0 * 4, k * 4, w * 4, n, H2D, s.cuda_stream)
    def d2h():
        for c0 in range(0, k, w):
            rt.cudaMemcpy2DAsync(h.data_ptr() + c0 * 4, k * 4, d.data_ptr() + c0 * 4, k * 4, w * 4, n, D2H, s.cuda_stream)
    print("2D width %d floats: H2D all chunks %.3f ms, D2H all chunks %.3f ms" % (w, timed(h2d), timed(d2h)))
# both directions at once (two streams)
s2 = torch.cuda.Stream()
def duplex():
    rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * k * 4, H2D, s.cuda_stream)
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), n * k * 4, D2H, s2.cuda_stream)
h2 = torch.empty((n, k), dtype=torch.float32).pin_memory()
def duplex2():
    rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n * k * 4, H2D, s.cuda_stream)
    rt.cudaMemcpyAsync(h2.data_ptr(), d.data_ptr(), n * k * 4, D2H, s2.cuda_stream)
print("duplex H2D+D2H concurrently %.3f ms" % timed(duplex2))
