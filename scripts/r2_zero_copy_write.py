"""Probe: can the SMs write C straight into pinned host memory at PCIe rate?  k_permute_rows (float4 gather/scatter of n x k
rows) with its output pointer in pinned host memory, alone and beside an H2D copy of the same size."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import flex_b200 as fx

n, k = 232965, 128
rp = torch.arange(n + 1, dtype=torch.int32, device="cuda")
c = torch.arange(n, dtype=torch.int32, device="cuda")
v = torch.ones(n, device="cuda")
dl = fx.DataLoader.from_device(n, n, rp.data_ptr(), c.data_ptr(), v.data_ptr(), k, "eye.csv")
src = torch.randn((n, k), device="cuda")
dst_dev = torch.empty((n, k), device="cuda")
dst_host = torch.empty((n, k), dtype=torch.float32).pin_memory()
up_host = torch.randn((n, k), dtype=torch.float32).pin_memory()
up_dev = torch.empty((n, k), device="cuda")
s2 = torch.cuda.Stream()
mb = n * k * 4 / 1e6


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


t = timed(lambda: dl.permute_rows(src.data_ptr(), dst_dev.data_ptr(), k)); print("kernel -> device memory %.3f ms (%.0f GB/s)" % (t, mb / t))
t = timed(lambda: dl.permute_rows(src.data_ptr(), dst_host.data_ptr(), k)); print("kernel -> pinned host memory %.3f ms (%.1f GB/s)" % (t, mb / t))
assert torch.equal(dst_host, src.cpu())
t = timed(lambda: dst_host.copy_(src, non_blocking=True)); print("cudaMemcpy D2H %.3f ms (%.1f GB/s)" % (t, mb / t))
def both():
    with torch.cuda.stream(s2):
        up_dev.copy_(up_host, non_blocking=True)
    dl.permute_rows(src.data_ptr(), dst_host.data_ptr(), k)
t = timed(both); print("kernel -> host beside an H2D copy of the same size %.3f ms (%.1f GB/s each way)" % (t, mb / t))
def both2():
    with torch.cuda.stream(s2):
        up_dev.copy_(up_host, non_blocking=True)
    dst_host.copy_(src, non_blocking=True)
t = timed(both2); print("memcpy D2H beside an H2D copy %.3f ms (%.1f GB/s each way)" % (t, mb / t))
