"""tPre / tElap of the Flex formats (tile, seg, pillar) beside aspt and tcw on one synthetic shape:
python scripts/r2_flex_tpre.py <workload> <k>   (prints one line per format; GPU only)"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

import flex_b200 as fx
from flex_b200 import synth

name, k = sys.argv[1], int(sys.argv[2])
rp, c, v = synth.generate(name, device="cuda")
n, nnz = rp.numel() - 1, c.numel()
dl = fx.DataLoader.from_arrays(rp.cpu().numpy().astype(np.uint32), c.cpu().numpy().astype(np.uint32), v.cpu().numpy(), k, name + ".csv")
B = synth.dense_B(n, k, device="cuda")
C = torch.empty((n, k), device="cuda")
for fmt in ("pillar", "seg", "tile", "aspt", "tcw"):
    try:
        t0 = time.perf_counter()
        mat = fx.Mat(dl, fmt=fmt, tm=4, tn=4) if fmt in ("pillar", "seg", "tile") else fx.Mat(dl, fmt=fmt)
        wall = (time.perf_counter() - t0) * 1e3
        pre = [mat.tPre_ms] + [mat.rebuild() for _ in range(3)]
        for _ in range(3):
            mat.spmm(B.data_ptr(), C.data_ptr(), k)
        ts = [mat.spmm(B.data_ptr(), C.data_ptr(), k, timed=True) for _ in range(10)]
        print(f"{name} k={k} {fmt:7s} tPre first {pre[0]:8.3f} ms, rebuilds {min(pre[1:]):8.3f} ms (wall of first build {wall:8.1f}), tElap {np.median(ts):7.4f} ms, "
              f"GFLOP/s {2 * nnz * k / np.median(ts) / 1e6:9.1f}", flush=True)
        mat.free()
    except Exception as e:  # noqa: BLE001
        print(f"{name} k={k} {fmt}: {e}", flush=True)
