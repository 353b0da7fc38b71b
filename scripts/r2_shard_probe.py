"""One 1/G row-panel shard on ONE GPU (what a rank of a G-GPU run does): step time under FLEX_SIDE / FLEX_SPLIT settings and
the per-kernel times.  python scripts/r2_shard_probe.py <workload> <G> <k> [shard index]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import flex_b200 as fx
from flex_b200 import synth
from flex_b200.shard import panel_shards

w, G, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
si = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rp, c, v = synth.generate(w, device="cuda")
n, nnz = rp.numel() - 1, c.numel()
rp32, c32 = rp.int(), c.int()
dl = fx.DataLoader.from_device(n, nnz, rp32.data_ptr(), c32.data_ptr(), v.data_ptr(), k, w + ".csv")
rph = rp.cpu().numpy()
B = synth.dense_B(n, k, device="cuda")
lo, hi = panel_shards(rph, G)[si]
mat = fx.Mat(dl, fmt="tcw", row_begin=lo, row_end=hi)
C = torch.empty((hi - lo, k), device="cuda")
st = torch.cuda.current_stream().cuda_stream
flops = 2.0 * int(rph[hi] - rph[lo]) * k
print(f"{w} shard {si}/{G} rows [{lo},{hi}) nnz {int(rph[hi] - rph[lo])} info {mat.tcw_info()}", flush=True)
print("kernel times", {a: round(b, 4) for a, b in mat.kernel_times(B.data_ptr(), C.data_ptr(), k, stream=st).items()}, flush=True)


def run(label):
    for _ in range(5):
        mat.spmm(B.data_ptr(), C.data_ptr(), k, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(50):
            mat.spmm(B.data_ptr(), C.data_ptr(), k, stream=st)
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    print(f"{label:28s} {best:.4f} ms  {flops / best / 1e9:8.2f} TFLOP/s-equivalent of the shard", flush=True)
    return C.clone()


ref = None
for split in ("", "1", "2", "3"):
    if split:
        os.environ["FLEX_SPLIT"] = split
    else:
        os.environ.pop("FLEX_SPLIT", None)
    out = run(f"FLEX_SIDE={os.environ.get('FLEX_SIDE', 'auto')} split={split or 'auto'}")
    if ref is None:
        ref = out
    assert torch.equal(out, ref)
