"""Times every shard of a G-way row-panel split on ONE GPU (what each rank of a G-GPU run would do)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flex_b200 as fx
from flex_b200 import synth
from flex_b200.shard import panel_shards
w, G, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 128
fmt = sys.argv[4] if len(sys.argv) > 4 else "tcw"
only_first = int(os.environ.get("ONLY_FIRST", "0"))
rp, c, v = synth.generate(w, device="cuda")
n, nnz = rp.numel() - 1, c.numel()
dl = fx.DataLoader.from_device(n, nnz, rp.int().data_ptr(), c.int().data_ptr(), v.data_ptr(), k, w + ".csv")
rph = rp.cpu().numpy()
B = synth.dense_B(n, k, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for si, (lo, hi) in enumerate(panel_shards(rph, G)):
    if only_first and si >= only_first:
        break
    mat = fx.Mat(dl, fmt=fmt, row_begin=lo, row_end=hi)
    C = torch.empty((hi - lo, k), device="cuda")
    for _ in range(5):
        mat.spmm(B.data_ptr(), C.data_ptr(), k, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        mat.spmm(B.data_ptr(), C.data_ptr(), k, stream=st)
    e1.record(); e1.synchronize()
    print(f"rows [{lo},{hi}) nnz={int(rph[hi]-rph[lo])} ms={e0.elapsed_time(e1)/20:.4f} tPre={mat.tPre_ms:.3f}", flush=True)
    mat.free()
