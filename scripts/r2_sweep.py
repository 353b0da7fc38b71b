"""Round-2 sweep: one synthetic graph generated once on the GPU, many tensor-window plans timed back to back.

    python scripts/r2_sweep.py reddit 128 "4:256:224:1024" "4:1024:96:256" ...     (T:W:chunk_cost:min_gain[:min_total])

Prints one line per plan: tElap (device events, 20 steps after 5 warm-ups), window share, listed columns, tPre, and the
largest row-normwise difference from the FX_FMT_ASPT result of the same process (all-FMA fp32), as a sanity check of the
tensor path -- the parity tests proper are under tests/.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import flex_b200 as fx
from flex_b200 import synth


def timeit(mat, B, Cd, k, steps=int(os.environ.get('STEPS', 20)), warm=int(os.environ.get('WARM', 5))):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(warm):
        mat.spmm(B.data_ptr(), Cd.data_ptr(), k, stream=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        mat.spmm(B.data_ptr(), Cd.data_ptr(), k, stream=st)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    name, k = sys.argv[1], int(sys.argv[2])
    plans = sys.argv[3:]
    dev = torch.device("cuda", 0)
    rp, c, v = synth.generate(name, device=dev)
    n, nnz = rp.numel() - 1, c.numel()
    rp32, c32 = rp.to(torch.int32), c.to(torch.int32)
    dl = fx.DataLoader.from_device(n, nnz, rp32.data_ptr(), c32.data_ptr(), v.data_ptr(), k, name + ".csv")
    if os.environ.get("ORDER"):  # ORDER=deg|rcm|gor|rbt: the same sweep on the reordered matrix
        import numpy as np
        dl = fx.DataLoader.from_arrays(rp.cpu().numpy().astype(np.uint32), c.cpu().numpy().astype(np.uint32), v.cpu().numpy(), k, name + ".csv")
        dl = dl.reorder({"deg": fx.FX_ORDER_DEG, "rcm": fx.FX_ORDER_RCM, "gor": fx.FX_ORDER_GOR, "rbt": fx.FX_ORDER_RBT}[os.environ["ORDER"]])
    B = synth.dense_B(n, k, device=dev)
    Cref = torch.empty((n, k), dtype=torch.float32, device=dev)
    Cd = torch.empty((n, k), dtype=torch.float32, device=dev)
    m0 = fx.Mat(dl, fmt="aspt")
    t0 = timeit(m0, B, Cref, k)
    print(f"{name} k={k} aspt: {t0:.4f} ms  tPre {m0.tPre_ms:.2f}", flush=True)
    rownorm = Cref.abs().amax(dim=1).clamp_min(1.0)
    m0.free()
    for pl in plans:
        f = [int(x) for x in pl.split(":")]
        T, W, cost, gain = f[:4]
        mt = f[4] if len(f) > 4 else 0
        mat = fx.Mat(dl, fmt="tcw", tc_threshold=T, tc_width=W, tc_chunk_cost=cost, tc_min_gain=gain, tc_min_total=mt)
        tpre = min([mat.tPre_ms] + [mat.rebuild() for _ in range(2)])
        info = mat.tcw_info()
        Cd.zero_()
        t = timeit(mat, B, Cd, k)
        err = ((Cd - Cref).abs().amax(dim=1) / rownorm).max().item()
        print(f"  T={T} W={W} cost={cost} gain={gain}: {t:.4f} ms  win {info['win_nnz'] / nnz:.3f}  cols {info['listed_columns']}"
              f" ({info['listed_columns'] / nnz:.4f} of nnz)  ntc {info['ntc']}  tPre {tpre:.2f}  maxdiff {err:.2e}", flush=True)
        mat.free()


if __name__ == "__main__":
    main()
