import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import flex_b200 as fx
from util import random_csr, rand_dense
from oracle import orc
n, k = 1100, 64
rp, c, v = random_csr(n, 9, 5, hubs=1, blocks=6)
dl = fx.DataLoader.from_arrays(rp, c, v, k)
B = rand_dense(n, k, 4)
gold = orc.spmm_ref(rp, c, v, B)
for (T, W, g) in [(8, 1024, 64), (8, 512, 64), (4, 1024, 64)]:
    mat = fx.Mat(dl, fmt="tcw", tc_threshold=T, tc_width=W, tc_min_gain=g)
    e = mat.export_tcw()
    Bd = torch.from_numpy(B).cuda(); Cd = torch.full((n, k), float("nan"), device="cuda")
    for rep in range(6):
        mat.spmm(Bd.data_ptr(), Cd.data_ptr(), k); torch.cuda.synchronize()
        res = Cd.cpu().numpy()
        bad = np.abs(res - gold) > 1e-5 * np.maximum(1, np.abs(gold).max(1, keepdims=True))
        rows, cols = np.nonzero(bad)
        print(T, W, "rep", rep, "bad", bad.sum(), "rows", np.unique(rows)[:10], "cols", np.unique(cols)[:10], "ncol", e["tc_ncol"],
              "maxerr", np.abs(res - gold).max())
    mat.free()

