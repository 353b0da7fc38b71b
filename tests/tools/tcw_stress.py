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
Bd = torch.from_numpy(B).cuda()
nfail = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    T, W, g = [(8, 1024, 64), (8, 512, 64), (4, 1024, 64)][it % 3]
    mat = fx.Mat(dl, fmt="tcw", tc_threshold=T, tc_width=W, tc_min_gain=g)
    e = mat.export_tcw()
    CH = W // 32
    for rep in range(3):
        Cd = torch.full((n, k), float("nan"), device="cuda")
        mat.spmm(Bd.data_ptr(), Cd.data_ptr(), k); torch.cuda.synchronize()
        res = Cd.cpu().numpy()
        bad = np.abs(res - gold) > 1e-5 * np.maximum(1, np.abs(gold).max(1, keepdims=True))
        if not bad.any():
            continue
        nfail += 1
        for r in np.unique(np.nonzero(bad)[0])[:3]:
            p, rr = r // 128, r % 128
            d = res[r] - gold[r]
            msg = []
            for ch in range(CH):
                lo, hi = e["win_cptr"][p * CH + ch], e["win_cptr"][p * CH + ch + 1]
                code = e["win_code"][lo:hi].astype(np.int64); val = e["win_val"][lo:hi]
                for j in np.nonzero((code >> 5) == rr)[0]:
                    contrib = val[j] * B[e["tc_cols"][p, ch * 32 + (code[j] & 31)]]
                    for sgn, nm in ((-1, "missing"), (1, "doubled")):
                        if np.abs(d - sgn * contrib).max() < 1e-4:
                            msg.append("WIN chunk %d idx %d/%d slot %d %s" % (ch, j, hi - lo, code[j] & 31, nm))
            q0, q1 = e["rest_rowptr"][r], e["rest_rowptr"][r + 1]
            for j in range(q0, q1):
                contrib = e["rest_val"][j] * B[e["rest_col"][j]]
                for sgn, nm in ((-1, "missing"), (1, "doubled")):
                    if np.abs(d - sgn * contrib).max() < 1e-4:
                        msg.append("REST idx %d of [%d,%d) %s" % (j - q0, 0, q1 - q0, nm))
            print("it", it, (T, W), "rep", rep, "row", r, "panel", p, "rr", rr, "restlen", q1 - q0, "|d|max %.3g" % np.abs(d).max(), msg or "unexplained")
    mat.free()
print("failures", nfail)
