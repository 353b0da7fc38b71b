"""Shared-memory bank conflicts of the tensor-window kernel's A scatter under different orders of the nz inside a chunk
(numpy, CPU only; uses the plan restatement oracle/tcw.py, so it is analysis tooling like tests/, not product code)."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from flex_b200 import synth
from oracle import tcw
rp, col, val = synth.generate("reddit")
rp = rp.numpy().astype(np.int64); col = col.numpy(); val = val.numpy()
NP = 160
pl = tcw.plan(rp, col, val, row_begin=0, row_end=NP * 128, min_total=-1)
code = pl["win_code"].astype(np.int64); cptr = pl["win_cptr"]
def wavefronts(codes):
    tot = 0; n = 0
    for s in range(0, codes.size, 32):
        b = codes[s:s + 32] & 31
        tot += np.bincount(b, minlength=32).max(); n += 1
    return tot, n
res = {"builder (row, pos)": [0, 0], "column-major (pos, row)": [0, 0], "bank round-robin": [0, 0]}
nchunks = 0
for i in range(cptr.size - 1):
    c = code[cptr[i]:cptr[i + 1]]
    if c.size == 0: continue
    nchunks += 1
    r, kk = tcw.tile_word_inv(c)
    t, n = wavefronts(c); res["builder (row, pos)"][0] += t; res["builder (row, pos)"][1] += n
    o = np.lexsort((r, kk)); t, n = wavefronts(c[o]); res["column-major (pos, row)"][0] += t; res["column-major (pos, row)"][1] += n
    b = c & 31
    o = np.argsort(b, kind="stable"); bs = b[o]
    rank = np.arange(c.size) - np.searchsorted(bs, bs, side="left")
    o2 = o[np.lexsort((bs, rank))]
    t, n = wavefronts(c[o2]); res["bank round-robin"][0] += t; res["bank round-robin"][1] += n
print("panels", NP, "chunks", nchunks, "window nz", code.size, "nz per chunk %.0f" % (code.size / nchunks))
for k, (t, n) in res.items():
    print("%-28s wavefronts per 32-nz store %.2f" % (k, t / n))
