"""How much of each 128-row panel's nz lives in its W heaviest columns (candidate tensor-core windows)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from flex_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "reddit"
t = time.time()
rp, col, val = synth.generate(name)
rp = rp.numpy(); col = col.numpy()
n = rp.size - 1
print(name, "n", n, "nnz", col.size, "gen %.1fs" % (time.time() - t))
npanel = (n + 127) // 128
tot = col.size
res = {W: [0, 0, 0] for W in (64, 128, 256, 512)}
contig = 0
for p in range(npanel):
    lo, hi = rp[p * 128], rp[min(n, p * 128 + 128)]
    c = col[lo:hi]
    u, cnt = np.unique(c, return_counts=True)
    cnt_sorted = np.sort(cnt)[::-1]
    for W in res:
        top = cnt_sorted[:W]
        top = top[top >= 2]
        res[W][0] += int(top.sum())            # nz captured
        res[W][1] += int(top.size)             # B rows staged
        res[W][2] += int(top.sum() - top.size) # gathers saved
    # best contiguous 256-column window
    if c.size:
        cs = np.sort(c)
        j = np.searchsorted(cs, cs + 256, side="left")
        contig += int((j - np.arange(cs.size)).max())
for W, (cap, rows, saved) in res.items():
    print("W=%4d: nz captured %.1f%%  B rows staged %.2f%% of nnz  net gathers saved %.1f%%" % (W, 100 * cap / tot, 100 * rows / tot, 100 * saved / tot))
print("best contiguous 256-col window captures %.1f%% of nnz" % (100 * contig / tot))
