"""Regenerates tests/golden/*.npz from the reference itself (oracle/_ref/flexref = the reference's
host sources compiled for the CPU by oracle/ref_build.sh).  Run where /root/reference is mounted:

    python tests/golden/make_golden.py

The fixtures let tests/test_ref_pin.py pin the oracle on machines without /root/reference."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from test_ref_pin import small_graph  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref.build()
    assert ref.available(), "flexref could not be built (no /root/reference?)"
    tmp = tempfile.mkdtemp()
    graphs = [("a_mat", os.path.join(ROOT, "data", "a_mat.csv"))]
    for name, (n, deg, seed, sym) in {"rnd300": (300, 6, 1, False), "sym500": (500, 5, 2, True)}.items():
        rp, c, v = small_graph(n, deg, seed, sym)
        p = os.path.join(tmp, name + ".csv")
        ref.write_csv(p, rp, c, v)
        graphs.append((name, p))
    # pubmed is 2.6 MB of CSV: only its (small) ordering and F5 outputs are stored
    calls = []
    for name, path in graphs:
        calls.append((name, path, "load", ()))
        for kind in ("deg", "rcm", "gor"):
            calls += [(name, path, "rank", (kind,)), (name, path, "order", (kind,))]
        for kind in ("dfs", "rbt"):
            calls.append((name, path, "order", (kind,)))
        for tm in (2, 4, 8, 16):
            calls.append((name, path, "seg", (tm,)))
        for tm, tn in ((2, 2), (4, 4), (8, 4), (16, 4), (4, 32)):
            for major in "RC":
                calls.append((name, path, "tile", (tm, tn, major)))
        for n_sm in (2, 8, 148):
            calls.append((name, path, "diag", (4, n_sm)))
    pub = os.path.join(ROOT, "data", "pubmed.csv")
    for kind in ("deg", "rcm", "gor"):
        calls.append(("pubmed", pub, "rank", (kind,)))
    total = 0
    for name, path, cmd, args in calls:
        try:
            r = ref.run(cmd, path, *args)
        except RuntimeError as e:  # the reference asserted: recorded as absence
            print("skip", name, cmd, args, str(e)[:60])
            continue
        f = os.path.join(OUT, f"{name}_{cmd}_{'_'.join(map(str, args))}.npz")
        np.savez_compressed(f, **r)
        total += os.path.getsize(f)
    print(f"wrote {len(calls)} fixtures, {total / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
