"""AXW (cusp.cu run1 / run2): the dense factor on tcgen05 (3xTF32) + the SpMM, against oracle/axw.py."""
import os

import numpy as np
import pytest

import flex_b200 as fx
from util import random_csr, rand_dense

pytestmark = pytest.mark.gpu


def run_axw(mat, X, W, rows, order):
    import torch
    Xd, Wd = torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()
    Cd = torch.full((rows, W.shape[1]), float("nan"), dtype=torch.float32, device="cuda")
    g, s = mat.axw(Xd.data_ptr(), Wd.data_ptr(), Cd.data_ptr(), X.shape[1], W.shape[1], order=order, timed=True)
    assert g >= 0 and s >= 0
    return Cd.cpu().numpy()


@pytest.mark.parametrize("fmt", ["aspt", "tcw"])
@pytest.mark.parametrize("k,c", [(128, 64), (128, 128), (64, 16), (32, 32), (96, 40)])
def test_axw_both_orders(orc, data_dir, fmt, k, c):
    from oracle import axw
    dl = fx.DataLoader(os.path.join(data_dir, "pubmed.csv"), 128)
    rp, col, v = (a.copy() for a in dl.host_csr())
    n = dl.n
    X, W = rand_dense(n, k, 1), rand_dense(k, c, 2) / np.float32(np.sqrt(k))
    exact = axw.axw_f64(rp, col, v, X, W)
    scale = 1e-5 * np.maximum(1.0, np.abs(exact).max(axis=1, keepdims=True))
    kw = dict(tc_min_total=-1) if fmt == "tcw" else {}
    mat = fx.Mat(dl, fmt=fmt, **kw)
    for order in (0, 1):
        got = run_axw(mat, X, W, n, order)
        ref32 = axw.axw_f32(rp, col, v, X, W, order)
        # 1e-5 row-normwise against the exact product, and no further from it than the fp32 restatement of the two library calls
        assert (np.abs(got - exact) <= scale).all(), (fmt, k, c, order, np.abs(got - exact).max())
        assert np.abs(got - exact).max() <= 2 * np.abs(ref32 - exact).max() + 1e-6
    mat.free()


def test_axw_planted_windows_and_shard(orc):
    """A matrix whose panels keep tensor windows, and a row-panel shard of it (the dense factor covers all of X, C the shard's rows)."""
    from oracle import axw
    n, k, c = 1900, 128, 64
    rp, col, v = random_csr(n, 6, 3, blocks=8, hubs=2)
    X, W = rand_dense(n, k, 5), rand_dense(k, c, 6) / np.float32(np.sqrt(k))
    exact = axw.axw_f64(rp, col, v, X, W)
    scale = 1e-5 * np.maximum(1.0, np.abs(exact).max(axis=1, keepdims=True))
    dl = fx.DataLoader.from_arrays(rp, col, v, k)
    for lo, hi in ((0, n), (512, 1408)):
        mat = fx.Mat(dl, fmt="tcw", row_begin=lo, row_end=hi, tc_min_total=-1)
        assert mat.tcw_info()["ntc"] > 0
        for order in (0, 1):
            got = run_axw(mat, X, W, hi - lo, order)
            assert (np.abs(got - exact[lo:hi]) <= scale[lo:hi]).all(), (lo, hi, order)
        mat.free()
