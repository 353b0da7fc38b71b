"""CPU checks of the AXW restatement (oracle/axw.py; cusp.cu:3-208 run1 / run2): the two multiplication orders agree with
each other (main.cu:41 `data.compare()`, the only check the reference makes) and with the fp64 product."""
import numpy as np
import pytest

from util import rand_dense, random_csr


@pytest.mark.parametrize("n,k,c", [(300, 32, 16), (257, 64, 64), (129, 16, 128)])
def test_orders_agree(orc, n, k, c):
    from oracle import axw
    rp, col, val = random_csr(n, 7, 3, hubs=1)
    X, W = rand_dense(n, k, 5), rand_dense(k, c, 6)
    a0 = axw.axw_f32(rp, col, val, X, W, order=0)
    a1 = axw.axw_f32(rp, col, val, X, W, order=1)
    gold = axw.axw_f64(rp, col, val, X, W)
    assert a0.shape == a1.shape == gold.shape == (n, c) and a0.dtype == a1.dtype == np.float32
    scale = np.maximum(1.0, np.abs(gold).max(axis=1, keepdims=True))
    assert (np.abs(a0 - gold) / scale).max() < 2e-5 and (np.abs(a1 - gold) / scale).max() < 2e-5
    assert (np.abs(a0.astype(np.float64) - a1) / scale).max() < 4e-5


def test_identity_weight_is_the_spmm(orc):
    from oracle import axw
    rp, col, val = random_csr(200, 5, 9)
    X = rand_dense(200, 32, 1)
    I = np.eye(32, dtype=np.float32)
    ref = orc.spmm_ref(rp, col, val, X)
    assert np.array_equal(axw.axw_f32(rp, col, val, X, I, order=0), ref)
    assert np.array_equal(axw.axw_f32(rp, col, val, X, I, order=1), ref)
