"""GPU parity tests of the tensor-window format (FX_FMT_TCW): the plan bit for bit against its CPU
restatement (oracle/tcw.py), the SpMM (tcgen05 window kernel + ASpT remainder) against the
reference-pinned SpMM oracle within the repo's row-normwise 1e-5 contract."""
import os
import sys

import numpy as np
import pytest

import flex_b200 as fx
from util import random_csr, rand_dense
from test_gpu_spmm import run_spmm, assert_close

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import tcw as tcw_oracle  # noqa: E402

pytestmark = pytest.mark.gpu

ARRAYS = ("tc_cols", "tc_ncol", "win_cptr", "win_code", "win_val", "rest_rowptr", "rest_col", "rest_val")


def assert_plan_equal(mat, rp, c, v, **kw):
    e = mat.export_tcw()
    o = tcw_oracle.plan(rp, c, v, **kw)
    for f in ("n", "nr", "npanel", "W", "T", "ntc", "win_nnz", "rest_nnz", "net_gain"):
        assert e[f] == o[f], (f, e[f], o[f])
    for f in ARRAYS:
        assert np.array_equal(e[f], o[f]), f
    return o


@pytest.mark.parametrize("k", [32, 64, 128, 256])
def test_planted_blocks(orc, k):
    n = 1500
    rp, c, v = random_csr(n, 6, 21, hubs=2, blocks=8)
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 3)
    gold = orc.spmm_ref(rp, c, v, B)
    mat = fx.Mat(dl, fmt="tcw", tc_min_total=-1)
    o = assert_plan_equal(mat, rp, c, v, min_total=0)
    assert o["ntc"] > 0 and o["win_nnz"] > 0
    rp2, c2, v2 = tcw_oracle.reassemble(o)
    assert np.array_equal(rp2, rp.astype(np.int64)) and np.array_equal(c2, c.astype(np.int64)) and np.array_equal(v2, v)
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    mat.free()


@pytest.mark.parametrize("T,W,min_gain,chunk_cost,min_total", [
    (2, 64, 1, 8, -1), (3, 32, 16, 40, -1), (8, 1024, 64, 224, -1), (4, 512, 100000, 224, -1), (4, 512, -1, -1, -1),
    (4, 512, 64, 100, 30000), (4, 512, 64, 100, 10**9), (4, 2048, 64, 100, -1), (4, 4096, 64, 100, -1)])
def test_plan_parameters(orc, T, W, min_gain, chunk_cost, min_total):
    n, k = 1100, 64
    rp, c, v = random_csr(n, 9, 5, hubs=1, blocks=6)
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 4)
    gold = orc.spmm_ref(rp, c, v, B)
    mat = fx.Mat(dl, fmt="tcw", tc_threshold=T, tc_width=W, tc_min_gain=min_gain, tc_chunk_cost=chunk_cost,
                 tc_min_total=min_total)
    o = assert_plan_equal(mat, rp, c, v, T=T, W=W, min_gain=max(0, min_gain), chunk_cost=max(0, chunk_cost),
                          min_total=max(0, min_total))
    if min_gain == 100000 or min_total == 10**9:
        assert o["ntc"] == 0  # nothing qualifies: the whole matrix goes through the remainder
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    mat.free()


@pytest.mark.parametrize("name", ["a_mat.csv", "pubmed.csv"])
def test_reference_matrices(orc, data_dir, name):
    path = os.path.join(data_dir, name)
    m = orc.csv_load(path)
    rp, c, v = m["rowptr"], m["col"], m["val"]
    n = rp.size - 1
    k = 128
    dl = fx.DataLoader(path, k)
    B = orc.rand_B(n, k)
    gold = orc.spmm_ref(rp, c, v, B)
    mat = fx.Mat(dl, fmt="tcw", tc_threshold=2, tc_min_gain=8, tc_chunk_cost=16, tc_min_total=-1)
    assert_plan_equal(mat, rp, c, v, T=2, min_gain=8, chunk_cost=16, min_total=0)
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    mat.free()


def test_row_shards_and_rebuild(orc):
    n, k = 2000, 128
    rp, c, v = random_csr(n, 8, 11, hubs=1, blocks=10)
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 9)
    gold = orc.spmm_ref(rp, c, v, B)
    parts = []
    for lo, hi in ((0, 640), (640, 1408), (1408, n)):
        mat = fx.Mat(dl, fmt="tcw", row_begin=lo, row_end=hi, tc_min_total=-1)
        assert_plan_equal(mat, rp, c, v, row_begin=lo, row_end=hi, min_total=0)
        mat.rebuild()
        assert_plan_equal(mat, rp, c, v, row_begin=lo, row_end=hi, min_total=0)
        parts.append(run_spmm(mat, B, hi - lo))
        mat.free()
    assert_close(orc, gold, np.concatenate(parts), rp)


def test_dense_panel(orc):
    """A fully dense 256 x 256 corner: every row of two panels shares every column."""
    n, k = 640, 128
    rng = np.random.default_rng(3)
    A = np.zeros((n, n), np.float32)
    A[:256, :256] = rng.random((256, 256)).astype(np.float32) * 2 - 1
    A[np.arange(n), np.arange(n)] = 1.0
    r, cidx = np.nonzero(A)
    rp = np.zeros(n + 1, np.uint32); np.add.at(rp, r + 1, 1); rp = np.cumsum(rp).astype(np.uint32)
    c = cidx.astype(np.uint32); v = A[r, cidx]
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 1)
    gold = orc.spmm_ref(rp, c, v, B)
    mat = fx.Mat(dl, fmt="tcw", tc_min_total=-1)
    o = assert_plan_equal(mat, rp, c, v, min_total=0)
    assert o["tc_ncol"][0] == 256 and o["tc_ncol"][1] == 256
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    mat.free()


@pytest.mark.parametrize("k", [5, 8, 36, 100, 132])
def test_odd_feature_widths(orc, k):
    """k not a multiple of 32: the last column block of the tensor kernel and of the row kernel is partial;
    k not a multiple of 4 takes the raw CSR path (same as FX_FMT_ASPT)."""
    n = 900
    rp, c, v = random_csr(n, 8, 77, hubs=1, blocks=5)
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 2)
    gold = orc.spmm_ref(rp, c, v, B)
    mat = fx.Mat(dl, fmt="tcw", tc_min_total=-1)
    assert mat.tcw_info()["ntc"] > 0
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    out = mat.spmm_host(B)
    assert_close(orc, gold, out, rp)
    mat.free()


def test_tiles_in_the_remainder(orc, monkeypatch):
    """The remainder goes through the ASpT builder: force its shared-memory tile path on as well."""
    monkeypatch.setenv("FLEX_TILES", "1")
    n, k = 1500, 128
    rp, c, v = random_csr(n, 6, 21, hubs=2, blocks=8)
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 3)
    gold = orc.spmm_ref(rp, c, v, B)
    # a narrow window leaves dense column blocks in the remainder
    mat = fx.Mat(dl, fmt="tcw", tc_width=32, tc_min_total=-1, tc_min_gain=-1)
    assert_plan_equal(mat, rp, c, v, W=32, min_gain=0, min_total=0)
    res = run_spmm(mat, B, n)
    assert_close(orc, gold, res, rp)
    mat.free()
