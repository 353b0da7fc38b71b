"""GPU tests of the boundary's guard rails (round-2 review findings): calls wider than the build's k, device-resident CSR
that violates the builders' preconditions, pillar queues that no CTA owns, a non-finite feature row."""
import os

import numpy as np
import pytest

import flex_b200 as fx
from test_gpu_spmm import assert_close, run_spmm
from util import random_csr, rand_dense

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["aspt", "tcw"])
def test_call_wider_than_build_k(orc, fmt):
    """The per-handle scratch (512-chunk partial sums, window products) is sized for the build's k: a wider call must not
    write past it.  Rows of 512..1023 nz make the partial scratch as large as it gets relative to the arena."""
    n = 1500
    rng = np.random.default_rng(5)
    rows, cols = [], []
    for r in range(n):
        cnt = int(rng.integers(520, 900)) if r % 3 == 0 else int(rng.integers(2, 9))
        c = rng.choice(n, size=cnt, replace=False)
        rows.append(np.full(cnt, r)); cols.append(np.sort(c))
    r, c = np.concatenate(rows), np.concatenate(cols).astype(np.uint32)
    rp = np.zeros(n + 1, np.uint32)
    np.add.at(rp, r + 1, 1)
    rp = np.cumsum(rp).astype(np.uint32)
    v = (rng.random(c.size).astype(np.float32) * 2 - 1)
    dl = fx.DataLoader.from_arrays(rp, c, v, 32)
    kw = dict(tc_min_total=-1) if fmt == "tcw" else {}
    mat = fx.Mat(dl, fmt=fmt, **kw)
    for k in (32, 64, 128):  # k = 64 and 128 are wider than the build's 32
        B = rand_dense(n, k, k)
        assert_close(orc, orc.spmm_ref(rp, c, v, B), run_spmm(mat, B, n), rp)
        assert_close(orc, orc.spmm_ref(rp, c, v, B), mat.spmm_host(B), rp)
    mat.free()


def test_device_csr_is_validated():
    """fx_csr_from_device applies the column checks of the host loaders (the GPU builders index counters with col[e] and
    assume ascending columns) and leaves an identity permutation behind."""
    import torch
    rp, c, v = random_csr(400, 6, 11)
    n, nnz = rp.size - 1, c.size
    d = lambda a, dt: torch.from_numpy(a.astype(dt)).cuda()
    rpd, cd, vd = d(rp, np.int32), d(c, np.int32), d(v, np.float32)
    dl = fx.DataLoader.from_device(n, nnz, rpd.data_ptr(), cd.data_ptr(), vd.data_ptr(), 32)
    assert np.array_equal(dl.vo_mp, np.arange(n))
    B = torch.rand((n, 32), device="cuda")
    S = torch.empty_like(B)
    dl.permute_rows(B.data_ptr(), S.data_ptr(), 32)  # identity map on the device too
    torch.cuda.synchronize()
    assert torch.equal(S, B)
    bad = c.copy()
    bad[int(rp[7])] = n + 3  # column out of range
    with pytest.raises(fx.FlexError, match="column index"):
        fx.DataLoader.from_device(n, nnz, rpd.data_ptr(), d(bad, np.int32).data_ptr(), vd.data_ptr(), 32)
    lo, hi = int(rp[9]), int(rp[10])
    assert hi - lo >= 2
    bad = c.copy()
    bad[lo], bad[lo + 1] = c[lo + 1], c[lo]  # unsorted row
    with pytest.raises(fx.FlexError, match="ascending"):
        fx.DataLoader.from_device(n, nnz, rpd.data_ptr(), d(bad, np.int32).data_ptr(), vd.data_ptr(), 32)
    bad = c.copy()
    bad[lo + 1] = c[lo]  # duplicate column
    with pytest.raises(fx.FlexError, match="ascending"):
        fx.DataLoader.from_device(n, nnz, rpd.data_ptr(), d(bad, np.int32).data_ptr(), vd.data_ptr(), 32)


@pytest.mark.parametrize("n_sm,k", [(148, 256), (400, 128), (400, 384)])
def test_pillar_queues_without_an_owner(orc, data_dir, n_sm, k):
    """More queues than SMs, or several feature chunks: queues that no CTA owns are drained by the sweep (the reference's
    single-slice design leaves them; rows of C would come out zero)."""
    dl = fx.DataLoader(os.path.join(data_dir, "pubmed.csv"), 32)
    rp, c, v = dl.host_csr()
    mat = fx.Mat(dl, fmt="pillar", tm=4, n_sm=n_sm)
    B = rand_dense(dl.n, k, 9)
    assert_close(orc, orc.spmm_ref(rp, c, v, B), run_spmm(mat, B, dl.n), rp)
    mat.free()


def test_tensor_windows_need_finite_features(orc):
    """Documented divergence (include/flexb200.h, FX_FMT_TCW): the dense window contraction multiplies the zeros of the
    A tile by every listed B row, so one Inf in a listed row of B reaches all 128 rows of the panel (0 * Inf = NaN), where
    the sparse semantics of the reference -- and FX_FMT_ASPT here -- touch only the rows that hold a nz in that column."""
    n, k = 640, 128
    rng = np.random.default_rng(3)
    A = (rng.random((n, n)) < 0.02).astype(np.float32) * (rng.random((n, n)).astype(np.float32) * 2 - 1)
    A[:128, :64] = rng.random((128, 64)).astype(np.float32) * 2 - 1
    A[5, 7] = 0.0  # row 5 holds no nz in column 7
    A[np.arange(n), np.arange(n)] = 1.0
    r, cidx = np.nonzero(A)
    rp = np.zeros(n + 1, np.uint32)
    np.add.at(rp, r + 1, 1)
    rp = np.cumsum(rp).astype(np.uint32)
    c, v = cidx.astype(np.uint32), A[r, cidx]
    dl = fx.DataLoader.from_arrays(rp, c, v, k)
    B = rand_dense(n, k, 1)
    B[7, 3] = np.inf
    gold = orc.spmm_ref(rp, c, v, B)
    assert np.isfinite(gold[5]).all()
    m1 = fx.Mat(dl, fmt="aspt")
    assert np.array_equal(np.isfinite(run_spmm(m1, B, n)), np.isfinite(gold))  # sparse semantics: same rows affected
    m1.free()
    m2 = fx.Mat(dl, fmt="tcw", tc_min_total=-1)
    assert m2.tcw_info()["ntc"] > 0
    res = run_spmm(m2, B, n)
    assert not np.isfinite(res[5, 3])  # the divergence, made explicit
    assert np.isfinite(res[200:]).all() == np.isfinite(gold[200:]).all()
    m2.free()
