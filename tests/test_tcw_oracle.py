"""CPU checks of the tensor-window plan's restatement (oracle/tcw.py): conservation, the selection rule's
invariants and the gates -- the specification the GPU builder is compared with bit for bit in test_gpu_tcw.py."""
import os
import sys

import numpy as np
import pytest

from util import random_csr

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import tcw  # noqa: E402


@pytest.mark.parametrize("kw", [dict(min_total=0), dict(T=2, W=64, min_gain=1, chunk_cost=8, min_total=0),
                                dict(T=3, W=32, min_gain=16, chunk_cost=40, min_total=0),
                                dict(min_gain=0, chunk_cost=0, min_total=0)])
def test_conservation_and_rule(kw):
    n = 1100
    rp, c, v = random_csr(n, 9, 5, hubs=1, blocks=6)
    pl = tcw.plan(rp, c, v, **kw)
    rp2, c2, v2 = tcw.reassemble(pl)
    assert np.array_equal(rp2, rp.astype(np.int64)) and np.array_equal(c2, c.astype(np.int64)) and np.array_equal(v2, v)
    T, W = pl["T"], pl["W"]
    CH = W // 32
    assert pl["win_nnz"] + pl["rest_nnz"] == c.size
    assert pl["win_cptr"][0] == 0 and pl["win_cptr"][-1] == pl["win_nnz"] and np.all(np.diff(pl["win_cptr"]) >= 0)
    for p in range(pl["npanel"]):
        ns = pl["tc_ncol"][p]
        cols = pl["tc_cols"][p]
        assert np.all(cols[ns:] == -1) and np.all(np.diff(cols[:ns]) > 0)           # ascending, -1 padded
        lo, hi = rp[p * 128], rp[min(n, p * 128 + 128)]
        u, cnt = np.unique(c[lo:hi], return_counts=True)
        listed = dict(zip(u.tolist(), cnt.tolist()))
        assert all(listed[int(x)] >= T for x in cols[:ns])                           # every listed column is a candidate
        if ns:
            worst = min(listed[int(x)] for x in cols[:ns])
            others = [k for col_, k in listed.items() if col_ not in set(cols[:ns].tolist())]
            full_chunks = ns % 32 == 0
            # no unlisted column beats a listed one unless the list was cut at a chunk boundary or at W
            if others and not full_chunks:
                assert max(others) <= worst
        # window nz of the panel sit in its chunks, ordered by (row, position) inside a chunk
        for ch in range(CH):
            a, b = pl["win_cptr"][p * CH + ch], pl["win_cptr"][p * CH + ch + 1]
            r_in, kk = tcw.tile_word_inv(pl["win_code"][a:b].astype(np.int64))
            key = r_in * 32 + kk
            assert np.all(np.diff(key) > 0)
            assert np.all(ch * 32 + kk < ns)


def test_gates():
    rp, c, v = random_csr(1100, 9, 5, hubs=1, blocks=6)
    assert tcw.plan(rp, c, v, min_gain=10**6, min_total=0)["ntc"] == 0     # no panel pays
    assert tcw.plan(rp, c, v, min_total=10**9)["ntc"] == 0                 # the matrix as a whole does not pay
    some = tcw.plan(rp, c, v, min_total=0)
    assert some["ntc"] > 0 and some["net_gain"] > 0
    # a higher chunk cost never lists more columns
    a = tcw.plan(rp, c, v, chunk_cost=64, min_total=0)["tc_ncol"]
    b = tcw.plan(rp, c, v, chunk_cost=512, min_total=0)["tc_ncol"]
    assert np.all(b <= a)


def test_tile_word_is_a_bijection():
    r, kk = np.meshgrid(np.arange(128), np.arange(32), indexing="ij")
    w = tcw.tile_word(r, kk)
    assert np.array_equal(np.sort(w.ravel()), np.arange(4096))
    r2, k2 = tcw.tile_word_inv(w)
    assert np.array_equal(r2, r) and np.array_equal(k2, kk)
    # K-major core matrices: 4 consecutive k of one row are 4 consecutive words, 8 rows of a core matrix 16 B apart
    assert tcw.tile_word(5, 1) - tcw.tile_word(5, 0) == 1 and tcw.tile_word(6, 0) - tcw.tile_word(5, 0) == 4
    assert tcw.tile_word(0, 4) - tcw.tile_word(0, 0) == 32 and tcw.tile_word(8, 0) - tcw.tile_word(0, 0) == 256


def test_shard_gate_scales_with_the_nz_share():
    """A row-panel shard answers the whole-matrix gate for its share of the nz (fx_api.cu fx_build, oracle/tcw.py plan):
    halves of a matrix whose windows pay keep theirs; with the absolute threshold neither half would."""
    n = 2048
    rp, c, v = random_csr(n, 9, 7, hubs=1, blocks=12)
    whole = tcw.plan(rp, c, v, min_total=0)
    gate = whole["net_gain"]                       # the whole matrix passes exactly at its own gain
    assert tcw.plan(rp, c, v, min_total=gate)["ntc"] == whole["ntc"] > 0
    assert tcw.plan(rp, c, v, min_total=gate + 1)["ntc"] == 0
    mid = 1024
    halves = [tcw.plan(rp, c, v, min_total=0, row_begin=a, row_end=b) for a, b in ((0, mid), (mid, n))]
    assert sum(h["net_gain"] for h in halves) == whole["net_gain"]     # panels are independent
    nnz = int(rp[-1])
    for (a, b), h in zip(((0, mid), (mid, n)), halves):
        share = int(rp[b]) - int(rp[a])
        scaled = gate * share // nnz
        got = tcw.plan(rp, c, v, min_total=gate, row_begin=a, row_end=b)
        assert (got["ntc"] > 0) == (h["net_gain"] >= scaled)
        # the same shard against an UNscaled threshold (what round 1 did) would drop its windows whenever its own gain is
        # below the whole matrix's -- which is always, for a proper shard of a matrix with windows in both halves
        assert h["net_gain"] < gate
    assert any(tcw.plan(rp, c, v, min_total=gate, row_begin=a, row_end=b)["ntc"] > 0 for a, b in ((0, mid), (mid, n)))
