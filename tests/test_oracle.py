"""CPU tests of the oracle itself (no GPU): facts of the reference's own fixtures and the
invariants the reference checks at run time."""
import os

import numpy as np
import pytest

from util import random_csr, rand_dense


def test_pubmed_facts(orc, data_dir):
    m = orc.csv_load(os.path.join(data_dir, "pubmed.csv"))
    # 19717 rows, 108365 nz, symmetric, class count 3 (DataLoader.cu:68-69)
    assert (m["n"], m["nnz"], m["c"]) == (19717, 108365, 3)
    assert not m["is_directed"] and m["n_edges_asymmetric"] == 0
    assert m["n_nodes_z_out"] == 0 and m["n_nodes_z_deg"] == 0
    deg = np.diff(m["rowptr"].astype(np.int64))
    assert deg.min() == 2 and deg.max() == 172
    a = orc.Aspt(m["rowptr"], m["col"], m["val"], 128)
    # no dense tile at either BW, vari = 55.06, ne/nc = 5 -> sparse_v2 regime
    assert a.num_dense == 0 and a.regime == 1 and abs(a.vari - 55.06) < 0.01
    assert np.array_equal(a.mcsr_e, np.concatenate([m["rowptr"].astype(np.int32),
                                                    np.full(a.nr - m["n"], m["nnz"], np.int32)]))


def test_a_mat(orc, data_dir):
    m = orc.csv_load(os.path.join(data_dir, "a_mat.csv"))
    assert (m["n"], m["nnz"]) == (48, 280)
    B = orc.rand_B(48, 8)
    C1 = orc.spmm_ref(m["rowptr"], m["col"], m["val"], B)
    # hand check of one element against plain numpy in float64
    r = 5
    s = sum(float(m["val"][e]) * float(B[m["col"][e], 3]) for e in range(m["rowptr"][r], m["rowptr"][r + 1]))
    assert abs(C1[r, 3] - s) < 1e-4


def test_rand_streams(orc):
    # glibc rand() seeded 1: first value 1804289383 -> 2*r/RAND_MAX-1 (DataLoader.cu:205)
    B = orc.rand_B(2, 2, "flex")
    assert abs(B[0, 0] - (2 * np.float32(1804289383) / np.float32(2147483647) - 1)) < 1e-7
    A = orc.rand_B(2, 2, "aspt")
    assert A[0, 0] == np.float32((1804289383 % 1048576) / 1048576)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_spmm_variants_agree(orc, seed):
    rp, c, v = random_csr(700, 9, seed, hubs=1)
    B = rand_dense(700, 20, seed)
    C1 = orc.spmm_ref(rp, c, v, B)
    C2, _ = orc.spmm_omp(rp, c, v, B)
    assert np.array_equal(C1, C2)
    C3 = orc.spmm_f64(rp, c, v, B)
    assert np.abs(C3 - C1).max() < 1e-3
    rows = np.array([0, 5, 699, 333], np.int64)
    assert np.array_equal(orc.spmm_rows(rows, rp, c, v, B), C1[rows])


@pytest.mark.parametrize("BW", [128, 256])
@pytest.mark.parametrize("seed", [3, 4])
def test_aspt_invariants(orc, BW, seed):
    n = 1000
    rp, c, v = random_csr(n, 6, seed, hubs=2, blocks=5)
    a = orc.Aspt(rp, c, v, BW)
    ne = len(c)
    # every nz appears exactly once with identical value (mat.cu:905-940 style check)
    assert np.array_equal(np.sort(a.perm), np.arange(ne))
    assert np.array_equal(a.csr_e, c[a.perm].astype(np.int32))
    assert np.array_equal(a.csr_ev, v[a.perm])
    assert np.all(np.diff(a.mcsr_e) >= 0) and a.mcsr_e[-1] == ne and a.mcsr_e[0] == 0
    assert a.mcsr_cnt[-1] - a.npanel == a.num_dense
    if BW == 128:
        assert a.num_dense > 0  # planted blocks must be found
    # tile columns: every nz of dense group g sits on a column listed in that tile's slot
    for p in range(a.npanel):
        delta = a.mcsr_cnt[p + 1] - a.mcsr_cnt[p]
        for g in range(delta - 1):
            tile = a.mcsr_cnt[p] - p + g
            assert a.baddr[tile] == p and a.saddr[tile] == g
            lst = a.mcsr_list[tile * BW:(tile + 1) * BW]
            assert (lst >= 0).sum() >= BW * 3 // 4
            for r in range(128):
                base = a.mcsr_cnt[p] * 128 + r * delta
                cols = a.csr_e[a.mcsr_e[base + g]:a.mcsr_e[base + g + 1]]
                assert np.all(lst[cols & (BW - 1)] == cols)
                assert np.all(np.diff(cols) > 0)  # stable: ascending inside a group
    B = rand_dense(n, 12, seed)
    C1 = orc.spmm_ref(rp, c, v, B)
    C2 = a.spmm(B)[:n]
    e = orc.check(C1, C2, rp)
    assert e["tight_count"] == 0 and e["aspt_count"] == 0
    # forced slot lists (a reference run's choice) reproduce the same structure when fed back
    a2 = orc.Aspt(rp, c, v, BW, forced_cnt=a.mcsr_cnt, forced_list=a.mcsr_list)
    assert np.array_equal(a2.mcsr_e, a.mcsr_e) and np.array_equal(a2.perm, a.perm)


def test_perm_apply(orc):
    n = 300
    rp, c, v = random_csr(n, 5, 7)
    rng = np.random.default_rng(1)
    rank = rng.permutation(n).astype(np.uint64)
    vo, rp2, c2, v2 = orc.perm_apply(rp, c, v, rank)
    assert np.array_equal(rank[vo], np.arange(n))
    B = rand_dense(n, 4, 1)
    C1 = orc.spmm_f64(rp, c, v, B)
    # permuted system: A'[rank[i], rank[j]] = A[i,j]  =>  C' = P A P^T (P B)
    C2 = orc.spmm_f64(rp2, c2, v2, B[vo])
    assert np.allclose(C2, C1[vo], atol=1e-12)
    for r in range(n):
        assert np.all(np.diff(c2[rp2[r]:rp2[r + 1]].astype(np.int64)) > 0)
    assert np.array_equal(orc.permute_rows(vo, B), B[vo])


def test_validators(orc):
    g = np.array([[1.0, 0.5, 2.0, 0.0]], np.float32)
    r = np.array([[1.0 + 1e-3, 0.5, 2.0 * 1.02, 0.0]], np.float32)
    e = orc.check(g, r, np.array([0, 3], np.uint32))
    assert e["flex_count"] == 2 and e["aspt_count"] == 1 and e["tight_count"] == 2 and e["gold_zeros"] == 1
    assert abs(e["max_tight"] - 0.02) < 1e-6  # 0.04 / ||gold row||_inf = 2
