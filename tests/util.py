"""Shared helpers for the parity tests (inputs; no oracle or product logic here)."""
import numpy as np


def random_csr(n, avg_deg, seed, diag=True, hubs=0, blocks=0):
    """Random square CSR with sorted unique columns.  blocks>0 plants dense column blocks so the
    ASpT builder finds dense tiles; hubs>0 adds a few very long rows."""
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    m = n * avg_deg
    rows.append(rng.integers(0, n, m))
    cols.append(rng.integers(0, n, m))
    if diag:
        rows.append(np.arange(n)); cols.append(np.arange(n))
    for h in range(hubs):
        r = int(rng.integers(0, n))
        c = rng.choice(n, size=min(n, 700 + 600 * h), replace=False)
        rows.append(np.full(len(c), r)); cols.append(c)
    for b in range(blocks):
        r0 = int(rng.integers(0, max(1, n - 128)))
        width = int(rng.integers(150, 400))
        c0 = int(rng.integers(0, max(1, n - width)))
        dens = rng.uniform(0.2, 0.6)
        rr, cc = np.nonzero(rng.random((min(128, n - r0), min(width, n - c0))) < dens)
        rows.append(rr + r0); cols.append(cc + c0)
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    key = np.unique(r * n + c)
    r, c = key // n, key % n
    rowptr = np.zeros(n + 1, np.uint32)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.uint32)
    val = (rng.random(len(c)).astype(np.float32) * 2 - 1)
    return rowptr, c.astype(np.uint32), val


def rand_dense(n, k, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, k)).astype(np.float32) * 2 - 1)
