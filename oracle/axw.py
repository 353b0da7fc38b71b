"""CPU restatement of the reference's AXW product (cusp.cu:3-208 run1 / run2, main.cu:22-79) -- TEST INFRASTRUCTURE ONLY.

run1: B = X*W (cublasSgemm), C = A*B (cusparseSpMM);  run2: B = A*X, C = B*W.  Both libraries are closed source, so what is
restated is the arithmetic they are specified to perform: fp32 products accumulated in fp32 (here: the reference CPU SpMM of
oracle/fx_oracle.c for the sparse factor, numpy's fp32 matmul for the dense one), plus an fp64 evaluation of the same
products as the exact result the 1e-5 contract is measured against.  Parity unpinned: cusp.cu does not compile in the
reference and holds no golden vectors; main.cu:41 only checks run1 against run2 (`data.compare()`).
"""
import numpy as np

from . import orc


def axw_f32(rowptr, col, val, X, W, order=0):
    X = np.ascontiguousarray(X, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    if order == 0:
        return orc.spmm_ref(rowptr, col, val, np.ascontiguousarray(X @ W))
    return np.ascontiguousarray(orc.spmm_ref(rowptr, col, val, X) @ W)


def axw_f64(rowptr, col, val, X, W):
    T = X.astype(np.float64) @ W.astype(np.float64)
    out = np.zeros((len(rowptr) - 1, W.shape[1]), np.float64)
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(np.asarray(rowptr, np.int64)))
    np.add.at(out, rows, val.astype(np.float64)[:, None] * T[np.asarray(col, np.int64)])
    return out
