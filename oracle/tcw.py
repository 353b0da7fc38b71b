"""CPU restatement of the tensor-window plan (FX_FMT_TCW) -- TEST INFRASTRUCTURE ONLY.

The format has no counterpart layout in the reference (it serves the purpose of the reference's
on-chip tiles, mat.cu:1345-1518 / ASpT dense tiles aspt/sspmm_128.cu:896-981, on tcgen05), so this
file is the specification the CUDA builder (flex_b200/csrc/fx_tcw_build.cu) is checked against bit
for bit, and its only numerical claim -- window part + remainder == the matrix -- is checked against
the reference-pinned SpMM oracle.  Only tests/ may import it.
"""
import numpy as np

BH = 128
CAND_CAP = 8192


def tile_word(r, kk):
    """4-byte word of element (r, kk) in the K-major 128 x 32 operand tile (8 x 16-byte core matrices,
    K-adjacent ones 128 bytes apart, 8-row groups 1024 bytes apart)."""
    return ((r >> 3) << 8) | ((kk >> 2) << 5) | ((r & 7) << 2) | (kk & 3)


def tile_word_inv(w):
    return ((w >> 8) << 3) | ((w >> 2) & 7), (((w >> 5) & 7) << 2) | (w & 3)


def select_columns(panel_cols, T, W, min_gain, chunk_cost):
    """(columns of one panel's window, ascending; its net gain G_p).  Empty array = no window."""
    none = (np.zeros(0, np.int64), 0)
    if panel_cols.size == 0:
        return none
    u, cnt = np.unique(panel_cols.astype(np.int64), return_counts=True)
    t_cur = T
    cand = cnt >= t_cur
    rounds = 0
    while cand.sum() > CAND_CAP and rounds < 6:
        t_cur *= 2
        cand = cnt >= t_cur
        rounds += 1
    if cand.sum() == 0 or cand.sum() > CAND_CAP:
        return none
    cu, cc = u[cand], cnt[cand]
    order = np.lexsort((cu, -cc))[:W]  # count descending, column ascending
    m = order.size
    gain, nch = 0, 0
    for j in range((m + 31) // 32):
        sj = int((cc[order[32 * j:32 * j + 32]] - 1).sum())  # B-row fetches chunk j saves
        if sj < chunk_cost:
            break
        gain += sj - chunk_cost
        nch += 1
    if gain < min_gain or nch == 0:
        return none
    return np.sort(cu[order[:min(m, 32 * nch)]]), gain


def plan(rowptr, col, val, T=4, W=256, min_gain=1024, chunk_cost=224, min_total=1000000, row_begin=0, row_end=None):
    rowptr = np.asarray(rowptr, np.int64)
    n_all = rowptr.size - 1
    row_end = n_all if row_end is None or (row_begin == 0 and row_end == 0) else row_end
    n = row_end - row_begin
    nr = (n + BH - 1) // BH * BH
    npanel = nr // BH
    base = rowptr[row_begin]
    rp = rowptr[row_begin:row_end + 1] - base
    col = np.asarray(col)[base:rowptr[row_end]].astype(np.int64)
    val = np.asarray(val, np.float32)[base:rowptr[row_end]]
    sel = [select_columns(col[rp[p * BH]:rp[min(n, p * BH + BH)]], T, W, min_gain, chunk_cost) for p in range(npanel)]
    net_gain = sum(g for _, g in sel)
    # a row-panel shard answers the whole-matrix gate for its share of the nz (integer floor, as fx_build does)
    nnz_all, nnz_loc = int(rowptr[-1]), int(rowptr[row_end] - base)
    if 0 < nnz_loc < nnz_all:
        min_total = min_total * nnz_loc // nnz_all
    if net_gain < min_total:  # the windows together do not pay for a second kernel
        sel = [(np.zeros(0, np.int64), 0)] * npanel
    assert W % 32 == 0
    CH = W // 32
    tc_cols = np.full((npanel, W), -1, np.int32)
    tc_ncol = np.zeros(npanel, np.int32)
    win_cptr = np.zeros(npanel * CH + 1, np.int32)
    rest_rowptr = np.zeros(n + 1, np.uint32)
    wc, wv, rc, rv = [], [], [], []
    for p in range(npanel):
        r0, r1 = p * BH, min(n, p * BH + BH)
        lb, ub = rp[r0], rp[r1]
        S = sel[p][0]
        tc_ncol[p] = S.size
        tc_cols[p, :S.size] = S
        pr, ps, pv = [], [], []  # window nz of the panel: row in panel, list position, value (row order)
        for r in range(r0, r1):
            c = col[rp[r]:rp[r + 1]]
            v = val[rp[r]:rp[r + 1]]
            pos = np.searchsorted(S, c)
            inw = (pos < S.size)
            inw[inw] = S[pos[inw]] == c[inw]
            pr.append(np.full(int(inw.sum()), r - r0, np.int64)); ps.append(pos[inw].astype(np.int64)); pv.append(v[inw])
            rc.append(c[~inw].astype(np.uint32)); rv.append(v[~inw])
            rest_rowptr[r + 1] = rest_rowptr[r] + int((~inw).sum())
        pr, ps, pv = np.concatenate(pr), np.concatenate(ps), np.concatenate(pv)
        # chunk-major, then (row, position): a stable sort by chunk keeps the (row, position) order
        o = np.argsort(ps // 32, kind="stable")
        wc.append(tile_word(pr[o], ps[o] & 31).astype(np.uint16)); wv.append(pv[o])
        cnt = np.bincount(ps // 32, minlength=CH)
        win_cptr[p * CH + 1:(p + 1) * CH + 1] = win_cptr[p * CH] + np.cumsum(cnt)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    out = dict(n=n, nr=nr, npanel=npanel, W=W, T=T, min_gain=min_gain, ntc=int((tc_ncol > 0).sum()), net_gain=net_gain,
               tc_cols=tc_cols, tc_ncol=tc_ncol, win_cptr=win_cptr, win_code=cat(wc, np.uint16),
               win_val=cat(wv, np.float32), rest_rowptr=rest_rowptr, rest_col=cat(rc, np.uint32),
               rest_val=cat(rv, np.float32))
    out["win_nnz"] = int(out["win_code"].size)
    out["rest_nnz"] = int(out["rest_col"].size)
    return out


def reassemble(pl):
    """CSR (rowptr, col, val) rebuilt from the two parts: must equal the input matrix."""
    n, W = pl["n"], pl["W"]
    CH = W // 32
    # window nz back to (global row, column, value)
    e = np.arange(pl["win_nnz"])
    chunk = np.searchsorted(pl["win_cptr"], e, side="right") - 1
    p, ch = chunk // CH, chunk % CH
    code = pl["win_code"].astype(np.int64)
    r_in, kk = tile_word_inv(code)
    wrow = p * BH + r_in
    wcol = pl["tc_cols"][p, ch * 32 + kk].astype(np.int64)
    rrow = np.repeat(np.arange(n), np.diff(pl["rest_rowptr"].astype(np.int64)))
    rows = np.concatenate([wrow, rrow])
    cols = np.concatenate([wcol, pl["rest_col"].astype(np.int64)])
    vals = np.concatenate([pl["win_val"], pl["rest_val"]])
    o = np.lexsort((cols, rows))
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), cols[o], vals[o]
