"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (flex_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class OrcCsr(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz", C.c_int64), ("rowptr", C.POINTER(C.c_uint32)),
                ("col", C.POINTER(C.c_uint32)), ("val", C.POINTER(C.c_float)),
                ("uni_nb", C.c_int64), ("c", C.c_int)]


class OrcCensus(C.Structure):
    _fields_ = [("is_directed", C.c_int), ("n_edges_one_way", C.c_int64),
                ("n_edges_asymmetric", C.c_int64), ("n_nodes_z_out", C.c_int),
                ("n_nodes_z_in", C.c_int), ("n_nodes_z_deg", C.c_int)]


class OrcErrs(C.Structure):
    _fields_ = [("flex_count", C.c_int64), ("aspt_count", C.c_int64), ("tight_count", C.c_int64),
                ("gold_zeros", C.c_int64), ("max_err", C.c_double), ("max_tight", C.c_double),
                ("aspt_pct", C.c_double)]


class OrcAspt(C.Structure):
    _fields_ = [("n", C.c_int), ("nr", C.c_int), ("npanel", C.c_int), ("ne", C.c_int),
                ("BH", C.c_int), ("BW", C.c_int), ("num_dense", C.c_int), ("any_flag", C.c_int),
                ("mcsr_chk", C.POINTER(C.c_int)), ("mcsr_cnt", C.POINTER(C.c_int)),
                ("mcsr_e", C.POINTER(C.c_int)), ("mcsr_list", C.POINTER(C.c_int)),
                ("baddr", C.POINTER(C.c_int)), ("saddr", C.POINTER(C.c_int)),
                ("key2", C.POINTER(C.c_int)), ("perm", C.POINTER(C.c_int)),
                ("csr_e", C.POINTER(C.c_int)), ("csr_ev", C.POINTER(C.c_float)),
                ("S1", C.c_int64), ("S2", C.c_int64), ("avg", C.c_double), ("vari", C.c_double),
                ("special_p", C.c_int), ("special", C.POINTER(C.c_int)),
                ("special2", C.POINTER(C.c_int)), ("regime", C.c_int)]


class OrcFlexTile(C.Structure):
    _fields_ = [("m", C.c_int), ("tm", C.c_int), ("tn", C.c_int), ("ntiles", C.c_int), ("npanels", C.c_int),
                ("nnz", C.c_int), ("tileRowPtr", C.POINTER(C.c_uint32)), ("tileNnz", C.POINTER(C.c_uint32)),
                ("nnzTile", C.POINTER(C.c_int)), ("bitMap", C.POINTER(C.c_int)),
                ("tileColIdx", C.POINTER(C.c_uint32)), ("rcOffset", C.POINTER(C.c_int)),
                ("newVals", C.POINTER(C.c_float))]


class OrcSeg(C.Structure):
    _fields_ = [("m", C.c_int), ("tm", C.c_int), ("nnz", C.c_int), ("nsegs", C.c_int), ("rows_total", C.c_int),
                ("npanels", C.c_int), ("alpha_rowPtr", C.POINTER(C.c_uint32)),
                ("alpha_colIdx", C.POINTER(C.c_uint32)), ("alpha_vals", C.POINTER(C.c_float)),
                ("pillar_rowPtr", C.POINTER(C.c_uint32)), ("segVoMap", C.POINTER(C.c_uint32)),
                ("segs_per_panel", C.POINTER(C.c_int)), ("segPtr", C.POINTER(C.c_uint32)),
                ("segNzRCIdx", C.POINTER(C.c_uint32)), ("segVals", C.POINTER(C.c_float)),
                ("segVoMapPad", C.POINTER(C.c_uint32)), ("seg_rowPtr", C.POINTER(C.c_int)),
                ("segNzCV", C.POINTER(C.c_float))]


class OrcPillar(C.Structure):
    _fields_ = [("m", C.c_int), ("nnz", C.c_int), ("n_sm", C.c_int), ("n_segs", C.c_int), ("rows_total", C.c_int),
                ("warps_with_weights", C.c_int), ("alpha_rowPtr", C.POINTER(C.c_uint32)),
                ("alpha_colIdx", C.POINTER(C.c_uint32)), ("alpha_vals", C.POINTER(C.c_float)),
                ("alpha_pillar_rowPtr", C.POINTER(C.c_uint32)), ("alpha_pillarIdx", C.POINTER(C.c_uint32)),
                ("segVoMap", C.POINTER(C.c_uint32)), ("empty_wp_p", C.c_float), ("band_nz_p", C.c_float)]


def build():
    """Compile oracle/*.c into liborc.so (gcc).  Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    L.orc_csv_load.argtypes = [C.c_char_p, C.POINTER(OrcCsr)]
    L.orc_csr_free.argtypes = [C.POINTER(OrcCsr)]
    L.orc_census.argtypes = [C.POINTER(OrcCsr), C.POINTER(OrcCensus)]
    L.orc_rand_B_flex.argtypes = [C.c_int64, C.c_int, f32p]
    L.orc_rand_B_aspt.argtypes = [C.c_int64, C.c_int, f32p]
    L.orc_spmm_ref.argtypes = [C.c_int64, u32p, u32p, f32p, f32p, C.c_int, f32p]
    L.orc_spmm_omp.argtypes = [C.c_int64, u32p, u32p, f32p, f32p, C.c_int, f32p, C.c_int]
    L.orc_spmm_omp.restype = C.c_int
    L.orc_spmm_f64.argtypes = [C.c_int64, u32p, u32p, f32p, f32p, C.c_int, f64p, C.c_void_p]
    L.orc_spmm_rows.argtypes = [i64p, C.c_int64, u32p, u32p, f32p, f32p, C.c_int, f32p]
    L.orc_check.argtypes = [f32p, f32p, C.c_int64, C.c_int, C.c_void_p, C.POINTER(OrcErrs)]
    L.orc_perm_apply.argtypes = [C.c_int64, u32p, u32p, f32p, u64p, i32p, u32p, u32p, f32p]
    L.orc_permute_rows.argtypes = [C.c_int64, C.c_int, i32p, f32p, f32p]
    L.orc_aspt_build.argtypes = [C.c_int, u32p, u32p, f32p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.POINTER(OrcAspt)]
    L.orc_aspt_free.argtypes = [C.POINTER(OrcAspt)]
    L.orc_aspt_spmm.argtypes = [C.POINTER(OrcAspt), f32p, C.c_int, f32p]
    L.orc_num_threads.restype = C.c_int
    L.orc_order_deg.argtypes = [C.c_int64, u32p, u32p, C.c_int, u64p]
    L.orc_order_rcm.argtypes = [C.c_int64, u32p, u32p, u64p]
    L.orc_order_gorder.argtypes = [C.c_int64, u32p, u32p, C.c_int, u64p]
    L.orc_order_gorder.restype = C.c_int
    L.orc_order_dfs.argtypes = [C.c_int64, u32p, u32p, u64p]
    L.orc_order_rabbit.argtypes = [C.c_int64, u32p, u32p, C.c_int, i32p]
    L.orc_order_rabbit.restype = C.c_int
    L.orc_flex_tile_build.argtypes = [C.c_int, u32p, u32p, f32p, C.c_int, C.c_int, C.c_int, C.POINTER(OrcFlexTile)]
    L.orc_flextile_free.argtypes = [C.POINTER(OrcFlexTile)]
    L.orc_flextile_spmm.argtypes = [C.POINTER(OrcFlexTile), f32p, C.c_int, f32p]
    L.orc_seg_build.argtypes = [C.c_int, u32p, u32p, f32p, i32p, C.c_int, C.c_int, C.POINTER(OrcSeg)]
    L.orc_seg_free.argtypes = [C.POINTER(OrcSeg)]
    L.orc_sm_buckets.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]
    L.orc_diag_tiling.argtypes = [C.c_int, u32p, u32p, f32p, i32p, C.c_int, C.c_int, C.POINTER(OrcPillar)]
    L.orc_pillar_free.argtypes = [C.POINTER(OrcPillar)]
    L.orc_alpha_spmm.argtypes = [C.c_int, u32p, u32p, f32p, u32p, C.c_int64, f32p, C.c_int, f32p]
    _LIB = L
    return L


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def csv_load(path):
    m = OrcCsr()
    rc = lib().orc_csv_load(path.encode(), C.byref(m))
    if rc != 0:
        raise IOError(f"orc_csv_load({path}) -> {rc}")
    cen = OrcCensus()
    lib().orc_census(C.byref(m), C.byref(cen))
    out = dict(n=m.n, nnz=m.nnz, rowptr=_arr(m.rowptr, m.n + 1, np.uint32),
               col=_arr(m.col, m.nnz, np.uint32), val=_arr(m.val, m.nnz, np.float32),
               uni_nb=m.uni_nb, c=m.c, is_directed=bool(cen.is_directed),
               n_edges_one_way=cen.n_edges_one_way, n_edges_asymmetric=cen.n_edges_asymmetric,
               n_nodes_z_out=cen.n_nodes_z_out, n_nodes_z_in=cen.n_nodes_z_in,
               n_nodes_z_deg=cen.n_nodes_z_deg)
    lib().orc_csr_free(C.byref(m))
    return out


def rand_B(n, k, kind="flex"):
    out = np.empty(n * k, np.float32)
    (lib().orc_rand_B_flex if kind == "flex" else lib().orc_rand_B_aspt)(n, k, out)
    return out.reshape(n, k)


def _csr(rowptr, col, val):
    return (np.ascontiguousarray(rowptr, np.uint32), np.ascontiguousarray(col, np.uint32),
            np.ascontiguousarray(val, np.float32))


def spmm_ref(rowptr, col, val, B):
    rowptr, col, val = _csr(rowptr, col, val)
    n, k = len(rowptr) - 1, B.shape[1]
    Cm = np.empty((n, k), np.float32)
    lib().orc_spmm_ref(n, rowptr, col, val, np.ascontiguousarray(B, np.float32).ravel(), k, Cm.ravel())
    return Cm


def spmm_omp(rowptr, col, val, B, threads=0, out=None):
    rowptr, col, val = _csr(rowptr, col, val)
    n, k = len(rowptr) - 1, B.shape[1]
    Cm = out if out is not None else np.empty((n, k), np.float32)
    used = lib().orc_spmm_omp(n, rowptr, col, val, np.ascontiguousarray(B, np.float32).ravel(), k,
                              Cm.ravel(), threads)
    return Cm, used


def spmm_f64(rowptr, col, val, B, with_abs=False):
    rowptr, col, val = _csr(rowptr, col, val)
    n, k = len(rowptr) - 1, B.shape[1]
    Cm = np.empty((n, k), np.float64)
    Ca = np.empty((n, k), np.float64) if with_abs else None
    lib().orc_spmm_f64(n, rowptr, col, val, np.ascontiguousarray(B, np.float32).ravel(), k, Cm.ravel(),
                       Ca.ctypes.data if with_abs else None)
    return (Cm, Ca) if with_abs else Cm


def spmm_rows(rows, rowptr, col, val, B):
    rowptr, col, val = _csr(rowptr, col, val)
    rows = np.ascontiguousarray(rows, np.int64)
    k = B.shape[1]
    Cm = np.empty((len(rows), k), np.float32)
    lib().orc_spmm_rows(rows, len(rows), rowptr, col, val, np.ascontiguousarray(B, np.float32).ravel(), k,
                        Cm.ravel())
    return Cm


def check(gold, res, rowptr=None):
    gold = np.ascontiguousarray(gold, np.float32)
    res = np.ascontiguousarray(res, np.float32)
    n, k = gold.shape
    e = OrcErrs()
    rp = None
    if rowptr is not None:
        rp_arr = np.ascontiguousarray(rowptr, np.uint32)
        rp = rp_arr.ctypes.data
    lib().orc_check(gold.ravel(), res.ravel(), n, k, rp, C.byref(e))
    return {f: getattr(e, f) for f, _ in OrcErrs._fields_}


def perm_apply(rowptr, col, val, rank):
    rowptr, col, val = _csr(rowptr, col, val)
    n, nnz = len(rowptr) - 1, len(col)
    vo = np.empty(n, np.int32)
    rp = np.empty(n + 1, np.uint32)
    c2 = np.empty(nnz, np.uint32)
    v2 = np.empty(nnz, np.float32)
    lib().orc_perm_apply(n, rowptr, col, val, np.ascontiguousarray(rank, np.uint64), vo, rp, c2, v2)
    return vo, rp, c2, v2


def permute_rows(vo_mp, B):
    n, k = B.shape
    out = np.empty((n, k), np.float32)
    lib().orc_permute_rows(n, k, np.ascontiguousarray(vo_mp, np.int32),
                           np.ascontiguousarray(B, np.float32).ravel(), out.ravel())
    return out


def order(kind, rowptr, col, window=3):
    """rank[old] = new for kind in deg|rcm|gor (DataLoaderDeg/Rcm/Gorder use deg DESC, rcm, gorder w=3)."""
    rowptr = np.ascontiguousarray(rowptr, np.uint32)
    col = np.ascontiguousarray(col, np.uint32)
    n = len(rowptr) - 1
    rank = np.empty(n, np.uint64)
    if kind == "deg":
        lib().orc_order_deg(n, rowptr, col, 1, rank)
    elif kind == "rcm":
        lib().orc_order_rcm(n, rowptr, col, rank)
    elif kind == "dfs":
        lib().orc_order_dfs(n, rowptr, col, rank)
    elif kind == "rbt":
        raise ValueError("use order_rabbit (it yields vo_mp and needs is_directed)")
    else:
        if lib().orc_order_gorder(n, rowptr, col, window, rank) != 0:
            raise ValueError("gorder: isolated vertex")
    return rank


def order_rabbit(rowptr, col, is_directed):
    """vo_mp[new] = old of DataLoaderRabbit (DataLoader.cu:455-655)."""
    rowptr = np.ascontiguousarray(rowptr, np.uint32)
    col = np.ascontiguousarray(col, np.uint32)
    n = len(rowptr) - 1
    vo = np.empty(n, np.int32)
    if lib().orc_order_rabbit(n, rowptr, col, int(bool(is_directed)), vo) != 0:
        raise ValueError("rabbit: no progress")
    return vo


class Aspt:
    """Canonical ASpT tile metadata as numpy arrays (copied out of the C struct)."""

    def __init__(self, rowptr, col, val, BW=128, forced_cnt=None, forced_list=None):
        rowptr, col, val = _csr(rowptr, col, val)
        self._t = OrcAspt()
        n = len(rowptr) - 1
        fc = np.ascontiguousarray(forced_cnt, np.int32) if forced_cnt is not None else None
        fl = np.ascontiguousarray(forced_list, np.int32) if forced_list is not None else None
        rc = lib().orc_aspt_build(n, rowptr, col, val, BW, fc.ctypes.data if fc is not None else None,
                                  fl.ctypes.data if fl is not None else None, C.byref(self._t))
        assert rc == 0
        t = self._t
        for f in ("n", "nr", "npanel", "ne", "BH", "BW", "num_dense", "any_flag", "S1", "S2", "avg",
                  "vari", "special_p", "regime"):
            setattr(self, f, getattr(t, f))
        nd, npn, ne = t.num_dense, t.npanel, t.ne
        self.mcsr_chk = _arr(t.mcsr_chk, npn, np.int32)
        self.mcsr_cnt = _arr(t.mcsr_cnt, npn + 1, np.int32)
        self.mcsr_e = _arr(t.mcsr_e, t.BH * (nd + npn) + 1, np.int32)
        self.mcsr_list = _arr(t.mcsr_list, t.BW * nd, np.int32)
        self.baddr = _arr(t.baddr, nd, np.int32)
        self.saddr = _arr(t.saddr, nd, np.int32)
        self.key2 = _arr(t.key2, ne, np.int32)
        self.perm = _arr(t.perm, ne, np.int32)
        self.csr_e = _arr(t.csr_e, ne, np.int32)
        self.csr_ev = _arr(t.csr_ev, ne, np.float32)
        self.special = _arr(t.special, t.special_p, np.int32) if t.special_p else np.zeros(0, np.int32)
        self.special2 = _arr(t.special2, t.special_p, np.int32) if t.special_p else np.zeros(0, np.int32)

    def spmm(self, B):
        k = B.shape[1]
        out = np.empty((self.nr, k), np.float32)
        lib().orc_aspt_spmm(C.byref(self._t), np.ascontiguousarray(B, np.float32).ravel(), k, out.ravel())
        return out

    def __del__(self):
        try:
            lib().orc_aspt_free(C.byref(self._t))
        except Exception:
            pass


def flex_tile(rowptr, col, val, tm, tn, cmajor):
    """F1 (mat.cu:1345-1518) -> dict of arrays; raises ValueError on an empty row."""
    rowptr, col, val = _csr(rowptr, col, val)
    t = OrcFlexTile()
    if lib().orc_flex_tile_build(len(rowptr) - 1, rowptr, col, val, tm, tn, int(cmajor), C.byref(t)) != 0:
        raise ValueError("empty row")
    nt, nnz = t.ntiles, t.nnz
    out = dict(ntiles=nt, npanels=t.npanels, tileRowPtr=_arr(t.tileRowPtr, t.npanels + 1, np.uint32),
               tileNnz=_arr(t.tileNnz, nt + 1, np.uint32), nnzTile=_arr(t.nnzTile, nt, np.int32),
               bitMap=_arr(t.bitMap, nt, np.int32), tileColIdx=_arr(t.tileColIdx, nt, np.uint32),
               rcOffset=_arr(t.rcOffset, nnz, np.int32), newVals=_arr(t.newVals, nnz, np.float32))
    lib().orc_flextile_free(C.byref(t))
    return out


def flextile_spmm(rowptr, col, val, tm, tn, cmajor, B):
    rowptr, col, val = _csr(rowptr, col, val)
    t = OrcFlexTile()
    assert lib().orc_flex_tile_build(len(rowptr) - 1, rowptr, col, val, tm, tn, int(cmajor), C.byref(t)) == 0
    out = np.empty((len(rowptr) - 1, B.shape[1]), np.float32)
    lib().orc_flextile_spmm(C.byref(t), np.ascontiguousarray(B, np.float32).ravel(), B.shape[1], out.ravel())
    lib().orc_flextile_free(C.byref(t))
    return out


def seg(rowptr, col, val, vo_mp, tm, nnz_limit=128):
    """F2/F3 (mat.cu:1192-1269) -> dict of arrays."""
    rowptr, col, val = _csr(rowptr, col, val)
    s = OrcSeg()
    if lib().orc_seg_build(len(rowptr) - 1, rowptr, col, val, np.ascontiguousarray(vo_mp, np.int32), tm, nnz_limit,
                           C.byref(s)) != 0:
        raise ValueError("empty row")
    S, nnz, R = s.nsegs, s.nnz, s.rows_total
    out = dict(nsegs=S, rows_total=R, npanels=s.npanels, alpha_rowPtr=_arr(s.alpha_rowPtr, R + 1, np.uint32),
               alpha_colIdx=_arr(s.alpha_colIdx, nnz, np.uint32), alpha_vals=_arr(s.alpha_vals, nnz, np.float32),
               alpha_pillar_rowPtr=_arr(s.pillar_rowPtr, S + 1, np.uint32), segVoMap=_arr(s.segVoMap, R, np.uint32),
               segs_per_panel=_arr(s.segs_per_panel, s.npanels, np.int32), segPtr=_arr(s.segPtr, S + 1, np.uint32),
               segNzRCIdx=_arr(s.segNzRCIdx, 2 * nnz, np.uint32), segVals=_arr(s.segVals, nnz, np.float32),
               segVoMapPad=_arr(s.segVoMapPad, S * tm, np.uint32), seg_rowPtr=_arr(s.seg_rowPtr, S * (tm + 1), np.int32),
               segNzCV=_arr(s.segNzCV, 2 * nnz, np.float32))
    lib().orc_seg_free(C.byref(s))
    return out


def sm_buckets(n_sm, nsegs, segs_per_panel):
    spp = np.ascontiguousarray(segs_per_panel, np.int32)
    nx = np.empty(n_sm + 1, np.int32)
    tl = np.empty(n_sm + 1, np.int32)
    lib().orc_sm_buckets(n_sm, nsegs, len(spp), spp, nx, tl)
    return nx, tl


def diag_tiling(rowptr, col, val, vo_mp, tm, n_sm):
    """F5 (mat.cu:680-903) -> dict of arrays; raises ValueError(code) where the reference asserts."""
    rowptr, col, val = _csr(rowptr, col, val)
    p = OrcPillar()
    rc = lib().orc_diag_tiling(len(rowptr) - 1, rowptr, col, val, np.ascontiguousarray(vo_mp, np.int32), tm, n_sm,
                               C.byref(p))
    if rc != 0:
        raise ValueError(rc)
    R, nnz = p.rows_total, p.nnz
    out = dict(n_segs=p.n_segs, rows_total=R, warps_with_weights=p.warps_with_weights, empty_wp_p=p.empty_wp_p,
               band_nz_p=p.band_nz_p, alpha_rowPtr=_arr(p.alpha_rowPtr, R + 1, np.uint32),
               alpha_colIdx=_arr(p.alpha_colIdx, nnz, np.uint32), alpha_vals=_arr(p.alpha_vals, nnz, np.float32),
               alpha_pillar_rowPtr=_arr(p.alpha_pillar_rowPtr, p.n_segs + 1, np.uint32),
               alpha_pillarIdx=_arr(p.alpha_pillarIdx, n_sm + 2, np.uint32), segVoMap=_arr(p.segVoMap, R, np.uint32))
    lib().orc_pillar_free(C.byref(p))
    return out


def alpha_spmm(alpha, m, shadowB):
    k = shadowB.shape[1]
    out = np.empty((m, k), np.float32)
    lib().orc_alpha_spmm(alpha["rows_total"], alpha["alpha_rowPtr"], alpha["alpha_colIdx"], alpha["alpha_vals"],
                         alpha["segVoMap"], m, np.ascontiguousarray(shadowB, np.float32).ravel(), k, out.ravel())
    return out
