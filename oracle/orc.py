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
