"""flex_b200 -- host-side mirror of guohaoqiang/Flex's driver surface over libflexb200.so.

The product is the CUDA library (flex_b200/csrc -> flex_b200/libflexb200.so, C ABI in
include/flexb200.h).  This module is the thin ctypes binding the tests and bench.py use; its
classes carry the reference's names and meaning:

    DataLoader(path, k)                  DataLoader.cuh:21-112   (CSV -> CSR -> HBM)
    DataLoaderDeg/Rcm/Gorder(dl)         DataLoader.cuh:128-145  (reordered copy, vo_mp[new]=old)
    Mat(dl, fmt=...)                     mat.cuh:67-229          (tile-format build on the GPU, tPre)
    flex_spmm(A, B, k)                   run() flex.cu:4560 / process() aspt/sspmm_128.cu:1089
                                         -> tPre / tElap / GFlops / Errs report

There is no CPU path: if the shared library is missing or no sm_100 device is visible, the
calls raise.  Nothing here imports the oracle.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libflexb200.so")
_lib = None

FX_ORDER_OVO, FX_ORDER_DEG, FX_ORDER_RCM, FX_ORDER_GOR, FX_ORDER_DFS, FX_ORDER_RBT = 0, 1, 2, 3, 4, 5
FX_FMT_CSR, FX_FMT_ASPT, FX_FMT_TILE, FX_FMT_SEG, FX_FMT_PILLAR, FX_FMT_TCW = 0, 1, 2, 3, 4, 5
_FMT = {"csr": FX_FMT_CSR, "aspt": FX_FMT_ASPT, "tile": FX_FMT_TILE, "seg": FX_FMT_SEG,
        "pillar": FX_FMT_PILLAR, "tcw": FX_FMT_TCW}


class FlexError(RuntimeError):
    pass


class MatrixInfo(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("nnz", C.c_int64), ("dim", C.c_int64),
                ("c", C.c_int64), ("uni_nb", C.c_int64), ("is_directed", C.c_int32),
                ("n_nodes_z_out", C.c_int32), ("n_nodes_z_in", C.c_int32),
                ("n_nodes_z_deg", C.c_int32), ("n_edges_one_way", C.c_int64),
                ("n_edges_asymmetric", C.c_int64), ("order", C.c_int32),
                ("graph_name", C.c_char * 64), ("order_abbr", C.c_char * 4)]


class BuildOpts(C.Structure):
    _fields_ = [("format", C.c_int32), ("tm", C.c_int32), ("tn", C.c_int32), ("bw", C.c_int32),
                ("nnz_limit", C.c_int32), ("n_sm", C.c_int32), ("row_begin", C.c_int32),
                ("row_end", C.c_int32), ("cmajor", C.c_int32), ("tc_threshold", C.c_int32),
                ("tc_width", C.c_int32), ("tc_min_gain", C.c_int32), ("tc_chunk_cost", C.c_int32),
                ("tc_min_total", C.c_int32), ("reserved", C.c_int32 * 2)]


class AsptArrays(C.Structure):
    _fields_ = [("n", C.c_int32), ("nr", C.c_int32), ("npanel", C.c_int32), ("ne", C.c_int32),
                ("BH", C.c_int32), ("BW", C.c_int32), ("num_dense", C.c_int32),
                ("any_flag", C.c_int32), ("regime", C.c_int32), ("special_p", C.c_int32),
                ("S1", C.c_int64), ("S2", C.c_int64), ("avg", C.c_double), ("vari", C.c_double),
                ("mcsr_chk", C.POINTER(C.c_int32)), ("mcsr_cnt", C.POINTER(C.c_int32)),
                ("mcsr_e", C.POINTER(C.c_int32)), ("mcsr_list", C.POINTER(C.c_int32)),
                ("baddr", C.POINTER(C.c_int32)), ("saddr", C.POINTER(C.c_int32)),
                ("perm", C.POINTER(C.c_int32)), ("csr_e", C.POINTER(C.c_int32)),
                ("special", C.POINTER(C.c_int32)), ("special2", C.POINTER(C.c_int32)),
                ("csr_ev", C.POINTER(C.c_float))]


class TileArrays(C.Structure):
    _fields_ = [("m", C.c_int32), ("tm", C.c_int32), ("tn", C.c_int32), ("ntiles", C.c_int32),
                ("npanels", C.c_int32), ("nnz", C.c_int32), ("cmajor", C.c_int32),
                ("tileRowPtr", C.POINTER(C.c_uint32)), ("tileNnz", C.POINTER(C.c_uint32)),
                ("nnzTile", C.POINTER(C.c_int32)), ("bitMap", C.POINTER(C.c_int32)),
                ("tileColIdx", C.POINTER(C.c_uint32)), ("rcOffset", C.POINTER(C.c_int32)),
                ("newVals", C.POINTER(C.c_float))]


class SegArrays(C.Structure):
    _fields_ = [("m", C.c_int32), ("tm", C.c_int32), ("nnz", C.c_int32), ("nsegs", C.c_int32),
                ("rows_total", C.c_int32), ("npanels", C.c_int32), ("n_sm", C.c_int32),
                ("alpha_rowPtr", C.POINTER(C.c_uint32)), ("alpha_colIdx", C.POINTER(C.c_uint32)),
                ("alpha_pillar_rowPtr", C.POINTER(C.c_uint32)), ("segVoMap", C.POINTER(C.c_uint32)),
                ("alpha_vals", C.POINTER(C.c_float)), ("segs_per_panel", C.POINTER(C.c_int32)),
                ("segPtr", C.POINTER(C.c_uint32)), ("segNzRCIdx", C.POINTER(C.c_uint32)),
                ("segVoMapPad", C.POINTER(C.c_uint32)), ("segVals", C.POINTER(C.c_float)),
                ("segNzCV", C.POINTER(C.c_float)), ("seg_rowPtr", C.POINTER(C.c_int32)),
                ("next_seg", C.POINTER(C.c_int32)), ("grouped_tailSeg", C.POINTER(C.c_int32))]


class PillarArrays(C.Structure):
    _fields_ = [("m", C.c_int32), ("nnz", C.c_int32), ("n_sm", C.c_int32), ("n_segs", C.c_int32),
                ("rows_total", C.c_int32), ("warps_with_weights", C.c_int32),
                ("alpha_rowPtr", C.POINTER(C.c_uint32)), ("alpha_colIdx", C.POINTER(C.c_uint32)),
                ("alpha_pillar_rowPtr", C.POINTER(C.c_uint32)), ("alpha_pillarIdx", C.POINTER(C.c_uint32)),
                ("segVoMap", C.POINTER(C.c_uint32)), ("alpha_vals", C.POINTER(C.c_float)),
                ("empty_wp_p", C.c_float), ("band_nz_p", C.c_float), ("round1_on_gpu", C.c_int32)]


class TcwArrays(C.Structure):
    _fields_ = [("n", C.c_int32), ("nr", C.c_int32), ("npanel", C.c_int32), ("W", C.c_int32), ("T", C.c_int32),
                ("min_gain", C.c_int32), ("ntc", C.c_int32), ("dropped", C.c_int32), ("chunk_cost", C.c_int32),
                ("reserved0", C.c_int32),
                ("win_nnz", C.c_int64), ("rest_nnz", C.c_int64), ("min_total", C.c_int64), ("net_gain", C.c_int64),
                ("tc_cols", C.POINTER(C.c_int32)), ("tc_ncol", C.POINTER(C.c_int32)),
                ("win_cptr", C.POINTER(C.c_int32)), ("win_code", C.POINTER(C.c_uint16)),
                ("win_val", C.POINTER(C.c_float)), ("rest_rowptr", C.POINTER(C.c_uint32)),
                ("rest_col", C.POINTER(C.c_uint32)), ("rest_val", C.POINTER(C.c_float))]


class Report(C.Structure):
    _fields_ = [("tPre_ms", C.c_float), ("tElap_ms", C.c_float), ("gflops", C.c_double),
                ("tpre_over_telap", C.c_double), ("errs_flex", C.c_int64),
                ("errs_tight", C.c_int64), ("errs_aspt_pct", C.c_double), ("max_err", C.c_double)]


# every extern "C" symbol include/flexb200.h declares (tests check the library exports all)
ABI_SYMBOLS = [
    "fx_last_error", "fx_version", "fx_launch_count", "fx_device_sm_count", "fx_csr_load",
    "fx_csr_from_arrays", "fx_csr_from_device", "fx_mtx_load", "fx_csr_write_csv", "fx_csr_save_bin", "fx_csr_load_bin", "fx_matrix_get_info", "fx_matrix_host_csr",
    "fx_matrix_device_csr", "fx_matrix_free", "fx_rand_B", "fx_reorder", "fx_reorder_with_rank",
    "fx_permutation", "fx_permute_rows", "fx_unpermute_rows", "fx_build", "fx_rebuild",
    "fx_tiles_export_aspt", "fx_tiles_export_tile", "fx_tiles_export_seg", "fx_tiles_export_pillar", "fx_tiles_export_tcw", "fx_tiles_tcw_info", "fx_tiles_free", "fx_spmm", "fx_spmm_kernel_times", "fx_axw", "fx_spmm_host", "fx_set_device", "fx_panel_shards", "fx_comm_unique_id", "fx_comm_init", "fx_comm_free", "fx_comm_slice", "fx_spmm_sharded_host", "fx_check",
]


def lib():
    """Load libflexb200.so (never builds, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise FlexError(f"{_LIBPATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or make -C flex_b200/csrc); flex_b200 has no CPU fallback")
    L = C.CDLL(_LIBPATH)
    vp = C.c_void_p
    L.fx_last_error.restype = C.c_char_p
    L.fx_launch_count.restype = C.c_int64
    L.fx_device_sm_count.argtypes = [C.POINTER(C.c_int)]
    L.fx_csr_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.fx_csr_from_arrays.argtypes = [C.c_int64, C.c_int64, vp, vp, vp, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.fx_csr_from_device.argtypes = [C.c_int64, C.c_int64, vp, vp, vp, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.fx_mtx_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.fx_csr_write_csv.argtypes = [vp, C.c_char_p]
    L.fx_csr_save_bin.argtypes = [vp, C.c_char_p]
    L.fx_csr_load_bin.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.fx_matrix_get_info.argtypes = [vp, C.POINTER(MatrixInfo)]
    L.fx_matrix_host_csr.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.fx_matrix_device_csr.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.fx_matrix_free.argtypes = [vp]
    L.fx_matrix_free.restype = None
    L.fx_rand_B.argtypes = [C.c_int64, C.c_int, vp]
    L.fx_reorder.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.fx_reorder_with_rank.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.fx_permutation.argtypes = [vp, C.POINTER(vp)]
    L.fx_permute_rows.argtypes = [vp, vp, vp, C.c_int, vp]
    L.fx_unpermute_rows.argtypes = [vp, vp, vp, C.c_int, vp]
    L.fx_build.argtypes = [vp, C.POINTER(BuildOpts), C.POINTER(vp), C.POINTER(C.c_float)]
    L.fx_rebuild.argtypes = [vp, C.POINTER(C.c_float)]
    L.fx_tiles_export_aspt.argtypes = [vp, C.POINTER(AsptArrays)]
    L.fx_tiles_export_tile.argtypes = [vp, C.POINTER(TileArrays)]
    L.fx_tiles_export_seg.argtypes = [vp, C.POINTER(SegArrays)]
    L.fx_tiles_export_pillar.argtypes = [vp, C.POINTER(PillarArrays)]
    L.fx_tiles_export_tcw.argtypes = [vp, C.POINTER(TcwArrays)]
    L.fx_tiles_tcw_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.fx_tiles_free.argtypes = [vp]
    L.fx_tiles_free.restype = None
    L.fx_spmm.argtypes = [vp, vp, vp, C.c_int, vp, C.POINTER(C.c_float)]
    L.fx_spmm_kernel_times.argtypes = [vp, vp, vp, C.c_int, vp, C.POINTER(C.c_float)]
    L.fx_axw.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.fx_spmm_host.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.fx_set_device.argtypes = [C.c_int]
    L.fx_panel_shards.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.fx_comm_unique_id.argtypes = [C.c_char_p]
    L.fx_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.fx_comm_free.argtypes = [vp]
    L.fx_comm_free.restype = None
    L.fx_comm_slice.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.fx_spmm_sharded_host.argtypes = [vp, vp, vp, vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.fx_check.argtypes = [vp, vp, C.c_int64, C.c_int, vp, C.POINTER(Report)]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise FlexError(f"libflexb200 error {rc}: {lib().fx_last_error().decode()}")


def launch_count():
    return int(lib().fx_launch_count())


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))


class DataLoader:
    """CSR matrix on host + device.  DataLoader(path, k) mirrors DataLoader.cu:9-124."""

    order_abbr = "OVO"

    def __init__(self, path=None, k=None, *, _handle=None, _parent=None):
        self._h = C.c_void_p()
        self._parent = _parent  # reordered loaders alias the parent's B (DataLoader.cu:663)
        if _handle is not None:
            self._h = _handle
        else:
            _ck(lib().fx_csr_load(os.fsencode(path), int(k), C.byref(self._h)))
        self._info = None

    @classmethod
    def from_arrays(cls, rowptr, col, val, k, name="arrays.csv"):
        rowptr = np.ascontiguousarray(rowptr, np.uint32)
        col = np.ascontiguousarray(col, np.uint32)
        val = np.ascontiguousarray(val, np.float32)
        h = C.c_void_p()
        _ck(lib().fx_csr_from_arrays(len(rowptr) - 1, len(col), rowptr.ctypes.data, col.ctypes.data,
                                     val.ctypes.data, int(k), name.encode(), C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def from_mtx(cls, path, k):
        """Matrix Market file, converted as data/SuiteSparse/mtx2csr.cc + DataLoader would."""
        h = C.c_void_p()
        _ck(lib().fx_mtx_load(os.fsencode(path), int(k), C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def from_bin(cls, path, k=0):
        h = C.c_void_p()
        _ck(lib().fx_csr_load_bin(os.fsencode(path), int(k), C.byref(h)))
        return cls(_handle=h)

    def save_bin(self, path):
        _ck(lib().fx_csr_save_bin(self._h, os.fsencode(path)))

    def write_csv(self, path):
        _ck(lib().fx_csr_write_csv(self._h, os.fsencode(path)))

    @classmethod
    def from_device(cls, n, nnz, rowptr_ptr, col_ptr, val_ptr, k, name="device.csv"):
        """CSR already in HBM (uint32 rowptr/col, float32 val device pointers)."""
        h = C.c_void_p()
        _ck(lib().fx_csr_from_device(int(n), int(nnz), rowptr_ptr, col_ptr, val_ptr, int(k), name.encode(),
                                     C.byref(h)))
        return cls(_handle=h)

    @property
    def info(self):
        if self._info is None:
            i = MatrixInfo()
            _ck(lib().fx_matrix_get_info(self._h, C.byref(i)))
            self._info = i
        return self._info

    m = property(lambda s: s.info.m)
    n = property(lambda s: s.info.n)
    nnz = property(lambda s: s.info.nnz)
    dim = property(lambda s: s.info.dim)
    graph_name = property(lambda s: s.info.graph_name.decode())
    vertex_order_abbr = property(lambda s: s.info.order_abbr.decode())

    def panel_shards(self, nranks):
        """[(lo, hi)] per rank: the library's row-panel sharding rule (fx_panel_shards)."""
        cuts = (C.c_int64 * (nranks + 1))()
        _ck(lib().fx_panel_shards(self._h, int(nranks), cuts))
        return [(int(cuts[r]), int(cuts[r + 1])) for r in range(nranks)]

    def host_csr(self):
        rp, c, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _ck(lib().fx_matrix_host_csr(self._h, C.byref(rp), C.byref(c), C.byref(v)))
        i = self.info
        return (_np(rp, i.n + 1, np.uint32), _np(c, i.nnz, np.uint32), _np(v, i.nnz, np.float32))

    rowPtr = property(lambda s: s.host_csr()[0])
    col = property(lambda s: s.host_csr()[1])
    vals = property(lambda s: s.host_csr()[2])

    @property
    def vo_mp(self):
        p = C.c_void_p()
        _ck(lib().fx_permutation(self._h, C.byref(p)))
        return _np(p, self.info.n, np.int32)

    def rand_B(self, k=None):
        """The reference's dense B: glibc rand() seeded 1 (DataLoader.cu:198-209)."""
        k = int(k or self.info.dim)
        out = np.empty((self.info.n, k), np.float32)
        _ck(lib().fx_rand_B(self.info.n, k, out.ctypes.data))
        return out

    def reorder(self, order):
        h = C.c_void_p()
        _ck(lib().fx_reorder(self._h, int(order), C.byref(h)))
        return DataLoader(_handle=h, _parent=self)

    def reorder_with_rank(self, rank, tag=FX_ORDER_OVO):
        rank = np.ascontiguousarray(rank, np.uint64)
        h = C.c_void_p()
        _ck(lib().fx_reorder_with_rank(self._h, rank.ctypes.data, int(tag), C.byref(h)))
        return DataLoader(_handle=h, _parent=self)

    def permute_rows(self, B_ptr, out_ptr, k, stream=None):
        _ck(lib().fx_permute_rows(self._h, B_ptr, out_ptr, int(k), stream))

    def unpermute_rows(self, C_ptr, out_ptr, k, stream=None):
        _ck(lib().fx_unpermute_rows(self._h, C_ptr, out_ptr, int(k), stream))

    def free(self):
        if self._h:
            lib().fx_matrix_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def DataLoaderDeg(dl):
    return dl.reorder(FX_ORDER_DEG)


def DataLoaderRcm(dl):
    return dl.reorder(FX_ORDER_RCM)


def DataLoaderGorder(dl):
    return dl.reorder(FX_ORDER_GOR)


def DataLoaderDFS(dl):
    return dl.reorder(FX_ORDER_DFS)


def DataLoaderRabbit(dl):
    return dl.reorder(FX_ORDER_RBT)


class Mat:
    """Tile format built on the GPU (Mat::Mat + csr2tile/transfer/launch_prep, mat.cuh:74-182;
    ASpT: the pre-process section of process(), aspt/sspmm_128.cu:1207-1333)."""

    def __init__(self, dl, fmt="aspt", tm=4, tn=4, bw=0, row_begin=0, row_end=0, n_sm=0, cmajor=0, nnz_limit=0,
                 tc_threshold=0, tc_width=0, tc_min_gain=0, tc_chunk_cost=0, tc_min_total=0):
        self.dl = dl
        self._h = C.c_void_p()
        o = BuildOpts()
        o.format = _FMT[fmt] if isinstance(fmt, str) else int(fmt)
        o.tm, o.tn, o.bw, o.n_sm, o.cmajor, o.nnz_limit = tm, tn, bw, n_sm, int(cmajor), nnz_limit
        o.row_begin, o.row_end = row_begin, row_end
        o.tc_threshold, o.tc_width, o.tc_min_gain = tc_threshold, tc_width, tc_min_gain
        o.tc_chunk_cost, o.tc_min_total = tc_chunk_cost, tc_min_total
        self.fmt = o.format
        self.row_begin = row_begin
        self.row_end = row_end if (row_begin or row_end) else dl.n
        t = C.c_float()
        _ck(lib().fx_build(dl._h, C.byref(o), C.byref(self._h), C.byref(t)))
        self.tPre_ms = t.value

    def rebuild(self):
        t = C.c_float()
        _ck(lib().fx_rebuild(self._h, C.byref(t)))
        self.tPre_ms = t.value
        return t.value

    def export_aspt(self):
        a = AsptArrays()
        _ck(lib().fx_tiles_export_aspt(self._h, C.byref(a)))
        nd, npn, ne = a.num_dense, a.npanel, a.ne
        out = {f: getattr(a, f) for f in ("n", "nr", "npanel", "ne", "BH", "BW", "num_dense", "any_flag",
                                          "regime", "special_p", "S1", "S2", "avg", "vari")}
        out.update(mcsr_chk=_np(a.mcsr_chk, npn, np.int32).copy(), mcsr_cnt=_np(a.mcsr_cnt, npn + 1, np.int32).copy(),
                   mcsr_e=_np(a.mcsr_e, a.BH * (nd + npn) + 1, np.int32).copy(),
                   mcsr_list=_np(a.mcsr_list, nd * a.BW, np.int32).copy(),
                   baddr=_np(a.baddr, nd, np.int32).copy(), saddr=_np(a.saddr, nd, np.int32).copy(),
                   perm=_np(a.perm, ne, np.int32).copy(), csr_e=_np(a.csr_e, ne, np.int32).copy(),
                   csr_ev=_np(a.csr_ev, ne, np.float32).copy(),
                   special=_np(a.special, a.special_p, np.int32).copy(),
                   special2=_np(a.special2, a.special_p, np.int32).copy())
        return out

    def export_tile(self):
        a = TileArrays()
        _ck(lib().fx_tiles_export_tile(self._h, C.byref(a)))
        nt, nnz = a.ntiles, a.nnz
        return dict(ntiles=nt, npanels=a.npanels, tileRowPtr=_np(a.tileRowPtr, a.npanels + 1, np.uint32).copy(),
                    tileNnz=_np(a.tileNnz, nt + 1, np.uint32).copy(), nnzTile=_np(a.nnzTile, nt, np.int32).copy(),
                    bitMap=_np(a.bitMap, nt, np.int32).copy(), tileColIdx=_np(a.tileColIdx, nt, np.uint32).copy(),
                    rcOffset=_np(a.rcOffset, nnz, np.int32).copy(), newVals=_np(a.newVals, nnz, np.float32).copy())

    def export_seg(self):
        a = SegArrays()
        _ck(lib().fx_tiles_export_seg(self._h, C.byref(a)))
        S, R, nnz, tm = a.nsegs, a.rows_total, a.nnz, a.tm
        return dict(nsegs=S, rows_total=R, npanels=a.npanels, n_sm=a.n_sm,
                    alpha_rowPtr=_np(a.alpha_rowPtr, R + 1, np.uint32).copy(),
                    alpha_colIdx=_np(a.alpha_colIdx, nnz, np.uint32).copy(),
                    alpha_vals=_np(a.alpha_vals, nnz, np.float32).copy(),
                    alpha_pillar_rowPtr=_np(a.alpha_pillar_rowPtr, S + 1, np.uint32).copy(),
                    segVoMap=_np(a.segVoMap, R, np.uint32).copy(),
                    segs_per_panel=_np(a.segs_per_panel, a.npanels, np.int32).copy(),
                    segPtr=_np(a.segPtr, S + 1, np.uint32).copy(), segNzRCIdx=_np(a.segNzRCIdx, 2 * nnz, np.uint32).copy(),
                    segVals=_np(a.segVals, nnz, np.float32).copy(), segVoMapPad=_np(a.segVoMapPad, S * tm, np.uint32).copy(),
                    seg_rowPtr=_np(a.seg_rowPtr, S * (tm + 1), np.int32).copy(),
                    segNzCV=_np(a.segNzCV, 2 * nnz, np.float32).copy(),
                    next_seg=_np(a.next_seg, a.n_sm + 1, np.int32).copy(),
                    grouped_tailSeg=_np(a.grouped_tailSeg, a.n_sm + 1, np.int32).copy())

    def export_pillar(self):
        a = PillarArrays()
        _ck(lib().fx_tiles_export_pillar(self._h, C.byref(a)))
        R, nnz = a.rows_total, a.nnz
        return dict(n_segs=a.n_segs, rows_total=R, warps_with_weights=a.warps_with_weights, n_sm=a.n_sm, round1_on_gpu=a.round1_on_gpu,
                    empty_wp_p=a.empty_wp_p, band_nz_p=a.band_nz_p,
                    alpha_rowPtr=_np(a.alpha_rowPtr, R + 1, np.uint32).copy(),
                    alpha_colIdx=_np(a.alpha_colIdx, nnz, np.uint32).copy(),
                    alpha_vals=_np(a.alpha_vals, nnz, np.float32).copy(),
                    alpha_pillar_rowPtr=_np(a.alpha_pillar_rowPtr, a.n_segs + 1, np.uint32).copy(),
                    alpha_pillarIdx=_np(a.alpha_pillarIdx, a.n_sm + 2, np.uint32).copy(),
                    segVoMap=_np(a.segVoMap, R, np.uint32).copy())

    def tcw_info(self):
        """Scalars of the tensor-window plan (no array copies)."""
        o = (C.c_int64 * 8)()
        _ck(lib().fx_tiles_tcw_info(self._h, o))
        return dict(zip(("npanel", "ntc", "win_nnz", "rest_nnz", "listed_columns", "net_gain", "W", "T"), [int(x) for x in o]))

    def export_tcw(self):
        a = TcwArrays()
        _ck(lib().fx_tiles_export_tcw(self._h, C.byref(a)))
        wn, rn = a.win_nnz, a.rest_nnz
        return dict(n=a.n, nr=a.nr, npanel=a.npanel, W=a.W, T=a.T, min_gain=a.min_gain, ntc=a.ntc, dropped=a.dropped,
                    chunk_cost=a.chunk_cost, min_total=a.min_total, net_gain=a.net_gain,
                    win_nnz=wn, rest_nnz=rn,
                    tc_cols=_np(a.tc_cols, a.npanel * a.W, np.int32).copy().reshape(a.npanel, a.W),
                    tc_ncol=_np(a.tc_ncol, a.npanel, np.int32).copy(),
                    win_cptr=_np(a.win_cptr, a.npanel * (a.W // 32) + 1, np.int32).copy(),
                    win_code=_np(a.win_code, wn, np.uint16).copy(), win_val=_np(a.win_val, wn, np.float32).copy(),
                    rest_rowptr=_np(a.rest_rowptr, a.n + 1, np.uint32).copy(),
                    rest_col=_np(a.rest_col, rn, np.uint32).copy(), rest_val=_np(a.rest_val, rn, np.float32).copy())

    def spmm(self, B_ptr, C_ptr, k, stream=None, timed=False):
        """C = A*B on device pointers.  timed=True returns tElap in ms (events + sync)."""
        t = C.c_float()
        _ck(lib().fx_spmm(self._h, B_ptr, C_ptr, int(k), stream, C.byref(t) if timed else None))
        return t.value if timed else None

    def kernel_times(self, B_ptr, C_ptr, k, stream=None):
        """One SpMM with events between its kernels: dict of ms for the tensor-window kernel, the 512-chunk kernel of
        long rows, the row kernel and the whole step."""
        ms = (C.c_float * 4)()
        _ck(lib().fx_spmm_kernel_times(self._h, B_ptr, C_ptr, int(k), stream, ms))
        return {"k_spmm_tc": ms[0], "k_spmm_special_cta": ms[1], "k_spmm_rows": ms[2], "step": ms[3]}

    def axw(self, X_ptr, W_ptr, C_ptr, k, c, order=0, stream=None, timed=False):
        """C = A*(X*W) (order 0, cusp.cu run1) or (A*X)*W (order 1, run2) on device pointers; timed=True returns (gemm_ms, spmm_ms)."""
        g, sp = C.c_float(), C.c_float()
        _ck(lib().fx_axw(self._h, X_ptr, W_ptr, C_ptr, int(k), int(c), int(order), stream,
                         C.byref(g) if timed else None, C.byref(sp) if timed else None))
        return (g.value, sp.value) if timed else None

    def spmm_host(self, B, out=None):
        """Host buffers in, host buffer out (H2D + kernels + D2H)."""
        B = np.ascontiguousarray(B, np.float32)
        k = B.shape[1]
        rows = self.row_end - self.row_begin
        if out is None:
            out = np.empty((rows, k), np.float32)
        tot, tk = C.c_float(), C.c_float()
        _ck(lib().fx_spmm_host(self._h, B.ctypes.data, out.ctypes.data, k, C.byref(tot), C.byref(tk)))
        self.last_total_ms, self.last_tElap_ms = tot.value, tk.value
        return out

    def free(self):
        if self._h:
            lib().fx_tiles_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Comm:
    """fx_comm: the NCCL communicator of the sharded host path (include/flexb200.h, L3b).  `broadcast_bytes(bytes_or_None)`
    is the host program's way of sending rank 0's 128-byte unique id to every rank (torch.distributed, MPI, ...)."""

    def __init__(self, nranks, rank, broadcast_bytes):
        uid = C.create_string_buffer(128)
        if rank == 0:
            _ck(lib().fx_comm_unique_id(uid))
        raw = broadcast_bytes(uid.raw if rank == 0 else None)
        assert len(raw) == 128
        self._h = C.c_void_p()
        self.nranks, self.rank = nranks, rank
        _ck(lib().fx_comm_init(nranks, rank, C.create_string_buffer(raw, 128), C.byref(self._h)))

    def slice(self, n):
        lo, hi = C.c_int64(), C.c_int64()
        _ck(lib().fx_comm_slice(self._h, int(n), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def spmm_sharded_host(self, mat, B_rows, C_local):
        """B_rows: this rank's rows of B, C_local: this rank's rows of C (C-contiguous float32 numpy arrays, ideally pinned).
        Returns (total_ms, tElap_ms) on this rank's device clock."""
        k = C_local.shape[1]
        tot, tk = C.c_float(), C.c_float()
        _ck(lib().fx_spmm_sharded_host(mat._h, self._h, B_rows.ctypes.data if B_rows.size else None, C_local.ctypes.data, k,
                                       C.byref(tot), C.byref(tk)))
        return tot.value, tk.value

    def free(self):
        if self._h:
            lib().fx_comm_free(self._h)
            self._h = C.c_void_p()


def check(gold, res, rowptr=None):
    """resCheck (flex.cu:4155-4213) + ASpT validator (aspt/sspmm_128.cu:1425-1446)."""
    gold = np.ascontiguousarray(gold, np.float32)
    res = np.ascontiguousarray(res, np.float32)
    rep = Report()
    rp = np.ascontiguousarray(rowptr, np.uint32) if rowptr is not None else None
    _ck(lib().fx_check(gold.ctypes.data, res.ctypes.data, gold.shape[0], gold.shape[1],
                       rp.ctypes.data if rp is not None else None, C.byref(rep)))
    return rep


def flex_spmm(A, B, k, fmt="aspt", gold=None):
    """The north-star call: build the tile format for DataLoader `A`, run C = A*B (host B[n,k]),
    return (C, report) with tPre, tElap, GFlops = 2*nnz*k/tElap and -- if `gold` is given -- Errs."""
    mat = Mat(A, fmt=fmt)
    Cm = mat.spmm_host(B)
    rep = Report()
    if gold is not None:
        rep = check(gold, Cm, A.rowPtr)
    rep.tPre_ms = mat.tPre_ms
    rep.tElap_ms = mat.last_tElap_ms
    rep.gflops = 2.0 * A.nnz * k / (mat.last_tElap_ms * 1e-3) / 1e9 if mat.last_tElap_ms > 0 else 0.0
    rep.tpre_over_telap = mat.tPre_ms / mat.last_tElap_ms if mat.last_tElap_ms > 0 else 0.0
    mat.free()
    return Cm, rep
