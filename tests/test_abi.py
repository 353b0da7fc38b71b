"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/flexb200.h
declares, its host-side logic (CSV parse, census, reordering) agrees with the oracle, and the
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import flex_b200 as fx
from conftest import HAS_GPU, ROOT
from util import random_csr


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "flexb200.h")).read()
    declared = set(re.findall(r"\b(fx_[A-Za-z_0-9]+)\s*\(", hdr))
    assert declared == set(fx.ABI_SYMBOLS), declared ^ set(fx.ABI_SYMBOLS)
    L = fx.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.fx_version() >= 100


def test_csv_load_matches_oracle(orc, data_dir):
    for name in ("a_mat.csv", "pubmed.csv"):
        path = os.path.join(data_dir, name)
        dl = fx.DataLoader(path, 32)
        m = orc.csv_load(path)
        rp, c, v = dl.host_csr()
        assert np.array_equal(rp, m["rowptr"]) and np.array_equal(c, m["col"]) and np.array_equal(v, m["val"])
        i = dl.info
        for f in ("uni_nb", "c", "n_edges_one_way", "n_edges_asymmetric", "n_nodes_z_out", "n_nodes_z_in",
                  "n_nodes_z_deg"):
            assert getattr(i, f) == m[f], f
        assert bool(i.is_directed) == m["is_directed"]
        assert dl.graph_name == name[:-4] and dl.vertex_order_abbr == "OVO"
        assert np.array_equal(dl.vo_mp, np.arange(dl.n))
    assert np.array_equal(dl.rand_B(4), orc.rand_B(dl.n, 4))


def test_bad_inputs(tmp_path):
    with pytest.raises(fx.FlexError):
        fx.DataLoader(str(tmp_path / "missing.csv"), 8)
    p = tmp_path / "bad.csv"
    p.write_text("0,2,3\n1,0,2\n1.0,2.0\n")  # 3 cols, 2 vals
    with pytest.raises(fx.FlexError):
        fx.DataLoader(str(p), 8)
    p.write_text("0,2,3\n1,1,2\n1.0,2.0,3.0\n")  # duplicate column in row 0 (DataLoader.cu:97 assert)
    with pytest.raises(fx.FlexError):
        fx.DataLoader(str(p), 8)
    # empty rows and a trailing newline-free file are fine
    p.write_text("0,0,2,2\n1,2\n0.5,-1.5")
    dl = fx.DataLoader(str(p), 8)
    assert (dl.n, dl.nnz) == (3, 2) and dl.info.n_nodes_z_out == 2


@pytest.mark.skipif(HAS_GPU, reason="only meaningful where no device exists")
def test_no_cpu_fallback(data_dir):
    dl = fx.DataLoader(os.path.join(data_dir, "a_mat.csv"), 8)
    with pytest.raises(fx.FlexError, match="no CUDA device|CUDA"):
        fx.Mat(dl, fmt="aspt")


def test_reorder_with_rank_matches_oracle(orc):
    n = 400
    rp, c, v = random_csr(n, 6, 11)
    dl = fx.DataLoader.from_arrays(rp, c, v, 8)
    rank = np.random.default_rng(5).permutation(n).astype(np.uint64)
    d2 = dl.reorder_with_rank(rank)
    vo, rp2, c2, v2 = orc.perm_apply(rp, c, v, rank)
    r, cc, vv = d2.host_csr()
    assert np.array_equal(d2.vo_mp, vo) and np.array_equal(r, rp2) and np.array_equal(cc, c2) and np.array_equal(vv, v2)
    with pytest.raises(fx.FlexError):
        dl.reorder_with_rank(np.zeros(n, np.uint64))


def test_header_is_plain_c_and_struct_sizes_match(tmp_path):
    """include/flexb200.h compiles as C11 (the boundary is a C ABI: no C++ in the signatures), and the ctypes mirrors of its
    structs in flex_b200/__init__.py have the sizes the C compiler gives them."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc on this box")
    names = {"fx_matrix_info": fx.MatrixInfo, "fx_build_opts": fx.BuildOpts, "fx_pillar_arrays": fx.PillarArrays, "fx_tcw_arrays": fx.TcwArrays,
             "fx_seg_arrays": fx.SegArrays, "fx_tile_arrays": fx.TileArrays, "fx_aspt_arrays": fx.AsptArrays, "fx_report": fx.Report}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "flexb200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    p = subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        n, size = line.split()
        assert C.sizeof(names[n]) == int(size), (n, C.sizeof(names[n]), size)
