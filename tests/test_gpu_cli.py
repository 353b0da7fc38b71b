"""The CLI `flex_b200/flexb200` (mirror of the reference's `./flex <csv> <k>` and `./sspmm_128 <csv> <k>`, main.cu:7-13,
aspt/sspmm_128.cu:1460): every format and a reordering through the C++ facade, with `--check` (the reference binary's own
validation step: CPU loop as gold, then the three error counts), and the forked multi-process path of `--gpus N`."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "flex_b200", "flexb200")
PUBMED = os.path.join(ROOT, "data", "pubmed.csv")


def run_cli(*args):
    assert os.path.exists(EXE), "flex_b200/flexb200 is built by __graft_entry__.build() / make -C flex_b200/csrc"
    p = subprocess.run([EXE, PUBMED, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout + p.stderr
    m = re.search(r"errs: (\d+) \(resCheck\)\s+([0-9.eE+-]+) % \(ASpT validator\)\s+(\d+) \(1e-5 row-normwise\)", out)
    assert m, out
    g = re.search(r"GFLOPS: ([0-9.]+)", out)
    assert g and float(g.group(1)) > 0, out
    return int(m.group(1)), float(m.group(2)), int(m.group(3)), out


@pytest.mark.parametrize("fmt", ["tcw", "aspt", "csr", "tile", "seg", "pillar"])
def test_cli_formats(fmt):
    res, pct, tight, out = run_cli(32, "--format", fmt, "--check")
    assert res == 0 and tight == 0 and pct < 0.01, out
    assert "t_pre/t_exe" in out


@pytest.mark.parametrize("order,fmt", [("rcm", "tcw"), ("deg", "seg"), ("gor", "pillar"), ("dfs", "aspt")])
def test_cli_reordered(order, fmt):
    res, pct, tight, out = run_cli(128, "--order", order, "--format", fmt, "--check")
    assert res == 0 and tight == 0 and pct < 0.01, out


def test_cli_gpus():
    """--gpus N: one forked process per GPU, the NCCL id through pipes, fx_spmm_sharded_host per rank (two ranks where the box
    has two GPUs; on a one-GPU box the same command line runs the single-process path)."""
    import torch
    n = min(2, torch.cuda.device_count())
    res, pct, tight, out = run_cli(128, "--gpus", n, "--check")
    assert res == 0 and tight == 0 and pct < 0.01, out
