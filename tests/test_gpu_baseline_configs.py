"""Full-size parity of the BASELINE.json configurations that tests/test_gpu_fullsize.py does not cover: flickr-shape + RCM,
yelp-shape k=128, yelp-shape + DEG / + Gorder, Amazon-shape k=128 -- each through the whole path of a reordered run
(reorder -> build -> gather B by vo_mp -> SpMM -> scatter C back) and checked IN THE ORIGINAL ORDER against the CPU oracle on
sampled rows, plus the size-independent properties (checksum of checksums in fp64, run-to-run bit identity).

The error report per configuration (gpurun_out/parity_report.jsonl, one line each; profiles/r2_parity_report.jsonl is a copy of the last run):
  * the reference's validators: resCheck misses (flex.cu:4155) and the ASpT 1 % check (aspt/sspmm_128.cu:1425);
  * this repo's contract, row-normwise: |d| <= 1e-5 * max(1, ||gold[row,:]||_inf);
  * the ELEMENTWISE relative count the north star words literally (|d| > 1e-5 * |gold|, gold != 0) -- for the GPU result
    AND for the reference's own CPU loop, both against an fp64 accumulation of the same products: elementwise relative
    error is dominated by cancellation (|gold| << the row's scale), where two correct fp32 summation orders differ;
  * the principled bound asserted here: the GPU's error against fp64 is no larger than twice the error of the reference's
    own fp32 CPU loop against fp64 (max and rms over the sampled elements; rms within four times where 3xTF32 tensor windows
    carry part of the nz -- see the comment at the assertion).
"""
import json
import os

import numpy as np
import pytest

import flex_b200 as fx
from flex_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDERS = {"ovo": None, "deg": fx.FX_ORDER_DEG, "rcm": fx.FX_ORDER_RCM, "gor": fx.FX_ORDER_GOR}

CONFIGS = [
    ("flickr", 128, "rcm"),   # BASELINE configs[1]
    ("yelp", 128, "ovo"),     # configs[3]
    ("yelp", 32, "deg"),      # configs[3]
    ("yelp", 128, "gor"),     # configs[3]
    ("reddit", 128, "ovo"),   # configs[2], the headline: 41 % of the nz through the tcgen05 3xTF32 windows
    ("reddit", 128, "deg"),   # the hubs-first order that closes most windows
    ("amazon", 128, "ovo"),   # configs[4]
]


def error_report(orc, rows, rph, ch, vh, Bh, got):
    """Error statistics of `got` (rows `rows` of C) and of the reference CPU loop, both against fp64."""
    deg = np.diff(rph.astype(np.int64))[rows]
    sub_rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    idx = np.concatenate([np.arange(rph[r], rph[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    sc, sv = ch[idx], vh[idx]
    gold = orc.spmm_rows(rows, rph, ch, vh, Bh)                  # the reference's fp32 loop, its summation order
    f64 = orc.spmm_f64(sub_rp, sc, sv, Bh)                       # the same products accumulated in fp64
    f64 = f64[0] if isinstance(f64, tuple) else f64
    chk = orc.check(gold, got, sub_rp)
    nz = gold != 0
    rel = lambda x, ref: np.abs(x.astype(np.float64) - ref)[nz] / np.abs(ref)[nz]
    e_gpu, e_cpu = np.abs(got.astype(np.float64) - f64), np.abs(gold.astype(np.float64) - f64)
    scale = 1e-5 * np.maximum(1.0, np.abs(f64).max(axis=1, keepdims=True))  # the 1e-5 contract, row-normwise, against the exact product
    rep = dict(rows=int(len(rows)), elements=int(gold.size),
               rownorm_1e5_misses_gpu_vs_f64=int((e_gpu > scale).sum()), rownorm_1e5_misses_cpu_vs_f64=int((e_cpu > scale).sum()),
               resCheck_misses=int(chk["flex_count"]), aspt_pct=float(chk["aspt_pct"]), rownorm_1e5_misses=int(chk["tight_count"]),
               elementwise_rel_1e5_gpu_vs_cpu=int((rel(got, gold.astype(np.float64)) > 1e-5).sum()),
               elementwise_rel_1e5_gpu_vs_f64=int((rel(got, f64) > 1e-5).sum()),
               elementwise_rel_1e5_cpu_vs_f64=int((rel(gold, f64) > 1e-5).sum()),
               max_abs_err_gpu_vs_f64=float(e_gpu.max()), max_abs_err_cpu_vs_f64=float(e_cpu.max()),
               rms_err_gpu_vs_f64=float(np.sqrt((e_gpu ** 2).mean())), rms_err_cpu_vs_f64=float(np.sqrt((e_cpu ** 2).mean())))
    return rep


@pytest.mark.parametrize("name,k,order", CONFIGS)
def test_baseline_config(orc, name, k, order):
    import torch
    rp, c, v = synth.generate(name, device="cuda")
    n, nnz = rp.numel() - 1, c.numel()
    assert (n, nnz) == synth.SHAPES[name][:2]
    rph, ch, vh = rp.cpu().numpy().astype(np.uint32), c.cpu().numpy().astype(np.uint32), v.cpu().numpy()
    if order == "ovo":
        rp32, c32 = rp.to(torch.int32), c.to(torch.int32)
        dl = fx.DataLoader.from_device(n, nnz, rp32.data_ptr(), c32.data_ptr(), v.data_ptr(), k, name + ".csv")
    else:
        dl = fx.DataLoader.from_arrays(rph, ch, vh, k, name + ".csv").reorder(ORDERS[order])
        vo = dl.vo_mp
        assert np.array_equal(np.sort(vo), np.arange(n))  # a permutation
    del rp, c, v
    mat = fx.Mat(dl, fmt="tcw")
    info = mat.tcw_info()
    B = synth.dense_B(n, k, device="cuda")
    Bh = B.cpu().numpy()
    C1 = torch.full((n, k), float("nan"), device="cuda")

    def run(out):
        if order == "ovo":
            mat.spmm(B.data_ptr(), out.data_ptr(), k)
        else:  # shadow_b = B gathered by vo_mp, C' = A'*shadow_b, C[vo_mp[row]] = C'[row]  (flex.cu:276,994)
            S, Cp = torch.empty_like(B), torch.empty_like(out)
            dl.permute_rows(B.data_ptr(), S.data_ptr(), k)
            mat.spmm(S.data_ptr(), Cp.data_ptr(), k)
            dl.unpermute_rows(Cp.data_ptr(), out.data_ptr(), k)
        torch.cuda.synchronize()

    run(C1)
    assert torch.isfinite(C1).all()
    # (1) sampled rows, ORIGINAL order, against the oracle -- heaviest rows included
    rng = np.random.default_rng(2)
    deg = np.diff(rph.astype(np.int64))
    nsample = 1500 if name == "amazon" else 3000
    rows = np.unique(np.concatenate([rng.integers(0, n, nsample), np.argsort(deg)[-16:], [0, n - 1]])).astype(np.int64)
    got = C1[torch.from_numpy(rows).cuda()].cpu().numpy()
    rep = error_report(orc, rows, rph, ch, vh, Bh, got)
    rep.update(config=f"{name}-shape k={k} order={order}", n=n, nnz=nnz, window_share=info["win_nnz"] / nnz, panels_with_window=info["ntc"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(rep) + "\n")
    assert rep["resCheck_misses"] == 0 and rep["aspt_pct"] < 0.01, rep
    # 1e-5 row-normwise against the exact (fp64) product: no miss.  Against the reference's fp32 CPU loop the same bound can
    # only hold where that loop is itself within it (Amazon-shape: 16 k-nz rows of U(-1,1) values, the sequential fp32 sum is
    # off by up to 3e-3 where the GPU's blocked sum is off by 2e-4): misses there are the gold's, and are bounded by them.
    assert rep["rownorm_1e5_misses_gpu_vs_f64"] == 0, rep
    assert rep["rownorm_1e5_misses"] <= rep["rownorm_1e5_misses_cpu_vs_f64"], rep
    # the GPU result is as close to the exact (fp64) product as the reference's own fp32 CPU loop, within a factor of two --
    # for the all-FMA kernels.  Where nz go through the tensor windows the 3xTF32 split (hi.hi + hi.lo + lo.hi, lo.lo dropped,
    # tensor-core accumulation) carries ~2^-22 per product against fp32's 2^-24: measured on Reddit-shape (41 % of the nz in
    # windows) rms 4.9e-8 against the CPU loop's 1.6e-8, max 3.8e-7 against 1.0e-6 -- bounded here by four times the CPU
    # loop's rms, and 25 times inside the 1e-5 contract asserted above.
    assert rep["max_abs_err_gpu_vs_f64"] <= 2.0 * rep["max_abs_err_cpu_vs_f64"] + 1e-7, rep
    assert rep["rms_err_gpu_vs_f64"] <= (4.0 if info["win_nnz"] else 2.0) * rep["rms_err_cpu_vs_f64"] + 1e-9, rep
    # (2) run-to-run bit identity (no atomics anywhere on the path)
    C2 = torch.empty_like(C1)
    run(C2)
    assert torch.equal(C2, C1)
    # (3) checksum of checksums in fp64: column sums of C equal (column sums of A) . B
    cd = torch.from_numpy(ch.astype(np.int64)).cuda()
    colsum_A = torch.zeros(n, dtype=torch.float64, device="cuda").index_add_(0, cd, torch.from_numpy(vh).cuda().double())
    lhs, rhs = C1.double().sum(0), colsum_A @ B.double()
    scale = (colsum_A.abs() @ B.double().abs()).clamp_min(1.0)
    assert ((lhs - rhs).abs() / scale).max().item() < 1e-5
    mat.free()
