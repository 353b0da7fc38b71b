#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: SpMM GFLOP/s = 2*nnz*k / tElap
(aspt/sspmm_128.cu:1406) on a synthetic graph of a README-named shape, with the HBM-roofline
fraction, the end-to-end (host buffers) number, the CPU baseline and the clocks seen.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload reddit|flickr|yelp|amazon|pubmed]
                    [--k 128] [--fmt aspt|csr] [--impl reference]

A "step" is one SpMM C = A*B over the whole matrix (every rank: its row-panel shard of A, B
replicated, no data-path collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_K = {"reddit": 128, "flickr": 128, "yelp": 128, "amazon": 128, "pubmed": 32}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("FLEX_WORKLOAD", "reddit"))
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--fmt", default="tcw", help="tcw (tensor windows + ASpT remainder, default) | aspt | csr | tile | seg | pillar")
    ap.add_argument("--tc-threshold", type=int, default=0)
    ap.add_argument("--tc-width", type=int, default=0)
    ap.add_argument("--tc-min-gain", type=int, default=0)
    ap.add_argument("--tc-chunk-cost", type=int, default=0)
    ap.add_argument("--tc-min-total", type=int, default=0)
    ap.add_argument("--order", default="ovo", choices=["ovo", "deg", "rcm", "gor", "dfs", "rbt"])
    ap.add_argument("--shuffle", action="store_true", help="hide the planted block order of the synthetic graph")
    ap.add_argument("--impl", default="flex_b200", choices=["flex_b200", "reference"])
    ap.add_argument("--allgather", action="store_true", help="also time the optional NCCL all-gather of C")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-amazon", action="store_true", help="skip the Amazon-shape k=128 sub-record (the north-star's scaling workload)")
    ap.add_argument("--cusparse", action="store_true", help="context number: torch.sparse (cuSPARSE) on the same GPU")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernels_sha():
    """Hash of the sources of the SpMM kernels and of the builders that lay out what they read."""
    import hashlib
    h = hashlib.sha256()
    for f in ("fx_spmm.cu", "fx_tc_kernel.cuh", "fx_aspt_build.cu", "fx_tcw_build.cu"):
        with open(os.path.join(ROOT, "flex_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(workload, k, fmt):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum over the step's kernels) from the committed ncu
    capture of this workload, profiles/r2_traffic.json -- reported only while the kernel sources are the ones captured."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("kernels_sha") != kernels_sha():
            return None
        return d.get("bytes", {}).get(f"{workload}:{k}:{fmt}")
    except Exception:
        return None


FP32_FMA_PEAK = 148 * 128 * 2 * 1.965e9   # 74.4 TFLOP/s: 148 SMs x 128 lanes x 2 flop x 1965 MHz
XBAR_PEAK = 148 * 64 * 1.93e9             # 18.3 TB/s: 64 B/clk per SM from L2, the rate k_spmm_special_cta sustains (ncu)


def kernel_note(fmt, tcw):
    tail = ("one step = all of them, timed together (kernel_ms: CUDA events between them, this run); the bound that binds is "
            "the L2->SM gather path (nnz*k*4 bytes), see DESIGN.md section 5")
    if fmt == "tcw" and tcw and tcw["ntc"]:
        share = 100.0 * tcw["win_nnz"] / max(1, tcw["win_nnz"] + tcw["rest_nnz"])
        return ("k_spmm_rows (remainder nz) + k_spmm_special_cta (512-nz chunks of long rows) + "
                f"k_spmm_tc (tcgen05 3xTF32 over the panels' shared columns, {share:.0f} % of the nz); " + tail)
    return "k_spmm_rows + k_spmm_special_cta (512-nz chunks of long rows); " + tail


def algorithmic_bytes(n, nnz, k):
    """CSR A (rowptr + col,val) + B read once + C written once."""
    return 4 * (n + 1) + 8 * nnz + 4 * n * k + 4 * n * k


def load_workload(name, k, device, shuffle=False):
    """Returns (rowptr int64, col int64, val f32) torch tensors on `device`."""
    import torch
    from flex_b200 import synth
    if name == "pubmed":
        import flex_b200 as fx
        dl = fx.DataLoader(os.path.join(ROOT, "data", "pubmed.csv"), k)
        rp, c, v = dl.host_csr()
        return (torch.from_numpy(rp.astype("int64")).to(device), torch.from_numpy(c.astype("int64")).to(device),
                torch.from_numpy(v.copy()).to(device))
    return synth.generate(name, device=device, shuffle=shuffle)


class ClockSampler(threading.Thread):
    """NVML clocks / throttle reasons every few ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        import statistics
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(rp, c, v, Bh, k, budget_s=10.0):
    """The reference's CPU SpMM (aspt/sspmm_128.cu:1415-1422) restated in oracle/ -- inner k-loop vectorised, rows over
    all host cores -- on the WHOLE matrix, repeated for about `budget_s` seconds."""
    import numpy as np
    from oracle import orc
    orc.build()
    n, nnz = len(rp) - 1, int(rp[-1])
    cores = int(orc.lib().orc_num_threads())
    out = np.empty((n, k), np.float32)
    orc.spmm_omp(rp, c, v, Bh, threads=cores, out=out)  # warm-up: thread pool, page faults of `out`
    passes, t0 = 0, time.perf_counter()
    while True:
        orc.spmm_omp(rp, c, v, Bh, threads=cores, out=out)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or passes >= 1000:
            break
    return {"value": 2.0 * nnz * k * passes / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": f"the whole workload ({nnz} nz) x {passes} passes after one warm-up pass, {dt:.2f} s; OpenMP row-parallel, "
                      f"vectorised (AVX-512/AVX2 clones) restatement of aspt/sspmm_128.cu:1415-1422, bit-identical to the scalar loop"}


def run_reference(args, k):
    """--impl reference: the reference's own CPU SpMM (oracle port, all host cores) over the WHOLE matrix per step; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import orc
    orc.build()
    rp, c, v = load_workload(args.workload, k, "cpu", args.shuffle)
    rp, c, v = rp.numpy().astype(np.uint32), c.numpy().astype(np.uint32), v.numpy()
    n, nnz = len(rp) - 1, len(c)
    from flex_b200 import synth
    Bh = synth.dense_B(n, k).numpy()
    cores = int(orc.lib().orc_num_threads())  # omp_get_num_procs(): torchrun's OMP_NUM_THREADS=1 does not apply
    out = np.empty((n, k), np.float32)
    for _ in range(max(1, args.warmup)):
        orc.spmm_omp(rp, c, v, Bh, threads=cores, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.spmm_omp(rp, c, v, Bh, threads=cores, out=out)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = 2.0 * nnz * k / dt / 1e9
    sample = (f"the whole workload ({nnz} nz) per step on {cores} host threads; OpenMP row-parallel, vectorised restatement of the "
              f"reference CPU SpMM aspt/sspmm_128.cu:1415-1422 (the reference embeds its CPU loop in a CUDA main(), so it cannot be "
              f"timed alone; built -O3 as aspt/h100_compile_GPU_SpMM_ASpT.sh:7 does)")
    line = {"impl": "reference", "metric": "SpMM GFLOP/s (2*nnz*k/tElap)", "value": val, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, k, n, nnz),
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, k, n, nnz):
    return {"workload": f"{args.workload}-shape" if args.workload != "pubmed" else "pubmed.csv", "n": n, "nnz": nnz,
            "k": k, "format": args.fmt, "order": args.order, "shuffled_ids": bool(args.shuffle),
            "l2_policy": "inputs larger than L2 (A+B+C bytes > 126 MB)" if algorithmic_bytes(n, nnz, k) > 126e6
            else "whole problem fits L2: L2 flushed by a 256 MB write between timed steps",
            # the same strings in both arms (the driver compares the configs): what differs is said per arm
            "arithmetic": "fp32 (reference arm: mul+add on the host cores; flex_b200 arm: FMA, tensor windows on tcgen05 with the 3xTF32 split)",
            "parallelism": "row-panel shards over the GPUs of the run, B replicated (reference arm: one host, all cores)"}


def e2e_sharded(dist, fx, mat, Bh, Ch, n, k, lo, hi, rank, world, dev, steps, sync_all):
    """End to end at N > 1 through the C ABI (fx_spmm_sharded_host, include/flexb200.h L3b): B is uploaded once per JOB --
    rank r sends rows r*ceil(n/N).. of it --, all-gathered over NVLink by NCCL, every rank multiplies its row-panel shard and
    copies its rows of C back, pipelined over column chunks.  fx_spmm_host on every rank pushes all of B through the host's
    PCIe root N times; the caller keeps that time as e2e.replicated_ms.  `Ch` holds the fx_spmm_host result of this rank's
    shard and is the check.  Returns (ms per step: host wall clock, max over ranks, or None; H2D bytes per rank; note)."""
    import torch
    try:
        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        comm = fx.Comm(world, rank, bcast)
        slo, shi = comm.slice(n)
        Bslice = Bh.numpy()[slo:shi]
        Ch2 = torch.empty((hi - lo, k), dtype=torch.float32)
        if dev.type == "cuda":
            Ch2 = Ch2.pin_memory()
        for _ in range(2):
            comm.spmm_sharded_host(mat, Bslice, Ch2.numpy())
        # same kernels and the same column chunks as fx_spmm_host: the results must agree closely
        same = torch.tensor([int(torch.allclose(Ch2, Ch, rtol=1e-4, atol=1e-4))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            comm.spmm_sharded_host(mat, Bslice, Ch2.numpy())  # collective; returns when this rank's C is in host memory
        wall = (time.perf_counter() - t0) * 1e3 / steps
        sh = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(sh, op=dist.ReduceOp.MAX)
        comm.free()
        if int(same.item()) != 1:
            return None, None, "sharded-input path disagreed with fx_spmm_host: not used"
        return sh.item(), int(4 * (shi - slo) * k), (
            "fx_spmm_sharded_host: B uploaded once per job (each rank copies its 1/N row slice from pinned host memory), "
            "ncclAllGather over NVLink, SpMM of the rank's row-panel shard, D2H of its rows of C, pipelined over two column "
            "chunks; bytes are per rank; result checked against fx_spmm_host on every rank (1e-4)")
    except Exception as ex:  # keep the replicated number
        return None, None, "sharded-input path failed: %r" % (ex,)


def timed_steps(step, steps, warmup, small, flush, dist, dev):
    """W untimed + exactly K timed steps bracketed by barrier + synchronize; device time by CUDA events on the launching
    stream, max over ranks.  Returns ms per step."""
    import torch

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if small:  # problem fits in L2: flush between steps and sum per-step event times
        ms = 0.0
        for _ in range(steps):
            flush.fill_(1.0)
            e0.record(); step(); e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
    else:
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
    sync_all()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps


def build_workload(fx, args, workload, k, dev, world, rank, order="ovo", shuffle=False, fmt=None):
    """Synthetic graph on the device -> DataLoader -> this rank's row-panel shard built in the requested format."""
    import numpy as np
    import torch
    from flex_b200.shard import panel_shards
    rp, c, v = load_workload(workload, k, dev, shuffle)
    n, nnz = rp.numel() - 1, c.numel()
    rp_host = rp.cpu().numpy()
    rp32, c32 = rp.to(torch.int32), c.to(torch.int32)
    if order != "ovo":
        dl0 = fx.DataLoader.from_arrays(rp_host.astype(np.uint32), c.cpu().numpy().astype(np.uint32), v.cpu().numpy(), k, workload + ".csv")
        dl = dl0.reorder({"deg": fx.FX_ORDER_DEG, "rcm": fx.FX_ORDER_RCM, "gor": fx.FX_ORDER_GOR, "dfs": fx.FX_ORDER_DFS, "rbt": fx.FX_ORDER_RBT}[order])
        rp_host = dl.rowPtr.astype(np.int64)
    else:
        dl = fx.DataLoader.from_device(n, nnz, rp32.data_ptr(), c32.data_ptr(), v.data_ptr(), k, workload + ".csv")
    shards = panel_shards(rp_host, world)
    lo, hi = shards[rank]
    mat = fx.Mat(dl, fmt=fmt or args.fmt, row_begin=lo, row_end=hi, tc_threshold=args.tc_threshold, tc_width=args.tc_width,
                 tc_min_gain=args.tc_min_gain, tc_chunk_cost=args.tc_chunk_cost, tc_min_total=args.tc_min_total)
    return dict(rp=rp, c=c, v=v, rp32=rp32, c32=c32, n=n, nnz=nnz, rp_host=rp_host, dl=dl, shards=shards, lo=lo, hi=hi, mat=mat)


def amazon_record(fx, args, dev, world, rank, dist):
    """The north-star's scaling workload next to the headline: Amazon-shape k=128, this run's GPUs, 5 timed steps after 3."""
    import torch
    from flex_b200 import synth
    k = 128
    w = build_workload(fx, args, "amazon", k, dev, world, rank)
    mat, n, nnz, lo, hi = w["mat"], w["n"], w["nnz"], w["lo"], w["hi"]
    tpre = min([mat.tPre_ms] + [mat.rebuild() for _ in range(2)])
    B = synth.dense_B(n, k, device=dev)
    Cd = torch.empty((hi - lo, k), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = timed_steps(lambda: mat.spmm(B.data_ptr(), Cd.data_ptr(), k, stream=stream), 5, 3, False, None, dist, dev)
    loc = torch.tensor([int(w["rp_host"][hi] - w["rp_host"][lo])], dtype=torch.int64, device=dev)
    per_rank = [loc.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, loc)
    tcw = mat.tcw_info() if args.fmt == "tcw" else None
    rec = {"workload": "amazon-shape", "n": n, "nnz": nnz, "k": k, "steps": 5, "warmup": 3, "ms_per_step": ms,
           "value": 2.0 * nnz * k / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "scaling": "strong",
           "hbm_frac": algorithmic_bytes(n, nnz, k) / (ms * 1e-3) / 1e9 / peaks()[0],
           "tPre_ms_rank0": tpre, "tensor_windows_rank0": tcw, "nnz_per_rank": [int(x.item()) for x in per_rank]}
    mat.free()
    return rec


def main():
    args = parse()
    k = args.k or DEFAULT_K.get(args.workload, 128)
    if args.impl == "reference":
        return run_reference(args, k)

    import numpy as np
    import torch
    import flex_b200 as fx
    from flex_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        print("bench.py --gpus N>1 must be launched by torch.distributed.run", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w = build_workload(fx, args, args.workload, k, dev, world, rank, args.order, args.shuffle)
    rp, c, v, n, nnz, rp_host, lo, hi, mat, shards = (w[x] for x in ("rp", "c", "v", "n", "nnz", "rp_host", "lo", "hi", "mat", "shards"))
    tpre = [mat.tPre_ms] + [mat.rebuild() for _ in range(3)]
    tcw = mat.tcw_info() if args.fmt == "tcw" else None
    B = synth.dense_B(n, k, device=dev)
    Cd = torch.empty((hi - lo, k), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    small = algorithmic_bytes(n, nnz, k) <= 126e6
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if small else None

    def step():
        mat.spmm(B.data_ptr(), Cd.data_ptr(), k, stream=stream)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = fx.launch_count()
    ms_per_step = timed_steps(step, args.steps, args.warmup, small, flush, dist, dev)
    launches = (fx.launch_count() - l0) * args.steps // (args.steps + max(args.warmup, 3))  # the timed steps' share
    # keep the device busy long enough for the clock sampler to see the kernel under load
    t_end = time.time() + 0.5
    while len(sampler.samples) < 50 and time.time() < t_end:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join()

    # per-kernel times of the same step, measured in THIS run with events between the kernels (5 steps)
    kernel_ms = None
    if args.fmt in ("tcw", "aspt") and k % 4 == 0:
        acc = {}
        for _ in range(5):
            if small:
                flush.fill_(1.0)
            for name, t in mat.kernel_times(B.data_ptr(), Cd.data_ptr(), k, stream=stream).items():
                acc[name] = acc.get(name, 0.0) + t / 5
        kernel_ms = acc

    # end to end through the public call with HOST buffers (pinned): H2D B, kernels, D2H C -- wall clock around the call
    Bh = torch.empty((n, k), dtype=torch.float32).pin_memory()
    Bh.copy_(B)
    Ch = torch.empty((hi - lo, k), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    mat.spmm_host(Bh.numpy(), out=Ch.numpy())
    sync_all()
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(e2e_steps):
        mat.spmm_host(Bh.numpy(), out=Ch.numpy())  # returns when C is in host memory
        dev_ms += mat.last_total_ms
    wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_t = torch.tensor([wall_ms, dev_ms / e2e_steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_dev_ms = e2e_t[0].item(), e2e_t[1].item()
    e2e_h2d, e2e_note, e2e_replicated_ms = int(4 * n * k), None, None
    if dist is not None:
        sh_ms, sh_h2d, e2e_note = e2e_sharded(dist, fx, mat, Bh, Ch, n, k, lo, hi, rank, world, dev, e2e_steps, sync_all)
        if sh_ms is not None:
            e2e_replicated_ms, e2e_ms, e2e_h2d = e2e_ms, sh_ms, sh_h2d

    ag_ms = None
    if args.allgather and dist is not None:
        from flex_b200.shard import gather_rows
        gather_rows(dist, Cd, shards, k)  # warm-up
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        gather_rows(dist, Cd, shards, k)
        a1.record()
        a1.synchronize()
        ag = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ag, op=dist.ReduceOp.MAX)
        ag_ms = ag.item()

    cusparse = None
    if args.cusparse and rank == 0:
        # context only, same protocol as the timed region (L2 flushed between steps when it fits)
        A = torch.sparse_csr_tensor(rp, c, v, size=(n, n))
        for _ in range(3):
            torch.sparse.mm(A, B)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(10):
            if small:
                flush.fill_(1.0)
            c0.record()
            torch.sparse.mm(A, B)
            c1.record(); c1.synchronize()
            tot += c0.elapsed_time(c1)
        cusparse = 2.0 * nnz * k / (tot / 10 * 1e-3) / 1e9
        del A

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(rp_host.astype(np.uint32), w["c32"].cpu().numpy().view(np.uint32), v.cpu().numpy(), Bh.numpy(), k)

    loc_nnz = int(rp_host[hi] - rp_host[lo])
    mat.free()
    del w, rp, c, v, B, Cd, Bh, Ch, flush, mat
    torch.cuda.empty_cache()

    amazon = None
    if args.workload != "amazon" and not args.no_amazon:
        try:
            amazon = amazon_record(fx, args, dev, world, rank, dist)
        except Exception as ex:  # the headline line must survive (e.g. out of memory on a shared box)
            amazon = {"workload": "amazon-shape", "failed": repr(ex)[:200]}

    if rank == 0:
        flops = 2.0 * nnz * k
        value = flops / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = peaks()
        # roofline of the step: every rank streams its share of A and C plus (at most) all of B
        abytes = 4 * (hi - lo + 1) + 8 * loc_nnz + 4 * n * k + 4 * (hi - lo) * k
        achieved = abytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "SpMM GFLOP/s (2*nnz*k/tElap)", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic" if args.workload != "pubmed" else "data/pubmed.csv",
            "config": workload_config(args, k, n, nnz),
            "tPre_ms": min(tpre), "tPre_over_tElap": min(tpre) / ms_per_step,
            "tensor_windows": tcw,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "traffic": measured_traffic(args.workload, k, args.fmt) if (world == 1 and args.order == "ovo" and not args.shuffle) else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes,
                         "kernel": kernel_note(args.fmt, tcw),
                         "kernel_ms": kernel_ms},
            "e2e": {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "device_ms_per_step": e2e_dev_ms, "timer": "host wall clock around the call (pinned host buffers in, C back in host memory)",
                    "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": int(4 * (hi - lo) * k)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        try:
            # The roofs this step can be held against: HBM (algorithmic bytes), fp32 FMA, and the bytes
            # the SMs must pull through the L2->SM crossbar -- one B row per gathered nz and per listed window column, tc_out
            # out and back, A, C -- at 64 B/clk per SM.  `restated` = the largest of the three lower bounds over the step time.
            g_nz = tcw["rest_nnz"] if (tcw and tcw.get("ntc")) else loc_nnz
            g_cols = tcw["listed_columns"] if (tcw and tcw.get("ntc")) else 0
            g_tc = 2 * tcw["ntc"] * 128 * k * 4 if (tcw and tcw.get("ntc")) else 0
            g_bytes = 4 * k * (g_nz + g_cols) + g_tc + 8 * loc_nnz + 4 * (hi - lo) * k
            t_step = ms_per_step * 1e-3
            bounds = {"hbm": abytes / (peak * 1e9), "fp32_fma": (2.0 * loc_nnz * k) / FP32_FMA_PEAK, "xbar": g_bytes / XBAR_PEAK}
            binding = max(bounds, key=bounds.get)
            line["roofline"]["gather_path"] = {"bytes": int(g_bytes), "achieved": g_bytes / t_step / 1e12,
                                               "ceiling": XBAR_PEAK / 1e12, "unit": "TB/s", "frac": g_bytes / t_step / XBAR_PEAK}
            line["roofline"]["restated"] = {"lower_bounds_us": {kk: vv * 1e6 for kk, vv in bounds.items()}, "binding": binding,
                                            "frac": bounds[binding] / t_step}
        except Exception:
            pass
        if e2e_note is not None:
            line["e2e"]["path"] = e2e_note
        if e2e_replicated_ms is not None:
            line["e2e"]["replicated_ms"] = e2e_replicated_ms
        if ag_ms is not None:
            line["allgather_ms"] = ag_ms
        if cusparse is not None:
            line["cusparse_context_gflops"] = cusparse
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if amazon is not None:
            line["amazon"] = amazon
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
