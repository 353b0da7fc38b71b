"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the row-panel sharding used by
`bench.py --gpus N`.  Each rank computes its shard with the CPU oracle standing in for the GPU
engine (the GPU engine's per-shard parity is tests/test_gpu_spmm.py::test_row_shards); the test
checks the partition, the nnz balance and that the gathered C equals the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flex_b200.shard import gather_rows, panel_shards
    from oracle import orc
    from util import random_csr, rand_dense
    n, k = 1500, 16
    rp, c, v = random_csr(n, 9, 5, hubs=3)
    B = rand_dense(n, k, 1)
    shards = panel_shards(rp, world)
    lo, hi = shards[rank]
    sub = (rp[lo:hi + 1] - rp[lo]).astype(np.uint32)
    local = orc.spmm_ref(sub, c[rp[lo]:rp[hi]], v[rp[lo]:rp[hi]], B) if hi > lo else np.zeros((0, k), np.float32)
    full = gather_rows(dist, torch.from_numpy(local), shards, k).numpy()
    ok = np.array_equal(full, orc.spmm_ref(rp, c, v, B))
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put((int(t.item()), shards, int(rp[-1])))
    dist.barrier()
    dist.destroy_process_group()


def _worker_host(rank, world, port, ret):
    """ShardedHostSpmm (bench.py's end-to-end path at N > 1): B arrives in row slices, one per rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flex_b200.shard import ShardedHostSpmm, gather_rows, panel_shards
    from oracle import orc
    from util import random_csr, rand_dense
    n, k = 1501, 8  # n not a multiple of the world size: the last slice is short
    rp, c, v = random_csr(n, 7, 11, hubs=2)
    B = rand_dense(n, k, 2)
    shards = panel_shards(rp, world)
    lo, hi = shards[rank]
    sub = (rp[lo:hi + 1] - rp[lo]).astype(np.uint32)

    def spmm(B_full, C_local):
        Bn = np.ascontiguousarray(B_full.numpy()[:n])
        if hi > lo:
            C_local.copy_(torch.from_numpy(orc.spmm_ref(sub, c[rp[lo]:rp[hi]], v[rp[lo]:rp[hi]], Bn)))

    run = ShardedHostSpmm(dist, n, k, rank, world, torch.device("cpu"), spmm, hi - lo)
    Ch = torch.empty((hi - lo, k), dtype=torch.float32)
    for _ in range(2):  # buffers are reused between steps
        run(torch.from_numpy(B[run.lo:run.hi]), Ch)
    full = gather_rows(dist, Ch, shards, k).numpy()
    ok = np.array_equal(full, orc.spmm_ref(rp, c, v, B)) and np.array_equal(run.B_full.numpy()[:n], B)
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.barrier()
    dist.destroy_process_group()


class _CpuMat:
    """Stand-in for flex_b200.Mat on the CPU ranks: the shard's multiply through the oracle."""

    def __init__(self, sub, c, v, n, rows):
        self.sub, self.c, self.v, self.n, self.rows = sub, c, v, n, rows


class _CpuFx:
    """Stand-in for the `flex_b200` module in bench.e2e_sharded: `Comm` with the surface of flex_b200.Comm (the C entry
    fx_spmm_sharded_host over NCCL), here flex_b200.shard.ShardedHostSpmm over gloo with the oracle as the multiply."""

    class Comm:
        def __init__(self, nranks, rank, broadcast_bytes):
            assert broadcast_bytes(b"id" if rank == 0 else None) == b"id"   # the ncclUniqueId hand-over
            self.nranks, self.rank, self.run = nranks, rank, None

        def slice(self, n):
            per = (n + self.nranks - 1) // self.nranks
            lo = min(n, self.rank * per)
            return lo, min(n, lo + per)

        def spmm_sharded_host(self, mat, B_rows, C_local):
            from flex_b200.shard import ShardedHostSpmm
            from oracle import orc
            k = C_local.shape[1]
            if self.run is None:
                def spmm(B_full, C_out):
                    if mat.rows:
                        C_out.copy_(torch.from_numpy(orc.spmm_ref(mat.sub, mat.c, mat.v, np.ascontiguousarray(B_full.numpy()[:mat.n]))))
                self.run = ShardedHostSpmm(dist, mat.n, k, self.rank, self.nranks, torch.device("cpu"), spmm, mat.rows)
            self.run(torch.from_numpy(np.ascontiguousarray(B_rows)), torch.from_numpy(C_local))

        def free(self):
            self.run = None


def _worker_bench(rank, world, port, ret):
    """bench.py's e2e leg at N > 1 (bench.e2e_sharded) end to end on gloo: names, shapes, the check and the reduction."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from flex_b200.shard import panel_shards
    from oracle import orc
    from util import random_csr, rand_dense
    n, k = 1333, 8
    rp, c, v = random_csr(n, 6, 13, hubs=2)
    B = rand_dense(n, k, 3)
    lo, hi = panel_shards(rp, world)[rank]
    sub = (rp[lo:hi + 1] - rp[lo]).astype(np.uint32)
    cs, vs = c[rp[lo]:rp[hi]], v[rp[lo]:rp[hi]]
    mat = _CpuMat(sub, cs, vs, n, hi - lo)
    Ch = torch.from_numpy(orc.spmm_ref(sub, cs, vs, B) if hi > lo else np.zeros((0, k), np.float32))
    ms, h2d, note = bench.e2e_sharded(dist, _CpuFx, mat, torch.from_numpy(B), Ch, n, k, lo, hi, rank, world, torch.device("cpu"),
                                      2, dist.barrier)
    if rank == 0:
        ret.put((ms, h2d, note))
    dist.barrier()
    dist.destroy_process_group()


def test_bench_e2e_sharded_gloo():
    from oracle import orc
    orc.build()
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29900 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker_bench, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    ms, h2d, note = ret.get(timeout=120)
    [p.join(60) for p in procs]
    assert ms is not None and ms > 0, note
    assert h2d == 4 * 667 * 8 and "once per job" in note


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_spmm_gloo(world):
    from oracle import orc
    orc.build()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29700 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    ok = ret.get(timeout=120)
    [p.join(60) for p in procs]
    assert ok == 1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_spmm_gloo(world):
    from oracle import orc
    orc.build()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    [p.start() for p in procs]
    ok, shards, nnz = ret.get(timeout=120)
    [p.join(60) for p in procs]
    assert ok == 1
    assert shards[0][0] == 0 and shards[-1][1] == 1500
    for (a, b), (c, d) in zip(shards[:-1], shards[1:]):
        assert b == c and b % 128 == 0


def test_panel_shards_balance():
    from flex_b200.shard import panel_shards
    rng = np.random.default_rng(0)
    deg = (rng.pareto(1.2, 200000) * 5 + 1).astype(np.int64)  # power-law rows
    rp = np.concatenate([[0], np.cumsum(deg)])
    for world in (1, 2, 4, 8):
        sh = panel_shards(rp, world)
        assert sh[0][0] == 0 and sh[-1][1] == 200000 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        nnz = np.array([rp[hi] - rp[lo] for lo, hi in sh])
        assert nnz.max() <= rp[-1] / world * 1.25 + deg.max() + 128 * deg.mean()
    # more ranks than panels: trailing ranks get empty shards, nothing is lost
    sh = panel_shards(np.arange(0, 301), 4)
    assert sh[0] == (0, 128) or sum(hi - lo for lo, hi in sh) == 300
    assert sum(hi - lo for lo, hi in sh) == 300
