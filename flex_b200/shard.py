"""Row-panel sharding of A across the GPUs of one box (no reference counterpart:
the reference is single-GPU).  Each rank owns a contiguous range of 128-row panels chosen so that
every rank holds about nnz/G nonzeros; B is replicated; rank r computes rows [lo_r, hi_r) of C with
the single-GPU engine (fx_build with row_begin/row_end).  No collective is on the data path; the
optional all-gather of C is a separate, separately timed step."""
import numpy as np

PANEL = 128  # ASpT panel height (aspt/sspmm_128.cu:33); shard edges sit on panel edges


def panel_shards(rowptr, world):
    """[(lo, hi)] per rank: contiguous, panel-aligned, nnz-balanced by prefix sums over rowPtr."""
    rowptr = np.asarray(rowptr)
    n = len(rowptr) - 1
    npanel = (n + PANEL - 1) // PANEL
    pstart = np.minimum(np.arange(npanel + 1, dtype=np.int64) * PANEL, n)
    pnnz = rowptr[pstart].astype(np.int64)
    total = int(pnnz[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(pnnz, total * r / world)))
    cuts.append(npanel)
    cuts = np.maximum.accumulate(np.minimum(np.array(cuts), npanel))
    return [(int(pstart[cuts[r]]), int(pstart[cuts[r + 1]])) for r in range(world)]


def my_shard(rowptr, rank, world):
    return panel_shards(rowptr, world)[rank]


def gather_rows(dist, local_rows, shards, k, device=None):
    """Optional output all-gather: every rank ends with the full C (rows in shard order).
    Shards differ in height, so each rank contributes a block padded to the tallest shard."""
    import torch
    hmax = max(hi - lo for lo, hi in shards)
    pad = torch.zeros((hmax, k), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    outs = [torch.empty_like(pad) for _ in shards]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, shards)], 0)


class ShardedHostSpmm:
    """CPU-testable restatement (torch.distributed, any backend) of the exchange that the PRODUCT path performs in C:
    `fx_spmm_sharded_host` (include/flexb200.h L3b, flex_b200/csrc/fx_shard.cu; `flex_b200.Comm` binds it) -- same slicing
    rule, same all-gather, same shard multiply.  tests/test_shard_gloo.py drives this class with gloo and the oracle; the
    GPU path is checked in bench.py against fx_spmm_host on every rank.

    C = A*B from HOST buffers on G ranks without sending B over PCIe G times.

    `fx_spmm_host` on every rank copies all of B to its GPU (B is replicated), so at G ranks the host
    feeds G * n*k*4 bytes through its PCIe root per SpMM and the end-to-end time grows with G (8 GPUs,
    Reddit-shape k=128: 5.96 ms against 3.81 ms on one).  Here rank r uploads only rows
    [r*ceil(n/G), (r+1)*ceil(n/G)) of B, the slices are all-gathered between the GPUs (NCCL over NVLink:
    the one real exchange step of the sharded path, SURVEY.md 8e "upload once per rank or ncclBroadcast"),
    every rank multiplies its row-panel shard and copies its own rows of C back.

    `spmm(B_full_tensor, C_local_tensor)` is the device multiply (Mat.spmm through the C ABI on the GPU;
    the CPU tests pass the oracle).  Buffers are allocated once; `__call__` is one step."""

    def __init__(self, dist, n, k, rank, world, device, spmm, n_local_rows):
        import torch
        self.dist, self.n, self.k, self.rank, self.world, self.spmm = dist, n, k, rank, world, spmm
        self.rows_per = (n + world - 1) // world
        self.lo = min(n, rank * self.rows_per)
        self.hi = min(n, self.lo + self.rows_per)
        self.B_full = torch.zeros((self.rows_per * world, k), dtype=torch.float32, device=device)
        self.stage = torch.zeros((self.rows_per, k), dtype=torch.float32, device=device)
        self.C_local = torch.empty((n_local_rows, k), dtype=torch.float32, device=device)

    def h2d_bytes(self):
        return 4 * (self.hi - self.lo) * self.k

    def __call__(self, B_host_slice, C_host_out):
        """B_host_slice: this rank's rows [lo, hi) of B (pinned host tensor); C_host_out: pinned, [n_local_rows, k]."""
        if self.hi > self.lo:
            self.stage[: self.hi - self.lo].copy_(B_host_slice, non_blocking=True)
        self.dist.all_gather_into_tensor(self.B_full, self.stage)
        self.spmm(self.B_full, self.C_local)
        C_host_out.copy_(self.C_local, non_blocking=True)
        return C_host_out
