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
