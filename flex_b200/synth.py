"""Deterministic synthetic graphs of the README's named shapes.

The reference repo ships only pubmed.csv and a_mat.csv; flickr/reddit/yelp/amazon are
absent (.MISSING_LARGE_BLOBS), so the benchmark inputs are generated here: degree-corrected
stochastic block model (power-law expected degrees, `ncomm` planted communities laid out
contiguously), symmetric pattern, self loop on every vertex (so every row is non-empty and
holds its diagonal: the Flex builders' preconditions mat.cu:1207,1359,718-727), columns
sorted and unique per row, exactly the named nnz.  All randomness comes from a counter-based
integer hash, so the same graph is produced on CPU and on GPU.

torch is used as an array library only (sort/unique/searchsorted on either device).
"""
import math

import torch

SHAPES = {
    # name: (n, nnz, exponent, ncomm, p_in, sub_size, p_sub, seed, values)
    # ncomm = the class counts DataLoader.cu:62-84 lists for the real datasets.
    "flickr": (89250, 989006, 2.5, 7, 0.4, 64, 0.3, 0xF11C, "gcn"),
    "reddit": (232965, 23446803, 2.0, 41, 0.4, 256, 0.4, 0x4EDD, "gcn"),
    "yelp": (716847, 13954819, 2.3, 100, 0.4, 128, 0.3, 0x7E19, "gcn"),
    "amazon": (1569960, 264339468, 2.1, 107, 0.4, 512, 0.4, 0xA3A2, "uniform"),
}

_M1 = -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687  # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685  # 0x94D049BB133111EB


def _mix(x):
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic, logical shifts emulated)."""
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)
    x = x + _M1
    x = (x ^ lsr(x, 30)) * _M2
    x = (x ^ lsr(x, 27)) * _M3
    return x ^ lsr(x, 31)


def _uniform(seed, stream, count, device, offset=0):
    """count doubles in [0,1) from counter = offset..offset+count-1."""
    ctr = torch.arange(offset, offset + count, dtype=torch.int64, device=device)
    h = _mix(ctr ^ _mix(torch.tensor(seed * 1315423911 + stream, dtype=torch.int64, device=device)))
    return ((h >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))


def generate(name=None, n=None, nnz=None, exponent=2.2, ncomm=16, p_in=0.4, sub_size=128, p_sub=0.3,
             seed=1, values="gcn", device="cpu", shuffle=False):
    """Returns (rowptr int64[n+1], col int64[nnz], val float32[nnz]) on `device`."""
    if name is not None:
        n, nnz, exponent, ncomm, p_in, sub_size, p_sub, seed, values = SHAPES[name]
    assert (nnz - n) % 2 == 0 and nnz >= n, "nnz must be n + 2*(undirected edges)"
    E = (nnz - n) // 2
    dev = torch.device(device)
    # community sizes: power-law-ish, contiguous id ranges
    cs = torch.arange(1, ncomm + 1, dtype=torch.float64, device=dev).pow(-0.6)
    bounds = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), torch.cumsum(cs, 0)])
    bounds = (bounds / bounds[-1] * n).round().to(torch.int64)
    bounds[-1] = n
    comm_of = torch.searchsorted(bounds[1:].contiguous(), torch.arange(n, device=dev), right=True)
    # expected-degree weights: Pareto with tail exponent `exponent`, capped at sqrt-ish of n*avgdeg
    u = _uniform(seed, 1, n, dev)
    w = (1.0 - u).clamp_min(1e-12).pow(-1.0 / (exponent - 1.0))
    w = w.clamp_max(max(4.0, math.sqrt(2.0 * E)))
    cw = torch.cumsum(w, 0)
    cw_lo = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), cw])[bounds[:-1]]
    cw_hi = cw[bounds[1:] - 1]
    pairs = torch.empty(0, dtype=torch.int64, device=dev)
    rnd = 0
    target = E
    while True:
        need = target - pairs.numel()
        if need <= 0:
            break
        batch = int(need * 1.25) + 1024
        off = rnd * (1 << 40)
        ua = _uniform(seed, 2, batch, dev, off)
        ub = _uniform(seed, 3, batch, dev, off)
        uc = _uniform(seed, 4, batch, dev, off)
        src = torch.searchsorted(cw, ua * cw[-1]).clamp_max(n - 1)
        c = comm_of[src]
        lo, hi = cw_lo[c], cw_hi[c]
        tgt_in = lo + ub * (hi - lo)
        tgt = torch.where(uc < p_in, tgt_in, ub * cw[-1])
        # second level: a block of `sub_size` consecutive ids around the source (clipped to its
        # community) receives a p_sub share of the edges -> local clustering
        s_lo = torch.maximum((src // sub_size) * sub_size, bounds[c])
        s_hi = torch.minimum(s_lo + sub_size, bounds[c + 1])
        tgt = torch.where(uc >= 1.0 - p_sub, torch.zeros_like(tgt), tgt)
        dst = torch.searchsorted(cw, tgt).clamp_max(n - 1)
        # local edges land uniformly inside the block (neighbourhoods are not hub-dominated)
        dst_sub = (s_lo + (ub * (s_hi - s_lo).to(torch.float64)).to(torch.int64)).clamp_max(n - 1)
        dst = torch.where(uc >= 1.0 - p_sub, dst_sub, dst)
        del s_lo, s_hi, dst_sub
        a = torch.minimum(src, dst)
        b = torch.maximum(src, dst)
        keep = a != b
        key = (a * n + b)[keep]
        del ua, ub, uc, src, dst, a, b, c, lo, hi, tgt, tgt_in, keep
        pairs = torch.unique(torch.cat([pairs, key]))
        rnd += 1
        assert rnd < 64, "generator failed to reach the requested nnz (graph too dense?)"
    if pairs.numel() > target:  # drop the surplus with the largest hash
        h = _mix(pairs ^ seed)
        order = torch.argsort(h)
        pairs = pairs[order[:target]]
        del h, order
    a = pairs // n
    b = pairs % n
    del pairs
    diag = torch.arange(n, dtype=torch.int64, device=dev)
    if shuffle:  # hide the planted block order behind a random relabelling
        hp = _mix(diag ^ (seed + 77))
        relabel = torch.empty(n, dtype=torch.int64, device=dev)
        relabel[torch.argsort(hp)] = diag
        a, b = relabel[a], relabel[b]
    keys = torch.cat([a * n + b, b * n + a, diag * n + diag])
    del a, b
    keys = torch.sort(keys).values
    rows = keys // n
    col = keys % n
    del keys
    deg = torch.bincount(rows, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(deg, 0)
    if values == "gcn":  # 1/sqrt(d_i d_j): the GCN normalisation pubmed.csv carries
        dinv = deg.to(torch.float64).rsqrt()
        val = (dinv[rows] * dinv[col]).to(torch.float32)
    else:  # 2*U-1 as DataLoader.cu:45 does for amazon.csv
        val = (2.0 * _uniform(seed, 9, nnz, dev) - 1.0).to(torch.float32)
    assert col.numel() == nnz
    return rowptr, col, val


def dense_B(n, k, seed=1, device="cpu"):
    """B[n,k] = 2*U-1 (the distribution of DataLoader.cu:205) from the counter hash."""
    return (2.0 * _uniform(seed, 1234, n * k, torch.device(device)) - 1.0).to(torch.float32).reshape(n, k)


def write_csv(path, rowptr, col, val):
    """3-line CSV in the reference's on-disk format (DataLoader.cu:19-53)."""
    import numpy as np
    with open(path, "w") as f:
        f.write(",".join(map(str, rowptr.tolist())) + "\n")
        f.write(",".join(map(str, col.tolist())) + "\n")
        f.write(",".join(np.format_float_positional(v, unique=True, trim="-") for v in
                         val.cpu().numpy()) + "\n")
