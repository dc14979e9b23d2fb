"""Candidate sharding across the GPUs of one box (one process per GPU).

Candidates are independent, so the M-point grid of `solve_lbfgs`
(reference solvers/lbfgs.py:45-51) is cut into contiguous blocks, one per rank;
the fit inputs (n x d observations, hyper-parameters) are tiny and every rank
refits redundantly.  The only exchange is the global incumbent: ONE all-reduce
(a sum over a [world, k, 3] tensor in which every rank fills only its own row),
after which each rank picks the maximum with `argmax`'s first-index tie rule.
`torch.distributed` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""

import numpy as np

_INT64_MAX = np.iinfo(np.int64).max


def shard_range(M, rank, world):
    """Contiguous block [lo, hi) of M items owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed():
    try:
        dist = _dist()
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except ImportError:
        return False


def _comm_device(device=None):
    import torch
    dist = _dist()
    if device is not None:
        return torch.device(device)
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_incumbents(vals, idxs, device=None, group=None):
    """Global (max value, lowest global index attaining it) for each of k incumbents (k = 1 for a
    scoring pass, k = #draws for Thompson) with ONE all-reduce: every rank writes its (value, index)
    pairs into its own row of a zero [world, k, 3] float64 tensor and the rows are summed (adding
    zeros is exact; indices are exact in float64 up to 2^53), after which every rank holds all
    candidates and takes the arg max with the first-index tie rule locally.  NaN never wins."""
    vals = np.array(vals, dtype=np.float64, ndmin=1)
    idxs = np.array(idxs, dtype=np.int64, ndmin=1)
    vals = np.where(np.isnan(vals), -np.inf, vals)
    if not is_distributed():
        return vals, idxs
    import torch
    dist = _dist()
    dev = _comm_device(device)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = len(vals)
    finite = np.isfinite(vals)
    buf = np.zeros((world, k, 3), dtype=np.float64)
    buf[rank, :, 0] = np.where(finite, vals, 0.0)          # +-inf would poison the sum: flag it instead
    buf[rank, :, 1] = idxs.astype(np.float64)
    buf[rank, :, 2] = np.where(finite, 0.0, np.where(vals > 0, 1.0, -1.0))
    t = torch.from_numpy(buf).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    allb = t.cpu().numpy()
    v = np.where(allb[:, :, 2] == 0.0, allb[:, :, 0], np.where(allb[:, :, 2] > 0, np.inf, -np.inf))
    ix = allb[:, :, 1].astype(np.int64)
    gmax = v.max(axis=0)
    cand = np.where(v == gmax[None, :], ix, _INT64_MAX)
    return gmax, cand.min(axis=0)


def reduce_incumbent(val, idx, device=None, group=None):
    """Scalar form of `reduce_incumbents`.  `idx` must already be a global index."""
    v, i = reduce_incumbents([val], [idx], device=device, group=group)
    return float(v[0]), int(i[0])


def gather_topk(vals, idxs, k, device=None, group=None):
    """Merge per-rank top-k lists (global indices) into the global top-k,
    value descending, lowest index first on ties (lbfgs.py:51 across shards)."""
    vals = np.array(vals, dtype=np.float64)
    idxs = np.array(idxs, dtype=np.int64)
    if is_distributed():
        import torch
        dist = _dist()
        dev = _comm_device(device)
        world = dist.get_world_size(group)
        pad = k - len(vals)
        if pad > 0:
            vals = np.concatenate([vals, np.full(pad, -np.inf)])
            idxs = np.concatenate([idxs, np.full(pad, _INT64_MAX)])
        tv = torch.from_numpy(vals[:k].copy()).to(dev)
        ti = torch.from_numpy(idxs[:k].copy()).to(dev)
        gv = torch.empty(world * k, dtype=torch.float64, device=dev)
        gi = torch.empty(world * k, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gv, tv, group=group)
        dist.all_gather_into_tensor(gi, ti, group=group)
        vals, idxs = gv.cpu().numpy(), gi.cpu().numpy()
    keep = idxs != _INT64_MAX
    vals, idxs = vals[keep], idxs[keep]
    order = np.lexsort((idxs, -vals))[:k]
    return vals[order], idxs[order]


class ShardedIndex(object):
    """Wrap a device-backed `ModelIndex` so that `best_of` scores this rank's
    block of the grid and merges the top-k over ranks.  Every rank gets the same
    answer, so the L-BFGS refinement that follows stays replicated."""

    fused = True

    def __init__(self, index, rank=None, world=None):
        dist = _dist()
        self.index = index
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world

    def __call__(self, X, grad=False):
        return self.index(X, grad=grad)

    def best_of(self, X, k):
        lo, hi = shard_range(len(X), self.rank, self.world)
        if hi > lo:
            idx, val = self.index.best_of(X[lo:hi], k)
            idx = idx + lo
        else:
            idx, val = np.zeros(0, dtype=np.int64), np.zeros(0)
        val, idx = gather_topk(val, idx, k)
        return idx, val

    def best_of_sobol(self, bounds, M, k, start=0):
        """Sharded form of `ModelIndex.best_of_sobol`: every rank generates and scores its own contiguous
        block of points [start, start + M) of the Sobol sequence on its device -- no candidate array exists
        on any host -- and the per-rank top-k lists (global sequence indices) are merged.  Returns
        (points (k, d), values (k,), sequence indices (k,)), identical on every rank."""
        from ._lib import sobol_points
        lo, hi = shard_range(int(M), self.rank, self.world)
        if hi > lo:
            _, val, idx = self.index.best_of_sobol(bounds, hi - lo, k, start=int(start) + lo)
        else:
            idx, val = np.zeros(0, dtype=np.int64), np.zeros(0)
        val, idx = gather_topk(val, idx, k)
        b = np.array(bounds, dtype=np.float64, ndmin=2)
        return sobol_points(b.shape[0], idx, b), val, idx


# ----------------------------------------------------------------------------------------------
# on-stream exchange: the incumbent never visits the host before the collective
# ----------------------------------------------------------------------------------------------

class _DevRecords(object):
    """Zero-copy view of `count` int64 words at a raw device address (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = dict(shape=(int(count),), typestr="<i8", data=(int(ptr), False), version=2)


def exchange_incumbents(ctx, record_ptr, k, group=None):
    """Global (values (k,), indices (k,)) from the packed device records `bo_score_incumbent` /
    `bo_thompson_incumbents` left in the handle `ctx`: ONE all-gather of 16 k bytes per rank, issued on the
    handle's own stream right behind the kernels that wrote the records (no host staging), then one merge
    kernel (`bo_incumbent_merge`) and a single read-back of the k results.  Without an initialised NCCL
    group the local records are merged alone."""
    if not is_distributed() or _dist().get_backend(group) != "nccl":
        if is_distributed():
            # CPU process groups (gloo, tests): records go through the host path
            val, idx = ctx.incumbent_merge(record_ptr, 1, k)
            return reduce_incumbents(val, idx, group=group)
        return ctx.incumbent_merge(record_ptr, 1, k)
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    with torch.cuda.device(dev), torch.cuda.stream(stream):
        mine = torch.as_tensor(_DevRecords(record_ptr, 2 * k), device=dev)
        allr = torch.empty(world * 2 * k, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allr, mine, group=group)
        out = ctx.incumbent_merge(allr.data_ptr(), world, k)      # same stream; synchronises before returning
    return out


class ShardedThompson(object):
    """BASELINE config 4 across ranks: every rank holds the same draws (same seed), evaluates its contiguous
    block of the candidates and the per-draw arg max is exchanged as ndraw packed records."""

    def __init__(self, batch, rank=None, world=None):
        self.batch = batch
        if rank is None or world is None:
            dist = _dist()
            rank, world = dist.get_rank(), dist.get_world_size()
        self.rank, self.world = rank, world

    def argmax(self, X):
        """(values (ndraw,), global indices (ndraw,)) over all ranks' blocks of the host array X."""
        lo, hi = shard_range(len(X), self.rank, self.world)
        ctx = self.batch._context()
        rec, nd = ctx.thompson_incumbents(hi - lo, np.ascontiguousarray(X[lo:hi]), offset=lo, flags=0)
        return exchange_incumbents(ctx, rec, nd)

    def argmax_device(self, M, xc_ptr, offset):
        """Same with this rank's block already on the device (`offset` = its first global index)."""
        ctx = self.batch._context()
        rec, nd = ctx.thompson_incumbents(M, xc_ptr, offset=offset)
        return exchange_incumbents(ctx, rec, nd)
