"""Candidate sharding across the GPUs of one box (one process per GPU).

Candidates are independent, so the M-point grid of `solve_lbfgs`
(reference solvers/lbfgs.py:45-51) is cut into contiguous blocks, one per rank;
the fit inputs (n x d observations, hyper-parameters) are tiny and every rank
refits redundantly.  The only exchange is the global incumbent: a MAX all-reduce
of the best score followed by a MIN all-reduce of the index among the ranks
that hold that score, which reproduces `argmax`'s first-index tie rule.
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


def reduce_incumbent(val, idx, device=None, group=None):
    """Global (max value, lowest global index attaining it) from per-rank bests.
    `idx` must already be a global index.  NaN scores never win."""
    if not is_distributed():
        return float(val), int(idx)
    import torch
    dist = _dist()
    dev = _comm_device(device)
    v = float(val)
    if v != v:
        v = -np.inf
    tv = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(tv, op=dist.ReduceOp.MAX, group=group)
    gmax = float(tv.item())
    ti = torch.tensor([int(idx) if v == gmax else _INT64_MAX], dtype=torch.int64, device=dev)
    dist.all_reduce(ti, op=dist.ReduceOp.MIN, group=group)
    return gmax, int(ti.item())


def reduce_incumbents(vals, idxs, device=None, group=None):
    """Vector form (e.g. one incumbent per Thompson draw): two all-reduces in total."""
    vals = np.array(vals, dtype=np.float64)
    idxs = np.array(idxs, dtype=np.int64)
    if not is_distributed():
        return vals, idxs
    import torch
    dist = _dist()
    dev = _comm_device(device)
    vals = np.where(np.isnan(vals), -np.inf, vals)
    tv = torch.from_numpy(vals.copy()).to(dev)
    dist.all_reduce(tv, op=dist.ReduceOp.MAX, group=group)
    gmax = tv.cpu().numpy()
    cand = np.where(vals == gmax, idxs, _INT64_MAX)
    ti = torch.from_numpy(cand).to(dev)
    dist.all_reduce(ti, op=dist.ReduceOp.MIN, group=group)
    return gmax, ti.cpu().numpy()


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
