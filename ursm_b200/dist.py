"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

Cells and bulk samples are independent given (A, alpha, pkappa, ptau) inside an
E-step, so each rank owns a contiguous block of rows and the only
data-path collective is a sum all-reduce of the packed statistics vector
(NCCL over NVLink on GPUs; gloo in the CPU tests of this host logic).
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def shard_bounds(n, r=None, w=None):
    """Contiguous block [lo, hi) of `n` rows owned by rank r of w (balanced to +-1)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, extra = divmod(int(n), w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def all_reduce_(t):
    """In-place sum all-reduce of a torch tensor (no-op for a single process)."""
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t)
    return t


def all_reduce_np(a):
    """Sum all-reduce of a small numpy array / scalar through the active backend."""
    d = _dist()
    a = np.asarray(a, dtype=np.float64)
    if d is None or d.get_world_size() == 1:
        return a
    import torch
    dev = "cuda" if d.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(a).copy()).to(dev)
    d.all_reduce(t)
    return t.cpu().numpy().reshape(a.shape)


def require_rows_everywhere(local_rows, what):
    """Every rank must own at least one row of `what`: the E-step of a rank without rows has
    nothing to launch, and an exception on that rank alone would leave the others waiting in
    their collectives.  One small all-reduce decides, and EVERY rank raises the same error."""
    empty = all_reduce_np(np.array([1.0 if int(local_rows) == 0 else 0.0]))[0]
    if empty > 0:
        raise ValueError("%d of %d ranks would hold no %s: use at most as many ranks as there are rows"
                         % (int(empty), world_size(), what))


def gather_rows(a, total_rows=None):
    """Concatenate per-rank row blocks (axis 0) on every rank."""
    d = _dist()
    a = np.ascontiguousarray(a)
    if d is None or d.get_world_size() == 1:
        return a
    out = [None] * d.get_world_size()
    d.all_gather_object(out, a)
    return np.concatenate(out, axis=0)
