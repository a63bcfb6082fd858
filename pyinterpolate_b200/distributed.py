"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on the GPU box, gloo in the
CPU tests).  The two paths shard differently:

  * variogram — pair tiles are independent and the per-lag sums are additive.  Every rank holds all
    points (24 MB at 1M points), evaluates the ``rank``-th of ``world`` interleaved slices of the
    tile-pair list (``shard=(rank, world)`` of ``pib_variogram_*``) and ONE small all-reduce (SUM) of
    ``[3, D, L]`` fp64 sums + ``[D, L]`` int64 counts finishes the job (counts reduce exactly).
  * kriging — unknown locations are independent: each rank takes a contiguous slice of the
    locations, no collective on the data path; results are gathered only if the caller asks.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_initialized():
    try:
        dist = _dist()
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def rank_world():
    if is_initialized():
        dist = _dist()
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _collective_device(device=None):
    """Where collective buffers must live: NCCL only moves CUDA tensors, so with that backend host arrays are staged
    on the current CUDA device (gloo takes CPU tensors)."""
    if device is not None:
        return device
    import torch
    dist = _dist()
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return None


def allreduce_variogram_sums(sums, device=None, group=None):
    """SUM-all-reduce a dict(sum_sq, sum_prod, sum_val, count[, pairs_evaluated]) of numpy arrays or torch tensors in
    ONE collective: the three fp64 sums, the pair counts and the evaluated-pairs counter travel in one fp64 buffer —
    the counts are integers below 2^53, so their fp64 sum is exact whatever the reduction order.  Returns the same
    kind of containers it was given."""
    import torch
    rank, world = rank_world()
    if world == 1:
        return sums
    dist = _dist()
    as_numpy = isinstance(sums["count"], np.ndarray)
    device = _collective_device(device)
    to_t = (lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))) if as_numpy else (lambda a: a.to(torch.float64))
    shape = tuple(sums["count"].shape)
    n = int(np.prod(shape))
    parts = [to_t(sums[k]).reshape(-1) for k in ("sum_sq", "sum_prod", "sum_val", "count")]
    # the evaluated-pairs counter rides along only when the caller asked for it: reading it back is a host synchronisation
    has_pairs = bool(sums.get("pairs_evaluated"))
    if has_pairs:
        parts.append(torch.tensor([float(int(sums["pairs_evaluated"]))], dtype=torch.float64, device=parts[0].device))
    buf = torch.cat(parts)
    if device is not None:
        buf = buf.to(device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    cnt = buf[3 * n:4 * n].round().to(torch.int64)
    out = dict(sum_sq=buf[:n].reshape(shape), sum_prod=buf[n:2 * n].reshape(shape), sum_val=buf[2 * n:3 * n].reshape(shape),
               count=cnt.reshape(shape), pairs_evaluated=int(round(float(buf[-1].item()))) if has_pairs else 0)
    if as_numpy:
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    return out


def shard_bounds(m, rank=None, world=None):
    """Contiguous, balanced ``[begin, end)`` slice of ``m`` locations for this rank."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(int(m), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local_rows, m, group=None):
    """All-gather the per-rank result rows (sliced with ``shard_bounds``) back into ``[m, ...]`` order."""
    import torch
    rank, world = rank_world()
    if world == 1:
        return local_rows
    dist = _dist()
    as_numpy = isinstance(local_rows, np.ndarray)
    t = torch.as_tensor(local_rows) if as_numpy else local_rows
    dev = _collective_device(None if as_numpy else t.device)
    if dev is not None:
        t = t.to(dev)
    sizes = [shard_bounds(m, r, world) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(sizes)])
    return out.cpu().numpy() if as_numpy else out


def experimental_variogram_sharded(ds, **kwargs):
    """``variogram.experimental_curves`` with this rank's slice of the tile pairs + the all-reduce."""
    from .variogram import experimental_curves
    rank, world = rank_world()
    return experimental_curves(ds, shard=(rank, world), reduce_fn=allreduce_variogram_sums, **kwargs)


def ordinary_kriging_sharded(theoretical_model, known_locations, unknown_locations, gather=True, **kwargs):
    """Each rank kriges its slice of ``unknown_locations``; ``gather=False`` returns only the local rows."""
    from . import core
    from .kriging import ordinary_kriging
    unknown = core.interpolation_points(unknown_locations)
    b, e = shard_bounds(len(unknown))
    local = ordinary_kriging(theoretical_model, known_locations, unknown[b:e], **kwargs)
    return gather_rows(local, len(unknown)) if gather else local
