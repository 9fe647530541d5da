"""Multi-GPU plumbing: one process per GPU (torchrun), chains sharded over ranks, no data-path collective.

The reference runs independent processes that never talk (tnco/parallel.py:330-341); the only exchange here is
the periodic / final *min-reduction of the best cost and broadcast of the winning tree* plus the gather of
per-run results, over torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        return dist if (dist.is_available() and dist.is_initialized()) else None
    except Exception:
        return None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def local_device(default=None):
    if default is not None:
        return int(default)
    return int(os.environ.get('LOCAL_RANK', '0'))


def shard(n_items, rank=None, size=None):
    """Contiguous block of run i -> rank floor(i*size/n); returns (lo, hi)."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    return (n_items * rank) // size, (n_items * (rank + 1)) // size


def _tensor(a):
    import torch
    d = _dist()
    t = torch.from_numpy(np.ascontiguousarray(a))
    if d and d.get_backend() == 'nccl':
        t = t.cuda(local_device())
    return t


def global_best(local_cost, local_payload):
    """all_reduce(MIN) of the best cost, then broadcast of the owner's payload (int32 array: the best tree
    [+ slices]).  Returns (cost, payload, owner_rank).  Identity without a process group."""
    d = _dist()
    if d is None:
        return float(local_cost), np.asarray(local_payload), 0
    import torch
    rank, size = world()
    c = _tensor(np.array([local_cost], np.float64))
    d.all_reduce(c, op=d.ReduceOp.MIN)
    best = float(c.cpu()[0])
    owner = _tensor(np.array([rank if float(local_cost) == best else size], np.int64))
    d.all_reduce(owner, op=d.ReduceOp.MIN)
    owner = int(owner.cpu()[0])
    payload = _tensor(np.asarray(local_payload, np.int32))
    d.broadcast(payload, src=owner)
    return best, payload.cpu().numpy(), owner


def all_gather_rows(local, n_total):
    """Concatenate per-rank row blocks (sharded with `shard`) into the full [n_total, ...] array on every rank."""
    d = _dist()
    local = np.ascontiguousarray(local)
    if d is None:
        return local
    import torch
    signed = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64, np.dtype(np.uint16): np.int16}
    if local.dtype in signed:  # collectives do not take unsigned types: ship the same bits as signed
        return all_gather_rows(local.view(signed[local.dtype]), n_total).view(local.dtype)
    rank, size = world()
    counts = [shard(n_total, r, size)[1] - shard(n_total, r, size)[0] for r in range(size)]
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = _tensor(pad)
    if all(c == mx for c in counts):
        # equal blocks (the usual case): gather into ONE tensor and read it back with one copy
        out = torch.empty((size * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        try:
            d.all_gather_into_tensor(out, t)
            return out.cpu().numpy()
        except (RuntimeError, NotImplementedError, AttributeError):
            pass  # backend without the flat form: the list form below
    outs = [torch.empty_like(t) for _ in range(size)]
    d.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def all_reduce_max(x):
    d = _dist()
    if d is None:
        return float(x)
    t = _tensor(np.array([x], np.float64))
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.cpu()[0])


def all_reduce_sum(x):
    d = _dist()
    if d is None:
        return float(x)
    t = _tensor(np.array([x], np.float64))
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.cpu()[0])
