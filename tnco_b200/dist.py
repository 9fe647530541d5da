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


_DEVICE = [None]


def set_device(device):
    """CUDA device the collectives of this process use (the engine's device: ``Optimizer(device=...)`` or LOCAL_RANK)."""
    _DEVICE[0] = None if device is None else int(device)


def _tensor(a):
    import torch
    d = _dist()
    t = torch.from_numpy(np.ascontiguousarray(a))
    if d and d.get_backend() == 'nccl':
        t = t.cuda(local_device(_DEVICE[0]))
    return t


KEY_ID_BITS = 24  # low bits of the packed key: global chain id (up to 16.7 M chains)


def pack_key(cost, chain_id):
    """Order-preserving 64-bit key of (cost, chain id) for ONE min-reduction: the high 40 bits of the
    fp64 bit pattern of a positive cost (monotone in the cost; 28 mantissa bits, 4e-9 relative) above the global chain
    id, so equal (truncated) costs resolve to the smaller chain id on every rank alike."""
    c = float(cost)
    if not (c >= 0.0) or c == float('inf'):
        return (1 << 63) - 1  # NaN / inf / negative: never wins
    bits = int(np.array([c], np.float64).view(np.uint64)[0])
    return ((bits >> KEY_ID_BITS) << KEY_ID_BITS) | (int(chain_id) & ((1 << KEY_ID_BITS) - 1))


def global_best(local_cost, local_chain_id, local_payload, payload_len=None):
    """ONE all_reduce(MIN) of the packed (cost, global chain id) key, then ONE broadcast from the winner's rank of
    {exact cost, payload} (payload: int32 array -- the best tree [+ slices]).  Returns (cost, payload, owner_rank,
    chain_id).  A rank without chains passes cost = inf.  Identity without a process group."""
    d = _dist()
    payload = np.asarray(local_payload, np.int32).reshape(-1)
    if d is None:
        return float(local_cost), payload, 0, int(local_chain_id)
    rank, size = world()
    mine = pack_key(local_cost, local_chain_id)
    k = _tensor(np.array([mine], np.int64))
    d.all_reduce(k, op=d.ReduceOp.MIN)
    best_key = int(k.cpu()[0])
    if best_key == (1 << 63) - 1:
        raise RuntimeError('global_best: no rank holds a finite cost')
    chain = best_key & ((1 << KEY_ID_BITS) - 1)
    n = int(payload_len if payload_len is not None else len(payload))
    owner = owner_of_chain(chain)
    assert 0 <= owner < size, 'no rank owns the winning chain'  
    msg = np.zeros(n + 2, np.int32)
    if rank == owner:
        assert mine == best_key, 'the winning key must come from the rank that owns the chain'
        msg[:2] = np.array([float(local_cost)], np.float64).view(np.int32)
        msg[2:] = payload
    t = _tensor(msg)
    d.broadcast(t, src=owner)
    out = t.cpu().numpy()
    return float(out[:2].view(np.float64)[0]), out[2:], owner, chain


_SHARD_TOTAL = [None]


def owner_of_chain(chain_id):
    """Rank owning global run `chain_id` under `shard` (needs the total set by `set_shard_total`)."""
    n = _SHARD_TOTAL[0]
    _, size = world()
    if n is None:
        raise RuntimeError('owner_of_chain: call set_shard_total(n_runs) first')
    for r in range(size):
        lo, hi = shard(n, r, size)
        if lo <= chain_id < hi:
            return r
    raise ValueError('chain id out of range')


def set_shard_total(n_items):
    _SHARD_TOTAL[0] = int(n_items)


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


def all_reduce_rows_sum(a):
    """Element-wise sum over ranks of an integer array (each element is filled by exactly one rank, zero elsewhere:
    the way the owners contribute the trees of the global top-k with one collective)."""
    d = _dist()
    a = np.ascontiguousarray(a)
    if d is None:
        return a
    signed = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64, np.dtype(np.uint16): np.int16}
    if a.dtype in signed:
        return all_reduce_rows_sum(a.view(signed[a.dtype])).view(a.dtype)
    t = _tensor(a)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.cpu().numpy()


def broadcast_int(value, src=0):
    """Rank `src`'s integer on every rank (< 2^62)."""
    d = _dist()
    if d is None:
        return int(value)
    t = _tensor(np.array([int(value)], np.int64))
    d.broadcast(t, src=src)
    return int(t.cpu()[0])


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
