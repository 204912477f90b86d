"""Multi-GPU sharding of independent units (folds, predictor sets, Shapley permutations).

One process per GPU (torchrun), ``torch.distributed`` for plumbing.  The path has no intra-fit
exchange: raw X, y are replicated, every rank fits its share of the units, and the only collective is
one all-gather of the score shards at the end (NCCL over NVLink for CUDA tensors, gloo on CPU in
tests).  Sharding applies to the outermost driver only (Shapley > Hierarchical > Validator); nested
drivers run locally."""
from __future__ import annotations

import contextlib
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

_depth = 0


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sharding_active() -> bool:
    return _depth == 0 and world()[1] > 1


@contextlib.contextmanager
def sharded_region():
    """Marks the outermost sharded loop; nested drivers see sharding_active() == False."""
    global _depth
    _depth += 1
    try:
        yield
    finally:
        _depth -= 1


def my_units(n_units: int) -> List[int]:
    """Round-robin assignment of unit indices to this rank."""
    rank, ws = world()
    return list(range(rank, n_units, ws))


def owner(unit: int) -> int:
    return unit % world()[1]


def _comm_device(t: torch.Tensor) -> torch.device:
    if dist.get_backend() == 'nccl':
        return t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_units(local: Dict[int, torch.Tensor], n_units: int) -> List[torch.Tensor]:
    """All ranks receive every unit's tensor (same shape/dtype on all ranks), in unit order."""
    rank, ws = world()
    if ws == 1:
        return [local[i] for i in range(n_units)]
    sample = next(iter(local.values())) if local else None
    # agree on shape / dtype via rank 0's (or the first non-empty rank's) sample
    meta = [None]
    metas = [None] * ws
    dist.all_gather_object(metas, None if sample is None else (tuple(sample.shape), str(sample.dtype)))
    meta = next(m for m in metas if m is not None)
    shape, dtype = meta[0], getattr(torch, meta[1].split('.')[-1])
    out_dev = sample.device if sample is not None else torch.device('cpu')
    per = (n_units + ws - 1) // ws
    cdev = _comm_device(sample) if sample is not None else (torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu'))
    buf = torch.zeros((per,) + shape, dtype=dtype, device=cdev)
    for slot, unit in enumerate(range(rank, n_units, ws)):
        buf[slot].copy_(local[unit])
    full = torch.empty((ws * per,) + shape, dtype=dtype, device=cdev)
    dist.all_gather_into_tensor(full, buf)
    full = full.reshape((ws, per) + shape)
    return [full[u % ws, u // ws].to(out_dev) for u in range(n_units)]


def broadcast_object(obj, src: int = 0):
    if world()[1] == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def lpt_assign(costs: Sequence[float], n_ranks: int) -> List[int]:
    """Longest-processing-time-first assignment of units to ranks: owner[u] for every unit.  Deterministic (ties by unit
    index), so every rank computes the same table without communicating."""
    load = [0.0] * n_ranks
    owner = [0] * len(costs)
    for u in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(n_ranks), key=lambda j: (load[j], j))
        owner[u] = r
        load[r] += costs[u]
    return owner


def gather_assigned(local: Dict[int, torch.Tensor], n_units: int) -> List[torch.Tensor]:
    """All ranks receive every unit's tensor for an arbitrary unit -> rank assignment: each rank writes its units into a
    zero buffer and the buffers are summed (x + 0 is exact, so the result is bit-identical to the owner's tensor)."""
    rank, ws = world()
    if ws == 1:
        return [local[i] for i in range(n_units)]
    sample = next(iter(local.values())) if local else None
    metas = [None] * ws
    dist.all_gather_object(metas, None if sample is None else (tuple(sample.shape), str(sample.dtype)))
    meta = next(m for m in metas if m is not None)
    shape, dtype = meta[0], getattr(torch, meta[1].split('.')[-1])
    out_dev = sample.device if sample is not None else torch.device('cpu')
    cdev = _comm_device(sample) if sample is not None else (torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu'))
    buf = torch.zeros((n_units,) + shape, dtype=dtype, device=cdev)
    for u, t in local.items():
        buf[u].copy_(t)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return [buf[u].to(out_dev) for u in range(n_units)]
