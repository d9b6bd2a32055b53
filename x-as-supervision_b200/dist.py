"""Sample sharding and the one exchange step of the path.

The batch shards by sample (what `DistributedSampler` + `batch_size // world_size` already does in
the reference, train.py:271-280).  Every (b,k) volume is independent through the head kernels; the
only cross-rank dependency is that `torch.min(torch.stack(...))` (model.py:114,162) picks ONE slot
from batch means.  The reference evaluates that min on the rank-local batch (scope 'local', no
collective).  Scope 'global' reproduces the single-process result on the global batch with one
all-reduce(SUM) of the `[4, NH]` fp32 partial sums — a latency-only message.

Works with any backend (`nccl` on the GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "global_batch", "reduce_partials", "select_slots"]


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `n` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _active(group) -> bool:
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def global_batch(local_batch: int, group=None) -> int:
    """Total number of samples over the group (the denominator of the batch means in 'global' scope).
    Equal shards are assumed — `DistributedSampler` pads/drops to guarantee them (train.py:274-278) —
    so this is host arithmetic, not a collective.  Use `global_batch_exact` for ragged shards."""
    if not _active(group):
        return int(local_batch)
    return int(local_batch) * dist.get_world_size(group)


def global_batch_exact(local_batch: int, group=None) -> int:
    """Sum of the ranks' local batch sizes (one int64 all-reduce + host read)."""
    if not _active(group):
        return int(local_batch)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([local_batch], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def reduce_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(SUM) of the per-hypothesis partial sums `[terms, NH]` (no-op without a group)."""
    if _active(group):
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def select_slots(partial: torch.Tensor, n_total: int, num_kp: int, w_mse: float, w_bone: Optional[float],
                 w_kp: Optional[float], w_kp2d: Optional[float]) -> Tuple[int, int, float, float]:
    """Host-side statement of what `xsup_reproj_select` computes in 'batch' mode from the reduced
    partial sums: (pseudo slot, symmetry slot or -1, pseudo loss, symmetry loss).  Used by the
    multi-rank CPU tests and for logging; the product path selects on the device."""
    p = partial.detach().to("cpu", torch.float64)
    n = float(n_total)
    mse = p[0] / (n * num_kp * 3)
    sm = int(torch.argmin(mse))
    use_sym = any(w is not None for w in (w_bone, w_kp, w_kp2d))
    if not use_sym:
        return sm, -1, float(w_mse * mse[sm]), 0.0
    sym = (w_bone or 0.0) * p[1] / (n * 4) + (w_kp or 0.0) * p[2] / (n * 6) + (w_kp2d or 0.0) * 1e2 * p[3] / (n * 4)
    ss = int(torch.argmin(sym))
    return sm, ss, float(w_mse * mse[sm]), float(sym[ss])
