"""Sample sharding and the one exchange step of the path.

The batch shards by sample (what `DistributedSampler` + `batch_size // world_size` already does in
the reference, train.py:271-280).  Every (b,k) volume is independent through the head kernels; the
only cross-rank dependency is that `torch.min(torch.stack(...))` (model.py:114,162) picks ONE slot
from batch means.  The reference evaluates that min on the rank-local batch (scope 'local', no
collective).  Scope 'global' reproduces the single-process result on the global batch with one
all-reduce(SUM) of the `[4, NH]` fp32 partial sums — a latency-only message.

Two transports for that message:
  * `PeerExchange` — the B200-native path: one tiny kernel of this library
    (`xsup_partial_allreduce`) publishes the sums into every peer's mailbox with NVLink P2P stores and
    a release flag, acquire-spins on its own mailbox and adds the slots in rank order.  It runs on the
    compute stream (no side stream; CUDA-graph replayable: the call sequence number is a device counter
    the kernel advances), measured 12 us median at 2 GPUs, 16-20 us at 8, and gives
    bit-identical sums on all ranks;
  * a `torch.distributed` process group (`nccl` on the GPUs — measured 13-20 us median for this
    message — and `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "is_active", "global_batch", "global_batch_exact", "reduce_partials", "select_slots", "PeerExchange"]


class PeerExchange:
    """NVLink peer-memory mailbox for the in-kernel all-reduce of the partial loss sums.

    Collective constructor (all ranks of `group`): allocates a zeroed mailbox in torch symmetric memory and
    rendezvous so that every rank holds the peer-mapped addresses of all mailboxes.  Pass the object as
    `group=` to `ops.integral_reproj_min_loss`."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi as cabi
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n = int(cabi.lib.xsup_xchg_floats(self.world))
        self.mailbox = symm_mem.empty(n, dtype=torch.float32, device=self.device)
        self.mailbox.zero_()
        self.handle = symm_mem.rendezvous(self.mailbox, self.group)
        self.peer_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                       # every mailbox is zeroed before anyone publishes
        # the call sequence number lives on the device and is advanced by the kernel, so a captured call can be replayed
        # word 0: sequence number; word 1: sticky error flag the kernel raises when a wait for a peer timed out
        self._words = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.seq = self._words[0:1]
        self.err = self._words[1:2]

    def descriptor(self):
        """The `xsup_xchg_t` of this mailbox (sequence number and error word live on the device)."""
        from . import _cabi as cabi
        return cabi.Xchg(self.peer_ptrs.data_ptr(), self.rank, self.world, 0, self.seq.data_ptr(), self.err.data_ptr())

    def check(self) -> None:
        """Raise if any exchange since construction timed out waiting for a peer (~10 s).  Reads one word back from the
        device (synchronises the stream), so call it at a step boundary you already synchronise on, not per kernel.
        After a timeout the ranks' sequence numbers are out of step: build a new PeerExchange."""
        if int(self.err.item()) != 0:
            raise RuntimeError("xsup_b200 PeerExchange: a peer did not arrive within the timeout; the sums of that step are NaN "
                               "and the mailboxes are out of step - rebuild the PeerExchange on all ranks")

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place SUM over the ranks of a small contiguous fp32 CUDA tensor (<= XSUP_XCHG_SLOT-1 = 1023 elements)."""
        from . import _cabi as cabi
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("PeerExchange.all_reduce_ needs a contiguous float32 CUDA tensor")
        x = self.descriptor()
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib.xsup_partial_allreduce(t.data_ptr(), t.numel(), x, cabi.stream_ptr(self.device)),
                       "xsup_partial_allreduce")
        return t


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `n` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def is_active(group) -> bool:
    """True when `group` spans more than one rank (a PeerExchange or an initialised torch.distributed group)."""
    return _active(group)


def _active(group) -> bool:
    if isinstance(group, PeerExchange):
        return group.world > 1
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _world(group) -> int:
    return group.world if isinstance(group, PeerExchange) else dist.get_world_size(group)


def global_batch(local_batch: int, group=None) -> int:
    """Total number of samples over the group (the denominator of the batch means in 'global' scope).
    Equal shards are assumed — `DistributedSampler` pads/drops to guarantee them (train.py:274-278) —
    so this is host arithmetic, not a collective.  Use `global_batch_exact` for ragged shards."""
    if not _active(group):
        return int(local_batch)
    return int(local_batch) * _world(group)


def global_batch_exact(local_batch: int, group=None) -> int:
    """Sum of the ranks' local batch sizes (one int64 all-reduce + host read)."""
    if not _active(group):
        return int(local_batch)
    if isinstance(group, PeerExchange):
        group = group.group
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([local_batch], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def reduce_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(SUM) of the per-hypothesis partial sums `[terms, NH]` (no-op without a group)."""
    if _active(group):
        if isinstance(group, PeerExchange):
            group.all_reduce_(partial)
        else:
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
