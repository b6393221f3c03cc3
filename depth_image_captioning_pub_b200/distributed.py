"""Multi-GPU plumbing for the decoder path (one process per GPU, torch.distributed over NCCL).

The path shards by batch:
  * decoding  -- images are independent: each rank decodes a contiguous slice of the batch with
    replicated weights and NO collective on the data path; an optional final all_gather collects
    the token ids.
  * training  -- data parallel: each rank runs the decoder on its own length-sorted sub-batch and
    the gradients of the 17 decoder tensors (plus any extra trainable tensors, e.g. the depth CNN)
    meet in ONE all-reduce of a flat fp32 buffer (19.3 MB for the decoder at V=10k), averaged.
The reference has no distributed code at all (config.py:68 pins a single 'cuda:0').
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sorted_batch(lengths: Sequence[int], rank: int, world: int) -> List[int]:
    """Indices of a length-sorted batch for `rank`, round-robin so that every rank's sub-batch is
    itself sorted descending (the bs_valid prefix property, depth_models.py:182) and the ranks get
    near-equal token counts."""
    return list(range(rank, len(lengths), world))


class FlatGradAllReduce:
    """One flat fp32 buffer for all gradients -> a single (NCCL) all-reduce per step.

    With `module=` (a decoder of this package) the library's backward writes the 17 parameter
    gradients straight into one flat buffer and `p.grad` aliases it, so the all-reduce (NCCL: AVG)
    runs in place: no gather / scatter copies.  Parameters whose gradients are not in that buffer
    (extra trainable tensors, or a step where the aliasing did not happen) take the copy path."""

    def __init__(self, params: Sequence[torch.Tensor], module=None):
        self.params = [p for p in params]
        self.module = module
        self._ready = None
        self._comm = None
        if module is not None:
            module.flat_grads = True
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def arm(self) -> None:
        """Call right before loss.backward(): the library then records three events while it enqueues the
        backward -- vocabulary-projection gradients final (before the time loop), everything but the
        encoder_att pair final, all final (before the dL/dF GEMM) -- and __call__ reduces the three matching
        slices of the flat buffer on a side stream from those points, overlapping the rest of backward."""
        if self.module is None or not torch.cuda.is_available():
            return
        from . import _lib
        if self._ready is None:
            self._ready = [torch.cuda.Event() for _ in range(3)]
            for ev in self._ready:
                ev.record()                         # instantiates the underlying cudaEvent_t
            self._comm = torch.cuda.Stream()
        self._armed = True
        _lib.load().dic_set_grads_ready_events(*[ev.cuda_event for ev in self._ready])

    def _buckets(self, flat: torch.Tensor):
        """(linear, middle, encoder_att) slices of the flat buffer, in the order their events fire."""
        sizes = [p.numel() for p in self.params]
        n = flat.numel()
        head, tail = sizes[0] + sizes[1], sizes[-2] + sizes[-1]
        return [flat[n - tail:], flat[head:n - tail], flat[:head]]

    def _in_place(self, average: bool, group) -> bool:
        m = self.module
        if m is None:
            return False
        engines = [e for e in getattr(m, "_engines", {}).values() if e.grad_flat is not None and e._grad_ptrs]
        if len(engines) != 1:
            return False
        e = engines[0]
        ptrs = e._grad_ptrs
        grads = [p.grad for p in self.params]
        if any(g is None or g.data_ptr() not in ptrs for g in grads) or len(grads) != len(ptrs):
            return False

        def reduce(t):
            if average and dist.get_backend(group) == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                if average:
                    t.mul_(1.0 / dist.get_world_size(group))
        if getattr(self, "_armed", False) and e.grad_flat.is_cuda and len(self.params) >= 5:
            self._armed = False
            main = torch.cuda.current_stream()
            with torch.cuda.stream(self._comm):
                for ev, part in zip(self._ready, self._buckets(e.grad_flat)):
                    self._comm.wait_event(ev)       # recorded by the library inside this step's backward
                    reduce(part)
            main.wait_stream(self._comm)                # the optimizer step comes after the all-reduce
        else:
            reduce(e.grad_flat)
        return True

    def __call__(self, average: bool = True, group=None) -> None:
        if self._in_place(average, group):
            return
        world = dist.get_world_size(group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def dp_loss_weight(local_count: int, group=None) -> float:
    """Weight that turns per-rank MEAN losses into the global mean under an averaging all-reduce:
    world * local_count / global_count (token-mean CE, depth_train.py:214)."""
    t = torch.tensor([float(local_count)])
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return dist.get_world_size(group) * float(local_count) / float(t.item())


def gather_tokens(tokens: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """All-gather equally sized [B/n, T] int64 token blocks (rank order = batch order)."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(tokens) for _ in range(world)]
    dist.all_gather(out, tokens.contiguous(), group=group)
    return torch.cat(out, dim=0)
