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


class PeerBuffer:
    """A flat fp32 buffer in peer-visible (symmetric) memory + the flag block of the library's NVLink
    all-reduce (csrc/dp_allreduce.cuh, dic_dp_allreduce).  Collective: every rank of `group` must construct it
    at the same point with the same n."""

    def __init__(self, n: int, device: torch.device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("the NVLink all-reduce serves one node (<= 8 ranks)")
        quantum = 4 * self.world
        self.n = n
        self.n_pad = (n + quantum - 1) // quantum * quantum
        flag_words = int(lib.dic_dp_flag_bytes()) // 4
        self.base = symm.empty(self.n_pad + flag_words, dtype=torch.float32, device=device)
        self.base.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.base, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._bufs = (C.c_void_p * self.world)(*ptrs)
        self._flags = (C.c_void_p * self.world)(*[p + 4 * self.n_pad for p in ptrs])
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        import os
        # NVLS (multimem.ld_reduce / multimem.st through the switch) wins from 4 ranks up, plain peer loads below:
        # 2 GPUs 48 us vs 70 us per 19.3 MB, 8 GPUs 80 us vs 66 us (profiles/r02_dp_allreduce_bench_n{2,8}.txt);
        # 8-GPU step 2.853 ms with it, 2.880 without, 2.857 with no exchange at all (profiles/r02_dp_modes_n8.txt)
        want = os.environ.get("DIC_DP_MULTIMEM", "auto")
        self.multicast = mc if (want == "1" or (want == "auto" and self.world >= 4)) else 0
        self.flat = self.base[:n]
        self.epoch = 0
        self.blocks = int(os.environ.get("DIC_DP_BLOCKS", "20"))      # = the SMs the dL/dF GEMM leaves free (kAllReduceSms)
        self._lib = lib
        dist.barrier(group)          # every rank's flags are zero before anyone's first all-reduce can write to them

    def all_reduce(self, average: bool = True) -> None:
        """In place on the current stream: flat <- (sum over ranks) (/ world)."""
        from . import _lib
        self.epoch += 1
        _lib.check(self._lib.dic_dp_allreduce(
            self.world, self.rank, self._bufs, self._flags, self.multicast, self.n_pad,
            (1.0 / self.world) if average else 1.0, self.epoch, self.blocks, _lib.stream_ptr(self.base.device)))


class FlatGradAllReduce:
    """One flat fp32 buffer for all gradients -> a single (NCCL) all-reduce per step.

    With `module=` (a decoder of this package) the library's backward writes the 17 parameter
    gradients straight into one flat buffer and `p.grad` aliases it, so the all-reduce (NCCL: AVG)
    runs in place: no gather / scatter copies.  Parameters whose gradients are not in that buffer
    (extra trainable tensors, or a step where the aliasing did not happen) take the copy path."""

    def __init__(self, params: Sequence[torch.Tensor], module=None, mode: Optional[str] = None, group=None):
        """mode: "p2p_overlap" (default on NCCL process groups: the library's own all-reduce over NVLink peer
        memory, dic_dp_allreduce, on a side stream next to the dL/dF GEMM, which leaves it 20 SMs), "p2p" (the same
        kernel after the backward on the main stream), "nccl" (three bucketed ncclAllReduce calls on a side stream),
        "none" (no exchange; diagnosis).  Measured on 2 B200 (profiles/r02_dp_modes_n2.txt), 2.825 ms/step alone:
        none 2.853, p2p 2.871, p2p_overlap 2.842, nccl 2.876 ms/step."""
        import os
        self.params = [p for p in params]
        self.module = module
        self.group = group
        self._ready = None
        self._comm = None
        self._peer: Optional[PeerBuffer] = None
        if mode is None:
            mode = os.environ.get("DIC_DP_MODE", "")
        if not mode:
            nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
            mode = "p2p_overlap" if (nccl and module is not None) else "nccl"
        self.mode = mode
        if module is not None:
            module.flat_grads = True
            if mode.startswith("p2p"):
                module.flat_alloc = self._alloc_peer
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def _alloc_peer(self, n: int, device: torch.device):
        """Engine hook: the flat gradient buffer comes from peer-visible memory.  A failure to set that up (no
        peer access, symmetric memory unavailable) is reported once and the NCCL path takes over."""
        try:
            self._peer = PeerBuffer(n, device, self.group)
        except Exception as e:          # noqa: BLE001  (every rank sees the same failure: the set-up is collective)
            import warnings
            warnings.warn(f"NVLink all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
            self._peer = None
            self.mode = "nccl"
            return None
        return self._peer.flat

    def arm(self) -> None:
        """Call right before loss.backward(): the library then records three events while it enqueues the
        backward -- vocabulary-projection gradients final (before the time loop), everything but the
        encoder_att pair final, all final (before the dL/dF GEMM) -- and __call__ reduces the three matching
        slices of the flat buffer on a side stream from those points, overlapping the rest of backward."""
        if self.module is None or not torch.cuda.is_available():
            return
        if self.mode in ("p2p", "plain"):
            return                                  # the reduce follows the backward on the same stream
        from . import _lib
        if self._ready is None:
            self._ready = [torch.cuda.Event() for _ in range(3)]
            for ev in self._ready:
                ev.record()                         # instantiates the underlying cudaEvent_t
            self._comm = torch.cuda.Stream()
        self._armed = True
        if self.mode == "p2p_overlap":
            _lib.load().dic_set_grads_ready_events(None, None, self._ready[2].cuda_event)
        else:
            _lib.load().dic_set_grads_ready_events(*[ev.cuda_event for ev in self._ready])

    def _buckets(self, flat: torch.Tensor, offsets):
        """(linear, middle, encoder_att) slices of the flat buffer, in the order their events fire."""
        head, tail = offsets[2], offsets[-2]
        return [flat[tail:], flat[head:tail], flat[:head]]

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
        peer = self._peer
        if peer is not None and peer.flat.data_ptr() == e.grad_flat.data_ptr() and self.mode.startswith("p2p"):
            if self.mode == "p2p_overlap" and getattr(self, "_armed", False):
                self._armed = False
                main = torch.cuda.current_stream()
                with torch.cuda.stream(self._comm):
                    self._comm.wait_event(self._ready[2])
                    peer.all_reduce(average)
                main.wait_stream(self._comm)
            else:
                peer.all_reduce(average)
        elif getattr(self, "_armed", False) and e.grad_flat.is_cuda and len(self.params) >= 5:
            self._armed = False
            main = torch.cuda.current_stream()
            with torch.cuda.stream(self._comm):
                for ev, part in zip(self._ready, self._buckets(e.grad_flat, e.grad_offsets)):
                    self._comm.wait_event(ev)       # recorded by the library inside this step's backward
                    reduce(part)
            main.wait_stream(self._comm)                # the optimizer step comes after the all-reduce
        else:
            reduce(e.grad_flat)
        return True

    def __call__(self, average: bool = True, group=None) -> None:
        if self.mode == "none":          # diagnosis only: no exchange at all (what does being one of N ranks cost?)
            return
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
