"""Fused multi-tensor AdamW through the CUDA library.

Drop-in for the reference's `torch.optim.AdamW(decoder.parameters() + depth_encoder.parameters(),
lr=1e-3)` (depth_train.py:136-137, stepped at :221): same constructor arguments, same update rule
(decoupled weight decay, bias-corrected moments, no amsgrad), same state_dict keys (`step`,
`exp_avg`, `exp_avg_sq`).  All fp32 CUDA parameters of one device are updated by one kernel
launch per 32 tensors (dic_adamw_step); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import DicError


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            by_dev = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise DicError("FusedAdamW updates float32 CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise DicError("FusedAdamW does not support sparse gradients")
                if not p.is_contiguous():
                    raise DicError("parameters must be contiguous")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                by_dev.setdefault((p.device, st["step"]), []).append((p, p.grad.contiguous(), st))
            b1, b2 = group["betas"]
            for (dev, step), items in by_dev.items():
                n = len(items)
                arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
                ps, gs = arr([i[0] for i in items]), arr([i[1] for i in items])
                ms, vs = arr([i[2]["exp_avg"] for i in items]), arr([i[2]["exp_avg_sq"] for i in items])
                sizes = (C.c_longlong * n)(*[i[0].numel() for i in items])
                with torch.cuda.device(dev):
                    _lib.check(lib.dic_adamw_step(n, ps, gs, ms, vs, sizes, float(group["lr"]), float(b1), float(b2),
                                                  float(group["eps"]), float(group["weight_decay"]), int(step),
                                                  _lib.stream_ptr(dev)))
        return loss
