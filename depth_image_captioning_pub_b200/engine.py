"""Host-side engine: device buffers, weight packing and the autograd bridge to libdic.so.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every
arithmetic step of the decoder path runs in the CUDA library behind include/dic.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DicError, Dims, Params, PARAM_FIELDS

PRECISIONS = {"fp32": _lib.DIC_F32, "float32": _lib.DIC_F32, "bf16": _lib.DIC_BF16, "bfloat16": _lib.DIC_BF16}


def batch_sizes_from_lengths(lengths: Sequence[int]) -> List[int]:
    """bs_valid per decoder step for caption lengths (incl. <start>) sorted descending
    (depth_models.py:170,182; util.py:95)."""
    dec = np.asarray([int(l) for l in lengths], dtype=np.int64) - 1
    if dec.size == 0 or dec.min() < 1:
        raise ValueError("every caption needs at least <start> and one target token")
    if np.any(dec[:-1] < dec[1:]):
        raise ValueError("lengths must be sorted in descending order (as collate_func does)")
    # number of captions still active at step t = #{b : dec[b] > t}
    counts = np.bincount(dec, minlength=int(dec.max()) + 1)
    active = dec.size - np.cumsum(counts)[:int(dec.max())]
    return [int(x) for x in active]


def _params_struct(tensors: Sequence[torch.Tensor]) -> Params:
    p = Params()
    for name, t in zip(PARAM_FIELDS, tensors):
        if t.dtype != torch.float32:
            raise DicError(f"parameter {name} must be float32, got {t.dtype}")
        setattr(p, name, _lib.ptr(t))
    return p


class Engine:
    """Per-(dims, precision, device) state: packed weights and reusable workspaces."""

    def __init__(self, L: int, D: int, A: int, E: int, H: int, V: int, precision: str, device: torch.device):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if device.type != "cuda":
            raise DicError("the decoder path runs on CUDA only (no CPU fallback)")
        self.lib = _lib.load()
        self.dims = Dims(L, D, A, E, H, V)
        self.dtype = PRECISIONS[precision]
        self.device = device
        nbytes = self.lib.dic_pack_bytes(C.byref(self.dims), self.dtype)
        if nbytes == 0:
            raise DicError(self.lib.dic_last_error().decode())
        self.pack = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self._pack_key = None
        self._ws: Dict[Tuple, torch.Tensor] = {}
        # optional: all parameter gradients written into ONE flat fp32 buffer (data-parallel training:
        # the NCCL all-reduce then runs on it in place, no gather / scatter copies)
        self.use_flat_grads = False
        self.grad_flat: Optional[torch.Tensor] = None
        self._grad_ptrs: Optional[set] = None      # data pointers of the views handed to autograd

    # ---- weights -------------------------------------------------------------------------
    def ensure_packed(self, params: Sequence[torch.Tensor], allow_cached: bool = False) -> None:
        """(Re)build the compute-layout weight pack.

        The pack is rebuilt on EVERY call by default: tensor version counters are not a safe
        change detector (fused optimizers such as AdamW(fused=True) update parameters without
        bumping them), and a stale pack would silently decode/train with old weights.  Callers
        that know the weights are frozen (inference loops) may pass allow_cached=True.
        """
        key = tuple((p.data_ptr(), p._version) for p in params)
        if allow_cached and key == self._pack_key:
            return
        self._pack_serial = getattr(self, "_pack_serial", 0) + 1
        with torch.cuda.device(self.device):
            ps = _params_struct([p.detach().contiguous() for p in params])
            _lib.check(self.lib.dic_pack_weights(C.byref(self.dims), self.dtype, C.byref(ps),
                                                 _lib.ptr(self.pack), _lib.stream_ptr(self.device)))
        self._pack_key = key

    # ---- workspaces ------------------------------------------------------------------------
    def train_workspace(self, B: int, T: int, fresh: bool) -> torch.Tensor:
        n = self.lib.dic_train_workspace_bytes(C.byref(self.dims), self.dtype, B, T)
        if n == 0:
            raise DicError(self.lib.dic_last_error().decode())
        if fresh:   # lives until the matching backward; never shared between graphs
            return torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._cached(("train", B, T), n)

    def decode_workspace(self, B: int, beam: int) -> torch.Tensor:
        n = self.lib.dic_decode_workspace_bytes(C.byref(self.dims), self.dtype, B, beam)
        if n == 0:
            raise DicError(self.lib.dic_last_error().decode())
        return self._cached(("decode", B, beam), n)

    def _cached(self, key, n):
        t = self._ws.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    # ---- raw calls ---------------------------------------------------------------------------
    def forward(self, attn_mode: int, f_rgb, f_depth, captions, batch_sizes: Sequence[int], u, temp: float,
                dropout_mask, ws, logits_dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
        d = self.dims
        B, T = f_rgb.shape[0], len(batch_sizes)
        total = int(sum(batch_sizes))
        logits = torch.empty(total, d.V, dtype=logits_dtype, device=self.device)
        # padded steps of ragged captions must read as zeros (depth_models.py:176,201); equal lengths write every row
        alloc = torch.zeros if batch_sizes[-1] < B else torch.empty
        alphas = alloc(B, T, d.L, dtype=torch.float32, device=self.device)
        bs = (C.c_int32 * T)(*batch_sizes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dic_decoder_forward_ex(
                C.byref(d), self.dtype, attn_mode, _lib.ptr(self.pack), _lib.ptr(f_rgb), _lib.ptr(f_depth),
                _lib.dtype_code(f_rgb), _lib.ptr(captions), captions.shape[1], bs, T, B, _lib.ptr(u),
                float(temp), _lib.ptr(dropout_mask), _lib.ptr(logits), _lib.dtype_code(logits), _lib.ptr(alphas),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self.device)))
        return logits, alphas

    def backward(self, attn_mode: int, f_rgb, f_depth, captions, batch_sizes, d_logits, d_alphas, alphas,
                 temp, dropout_mask, ws, param_shapes, need_dfeat: bool, params=None):
        """d_logits: float32, or the storage dtype of the mode (what the fused loss head writes)."""
        d = self.dims
        B, T = f_rgb.shape[0], len(batch_sizes)
        if self.use_flat_grads and not self._flat_buffer_is_live(params):
            grads = self._flat_grad_views(param_shapes)
        else:
            grads = [torch.empty(s, dtype=torch.float32, device=self.device) for s in param_shapes]
        d_feats = torch.empty(B, d.L, d.D, dtype=f_rgb.dtype, device=self.device) if need_dfeat else None
        gs = _params_struct(grads)
        bs = (C.c_int32 * T)(*batch_sizes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dic_decoder_backward_ex(
                C.byref(d), self.dtype, attn_mode, _lib.ptr(self.pack), _lib.ptr(f_rgb), _lib.ptr(f_depth),
                _lib.dtype_code(f_rgb), _lib.ptr(captions), captions.shape[1], bs, T, B, _lib.ptr(d_logits),
                _lib.dtype_code(d_logits), _lib.ptr(d_alphas), _lib.ptr(alphas), float(temp),
                _lib.ptr(dropout_mask), C.byref(gs), _lib.ptr(d_feats), _lib.ptr(ws), ws.numel(),
                _lib.stream_ptr(self.device)))
        return grads, d_feats

    def _flat_buffer_is_live(self, params) -> bool:
        """True when some p.grad still aliases the flat buffer (gradient accumulation, or
        zero_grad(set_to_none=False)): writing this backward's gradients into the buffer would overwrite the
        accumulated value and autograd's `p.grad += new` would then add the buffer to itself.  The backward
        falls back to freshly allocated gradients and autograd accumulates them INTO the flat buffer."""
        if self.grad_flat is None or params is None:
            return False
        lo = self.grad_flat.data_ptr()
        hi = lo + self.grad_flat.numel() * 4
        return any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params)

    def _flat_grad_views(self, param_shapes) -> List[torch.Tensor]:
        """Fresh view tensors into the persistent flat gradient buffer (autograd's AccumulateGrad takes
        them over as .grad without a copy, so p.grad aliases the flat buffer after backward)."""
        # every tensor starts on a 256-byte boundary: the kernels' 16-byte stores / vector red.add / AdamW's float4
        # path need aligned bases (full_att.bias has ONE element: packed back to back, every tensor behind it sat
        # on a 4-byte boundary and took the scalar paths -- 0.1 ms per step, profiles/r02_dp_timeline_*.txt)
        ALIGN = 64
        offs, o = [], 0
        for s in param_shapes:
            offs.append(o)
            o += (int(np.prod(s)) + ALIGN - 1) // ALIGN * ALIGN
        n = o
        if self.grad_flat is None or self.grad_flat.numel() != n:
            # a data-parallel wrapper may supply the buffer (peer-visible memory for the NVLink all-reduce)
            alloc = getattr(self, "flat_alloc", None)
            self.grad_flat = alloc(n, self.device) if alloc is not None else None
            if self.grad_flat is None:
                self.grad_flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad_offsets = offs
        views = []
        for s, o in zip(param_shapes, offs):
            views.append(self.grad_flat[o:o + int(np.prod(s))].view(s))
        # keep no reference to the views: AccumulateGrad only takes a gradient over without a copy when
        # nobody else holds it
        self._grad_ptrs = {v.data_ptr() for v in views}
        return views

    def caption_loss(self, logits, captions, batch_sizes, ignore_index: int, alphas, lam: float):
        """Fused CE + doubly-stochastic regulariser (dic_caption_loss).
        -> (loss [1] fp32, d_logits [N,V] in the storage dtype, d_alphas [B,T,L] fp32 or None)."""
        d = self.dims
        B, T = captions.shape[0], len(batch_sizes)
        N = int(sum(batch_sizes))
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        if _lib.dtype_code(logits) == self.dtype:
            d_logits = logits          # same dtype (fp32 mode, or bf16 logits in bf16 mode): in place
        else:
            d_logits = torch.empty(N, d.V, dtype=torch.bfloat16, device=self.device)
        use_reg = alphas is not None and lam != 0.0
        d_alphas = torch.empty_like(alphas) if use_reg else None
        n = self.lib.dic_caption_loss_workspace_bytes(N, B)
        ws = self._cached(("loss", N, B), n)
        bs = (C.c_int32 * T)(*batch_sizes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dic_caption_loss(
                C.byref(d), self.dtype, _lib.ptr(logits), _lib.dtype_code(logits), _lib.ptr(captions),
                captions.shape[1], bs, T, B,
                int(ignore_index), _lib.ptr(alphas) if use_reg else None, float(lam), _lib.ptr(loss),
                _lib.ptr(d_logits), _lib.ptr(d_alphas), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self.device)))
        return loss, d_logits, d_alphas

    def scale_loss_grads(self, grad_loss, d_logits, d_alphas):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dic_scale_loss_grads(
                self.dtype, _lib.ptr(grad_loss), _lib.ptr(d_logits), d_logits.numel(), _lib.ptr(d_alphas),
                d_alphas.numel() if d_alphas is not None else 0, _lib.stream_ptr(self.device)))

    def greedy(self, attn_mode: int, f_rgb, f_depth, start_id: int, max_len: int, u=None,
               want_alphas: bool = False, want_logits: bool = False):
        d = self.dims
        B = f_rgb.shape[0]
        tokens = torch.empty(B, max_len, dtype=torch.int64, device=self.device)
        alphas = torch.empty(max_len, B, d.L, dtype=torch.float32, device=self.device) if want_alphas else None
        logits = torch.empty(max_len, B, d.V, dtype=torch.float32, device=self.device) if want_logits else None
        ws = self.decode_workspace(B, 1)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dic_decode_greedy(
                C.byref(d), self.dtype, attn_mode, _lib.ptr(self.pack), _lib.ptr(f_rgb), _lib.ptr(f_depth),
                _lib.dtype_code(f_rgb), B, int(start_id), int(max_len), _lib.ptr(u), _lib.ptr(tokens),
                _lib.ptr(alphas), _lib.ptr(logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self.device)))
        return tokens, alphas, logits

    def beam(self, f_rgb, f_depth, start_id: int, end_id: int, beam: int, max_len: int, trace: bool = False,
             want_logits: bool = False):
        d = self.dims
        B = f_rgb.shape[0]
        dev = self.device
        tokens = torch.empty(B, max_len, dtype=torch.int64, device=dev)
        lengths = torch.empty(B, dtype=torch.int32, device=dev)
        scores = torch.empty(B, dtype=torch.float32, device=dev)
        back = toks = step_scores = lse = logits = None
        if trace:
            back = torch.empty(max_len, B, beam, dtype=torch.int32, device=dev)
            toks = torch.empty(max_len, B, beam, dtype=torch.int32, device=dev)
            step_scores = torch.empty(max_len, B, beam, dtype=torch.float32, device=dev)
            lse = torch.empty(max_len, B, beam, dtype=torch.float32, device=dev)
        if want_logits:
            logits = torch.empty(max_len, B * beam, d.V, dtype=torch.float32, device=dev)
        ws = self.decode_workspace(B, beam)
        with torch.cuda.device(dev):
            _lib.check(self.lib.dic_decode_beam(
                C.byref(d), self.dtype, _lib.ptr(self.pack), _lib.ptr(f_rgb), _lib.ptr(f_depth),
                _lib.dtype_code(f_rgb), B, int(beam), int(start_id), int(end_id), int(max_len), _lib.ptr(tokens),
                _lib.ptr(lengths), _lib.ptr(scores), _lib.ptr(back), _lib.ptr(toks), _lib.ptr(step_scores),
                _lib.ptr(lse), _lib.ptr(logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        out = dict(tokens=tokens, lengths=lengths, scores=scores)
        if trace:
            out.update(back=back, toks=toks, all_scores=step_scores, lse=lse)
        if want_logits:
            out["logits"] = logits
        return out


def _needs_grad(grad_enabled: bool, params, f_rgb, f_depth) -> bool:
    return bool(grad_enabled) and (any(p.requires_grad for p in params) or f_rgb.requires_grad
                                   or (f_depth is not None and f_depth.requires_grad))


class DecoderFunction(torch.autograd.Function):
    """Teacher-forced decoder forward/backward through dic_decoder_forward / _backward.

    Inputs: engine, grad_mode, captions, batch_sizes, u, temp, dropout_mask, f_rgb, f_depth, *17 params
    (grad_mode = (attn_mode, torch.is_grad_enabled() of the CALLER: grad mode is always off inside
    Function.forward, so it cannot be read here)).
    Outputs: packed logits [sum(bs), V], alphas [B, T, L].
    """

    @staticmethod
    def forward(ctx, engine: Engine, grad_mode, captions, batch_sizes, u, temp, dropout_mask, f_rgb,
                f_depth, *params):
        attn_mode, grad_enabled = grad_mode
        needs_grad = _needs_grad(grad_enabled, params, f_rgb, f_depth)
        engine.ensure_packed(params)
        # a graph that will be differentiated owns its workspace (saved activations) until its backward;
        # only no-grad forwards share the cached one
        ws = engine.train_workspace(f_rgb.shape[0], len(batch_sizes), fresh=needs_grad)
        logits, alphas = engine.forward(attn_mode, f_rgb, f_depth, captions, batch_sizes, u, temp,
                                        dropout_mask, ws)
        ctx.engine, ctx.attn_mode, ctx.batch_sizes, ctx.temp = engine, attn_mode, list(batch_sizes), temp
        ctx.ws = ws
        ctx.params = params if needs_grad else None
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.pack_key = engine._pack_key
        ctx.has_depth = f_depth is not None
        saved = [captions, f_rgb, alphas]
        for opt in (f_depth, u, dropout_mask):
            saved.append(opt if opt is not None else torch.empty(0, device=f_rgb.device))
        ctx.save_for_backward(*saved)
        return logits, alphas

    @staticmethod
    def backward(ctx, d_logits, d_alphas):
        engine: Engine = ctx.engine
        if engine._pack_key != ctx.pack_key:
            raise DicError("parameters were modified between forward and backward")
        captions, f_rgb, alphas, f_depth, u, dropout_mask = ctx.saved_tensors
        f_depth = f_depth if ctx.has_depth else None
        dropout_mask = dropout_mask if dropout_mask.numel() else None
        if d_logits is None:
            d_logits = torch.zeros(sum(ctx.batch_sizes), engine.dims.V, dtype=torch.float32, device=f_rgb.device)
        d_logits = d_logits.contiguous().float()
        if d_alphas is not None:
            d_alphas = d_alphas.contiguous().float()
        if ctx.ws is None:
            raise DicError("this graph was already differentiated: its workspace is released after the first "
                           "backward (retain_graph / double backward is not supported)")
        need_dfeat = ctx.needs_input_grad[7] or ctx.needs_input_grad[8]
        grads, d_feats = engine.backward(ctx.attn_mode, f_rgb, f_depth, captions, ctx.batch_sizes, d_logits,
                                         d_alphas, alphas, ctx.temp, dropout_mask, ctx.ws, ctx.param_shapes,
                                         need_dfeat, params=ctx.params)
        ctx.ws = None
        ctx.params = None
        d_rgb = d_feats if (ctx.needs_input_grad[7] and d_feats is not None) else None
        d_dep = None
        if ctx.has_depth and ctx.needs_input_grad[8] and d_feats is not None:
            d_dep = d_feats
        return (None, None, None, None, None, None, None, d_rgb, d_dep, *grads)


class CaptionLossFunction(torch.autograd.Function):
    """Teacher-forced forward + fused caption loss as ONE autograd node.

    loss = CE(packed logits, packed targets, ignore_index) + lam * mean((1 - sum_t alphas)^2)
    (depth_train.py:210-216).  The logits never reach the caller: the loss kernel turns them into
    d_logits (storage dtype) right away, and backward feeds those to dic_decoder_backward_ex.
    Inputs: engine, grad_mode, captions, batch_sizes, ignore_index, lam, u, temp, dropout_mask,
    f_rgb, f_depth, *17 params (grad_mode as in DecoderFunction).  Output: loss (0-d fp32).
    """

    @staticmethod
    def forward(ctx, engine: Engine, grad_mode, captions, batch_sizes, ignore_index, lam, u, temp,
                dropout_mask, f_rgb, f_depth, *params):
        attn_mode, grad_enabled = grad_mode
        needs_grad = _needs_grad(grad_enabled, params, f_rgb, f_depth)
        engine.ensure_packed(params)
        ws = engine.train_workspace(f_rgb.shape[0], len(batch_sizes), fresh=needs_grad)
        # bf16 mode keeps the [sum(bs), V] logits block in bf16 (within the mode's 2e-2 bound on the logits):
        # the loss head overwrites it with d_logits in place
        ldt = torch.bfloat16 if engine.dtype == _lib.DIC_BF16 else torch.float32
        logits, alphas = engine.forward(attn_mode, f_rgb, f_depth, captions, batch_sizes, u, temp,
                                        dropout_mask, ws, logits_dtype=ldt)
        loss, d_logits, d_alphas = engine.caption_loss(logits, captions, batch_sizes, ignore_index,
                                                       alphas if lam != 0.0 else None, lam)
        ctx.engine, ctx.attn_mode, ctx.batch_sizes, ctx.temp = engine, attn_mode, list(batch_sizes), temp
        ctx.ws = ws
        ctx.params = params if needs_grad else None
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.pack_key = engine._pack_key
        ctx.has_depth = f_depth is not None
        ctx.has_dalpha = d_alphas is not None
        saved = [captions, f_rgb, alphas, d_logits]
        for opt in (f_depth, u, dropout_mask, d_alphas):
            saved.append(opt if opt is not None else torch.empty(0, device=f_rgb.device))
        ctx.save_for_backward(*saved)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        engine: Engine = ctx.engine
        if engine._pack_key != ctx.pack_key:
            raise DicError("parameters were modified between forward and backward")
        captions, f_rgb, alphas, d_logits, f_depth, u, dropout_mask, d_alphas = ctx.saved_tensors
        f_depth = f_depth if ctx.has_depth else None
        dropout_mask = dropout_mask if dropout_mask.numel() else None
        d_alphas = d_alphas if ctx.has_dalpha else None
        if ctx.ws is None:
            # the saved d_logits / d_alphas were scaled in place and the workspace released by the first backward
            raise DicError("this graph was already differentiated (retain_graph / double backward is not "
                           "supported by the fused loss node)")
        g = grad_loss.detach().reshape(1).to(device=f_rgb.device, dtype=torch.float32).contiguous()
        engine.scale_loss_grads(g, d_logits, d_alphas)        # device-side no-op when the upstream gradient is 1
        need_dfeat = ctx.needs_input_grad[9] or ctx.needs_input_grad[10]
        grads, d_feats = engine.backward(ctx.attn_mode, f_rgb, f_depth, captions, ctx.batch_sizes, d_logits,
                                         d_alphas, alphas, ctx.temp, dropout_mask, ctx.ws, ctx.param_shapes,
                                         need_dfeat, params=ctx.params)
        ctx.ws = None
        ctx.params = None
        d_rgb = d_feats if (ctx.needs_input_grad[9] and d_feats is not None) else None
        d_dep = d_feats if (ctx.has_depth and ctx.needs_input_grad[10] and d_feats is not None) else None
        return (None,) * 9 + (d_rgb, d_dep, *grads)
