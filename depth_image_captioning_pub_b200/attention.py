"""Drop-in attention modules: same classes, constructor arguments, parameter names and
forward signatures as the reference's ``Captioning_models/attention.py``, computed by the
sm_100a kernels behind include/dic.h.

Inside the decoders these modules are parameter containers (the decoder-level entry points
fuse attention with the rest of the timestep); called on their own they run
``dic_attention_forward``.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import DicError, Dims

_PREC = {"fp32": _lib.DIC_F32, "bf16": _lib.DIC_BF16}


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, precision, u, temp, feats, h, ew, eb, dw, db, fw, fb):
        if not feats.is_cuda:
            raise DicError("attention runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        B, L, D = feats.shape
        A, H = ew.shape[0], dw.shape[1]
        dims = Dims(L, D, A, 8, H, 8)
        dt = _PREC[precision]
        dev = feats.device
        n = lib.dic_attention_workspace_bytes(C.byref(dims), dt, B)
        if n == 0:
            raise DicError(lib.dic_last_error().decode())
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        ctxv = torch.empty(B, D, dtype=torch.float32, device=dev)
        alpha = torch.empty(B, L, dtype=torch.float32, device=dev)
        args = [t.detach().contiguous().float() for t in (ew, eb, dw, db, fw, fb, feats, h)]
        with torch.cuda.device(dev):
            _lib.check(lib.dic_attention_forward(
                C.byref(dims), dt, mode, *[_lib.ptr(t) for t in args], B, _lib.ptr(u), float(temp),
                _lib.ptr(ctxv), _lib.ptr(alpha), _lib.ptr(ws), n, _lib.stream_ptr(dev)))
        return ctxv, alpha

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            "the stand-alone attention module is forward-only; train through the decoder modules "
            "(their backward covers the attention parameters)")


class Gumbel_softmax(nn.Module):
    """attention.py:6-48.  The uniform draw comes from the CPU generator exactly like the
    reference's ``torch.rand(batch_size, k)`` (attention.py:17,40) and is then moved to the device."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def draw(self, batch_size: int, device) -> torch.Tensor:
        return torch.rand(batch_size, self.k).to(device)

    def forward(self, logits, device, temp):
        raise DicError("Gumbel_softmax is fused into the attention kernel; call Hard_Attention instead")

    def Gumbel_maxtrick(self, logits, device):
        raise DicError("Gumbel_maxtrick is fused into the attention kernel; call Hard_Attention.Hard_sample")


class Soft_Attention(nn.Module):
    """attention.py:52-95: additive attention with a ReLU energy (attention.py:73,86-87)."""

    def __init__(self, dim_encoder: int, dim_decoder: int, dim_attention: int):
        super().__init__()
        self.encoder_att = nn.Linear(dim_encoder, dim_attention)
        self.decoder_att = nn.Linear(dim_decoder, dim_attention)
        self.full_att = nn.Linear(dim_attention, 1)
        self.relu = nn.ReLU(inplace=True)
        self.precision = "fp32"

    def _weights(self):
        return (self.encoder_att.weight, self.encoder_att.bias, self.decoder_att.weight,
                self.decoder_att.bias, self.full_att.weight, self.full_att.bias)

    def forward(self, encoder_out: torch.Tensor, decoder_hidden: torch.Tensor):
        return _AttnFn.apply(_lib.ATTN_SOFT, self.precision, None, 1.0, encoder_out, decoder_hidden,
                             *self._weights())


class Hard_Attention(nn.Module):
    """attention.py:99-167: Gumbel-softmax relaxation (training) and Gumbel-max one-hot sampling."""

    def __init__(self, dim_encoder: int, dim_decoder: int, dim_attention: int, k=196):
        super().__init__()
        self.encoder_att = nn.Linear(dim_encoder, dim_attention)
        self.decoder_att = nn.Linear(dim_decoder, dim_attention)
        self.full_att = nn.Linear(dim_attention, 1)
        self.relu = nn.ReLU(inplace=True)
        self.gumbel_softmax = Gumbel_softmax(k)
        self.precision = "fp32"

    def _weights(self):
        return (self.encoder_att.weight, self.encoder_att.bias, self.decoder_att.weight,
                self.decoder_att.bias, self.full_att.weight, self.full_att.bias)

    def forward(self, encoder_out: torch.Tensor, decoder_hidden: torch.Tensor, device: str,
                temp: torch.Tensor):
        u = self.gumbel_softmax.draw(encoder_out.shape[0], encoder_out.device)
        return _AttnFn.apply(_lib.ATTN_GUMBEL_SOFTMAX, self.precision, u, float(temp), encoder_out,
                             decoder_hidden, *self._weights())

    @torch.no_grad()
    def Hard_sample(self, encoder_out: torch.Tensor, decoder_hidden: torch.Tensor, device: str):
        u = self.gumbel_softmax.draw(encoder_out.shape[0], encoder_out.device)
        ctx, alpha = _AttnFn.apply(_lib.ATTN_GUMBEL_MAX, self.precision, u, 1.0, encoder_out,
                                   decoder_hidden, *self._weights())
        return ctx, alpha.to(torch.int64)   # F.one_hot returns int64 (attention.py:46)
