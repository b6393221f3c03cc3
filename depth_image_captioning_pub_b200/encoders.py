"""Drop-in for the reference depth CNN encoder.

``Depth_CNN_endoder`` keeps the reference's (misspelt) class name, constructor argument, sub-module names and order
(depth_models.py:12-47), so ``state_dict()`` has the same keys and shapes (``conv1.*``, ``bn1.*`` ... and their
``features.N.*`` aliases) and checkpoints written by depth_train.py:306-322 load unchanged.  The sub-modules only hold
parameters and buffers: forward and backward run in the CUDA library (csrc/depth_encoder.cuh,
dic_depth_encoder_forward / _backward) and the annotations come out as [B, 196, 2048] in the decoder's layout and,
in bf16 precision, its storage dtype.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import DicError


class _EncParams(C.Structure):
    _fields_ = [(f"{n}{i}", C.c_void_p) for i in (1, 2, 3) for n in ("conv_w", "conv_b", "bn_w", "bn_b", "bn_mean", "bn_var")]


def _struct(tensors) -> _EncParams:
    s = _EncParams()
    for (name, _), t in zip(_EncParams._fields_, tensors):
        setattr(s, name, _lib.ptr(t))
    return s


class _DepthEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imgs, module, *params):
        lib = _lib.load()
        dev = imgs.device
        B, _, Hi, Wi = imgs.shape
        dtype = _lib.DIC_BF16 if module.precision == "bf16" else _lib.DIC_F32
        feat_dtype = torch.bfloat16 if module.precision == "bf16" else torch.float32
        bns = (module.bn1, module.bn2, module.bn3)
        tensors = []
        for i in range(3):
            tensors += list(params[4 * i:4 * i + 4]) + [bns[i].running_mean, bns[i].running_var]
        nbytes = int(lib.dic_depth_encoder_workspace_bytes(B, Hi, Wi, dtype))
        if nbytes == 0:
            raise DicError(lib.dic_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        feats = torch.empty(B, 196, 2048, dtype=feat_dtype, device=dev)
        ps = _struct(tensors)
        training = bool(module.training)
        with torch.cuda.device(dev):
            _lib.check(lib.dic_depth_encoder_forward(
                dtype, int(training), B, Hi, Wi, _lib.ptr(imgs), C.byref(ps), float(module.bn1.momentum),
                float(module.bn1.eps), _lib.ptr(feats), _lib.dtype_code(feats), _lib.ptr(ws), nbytes,
                _lib.stream_ptr(dev)))
        ctx.module, ctx.ws, ctx.tensors, ctx.shape, ctx.dtype, ctx.training = module, ws, tensors, (B, Hi, Wi), dtype, training
        return feats

    @staticmethod
    def backward(ctx, d_feats):
        if not ctx.training:
            raise DicError("depth encoder: backward through eval-mode batch norm is not built (the reference trains "
                           "the encoder in train() mode, depth_train.py:196)")
        lib = _lib.load()
        B, Hi, Wi = ctx.shape
        d_feats = d_feats.contiguous()
        grads = [torch.empty_like(t) for t in ctx.tensors]
        ps, gs = _struct(ctx.tensors), _struct(grads)
        dev = d_feats.device
        with torch.cuda.device(dev):
            _lib.check(lib.dic_depth_encoder_backward(
                ctx.dtype, B, Hi, Wi, C.byref(ps), _lib.ptr(d_feats), _lib.dtype_code(d_feats), C.byref(gs),
                _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.stream_ptr(dev)))
        out = []
        for i in range(3):
            out += grads[6 * i:6 * i + 4]
        return (None, None, *out)


class Depth_CNN_endoder(nn.Module):
    """depth_models.py:12-56: depth map [B, 1, 224, 224] -> annotations [B, 196, 2048]."""

    def __init__(self, encoded_img_size: int):
        super().__init__()
        if encoded_img_size != 14:
            raise ValueError("the CUDA encoder builds AdaptiveAvgPool2d(14) as the exact 2 x 2 replication of the 7 x 7 "
                             "feature map (config.py:16 uses 14)")
        # same construction order as the reference: same seed -> same initial weights
        self.conv1 = nn.Conv2d(1, 128, 7, stride=3)
        self.bn1 = nn.BatchNorm2d(128)
        self.conv2 = nn.Conv2d(128, 512, 3)
        self.bn2 = nn.BatchNorm2d(512)
        self.conv3 = nn.Conv2d(512, 2048, 1)
        self.bn3 = nn.BatchNorm2d(2048)
        self.avg_pool = nn.AdaptiveAvgPool2d(encoded_img_size)
        self.max_pool = nn.MaxPool2d((3, 3))
        self.relu = nn.ReLU(inplace=True)
        self.features = nn.Sequential(self.conv1, self.bn1, self.relu, self.max_pool, self.conv2, self.bn2, self.relu,
                                      self.max_pool, self.conv3, self.bn3, self.relu, self.avg_pool)
        import os
        self.precision = os.environ.get("DIC_PRECISION", "fp32")

    def forward(self, depth_imgs: torch.Tensor):
        if not depth_imgs.is_cuda:
            raise DicError("depth_imgs must be a CUDA tensor: this encoder has no CPU fallback")
        if depth_imgs.dim() != 4 or depth_imgs.shape[1] != 1:
            raise ValueError(f"depth_imgs must be [B, 1, H, W], got {tuple(depth_imgs.shape)}")
        if self.precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        x = depth_imgs.to(torch.float32).contiguous()
        params = []
        for cv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            params += [cv.weight, cv.bias, bn.weight, bn.bias]
        out = _DepthEncoderFn.apply(x, self, *params)
        if self.training:
            for bn in (self.bn1, self.bn2, self.bn3):
                bn.num_batches_tracked.add_(1)          # nn.BatchNorm2d bookkeeping (unused with a fixed momentum)
        return out
