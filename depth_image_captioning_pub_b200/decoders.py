"""Drop-in decoder modules for the reference's Show-Attend-Tell decoders with depth fusion.

Same class names, constructor arguments, ``state_dict`` keys/shapes and method signatures
as the reference:

  CD_RNNDecoderWithSoftAttention / CD_RNNDecoderWithHardAttention
      Captioning_models/Depth_caption_model/depth_models.py:96-305, 522-789
  RNNDecoderWithSoftAttention / RNNDecoderWithHardAttention
      Captioning_models/Base_caption_model/base_caption_models.py:49-250, 257-508

so checkpoints written by the reference load unchanged and the reference's training /
evaluation loops (depth_train.py:207, depth_evaluation.py:164,336) can construct these classes
instead.  All arithmetic runs in the CUDA library behind include/dic.h; there is no CPU path
(CPU tensors raise ``DicError``).  The ``nn.Linear`` / ``nn.LSTMCell`` / ``nn.Embedding``
submodules exist to hold the parameters under the reference's names (and are constructed in
the reference's order so that the same ``torch.manual_seed`` gives the same initial weights);
their ``forward`` methods are never called.

Extensions that do not exist in the reference: ``precision`` attribute ("fp32" parity mode,
"bf16" tensor-core mode), ``beam_search`` (the reference is greedy only).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import nn
from torch.nn.utils.rnn import PackedSequence

from . import _lib
from ._lib import DicError, PARAM_KEYS
from .attention import Hard_Attention, Soft_Attention
from .engine import CaptionLossFunction, DecoderFunction, Engine, batch_sizes_from_lengths

DEFAULT_PRECISION = os.environ.get("DIC_PRECISION", "fp32")


class _DecoderBase(nn.Module):
    _hard = False

    def _build(self, dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size, dropout):
        self.vocab_size = vocab_size
        att_cls = Hard_Attention if self._hard else Soft_Attention
        # construction order = reference order (depth_models.py:113-135) -> identical seeded init
        self.attention = att_cls(dim_encoder, dim_decoder, dim_attention)
        self.embed = nn.Embedding(vocab_size, dim_embedding)
        self.dropout = nn.Dropout(dropout)
        self.decode_step = nn.LSTMCell(dim_embedding + dim_encoder, dim_decoder, bias=True)
        self.init_linear = nn.Linear(dim_encoder, dim_decoder * 2)
        self.f_beta = nn.Linear(dim_decoder, dim_encoder)
        self.linear = nn.Linear(dim_decoder, vocab_size)
        self._reset_parameters()
        self.precision = DEFAULT_PRECISION
        # decode calls re-pack the weights every time unless the caller declares them frozen
        self.cache_packed_weights = False
        self._dims =(dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size)
        self._engines: Dict = {}

    def _reset_parameters(self):
        # depth_models.py:140-143
        nn.init.uniform_(self.embed.weight, -0.1, 0.1)
        nn.init.uniform_(self.linear.weight, -0.1, 0.1)
        nn.init.constant_(self.linear.bias, 0)

    # ---- plumbing --------------------------------------------------------------------------
    def _param_list(self) -> List[torch.Tensor]:
        sd = dict(self.named_parameters())
        return [sd[k] for k in PARAM_KEYS]

    def _engine(self, L: int, device: torch.device) -> Engine:
        A, E, D, H, V = self._dims
        key = (L, self.precision, device.index if device.index is not None else torch.cuda.current_device())
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(L, D, A, E, H, V, self.precision, device)
            self._engines[key] = eng
        eng.use_flat_grads = bool(getattr(self, "flat_grads", False))
        eng.flat_alloc = getattr(self, "flat_alloc", None)
        return eng

    def _check_feats(self, features, depth_features):
        if not features.is_cuda:
            raise DicError("features must be CUDA tensors: this decoder has no CPU fallback")
        if features.dim() != 3 or features.shape[2] != self._dims[2]:
            raise ValueError(f"features must be [B, L, {self._dims[2]}], got {tuple(features.shape)}")
        if features.dtype not in (torch.float32, torch.bfloat16):
            raise DicError("features must be float32 or bfloat16")
        features = features.contiguous()
        if depth_features is not None:
            if depth_features.shape != features.shape or depth_features.dtype != features.dtype:
                raise ValueError("depth_features must match features in shape and dtype")
            depth_features = depth_features.contiguous()
        return features, depth_features

    def _dropout_mask(self, total: int, device) -> Optional[torch.Tensor]:
        p = self.dropout.p
        if not self.training or p == 0.0:
            return None
        H = self._dims[3]
        if p >= 1.0:
            return torch.zeros(total, H, device=device)
        # nn.Dropout on h (depth_models.py:197): the keep / (1 - p) mask of the device generator, drawn by ONE kernel
        # (torch's fused dropout on a cached tensor of ones) instead of rand, compare, cast and divide
        ones = getattr(self, "_ones_cache", None)
        if ones is None or ones.shape != (total, H) or ones.device != device:
            ones = torch.ones(total, H, device=device)
            self._ones_cache = ones
        return torch.nn.functional.dropout(ones, p, training=True)

    def _teacher_forced(self, attn_mode, features, depth_features, captions, lengths, u, temp, train_dropout):
        features, depth_features = self._check_feats(features, depth_features)
        bsz = batch_sizes_from_lengths(lengths)
        B = features.shape[0]
        if len(lengths) != B or captions.shape[0] != B:
            raise ValueError("features, captions and lengths disagree on the batch size")
        if bsz[0] != B:
            raise ValueError("internal: first step must cover the whole batch")
        if len(bsz) > _lib.MAX_STEPS:
            raise ValueError(f"captions longer than {_lib.MAX_STEPS} steps are not supported")
        captions = captions.to(device=features.device, dtype=torch.int64).contiguous()
        eng = self._engine(features.shape[1], features.device)
        total = sum(bsz)
        mask = self._dropout_mask(total, features.device) if train_dropout else None
        logits, alphas = DecoderFunction.apply(eng, (attn_mode, torch.is_grad_enabled()), captions, bsz, u, float(temp), mask, features,
                                               depth_features, *self._param_list())
        packed = PackedSequence(logits, torch.tensor(bsz, dtype=torch.int64))
        return packed, alphas

    def _teacher_forced_loss(self, attn_mode, features, depth_features, captions, lengths, u, temp, ignore_index,
                             lam):
        """forward + the training loop's loss (depth_train.py:210-216) as one fused call -> 0-d loss."""
        features, depth_features = self._check_feats(features, depth_features)
        bsz = batch_sizes_from_lengths(lengths)
        B = features.shape[0]
        if len(lengths) != B or captions.shape[0] != B:
            raise ValueError("features, captions and lengths disagree on the batch size")
        if len(bsz) > _lib.MAX_STEPS:
            raise ValueError(f"captions longer than {_lib.MAX_STEPS} steps are not supported")
        captions = captions.to(device=features.device, dtype=torch.int64).contiguous()
        eng = self._engine(features.shape[1], features.device)
        mask = self._dropout_mask(sum(bsz), features.device)
        if ignore_index is None:
            ignore_index = -100          # F.cross_entropy default
        return CaptionLossFunction.apply(eng, (attn_mode, torch.is_grad_enabled()), captions, bsz, int(ignore_index), float(lam), u,
                                         float(temp), mask, features, depth_features, *self._param_list())

    def _draw_u(self, rows: int, L: int, device) -> torch.Tensor:
        # one draw for all steps == the reference's per-step torch.rand(bs_valid, k) calls on the
        # CPU generator concatenated (attention.py:17,40).  Drawn into pinned memory and copied without a
        # host synchronisation, so the host RNG of step n+1 overlaps the device work of step n.
        if not torch.cuda.is_available():
            return torch.rand(rows, L).to(device)
        host = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
        torch.rand(rows, L, out=host)
        return host.to(device, non_blocking=True)

    @torch.no_grad()
    def _greedy(self, attn_mode, features, depth_features, word_to_id, max_length, want_alphas):
        features, depth_features = self._check_feats(features, depth_features)
        eng = self._engine(features.shape[1], features.device)
        eng.ensure_packed(self._param_list(), allow_cached=self.cache_packed_weights)
        u = None
        if attn_mode == _lib.ATTN_GUMBEL_MAX:
            u = self._draw_u(max_length * features.shape[0], features.shape[1], features.device)
        tokens, alphas, _ = eng.greedy(attn_mode, features, depth_features, word_to_id['<start>'], max_length,
                                       u=u, want_alphas=want_alphas)
        return tokens, alphas

    @torch.no_grad()
    def _beam(self, features, depth_features, word_to_id, beam, max_length, trace=False):
        features, depth_features = self._check_feats(features, depth_features)
        eng = self._engine(features.shape[1], features.device)
        eng.ensure_packed(self._param_list(), allow_cached=self.cache_packed_weights)
        return eng.beam(features, depth_features, word_to_id['<start>'], word_to_id['<end>'], beam, max_length,
                        trace=trace)


# ---------------------------------------------------------------------------------------------
# depth (add-fusion) decoders
# ---------------------------------------------------------------------------------------------
class CD_RNNDecoderWithSoftAttention(_DecoderBase):
    """depth_models.py:96-305."""

    def __init__(self, dim_attention: int, dim_embedding: int, dim_encoder: int, dim_decoder: int,
                 vocab_size: int, dropout: float = 0.5):
        super().__init__()
        self._build(dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size, dropout)

    def forward(self, features: torch.Tensor, depth_features: torch.Tensor, captions: torch.Tensor,
                lengths: list):
        """-> (PackedSequence of logits, alphas [B, Tmax, L])   depth_models.py:153-207"""
        return self._teacher_forced(_lib.ATTN_SOFT, features, depth_features, captions, lengths, None, 1.0, True)

    def forward_loss(self, features, depth_features, captions, lengths, ignore_index=None, lam: float = 0.7):
        """Extension: forward + `F.cross_entropy(outputs.data, packed targets, ignore_index)
        + lam * ((1 - alphas.sum(dim=1)) ** 2).mean()` (depth_train.py:210-216) fused; -> 0-d loss."""
        return self._teacher_forced_loss(_lib.ATTN_SOFT, features, depth_features, captions, lengths, None, 1.0,
                                         ignore_index, lam)

    def sample(self, features: torch.Tensor, depth_features: torch.Tensor, word_to_id: list, max_length=30):
        """-> (list[int], list[Tensor [1, L]])   depth_models.py:216-257"""
        tokens, alphas = self._greedy(_lib.ATTN_SOFT, features, depth_features, word_to_id, max_length, True)
        return tokens[0].tolist(), [alphas[t] for t in range(max_length)]

    def batch_sample(self, features: torch.Tensor, depth_features: torch.Tensor, word_to_id: list,
                     max_length=30):
        """-> np.ndarray [B, max_length] int64   depth_models.py:259-305 (one D2H copy, not one per step)"""
        tokens, _ = self._greedy(_lib.ATTN_SOFT, features, depth_features, word_to_id, max_length, False)
        return tokens.cpu().numpy().astype(np.int64)

    def beam_search(self, features, depth_features, word_to_id, beam: int = 5, max_length=30, trace=False):
        """Not in the reference.  -> dict(tokens [B,max_length] int64, lengths [B], scores [B])."""
        return self._beam(features, depth_features, word_to_id, beam, max_length, trace)


class CD_RNNDecoderWithHardAttention(_DecoderBase):
    """depth_models.py:522-789."""
    _hard = True

    def __init__(self, dim_attention: int, dim_embedding: int, dim_encoder: int, dim_decoder: int,
                 vocab_size: int, device: str, dropout: float = 0.5):
        super().__init__()
        self.device = device
        self._build(dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size, dropout)

    def forward(self, features: torch.Tensor, depth_features: torch.Tensor, captions: torch.Tensor,
                lengths: list, temp: torch.tensor):
        """Gumbel-softmax relaxation; -> PackedSequence only   depth_models.py:580-634"""
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        packed, _ = self._teacher_forced(_lib.ATTN_GUMBEL_SOFTMAX, features, depth_features, captions, lengths,
                                         u, float(temp), True)
        return packed

    def forward_loss(self, features, depth_features, captions, lengths, temp, ignore_index=None):
        """Extension: forward(temp) + cross entropy (hard attention has no regulariser,
        depth_train.py:530-532) fused; -> 0-d loss."""
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        return self._teacher_forced_loss(_lib.ATTN_GUMBEL_SOFTMAX, features, depth_features, captions, lengths, u,
                                         float(temp), ignore_index, 0.0)

    @torch.no_grad()
    def eval_forward(self, features: torch.Tensor, depth_features: torch.Tensor, captions: torch.Tensor,
                     lengths: list):
        """Gumbel-max one-hot attention; -> PackedSequence   depth_models.py:637-689"""
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        packed, _ = self._teacher_forced(_lib.ATTN_GUMBEL_MAX, features, depth_features, captions, lengths, u,
                                         1.0, True)
        return packed

    def sample(self, features: torch.Tensor, depth_features: torch.Tensor, word_to_id: list, max_length=30):
        """-> (list[int], list[int64 one-hot Tensor [1, L]])   depth_models.py:698-740"""
        tokens, alphas = self._greedy(_lib.ATTN_GUMBEL_MAX, features, depth_features, word_to_id, max_length,
                                      True)
        return tokens[0].tolist(), [alphas[t].to(torch.int64) for t in range(max_length)]

    def batch_sample(self, features: torch.Tensor, depth_features: torch.Tensor, word_to_id: list,
                     max_length=30):
        """-> np.ndarray [B, max_length] int64   depth_models.py:742-789"""
        tokens, _ = self._greedy(_lib.ATTN_GUMBEL_MAX, features, depth_features, word_to_id, max_length, False)
        return tokens.cpu().numpy().astype(np.int64)


# ---------------------------------------------------------------------------------------------
# base (Show-Attend-Tell) decoders: same kernels, no depth tensor
# ---------------------------------------------------------------------------------------------
class RNNDecoderWithSoftAttention(_DecoderBase):
    """base_caption_models.py:49-250."""

    def __init__(self, dim_attention: int, dim_embedding: int, dim_encoder: int, dim_decoder: int,
                 vocab_size: int, dropout: float = 0.5):
        super().__init__()
        self._build(dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size, dropout)

    def forward(self, features: torch.Tensor, captions: torch.Tensor, lengths: list):
        return self._teacher_forced(_lib.ATTN_SOFT, features, None, captions, lengths, None, 1.0, True)

    def forward_loss(self, features, captions, lengths, ignore_index=None, lam: float = 0.7):
        """Extension: forward + the loss of base_train.py:156-162 fused; -> 0-d loss."""
        return self._teacher_forced_loss(_lib.ATTN_SOFT, features, None, captions, lengths, None, 1.0,
                                         ignore_index, lam)

    def sample(self, features: torch.Tensor, word_to_id: list, max_length=30):
        tokens, alphas = self._greedy(_lib.ATTN_SOFT, features, None, word_to_id, max_length, True)
        return tokens[0].tolist(), [alphas[t] for t in range(max_length)]

    def batch_sample(self, features: torch.Tensor, word_to_id: list, max_length=30):
        tokens, _ = self._greedy(_lib.ATTN_SOFT, features, None, word_to_id, max_length, False)
        return tokens.cpu().numpy().astype(np.int64)

    def beam_search(self, features, word_to_id, beam: int = 5, max_length=30, trace=False):
        return self._beam(features, None, word_to_id, beam, max_length, trace)


class RNNDecoderWithHardAttention(_DecoderBase):
    """base_caption_models.py:257-508."""
    _hard = True

    def __init__(self, dim_attention: int, dim_embedding: int, dim_encoder: int, dim_decoder: int,
                 vocab_size: int, device: str, dropout: float = 0.5):
        super().__init__()
        self.device = device
        self._build(dim_attention, dim_embedding, dim_encoder, dim_decoder, vocab_size, dropout)

    def forward(self, features: torch.Tensor, captions: torch.Tensor, lengths: list, temp: torch.tensor):
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        packed, _ = self._teacher_forced(_lib.ATTN_GUMBEL_SOFTMAX, features, None, captions, lengths, u,
                                         float(temp), True)
        return packed

    def forward_loss(self, features, captions, lengths, temp, ignore_index=None):
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        return self._teacher_forced_loss(_lib.ATTN_GUMBEL_SOFTMAX, features, None, captions, lengths, u,
                                         float(temp), ignore_index, 0.0)

    @torch.no_grad()
    def eval_forward(self, features: torch.Tensor, captions: torch.Tensor, lengths: list):
        bsz = batch_sizes_from_lengths(lengths)
        u = self._draw_u(sum(bsz), features.shape[1], features.device)
        packed, _ = self._teacher_forced(_lib.ATTN_GUMBEL_MAX, features, None, captions, lengths, u, 1.0, True)
        return packed

    def sample(self, features: torch.Tensor, word_to_id: list, max_length=30):
        tokens, alphas = self._greedy(_lib.ATTN_GUMBEL_MAX, features, None, word_to_id, max_length, True)
        return tokens[0].tolist(), [alphas[t].to(torch.int64) for t in range(max_length)]

    def batch_sample(self, features: torch.Tensor, word_to_id: list, max_length=30):
        tokens, _ = self._greedy(_lib.ATTN_GUMBEL_MAX, features, None, word_to_id, max_length, False)
        return tokens.cpu().numpy().astype(np.int64)


# ---------------------------------------------------------------------------------------------
# depth (concat-fusion, "MD_") decoders: torch.cat((features, depth_features), dim=2) instead of the
# add (depth_models.py:376,433,478,851,907,963,1009), then exactly the base decoder at
# dim_encoder = mlp_dim_encoder (2048 + 32 = 2080 in config.py:19).  No reference script instantiates
# them; they are provided for completeness.  The concatenation is a plain
# copy done by torch (data movement, no arithmetic); everything else runs in the CUDA library.
# ---------------------------------------------------------------------------------------------
def _concat(features: torch.Tensor, depth_features: torch.Tensor) -> torch.Tensor:
    if not features.is_cuda:
        raise DicError("features must be CUDA tensors: this decoder has no CPU fallback")
    return torch.cat((features, depth_features.to(features.dtype)), dim=2)


class MD_RNNDecoderWithSoftAttention(RNNDecoderWithSoftAttention):
    """depth_models.py:309-517."""

    def __init__(self, dim_attention: int, dim_embedding: int, mlp_dim_encoder: int, dim_decoder: int,
                 vocab_size: int, dropout: float = 0.5):
        super().__init__(dim_attention, dim_embedding, mlp_dim_encoder, dim_decoder, vocab_size, dropout)

    def forward(self, features, depth_features, captions, lengths):   # noqa: D102
        return super().forward(_concat(features, depth_features), captions, lengths)

    def forward_loss(self, features, depth_features, captions, lengths, ignore_index=None, lam: float = 0.7):
        return super().forward_loss(_concat(features, depth_features), captions, lengths, ignore_index, lam)

    def sample(self, features, depth_features, word_to_id, max_length=30):
        return super().sample(_concat(features, depth_features), word_to_id, max_length)

    def batch_sample(self, features, depth_features, word_to_id, max_length=30):
        return super().batch_sample(_concat(features, depth_features), word_to_id, max_length)

    def beam_search(self, features, depth_features, word_to_id, beam: int = 5, max_length=30, trace=False):
        return super().beam_search(_concat(features, depth_features), word_to_id, beam, max_length, trace)


class MD_RNNDecoderWithHardAttention(RNNDecoderWithHardAttention):
    """depth_models.py:792-1049."""

    def __init__(self, dim_attention: int, dim_embedding: int, mlp_dim_encoder: int, dim_decoder: int,
                 vocab_size: int, device: str, dropout: float = 0.5):
        super().__init__(dim_attention, dim_embedding, mlp_dim_encoder, dim_decoder, vocab_size, device, dropout)

    def forward(self, features, depth_features, captions, lengths, temp):   # noqa: D102
        return super().forward(_concat(features, depth_features), captions, lengths, temp)

    def forward_loss(self, features, depth_features, captions, lengths, temp, ignore_index=None):
        return super().forward_loss(_concat(features, depth_features), captions, lengths, temp, ignore_index)

    def eval_forward(self, features, depth_features, captions, lengths):
        return super().eval_forward(_concat(features, depth_features), captions, lengths)

    def sample(self, features, depth_features, word_to_id, max_length=30):
        return super().sample(_concat(features, depth_features), word_to_id, max_length)

    def batch_sample(self, features, depth_features, word_to_id, max_length=30):
        return super().batch_sample(_concat(features, depth_features), word_to_id, max_length)
