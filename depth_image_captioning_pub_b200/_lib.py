"""ctypes binding of include/dic.h (libdic.so).

There is no CPU fallback and no pure-PyTorch path: if the CUDA library cannot be
loaded, importing the product modules works (so that CPU-only tooling can inspect them) but
every compute call raises ``DicError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdic.so")

DIC_F32, DIC_BF16 = 0, 1
ATTN_SOFT, ATTN_GUMBEL_SOFTMAX, ATTN_GUMBEL_MAX = 0, 1, 2
MAX_STEPS, MAX_BEAM = 128, 8


class DicError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("L", "D", "A", "E", "H", "V")]


PARAM_FIELDS = (
    "enc_att_w", "enc_att_b", "dec_att_w", "dec_att_b", "full_att_w", "full_att_b", "embed_w",
    "w_ih", "w_hh", "b_ih", "b_hh", "init_w", "init_b", "fbeta_w", "fbeta_b", "lin_w", "lin_b",
)
# state_dict keys in the same order (depth_models.py:106-135)
PARAM_KEYS = (
    "attention.encoder_att.weight", "attention.encoder_att.bias",
    "attention.decoder_att.weight", "attention.decoder_att.bias",
    "attention.full_att.weight", "attention.full_att.bias",
    "embed.weight",
    "decode_step.weight_ih", "decode_step.weight_hh", "decode_step.bias_ih", "decode_step.bias_hh",
    "init_linear.weight", "init_linear.bias",
    "f_beta.weight", "f_beta.bias",
    "linear.weight", "linear.bias",
)


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_SZ = C.c_size_t
_DP = C.POINTER(Dims)
_PP = C.POINTER(Params)
_IP = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/dic.h declaration by declaration
PROTOTYPES = {
    "dic_version": (_I, []),
    "dic_last_error": (C.c_char_p, []),
    "dic_pack_bytes": (_SZ, [_DP, _I]),
    "dic_pack_weights": (_I, [_DP, _I, _PP, _P, _P]),
    "dic_train_workspace_bytes": (_SZ, [_DP, _I, _I, _I]),
    "dic_decoder_forward": (_I, [_DP, _I, _I, _P, _P, _P, _I, _P, _I, _IP, _I, _I, _P, _F, _P, _P, _P,
                                 _P, _SZ, _P]),
    "dic_decoder_backward": (_I, [_DP, _I, _I, _P, _P, _P, _I, _P, _I, _IP, _I, _I, _P, _P, _P, _F, _P,
                                  _PP, _P, _P, _SZ, _P]),
    "dic_decoder_backward_ex": (_I, [_DP, _I, _I, _P, _P, _P, _I, _P, _I, _IP, _I, _I, _P, _I, _P, _P, _F, _P,
                                     _PP, _P, _P, _SZ, _P]),
    "dic_caption_loss_workspace_bytes": (_SZ, [_I, _I]),
    "dic_caption_loss": (_I, [_DP, _I, _P, _I, _P, _I, _IP, _I, _I, _I, _P, _F, _P, _P, _P, _P, _SZ, _P]),
    "dic_decoder_forward_ex": (_I, [_DP, _I, _I, _P, _P, _P, _I, _P, _I, _IP, _I, _I, _P, _F, _P, _P, _I, _P, _P,
                                    _SZ, _P]),
    "dic_scale_loss_grads": (_I, [_I, _P, _P, _SZ, _P, _SZ, _P]),
    "dic_decode_workspace_bytes": (_SZ, [_DP, _I, _I, _I]),
    "dic_decode_greedy": (_I, [_DP, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "dic_decode_beam": (_I, [_DP, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P,
                             _P, _SZ, _P]),
    "dic_attention_workspace_bytes": (_SZ, [_DP, _I, _I]),
    "dic_attention_forward": (_I, [_DP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _F, _P, _P, _P,
                                   _SZ, _P]),
    "dic_beam_select_workspace_bytes": (_SZ, [_I, _I]),
    "dic_beam_select": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "dic_row_lse": (_I, [_P, _I, _I, _P, _P]),
    "dic_gemm_nt": (_I, [_I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P]),
    "dic_gemm_ex": (_I, [_I, _I, _I, _I, _P, _I, C.c_longlong, C.c_longlong, _P, _I, C.c_longlong, C.c_longlong,
                         _P, _P, C.c_longlong, _I, _P]),
    "dic_gemm_nt_bf16": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, C.c_longlong, _P]),
    "dic_adamw_step": (_I, [_I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                            C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), _F, _F, _F, _F, _F, _I, _P]),
    "dic_dfeat_gemm": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "dic_launch_count": (C.c_longlong, []),
    "dic_profile_classes": (_I, []),
    "dic_profile_class_name": (C.c_char_p, [_I]),
    "dic_profile_enable": (None, [_I]),
    "dic_set_substreams": (None, [_I]),
    "dic_set_grads_ready_event": (None, [_P]),
    "dic_set_grads_ready_events": (None, [_P, _P, _P]),
    "dic_depth_encoder_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "dic_depth_encoder_forward": (_I, [_I, _I, _I, _I, _I, _P, _P, _F, _F, _P, _I, _P, _SZ, _P]),
    "dic_depth_encoder_backward": (_I, [_I, _I, _I, _I, _P, _P, _I, _P, _P, _SZ, _P]),
    "dic_dp_flag_bytes": (C.c_size_t, []),
    "dic_dp_allreduce": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, C.c_longlong, C.c_float, C.c_uint, C.c_int, _P]),
    "dic_trace_start": (_I, [_P, C.c_uint]),
    "dic_trace_stop": (_I, [C.POINTER(C.c_uint)]),
    "dic_profile_read": (_I, [C.POINTER(C.c_float), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
}

_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libdic.so (building it with nvcc first when it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except Exception as e:
            # A library that does not match the sources is never loaded silently: tests and benchmarks
            # would run an old binary.  DIC_ALLOW_STALE_LIB=1 is the explicit escape hatch (e.g. a box
            # without nvcc that was handed a prebuilt library of a different source revision).
            if not os.path.exists(LIB_PATH):
                raise DicError(f"libdic.so is missing and could not be built: {e}") from e
            if _build.is_stale() and os.environ.get("DIC_ALLOW_STALE_LIB") != "1":
                raise DicError(f"libdic.so is stale (sources changed) and the rebuild failed: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise DicError(f"{LIB_PATH} not found; run `python -m depth_image_captioning_pub_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dic_version() != 100:
        raise DicError(f"libdic.so version {lib.dic_version()} != 100; rebuild")
    _lib = lib
    return lib


def profile_read():
    """-> {class: (ms, launches, algorithmic_bytes)} since the last dic_profile_enable(1)/read."""
    lib = load()
    n = lib.dic_profile_classes()
    ms = (C.c_float * n)()
    cnt = (C.c_longlong * n)()
    by = (C.c_double * n)()
    check(lib.dic_profile_read(ms, cnt, by))
    return {lib.dic_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i]), float(by[i])) for i in range(n)}


def check(rc: int) -> None:
    if rc != 0:
        msg = load().dic_last_error()
        raise DicError(f"libdic error {rc}: {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DicError("tensor is not on a CUDA device (this path has no CPU fallback)")
    if not t.is_contiguous():
        raise DicError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DIC_F32
    if t.dtype == torch.bfloat16:
        return DIC_BF16
    raise DicError(f"unsupported dtype {t.dtype} (float32 or bfloat16)")
