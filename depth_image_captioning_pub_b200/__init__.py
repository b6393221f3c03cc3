"""B200-native (sm_100a) decoder hot path of Kyo-suke-S/Depth_image_captioning_pub.

Drop-in modules (same names / signatures / state_dict keys as the reference):
  attention:  Soft_Attention, Hard_Attention, Gumbel_softmax
  decoders:   CD_RNNDecoderWithSoftAttention, CD_RNNDecoderWithHardAttention,
              RNNDecoderWithSoftAttention, RNNDecoderWithHardAttention
The arithmetic lives in csrc/ (hand-written CUDA behind the C ABI in include/dic.h).
"""
from ._lib import DicError  # noqa: F401
from .attention import Gumbel_softmax, Hard_Attention, Soft_Attention  # noqa: F401
from .decoders import (CD_RNNDecoderWithHardAttention, CD_RNNDecoderWithSoftAttention,  # noqa: F401
                       MD_RNNDecoderWithHardAttention, MD_RNNDecoderWithSoftAttention,
                       RNNDecoderWithHardAttention, RNNDecoderWithSoftAttention)

from .encoders import Depth_CNN_endoder  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__all__ = [
    "FusedAdamW", "Depth_CNN_endoder",
    "DicError", "Gumbel_softmax", "Hard_Attention", "Soft_Attention",
    "CD_RNNDecoderWithHardAttention", "CD_RNNDecoderWithSoftAttention",
    "MD_RNNDecoderWithHardAttention", "MD_RNNDecoderWithSoftAttention",
    "RNNDecoderWithHardAttention", "RNNDecoderWithSoftAttention",
]
