"""CPU restatement of the reference depth CNN encoder (TEST INFRASTRUCTURE ONLY: imported by tests/ and the
golden generator, never by the product package).

Follows Depth_CNN_endoder (/root/reference/Captioning_models/Depth_caption_model/depth_models.py:12-56) operation by
operation on a state_dict-keyed dict, with torch.nn.functional calls instead of module objects:
  conv1 (:18) bn1 (:19) relu maxpool3 (:33,:34,:38) conv2 (:20) bn2 (:21) relu maxpool3 conv3 (:22) bn3 (:23) relu
  AdaptiveAvgPool2d(encoded_img_size) (:31) ; permute(0,2,3,1).flatten(1,2) (:52).
Pinned against the unmodified reference module by oracle/make_golden_encoder.py -> tests/golden/depth_encoder.npz
(tests/test_oracle_golden_encoder.py).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

KEYS = ["conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias", "conv2.weight", "conv2.bias", "bn2.weight", "bn2.bias",
        "conv3.weight", "conv3.bias", "bn3.weight", "bn3.bias"]


def encoder_forward(sd: Dict[str, torch.Tensor], depth_imgs: torch.Tensor, training: bool = True,
                    encoded_img_size: int = 14, momentum: float = 0.1, eps: float = 1e-5) -> torch.Tensor:
    """depth_models.py:49-56.  sd holds the 12 parameters and, per bn, running_mean / running_var (updated in
    place in training mode exactly like nn.BatchNorm2d)."""
    x = depth_imgs
    for i, stride in ((1, 3), (2, 1), (3, 1)):
        x = F.conv2d(x, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], stride=stride)
        x = F.batch_norm(x, sd[f"bn{i}.running_mean"], sd[f"bn{i}.running_var"], sd[f"bn{i}.weight"], sd[f"bn{i}.bias"],
                         training=training, momentum=momentum, eps=eps)
        x = F.relu(x)
        if i < 3:
            x = F.max_pool2d(x, 3)
    x = F.adaptive_avg_pool2d(x, encoded_img_size)
    return x.permute(0, 2, 3, 1).flatten(1, 2)


def make_inputs(B: int, seed: int, size: int = 224) -> torch.Tensor:
    """Depth maps with exactly representable values (multiples of 1/256 in [0, 1))."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, 1, size, size), generator=g).float() / 256.0


def projection(shape, seed: int) -> torch.Tensor:
    """Fixed +-1 pattern used as dL/dF: the loss is sum(feats * projection)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, 2, shape, generator=g).float() * 2.0 - 1.0)


def make_weights(seed: int = 700, bn_seed: int = 703) -> Dict[str, torch.Tensor]:
    """The reference module's own initial state for `seed` (layers constructed in the reference's order,
    depth_models.py:18-23, so the default torch initialisers draw the same numbers), then the BN affine parameters
    of the golden generator.  -> state_dict-keyed dict incl. running statistics."""
    from torch import nn
    torch.manual_seed(seed)
    conv1 = nn.Conv2d(1, 128, 7, stride=3)
    bn1 = nn.BatchNorm2d(128)
    conv2 = nn.Conv2d(128, 512, 3)
    bn2 = nn.BatchNorm2d(512)
    conv3 = nn.Conv2d(512, 2048, 1)
    bn3 = nn.BatchNorm2d(2048)
    g = torch.Generator().manual_seed(bn_seed)
    sd = {}
    with torch.no_grad():
        for i, (cv, bn) in enumerate(((conv1, bn1), (conv2, bn2), (conv3, bn3)), start=1):
            bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) + 0.5)
            bn.bias.copy_(torch.rand(bn.bias.shape, generator=g) - 0.5)
        for i, (cv, bn) in enumerate(((conv1, bn1), (conv2, bn2), (conv3, bn3)), start=1):
            sd[f"conv{i}.weight"] = cv.weight.detach().clone()
            sd[f"conv{i}.bias"] = cv.bias.detach().clone()
            sd[f"bn{i}.weight"] = bn.weight.detach().clone()
            sd[f"bn{i}.bias"] = bn.bias.detach().clone()
            sd[f"bn{i}.running_mean"] = bn.running_mean.clone()
            sd[f"bn{i}.running_var"] = bn.running_var.clone()
    return sd
