"""Golden vectors for the depth CNN encoder from the UNMODIFIED reference module.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_encoder.py

Stores (tests/golden/depth_encoder.npz): the module's initial state_dict for seed 700 as float32 arrays of its SMALL
tensors plus seeds to regenerate the rest, a strided sub-sample of the training-mode and eval-mode outputs for B = 2
depth maps (seed 701), the updated running statistics, and for loss = sum(feats * projection(seed 702)) each of the 12
parameter gradients as (norm, first 64 elements)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("DIC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))
from Captioning_models.Depth_caption_model.depth_models import Depth_CNN_endoder  # noqa: E402
from oracle import depth_encoder_oracle as EO  # noqa: E402


def main():
    torch.manual_seed(700)
    m = Depth_CNN_endoder(14).train()
    # non-trivial BN affine parameters (the default init is weight 1, bias 0)
    g = torch.Generator().manual_seed(703)
    with torch.no_grad():
        for bn in (m.bn1, m.bn2, m.bn3):
            bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) + 0.5)
            bn.bias.copy_(torch.rand(bn.bias.shape, generator=g) - 0.5)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = EO.make_inputs(2, 701)
    feats = m(x)
    proj = EO.projection(feats.shape, 702)
    (feats * proj).sum().backward()
    rec = {"keys": np.array(sorted(sd0.keys()))}
    rec["train_sub"] = feats.detach()[:, ::7, ::64].numpy()
    rec["train_sum"] = np.array([float(feats.detach().double().sum()), float(feats.detach().double().abs().sum())])
    for k in EO.KEYS:
        gk = dict(m.named_parameters())[k].grad
        rec["gnorm." + k] = np.array([float(gk.double().norm())])
        rec["ghead." + k] = gk.flatten()[:64].numpy()
    for i in (1, 2, 3):
        rec[f"rm{i}"] = getattr(m, f"bn{i}").running_mean.numpy().copy()
        rec[f"rv{i}"] = getattr(m, f"bn{i}").running_var.numpy().copy()
    m.eval()
    with torch.no_grad():
        fe = m(x)
    rec["eval_sub"] = fe[:, ::7, ::64].numpy()
    # the small tensors of the initial state (the big ones are pinned through the outputs above)
    for k, v in sd0.items():
        if v.numel() <= 2048:
            rec["sd0." + k] = v.numpy()
    rec["sd0_sums"] = np.array([float(sd0[k].double().sum()) for k in EO.KEYS])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "depth_encoder.npz"), **rec)
    print("wrote depth_encoder.npz", {k: getattr(v, "shape", None) for k, v in rec.items() if k.startswith(("train", "eval"))})


if __name__ == "__main__":
    main()
