"""bf16 annotation tensors straight from an (autocast) encoder: forward, decode and dL/dF in bf16."""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from oracle import decoder_oracle as O
from test_gpu_parity import CASES, build_module, make_case, relmax

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["small_ragged", "ref_dims"])
@pytest.mark.parametrize("depth", [True, False])
def test_bf16_features(case, depth, cuda_device):
    cfg = dict(CASES[case])
    lengths, V = cfg["lengths"], cfg["V"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    Fr16, Fd16 = F_rgb.to(torch.bfloat16), F_dep.to(torch.bfloat16)
    # oracle on the bf16-rounded annotations, fp64 arithmetic
    wo = {k: v.double() for k, v in w.items()}
    Fr = Fr16.double().requires_grad_(True)
    Fd = Fd16.double().requires_grad_(True) if depth else None
    lo, _, ao = O.decoder_forward(wo, Fr, Fd, caps, lengths, hoist=True)
    O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao).backward()
    cls = P.CD_RNNDecoderWithSoftAttention if depth else P.RNNDecoderWithSoftAttention
    m = build_module(cls, w, cuda_device, "bf16").eval()
    fr = Fr16.to(cuda_device).requires_grad_(True)
    fd = Fd16.to(cuda_device).requires_grad_(True)
    feats = (fr, fd) if depth else (fr,)
    out, alphas = m(*feats, caps.to(cuda_device), lengths)
    assert relmax(out.data.detach().cpu(), lo.detach()) <= 2e-2
    tg = O.pack_targets(caps, lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    assert fr.grad.dtype == torch.bfloat16 and fr.grad.shape == fr.shape
    ref = Fr.grad.numpy()
    assert np.abs(fr.grad.double().cpu().numpy() - ref).max() <= 5e-2 * np.abs(ref).max()
    if depth:
        assert torch.equal(fr.grad, fd.grad)       # add-fusion: both inputs get the same gradient
    voc = O.synthetic_vocab(V)
    toks = m.batch_sample(*(f.detach() for f in feats), voc, max_length=5)
    assert toks.shape == (cfg["B"], 5)
