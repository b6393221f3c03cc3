"""Concat-fusion (MD_) decoders against golden vectors from the reference's MD_ modules."""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from conftest import load_golden
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O
from test_gpu_parity import relmax

pytestmark = pytest.mark.gpu


def _module(cls, w, dev, extra=()):
    A, D = w["attention.encoder_att.weight"].shape
    V, E = w["embed.weight"].shape
    H = w["decode_step.weight_hh"].shape[1]
    m = cls(A, E, D, H, V, *extra)
    m.load_state_dict(w)
    return m.to(dev).eval()


def test_md_soft_golden(cuda_device):
    rec, w, g = load_golden("md_soft")
    m = _module(P.MD_RNNDecoderWithSoftAttention, w, cuda_device)
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device).requires_grad_(True)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device).requires_grad_(True)
    caps = torch.from_numpy(rec["captions"]).to(cuda_device)
    lengths = rec["lengths"].tolist()
    out, alphas = m(F_rgb, F_dep, caps, lengths)
    assert relmax(out.data.detach().cpu(), rec["logits"]) <= 1e-4
    assert np.abs(alphas.detach().cpu().numpy() - rec["alphas"]).max() <= 1e-5
    V = out.data.shape[1]
    tg = O.pack_targets(caps.cpu(), lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    grads = dict(m.named_parameters())
    for k in _lib.PARAM_KEYS:
        ref = g[k].numpy()
        tol = 2e-4 * max(np.abs(ref).max(), 1e-3) + 1e-7
        assert np.abs(grads[k].grad.cpu().numpy() - ref).max() <= tol, k
    for got, ref in ((F_rgb.grad, rec["g_F_rgb"]), (F_dep.grad, rec["g_F_dep"])):
        assert got.shape == ref.shape
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-8
    toks = m.batch_sample(F_rgb.detach(), F_dep.detach(), O.synthetic_vocab(V), max_length=rec["greedy"].shape[1])
    np.testing.assert_array_equal(toks, rec["greedy"])


def test_md_hard_golden(cuda_device):
    rec, w, _ = load_golden("md_hard")
    m = _module(P.MD_RNNDecoderWithHardAttention, w, cuda_device, ("cuda:0",))
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device)
    caps = torch.from_numpy(rec["captions"]).to(cuda_device)
    lengths = rec["lengths"].tolist()
    V = w["linear.weight"].shape[0]
    torch.manual_seed(703)
    ev = m.eval_forward(F_rgb, F_dep, caps, lengths)
    assert relmax(ev.data.cpu(), rec["eval_logits"]) <= 1e-4
    torch.manual_seed(704)
    toks = m.batch_sample(F_rgb, F_dep, O.synthetic_vocab(V), max_length=rec["greedy"].shape[1])
    np.testing.assert_array_equal(toks, rec["greedy"])
