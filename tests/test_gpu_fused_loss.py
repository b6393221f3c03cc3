"""Fused caption-loss head (forward_loss) against (1) the golden loss / gradients
computed by the unmodified reference modules + the training loop's loss expression
(depth_train.py:210-216) and (2) the module's own forward() + torch loss on seeded inputs."""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from conftest import load_golden
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O
from test_gpu_parity import CASES, build_module, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["depth_soft", "base_soft", "depth_soft_peaked"])
def test_golden_forward_loss(name, cuda_device):
    rec, w, g = load_golden(name)
    depth = bool(int(rec["depth"]))
    cls = P.CD_RNNDecoderWithSoftAttention if depth else P.RNNDecoderWithSoftAttention
    m = build_module(cls, w, cuda_device).eval()
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device).requires_grad_(True)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device).requires_grad_(True)
    caps = torch.from_numpy(rec["captions"]).to(cuda_device)
    lengths = rec["lengths"].tolist()
    V = w["embed.weight"].shape[0]
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    loss = m.forward_loss(*feats, caps, lengths, ignore_index=V - 1, lam=0.7)
    assert loss.dim() == 0
    assert abs(float(loss.detach()) - float(rec["loss"])) <= 1e-5
    loss.backward()
    grads = dict(m.named_parameters())
    for k in _lib.PARAM_KEYS:
        ref = g[k].numpy()
        got = grads[k].grad.cpu().numpy()
        tol = 2e-4 * max(np.abs(ref).max(), 1e-3) + 1e-7
        assert np.abs(got - ref).max() <= tol, (k, np.abs(got - ref).max(), tol)
    assert np.abs(F_rgb.grad.cpu().numpy() - rec["g_F_rgb"]).max() <= 2e-4 * np.abs(rec["g_F_rgb"]).max() + 1e-8
    if depth:
        assert np.abs(F_dep.grad.cpu().numpy() - rec["g_F_dep"]).max() <= 2e-4 * np.abs(rec["g_F_dep"]).max() + 1e-8


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_loss_matches_unfused(case, precision, cuda_device):
    """forward_loss == forward() + the torch loss expression, gradients included; an upstream
    gradient != 1 (loss * 3) goes through the device-side scale kernel."""
    cfg = dict(CASES[case])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    V = cfg["V"]
    caps_g = caps.to(cuda_device)

    def run(fused, scale):
        m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, precision).eval()
        Fr = F_rgb.to(cuda_device).requires_grad_(True)
        Fd = F_dep.to(cuda_device).requires_grad_(True)
        if fused:
            loss = m.forward_loss(Fr, Fd, caps_g, lengths, ignore_index=V - 1, lam=0.7)
        else:
            out, alphas = m(Fr, Fd, caps_g, lengths)
            tg = O.pack_targets(caps, lengths).to(cuda_device)
            loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
            loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
        (loss * scale).backward()
        gr = {k: p.grad.double().cpu().numpy() for k, p in m.named_parameters()}
        gr["dF"] = Fd.grad.double().cpu().numpy()
        return float(loss.detach()), gr

    for scale in (1.0, 3.0):
        l_ref, g_ref = run(False, scale)
        l_fus, g_fus = run(True, scale)
        # bf16 mode: the fused step keeps the logits block in bf16 (2^-9 relative rounding per logit, inside
        # the mode's 2e-2 bound), so its loss differs from the loss on the fp32 copy of the same logits
        ltol = 2e-6 if precision == "fp32" else 2e-3
        assert abs(l_ref - l_fus) <= ltol * max(1.0, abs(l_ref))
        # bf16 mode: d_logits is rounded to bf16 either way (same values), fp32: identical math up to
        # the reduction order of the log-sum-exp
        tol = 2e-5 if precision == "fp32" else 2e-2
        for k in g_ref:
            ref, got = g_ref[k], g_fus[k]
            assert np.isfinite(got).all(), k
            if k == "attention.full_att.bias":
                assert np.abs(got - ref).max() <= 1e-6
                continue
            assert np.abs(got - ref).max() <= tol * np.abs(ref).max() + 1e-9, (k, scale)


def test_forward_loss_ignore_index_and_hard(cuda_device):
    """<null> targets inside the valid length are ignored (mean over the rest); hard attention
    (Gumbel-softmax) has CE only."""
    cfg = dict(CASES["small_ragged"])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    V = cfg["V"]
    caps = caps.clone()
    caps[0, 2] = V - 1          # a <null> in the middle of a caption
    caps[1, 1] = V - 1
    caps_g = caps.to(cuda_device)
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).eval()
    Fr, Fd = F_rgb.to(cuda_device), F_dep.to(cuda_device)
    out, alphas = m(Fr, Fd, caps_g, lengths)
    tg = O.pack_targets(caps, lengths).to(cuda_device)
    ref = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1) + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    got = m.forward_loss(Fr, Fd, caps_g, lengths, ignore_index=V - 1, lam=0.7)
    assert abs(float(ref) - float(got)) <= 2e-6 * max(1.0, abs(float(ref)))

    mh = build_module(P.CD_RNNDecoderWithHardAttention, w, cuda_device, extra=("cuda:0",)).eval()
    torch.manual_seed(5)
    out = mh(Fr, Fd, caps_g, lengths, torch.tensor(0.8))
    ref = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    torch.manual_seed(5)
    got = mh.forward_loss(Fr, Fd, caps_g, lengths, torch.tensor(0.8), ignore_index=V - 1)
    assert abs(float(ref) - float(got)) <= 2e-6 * max(1.0, abs(float(ref)))
