"""Autograd-graph lifetime rules of the fused decoder nodes:
  * two graphs with the same (B, T) alive at once own separate workspaces (saved activations), so
    `loss = f(a) + f(b)` and forward-validate-backward orders give the right gradients;
  * flat-gradient mode (DP all-reduce buffer) keeps accumulating correctly when p.grad is not reset to None;
  * a second backward through the same node raises instead of silently re-scaling saved gradients.
"""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import DicError
from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(L=50, D=64, A=32, E=16, H=32, V=101)
LENGTHS = [9, 9, 7, 4, 3, 2]


def make(seed):
    B = len(LENGTHS)
    g = torch.Generator().manual_seed(seed)
    F_rgb = torch.rand(B, CFG["L"], CFG["D"], generator=g)
    F_dep = torch.rand(B, CFG["L"], CFG["D"], generator=g)
    V = CFG["V"]
    caps = torch.full((B, max(LENGTHS)), V - 1, dtype=torch.int64)
    for b, n in enumerate(LENGTHS):
        caps[b, 0] = V - 4
        caps[b, 1:n - 1] = torch.randint(0, V - 4, (n - 2,), generator=g)
        caps[b, n - 1] = V - 3
    return F_rgb, F_dep, caps


def module(dev, w, flat=False):
    m = P.CD_RNNDecoderWithSoftAttention(CFG["A"], CFG["E"], CFG["D"], CFG["H"], CFG["V"])
    m.load_state_dict(w)
    m.precision = "fp32"
    m.flat_grads = flat
    return m.to(dev).eval()


def oracle_grads(w, batches):
    wo = {k: v.clone().double().requires_grad_(True) for k, v in w.items()}
    total = 0.0
    for F_rgb, F_dep, caps in batches:
        lo, _, ao = O.decoder_forward(wo, F_rgb.double(), F_dep.double(), caps, LENGTHS, hoist=True)
        total = total + O.caption_loss(lo, O.pack_targets(caps, LENGTHS), CFG["V"] - 1, ao, 0.7)
    total.backward()
    return {k: v.grad.numpy() for k, v in wo.items()}


def check(m, ref, scale=1.0):
    for k, p in m.named_parameters():
        got = p.grad.double().cpu().numpy()
        want = ref[k] * scale
        tol = 1e-4 * max(np.abs(want).max(), 1e-3) + 1e-7
        assert np.abs(got - want).max() <= tol, (k, float(np.abs(got - want).max()), tol)


@pytest.mark.parametrize("fused", [False, True])
def test_two_graphs_alive_at_once(fused, cuda_device):
    dev = cuda_device
    w = O.make_weights(CFG["A"], CFG["E"], CFG["D"], CFG["H"], CFG["V"], seed=31)
    a, b = make(32), make(33)
    m = module(dev, w)

    def loss_of(batch):
        F_rgb, F_dep, caps = (t.to(dev) for t in batch)
        if fused:
            return m.forward_loss(F_rgb, F_dep, caps, LENGTHS, ignore_index=CFG["V"] - 1, lam=0.7)
        out, alphas = m(F_rgb, F_dep, caps, LENGTHS)
        tg = O.pack_targets(batch[2], LENGTHS).to(dev)
        return (torch.nn.functional.cross_entropy(out.data, tg, ignore_index=CFG["V"] - 1)
                + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean())
    la = loss_of(a)                      # graph 1 (same B, T as graph 2)
    lb = loss_of(b)                      # graph 2: must not overwrite graph 1's saved activations
    with torch.no_grad():                # a validation forward in between shares only the cached workspace
        loss_of(b)
    (la + lb).backward()
    check(m, oracle_grads(w, [a, b]))


def test_flat_grads_accumulate(cuda_device):
    dev = cuda_device
    w = O.make_weights(CFG["A"], CFG["E"], CFG["D"], CFG["H"], CFG["V"], seed=41)
    a, b = make(42), make(43)
    m = module(dev, w, flat=True)

    def step(batch):
        F_rgb, F_dep, caps = (t.to(dev) for t in batch)
        m.forward_loss(F_rgb, F_dep, caps, LENGTHS, ignore_index=CFG["V"] - 1, lam=0.7).backward()
    step(a)
    eng = next(iter(m._engines.values()))
    flat = eng.grad_flat
    assert flat is not None
    p0 = next(m.parameters())
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    assert lo <= p0.grad.data_ptr() < hi          # p.grad aliases the flat buffer after the first backward
    step(b)                                       # accumulation: p.grad (still aliasing) += new gradients
    check(m, oracle_grads(w, [a, b]))
    m.zero_grad(set_to_none=False)                # zeroed in place: the alias survives
    step(a)
    check(m, oracle_grads(w, [a]))
    m.zero_grad(set_to_none=True)                 # the documented fast path
    step(b)
    check(m, oracle_grads(w, [b]))


def test_second_backward_raises(cuda_device):
    dev = cuda_device
    w = O.make_weights(CFG["A"], CFG["E"], CFG["D"], CFG["H"], CFG["V"], seed=51)
    F_rgb, F_dep, caps = (t.to(dev) for t in make(52))
    m = module(dev, w)
    loss = m.forward_loss(F_rgb, F_dep, caps, LENGTHS, ignore_index=CFG["V"] - 1, lam=0.7)
    loss.backward(retain_graph=True)
    with pytest.raises((DicError, RuntimeError)):
        loss.backward()
