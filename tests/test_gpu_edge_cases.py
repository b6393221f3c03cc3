"""Edge cases of the CUDA path against the oracle: degenerate batch / length, long captions
(multi-tile post-loop kernels), non-reference dimensions (generic kernel paths, column tails),
extreme raggedness, and loud failures on malformed input."""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O
from test_gpu_parity import build_module, make_case, relmax

pytestmark = pytest.mark.gpu

EDGE = {
    "single_image_single_step": dict(B=1, L=196, D=64, A=32, E=16, H=32, V=61, lengths=[2], seed=31),
    "long_captions_T40": dict(B=3, L=40, D=64, A=32, E=16, H=32, V=97, lengths=[41, 35, 12], seed=32),
    "concat_dims_D2080": dict(B=2, L=196, D=2080, A=128, E=128, H=128, V=300, lengths=[5, 4], seed=33),
    "odd_dims": dict(B=4, L=49, D=328, A=136, E=24, H=72, V=211, lengths=[6, 6, 3, 2], seed=34),
    "very_ragged": dict(B=5, L=30, D=64, A=32, E=16, H=32, V=71, lengths=[21, 3, 2, 2, 2], seed=35),
}


@pytest.mark.parametrize("case", list(EDGE))
def test_edge_forward_backward_fp32(case, cuda_device):
    cfg = dict(EDGE[case])
    lengths, V = cfg["lengths"], cfg["V"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    wo = {k: v.clone().double().requires_grad_(True) for k, v in w.items()}
    Fr = F_rgb.clone().double().requires_grad_(True)
    Fd = F_dep.clone().double().requires_grad_(True)
    lo, bsz, ao = O.decoder_forward(wo, Fr, Fd, caps, lengths, hoist=True)
    O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao).backward()
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, "fp32").eval()
    fr = F_rgb.to(cuda_device).requires_grad_(True)
    fd = F_dep.to(cuda_device).requires_grad_(True)
    out, alphas = m(fr, fd, caps.to(cuda_device), lengths)
    assert out.batch_sizes.tolist() == bsz
    assert tuple(alphas.shape) == (cfg["B"], len(bsz), cfg["L"])
    assert relmax(out.data.detach().cpu(), lo.detach()) <= 1e-4
    assert np.abs(alphas.detach().cpu().numpy() - ao.detach().numpy()).max() <= 1e-5
    tg = O.pack_targets(caps, lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    grads = dict(m.named_parameters())
    for k in _lib.PARAM_KEYS:
        ref = wo[k].grad.numpy()
        got = grads[k].grad.double().cpu().numpy()
        tol = 1e-6 if k == "attention.full_att.bias" else 1e-4 * np.abs(ref).max() + 1e-9
        assert np.abs(got - ref).max() <= tol, (k, float(np.abs(got - ref).max()), tol)
    ref = Fd.grad.numpy()
    assert np.abs(fd.grad.double().cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-12


@pytest.mark.parametrize("case", ["long_captions_T40", "odd_dims", "concat_dims_D2080"])
def test_edge_bf16_and_decode(case, cuda_device):
    cfg = dict(EDGE[case])
    lengths, V = cfg["lengths"], cfg["V"]
    if cfg["E"] % 8 or cfg["H"] % 8 or cfg["A"] % 8:
        pytest.skip("bf16 mode needs A, E, H multiples of 8")
    w, F_rgb, F_dep, caps = make_case(**cfg)
    lo, _, _ = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, hoist=True)
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, "bf16").eval()
    fr, fd = F_rgb.to(cuda_device).requires_grad_(True), F_dep.to(cuda_device).requires_grad_(True)
    out, alphas = m(fr, fd, caps.to(cuda_device), lengths)
    assert relmax(out.data.detach().cpu(), lo) <= 2e-2
    (out.data.float().square().mean() + alphas.sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    assert torch.isfinite(fd.grad).all()
    voc = O.synthetic_vocab(V)
    # fp32 decode against the oracle, including max_length = 1
    m32 = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, "fp32").eval()
    for T in (1, 6):
        ref, _, logits_o = O.greedy_decode(w, F_rgb, F_dep, V - 4, T, hoist=True)
        top2 = torch.stack(logits_o).topk(2, dim=-1).values
        if float((top2[..., 0] - top2[..., 1]).min()) > 1e-5:
            got = m32.batch_sample(F_rgb.to(cuda_device), F_dep.to(cuda_device), voc, max_length=T)
            np.testing.assert_array_equal(got, ref.numpy())
    res = m32.beam_search(F_rgb.to(cuda_device), F_dep.to(cuda_device), voc, beam=3, max_length=4)
    refb = O.beam_search(w, F_rgb, F_dep, V - 4, V - 3, 3, 4)
    np.testing.assert_array_equal(res["tokens"].cpu().numpy(), refb["tokens"].numpy())


def test_hard_attention_bf16_runs(cuda_device):
    cfg = dict(EDGE["long_captions_T40"], L=196)
    w, F_rgb, F_dep, caps = make_case(**cfg)
    m = build_module(P.CD_RNNDecoderWithHardAttention, w, cuda_device, "bf16", extra=("cuda:0",)).train()
    fd = F_dep.to(cuda_device).requires_grad_(True)
    torch.manual_seed(1)
    out = m(F_rgb.to(cuda_device), fd, caps.to(cuda_device), cfg["lengths"], torch.tensor(0.7))
    out.data.float().mean().backward()
    assert torch.isfinite(out.data).all() and torch.isfinite(fd.grad).all()
    toks = m.eval().batch_sample(F_rgb.to(cuda_device), fd.detach(), O.synthetic_vocab(cfg["V"]), max_length=3)
    assert toks.shape == (cfg["B"], 3)


def test_malformed_inputs_fail_loudly(cuda_device):
    cfg = dict(EDGE["very_ragged"])
    w, F_rgb, F_dep, caps = make_case(**cfg)
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).eval()
    fr, fd, cp = F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device)
    with pytest.raises(ValueError):
        m(fr, fd, cp, [3, 21, 2, 2, 2])                 # not sorted descending
    with pytest.raises(ValueError):
        m(fr, fd, cp, [21, 3, 2, 2, 1])                 # a caption with no target token
    with pytest.raises(ValueError):
        m(fr[:, :, :32].contiguous(), fd[:, :, :32].contiguous(), cp, cfg["lengths"])   # wrong dim_encoder
    with pytest.raises(ValueError):
        m(fr, fd[:3], cp, cfg["lengths"])               # depth batch mismatch
    with pytest.raises(P.DicError):
        m.beam_search(fr, fd, O.synthetic_vocab(cfg["V"]), beam=9, max_length=3)       # beam > DIC_MAX_BEAM
    with pytest.raises(P.DicError):
        m(fr.double(), fd.double(), cp, cfg["lengths"])  # unsupported dtype
    with pytest.raises((P.DicError, ValueError)):
        m(fr, fd, cp[:, :5].contiguous(), cfg["lengths"])   # captions shorter than the lengths say


def test_flat_grad_buffer_aliases_param_grads(cuda_device):
    """flat_grads: backward writes the 17 parameter gradients into one flat buffer and p.grad aliases
    it (what FlatGradAllReduce(module=...) all-reduces in place); values equal the default path."""
    from test_gpu_parity import CASES, build_module, make_case
    cfg = dict(CASES["small_ragged"])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    V = cfg["V"]

    def run(flat):
        m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).eval()
        m.flat_grads = flat
        loss = m.forward_loss(F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device), lengths,
                              ignore_index=V - 1)
        loss.backward()
        return m

    ref, got = run(False), run(True)
    eng = next(iter(got._engines.values()))
    assert eng.grad_flat is not None
    lo, hi = eng.grad_flat.data_ptr(), eng.grad_flat.data_ptr() + eng.grad_flat.numel() * 4
    n = 0
    for (k, p), (_, q) in zip(got.named_parameters(), ref.named_parameters()):
        assert lo <= p.grad.data_ptr() < hi, k
        # (weight gradients use atomic split-K: equal up to the fp32 summation order)
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7 + 1e-5 * float(q.grad.abs().max())), k
        assert p.grad.data_ptr() % 256 == 0, k      # every tensor on its own 256-byte boundary (vector stores / red.add)
        n += (p.numel() + 63) // 64 * 64
    assert n == eng.grad_flat.numel()
    # a second step re-uses the buffer after zero_grad(set_to_none=True)
    got.zero_grad(set_to_none=True)
    got.forward_loss(F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device), lengths,
                     ignore_index=V - 1).backward()
    for (k, p), (_, q) in zip(got.named_parameters(), ref.named_parameters()):
        assert lo <= p.grad.data_ptr() < hi, k
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7 + 1e-5 * float(q.grad.abs().max())), k


def test_fused_adamw_matches_torch(cuda_device):
    """dic_adamw_step == torch.optim.AdamW over several steps (fp32 rounding), odd sizes included."""
    torch.manual_seed(0)
    shapes = [(128, 2048), (10000, 128), (513,), (1,), (7, 3), (2048,)]
    ref = [torch.nn.Parameter(torch.randn(s, device=cuda_device)) for s in shapes]
    got = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2)
    o_got = P.FusedAdamW(got, lr=1e-3, weight_decay=1e-2)
    for step in range(5):
        for a, b in zip(ref, got):
            g = torch.randn_like(a) * (10.0 ** (step - 2))
            a.grad = g.clone()
            b.grad = g.clone()
        o_ref.step()
        o_got.step()
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (a - b).abs().max()
    sd = o_got.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    cpu_p = torch.nn.Parameter(torch.zeros(3))
    cpu_p.grad = torch.ones(3)
    with pytest.raises(P.DicError):          # no CPU fallback
        P.FusedAdamW([cpu_p]).step()


def test_cluster_fused_lstm_step_matches_default_path(cuda_device):
    """DIC_FUSED_GATES=1 (gates GEMM + split-K reduction over distributed shared memory + LSTM pointwise in
    one cluster kernel, csrc/gates_lstm.cuh) gives the same logits / tokens as the default two-kernel path.
    The switch is read once per process, so the fused run happens in a subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
import depth_image_captioning_pub_b200 as P
from oracle import decoder_oracle as O
A, E, D, H, V, L, B, T = 128, 128, 2048, 128, 1000, 196, 130, 6
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V); m.load_state_dict(O.make_weights(A, E, D, H, V, seed=5)); m.precision = "bf16"
m = m.cuda().eval()
g = torch.Generator().manual_seed(6)
Fr, Fd = torch.rand(B, L, D, generator=g).cuda(), torch.rand(B, L, D, generator=g).cuda()
caps = torch.randint(0, V - 4, (B, T + 1), generator=g).cuda(); caps[:, 0] = V - 4
out, alphas = m(Fr, Fd, caps, [T + 1] * B)
toks = m.batch_sample(Fr, Fd, O.synthetic_vocab(V), max_length=T)
torch.save({"logits": out.data.float().cpu(), "alphas": alphas.cpu(), "toks": torch.from_numpy(toks)}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    res = {}
    for flag in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            env = dict(os.environ, DIC_FUSED_GATES=flag)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            res[flag] = torch.load(f.name)
    a, b = res["0"], res["1"]
    assert float((a["logits"] - b["logits"]).abs().max()) <= 2e-2 * float(a["logits"].abs().max())
    assert float((a["alphas"] - b["alphas"]).abs().max()) <= 2e-3
    assert (a["toks"] == b["toks"]).float().mean() >= 0.98
