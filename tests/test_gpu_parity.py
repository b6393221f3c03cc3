"""GPU parity tests: the CUDA path (through the drop-in modules and the C ABI) against
(1) golden vectors computed by the unmodified reference modules and (2) the CPU oracle on
seeded inputs.  Tolerances (BASELINE.json north_star):
  fp32 mode: logits max|d|/max|logit| <= 1e-4, alpha max|d| <= 1e-5, tokens identical
  bf16 mode: logits <= 2e-2
  top-k / backpointers: bit-exact given identical logits
"""
import ctypes as C

import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from conftest import load_golden, split_steps
from depth_image_captioning_pub_b200 import _lib
from depth_image_captioning_pub_b200.engine import Engine, batch_sizes_from_lengths
from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu

LOGIT_TOL_F32 = 1e-4
ALPHA_TOL_F32 = 1e-5
LOGIT_TOL_BF16 = 2e-2


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def build_module(cls, w, dev, precision="fp32", extra=()):
    A, D = w["attention.encoder_att.weight"].shape
    V, E = w["embed.weight"].shape
    H = w["decode_step.weight_hh"].shape[1]
    m = cls(A, E, D, H, V, *extra)
    m.load_state_dict(w)
    m.precision = precision
    return m.to(dev)


def w2i(V):
    return O.synthetic_vocab(V)


# ------------------------------------------------------------------------------------------
# golden vectors from the reference modules
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["depth_soft", "base_soft", "depth_soft_peaked"])
def test_golden_soft_forward_backward(name, cuda_device):
    rec, w, g = load_golden(name)
    depth = bool(int(rec["depth"]))
    cls = P.CD_RNNDecoderWithSoftAttention if depth else P.RNNDecoderWithSoftAttention
    m = build_module(cls, w, cuda_device).eval()
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device).requires_grad_(True)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device).requires_grad_(True)
    caps = torch.from_numpy(rec["captions"]).to(cuda_device)
    lengths = rec["lengths"].tolist()
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    out, alphas = m(*feats, caps, lengths)
    assert out.batch_sizes.tolist() == rec["batch_sizes"].tolist()
    assert relmax(out.data.detach().cpu(), rec["logits"]) <= LOGIT_TOL_F32
    assert np.abs(alphas.detach().cpu().numpy() - rec["alphas"]).max() <= ALPHA_TOL_F32
    V = out.data.shape[1]
    tg = O.pack_targets(caps.cpu(), lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()       # depth_train.py:214-216
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


@pytest.mark.parametrize("name", ["depth_soft", "base_soft", "depth_soft_peaked"])
def test_golden_soft_greedy(name, cuda_device):
    rec, w, _ = load_golden(name)
    depth = bool(int(rec["depth"]))
    cls = P.CD_RNNDecoderWithSoftAttention if depth else P.RNNDecoderWithSoftAttention
    m = build_module(cls, w, cuda_device).eval()
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device)
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    V = w["linear.weight"].shape[0]
    T = rec["greedy"].shape[1]
    toks = m.batch_sample(*feats, w2i(V), max_length=T)
    assert toks.dtype == np.int64 and toks.shape == rec["greedy"].shape
    np.testing.assert_array_equal(toks, rec["greedy"])
    p1, a1 = m.sample(*(f[:1] for f in feats), w2i(V), max_length=T)
    assert p1 == rec["sample_tokens"].tolist()
    assert len(a1) == T and tuple(a1[0].shape) == (1, F_rgb.shape[1])
    got = torch.cat(a1).cpu().numpy()
    assert np.abs(got - rec["sample_alphas"]).max() <= ALPHA_TOL_F32


@pytest.mark.parametrize("name", ["depth_hard", "base_hard"])
def test_golden_hard_paths(name, cuda_device):
    rec, w, g = load_golden(name)
    depth = bool(int(rec["depth"]))
    seed = 400 if depth else 500      # oracle/make_golden.py
    cls = P.CD_RNNDecoderWithHardAttention if depth else P.RNNDecoderWithHardAttention
    m = build_module(cls, w, cuda_device, extra=("cuda:0",)).eval()
    F_rgb = torch.from_numpy(rec["F_rgb"]).to(cuda_device).requires_grad_(True)
    F_dep = torch.from_numpy(rec["F_dep"]).to(cuda_device).requires_grad_(True)
    caps = torch.from_numpy(rec["captions"]).to(cuda_device)
    lengths = rec["lengths"].tolist()
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    V = w["linear.weight"].shape[0]
    # Gumbel-softmax forward + backward; same CPU-generator seed as the reference run
    torch.manual_seed(seed + 2)
    out = m(*feats, caps, lengths, torch.tensor(float(rec["temp"])))
    assert relmax(out.data.detach().cpu(), rec["logits"]) <= LOGIT_TOL_F32
    tg = O.pack_targets(caps.cpu(), lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss.backward()
    grads = dict(m.named_parameters())
    for k in _lib.PARAM_KEYS:
        ref = g[k].numpy()
        got = grads[k].grad.cpu().numpy()
        tol = 2e-4 * max(np.abs(ref).max(), 1e-3) + 1e-7
        assert np.abs(got - ref).max() <= tol, (k, np.abs(got - ref).max(), tol)
    assert np.abs(F_rgb.grad.cpu().numpy() - rec["g_F_rgb"]).max() <= 2e-4 * np.abs(rec["g_F_rgb"]).max() + 1e-8
    fs = tuple(f.detach() for f in feats)
    torch.manual_seed(seed + 3)
    ev = m.eval_forward(*fs, caps, lengths)
    assert relmax(ev.data.cpu(), rec["eval_logits"]) <= LOGIT_TOL_F32
    T = rec["greedy"].shape[1]
    torch.manual_seed(seed + 4)
    toks = m.batch_sample(*fs, w2i(V), max_length=T)
    np.testing.assert_array_equal(toks, rec["greedy"])
    torch.manual_seed(seed + 5)
    p1, a1 = m.sample(*(f[:1] for f in fs), w2i(V), max_length=T)
    assert p1 == rec["sample_tokens"].tolist()
    assert a1[0].dtype == torch.int64
    np.testing.assert_array_equal(torch.cat(a1).cpu().numpy(), rec["sample_alphas"])


# ------------------------------------------------------------------------------------------
# oracle comparisons on seeded inputs (reference dims L=196 D=2048 A=E=H=128 V=10000 included)
# ------------------------------------------------------------------------------------------
def make_case(B, L, D, A, E, H, V, lengths, seed, peak=1.0):
    w = O.make_weights(A, E, D, H, V, seed=seed)
    w["attention.full_att.weight"] = w["attention.full_att.weight"] * peak
    g = torch.Generator().manual_seed(seed + 1)
    F_rgb = torch.rand(B, L, D, generator=g)
    F_dep = torch.rand(B, L, D, generator=g)
    voc = O.synthetic_vocab(V)
    caps = torch.full((B, max(lengths)), voc["<null>"], dtype=torch.int64)
    for b, n in enumerate(lengths):
        caps[b, 0] = voc["<start>"]
        caps[b, 1:n - 1] = torch.randint(0, V - 4, (n - 2,), generator=g)
        caps[b, n - 1] = voc["<end>"]
    return w, F_rgb, F_dep, caps


CASES = {
    "small_ragged": dict(B=6, L=50, D=64, A=32, E=16, H=32, V=101, lengths=[9, 9, 7, 4, 3, 2], seed=7),
    "mid_uniform": dict(B=5, L=196, D=256, A=64, E=32, H=64, V=1000, lengths=[6] * 5, seed=8),
    "ref_dims": dict(B=4, L=196, D=2048, A=128, E=128, H=128, V=10000, lengths=[8, 7, 7, 3], seed=9),
    "ref_dims_peaked": dict(B=3, L=196, D=2048, A=128, E=128, H=128, V=10000, lengths=[5, 4, 4], seed=10, peak=50.0),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_backward_vs_oracle(case, precision, cuda_device):
    cfg = dict(CASES[case])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    V = cfg["V"]
    # oracle in fp64 (hoisted form).  The gradients of the attention projections pass through
    # softmax-backward and a ReLU mask and are cancellation dominated on iid-random annotations:
    # the CPU fp32 oracle itself is only good to ~2.5e-3 there (scripts/grad_noise.py), so the
    # reference for gradients is fp64.
    wo = {k: v.clone().double().requires_grad_(True) for k, v in w.items()}
    Fr = F_rgb.clone().double().requires_grad_(True)
    Fd = F_dep.clone().double().requires_grad_(True)
    lo, bsz, ao = O.decoder_forward(wo, Fr, Fd, caps, lengths, hoist=True)
    loss_o = O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao)
    loss_o.backward()
    # CUDA path
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, precision).eval()
    Fr_g = F_rgb.to(cuda_device).requires_grad_(True)
    Fd_g = F_dep.to(cuda_device).requires_grad_(True)
    out, alphas = m(Fr_g, Fd_g, caps.to(cuda_device), lengths)
    assert out.batch_sizes.tolist() == bsz
    peaked = cfg.get("peak", 1.0) != 1.0
    if precision == "fp32":
        ltol, atol, gtol, gtol_att = LOGIT_TOL_F32, ALPHA_TOL_F32, 5e-5, 5e-5
    else:
        # bf16 storage of annotations / att1: the spec bounds the logits (2e-2).  alpha and the
        # gradients are reported bounds; the attention-projection gradients see ~0.4% of the ReLU
        # masks flip under bf16 rounding of att1, which the softmax cancellation amplifies.
        ltol, atol, gtol, gtol_att = LOGIT_TOL_BF16, (2e-2 if peaked else 2e-3), 5e-2, 0.3     # measured worst case 0.24 (peaked, B = 3); profiles/r02_mask_flip_study.txt
    assert relmax(out.data.detach().cpu(), lo.detach()) <= ltol
    assert np.abs(alphas.detach().cpu().numpy() - ao.detach().numpy()).max() <= atol
    tg = O.pack_targets(caps, lengths).to(cuda_device)
    loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    grads = dict(m.named_parameters())
    bad = []
    for k in _lib.PARAM_KEYS:
        ref = wo[k].grad.numpy()
        got = grads[k].grad.double().cpu().numpy()
        assert np.isfinite(got).all(), k
        if k == "attention.full_att.bias":       # exactly zero by softmax shift invariance
            tol = 1e-6
        else:
            att = k.startswith("attention.encoder_att") or k.startswith("attention.decoder_att")
            tol = (gtol_att if att else gtol) * np.abs(ref).max() + 1e-9
        err = float(np.abs(got - ref).max())
        if precision == "bf16" and k.startswith("attention."):
            print(f"[bf16 grad] {case} {k}: max-abs error / max|ref| = {err / (np.abs(ref).max() + 1e-30):.4f}")
        if err > tol:
            bad.append((k, err, tol))
    for name, got, ref in (("d_features", Fr_g.grad, Fr.grad), ("d_depth_features", Fd_g.grad, Fd.grad)):
        ref = ref.numpy()
        err = float(np.abs(got.double().cpu().numpy() - ref).max())
        tol = gtol * np.abs(ref).max() + 1e-12
        if err > tol:
            bad.append((name, err, tol))
    assert not bad, bad


def test_train_mode_dropout_mask_path(cuda_device):
    """Explicit dropout mask through the engine == oracle with the same mask."""
    cfg = dict(CASES["small_ragged"])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    bsz = batch_sizes_from_lengths(lengths)
    g = torch.Generator().manual_seed(3)
    mask = (torch.rand(sum(bsz), cfg["H"], generator=g) >= 0.5).float() / 0.5
    lo, _, ao = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, dropout_masks=split_steps(mask, bsz), hoist=True)
    eng = Engine(cfg["L"], cfg["D"], cfg["A"], cfg["E"], cfg["H"], cfg["V"], "fp32", cuda_device)
    params = [w[k].to(cuda_device) for k in _lib.PARAM_KEYS]
    eng.ensure_packed(params)
    ws = eng.train_workspace(cfg["B"], len(bsz), fresh=True)
    logits, alphas = eng.forward(_lib.ATTN_SOFT, F_rgb.to(cuda_device), F_dep.to(cuda_device),
                                 caps.to(cuda_device), bsz, None, 1.0, mask.to(cuda_device), ws)
    assert relmax(logits.cpu(), lo) <= LOGIT_TOL_F32
    assert np.abs(alphas.cpu().numpy() - ao.numpy()).max() <= ALPHA_TOL_F32
    # module-level train mode draws its own mask: only check it runs and differs from eval
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).train()
    out_t, _ = m(F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device), lengths)
    out_e, _ = m.eval()(F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device), lengths)
    assert torch.isfinite(out_t.data).all() and not torch.allclose(out_t.data, out_e.data)


@pytest.mark.parametrize("case", ["small_ragged", "ref_dims"])
@pytest.mark.parametrize("attn", ["gumbel_softmax", "gumbel_max"])
def test_hard_vs_oracle(case, attn, cuda_device):
    cfg = dict(CASES[case])
    lengths = cfg["lengths"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    bsz = batch_sizes_from_lengths(lengths)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(sum(bsz), cfg["L"], generator=g)
    temp = 0.7
    lo, _, _ = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, attn=attn, u_steps=split_steps(u, bsz),
                                 temp=torch.tensor(temp), hoist=True)
    eng = Engine(cfg["L"], cfg["D"], cfg["A"], cfg["E"], cfg["H"], cfg["V"], "fp32", cuda_device)
    eng.ensure_packed([w[k].to(cuda_device) for k in _lib.PARAM_KEYS])
    ws = eng.train_workspace(cfg["B"], len(bsz), fresh=True)
    mode = _lib.ATTN_GUMBEL_SOFTMAX if attn == "gumbel_softmax" else _lib.ATTN_GUMBEL_MAX
    logits, alphas = eng.forward(mode, F_rgb.to(cuda_device), F_dep.to(cuda_device), caps.to(cuda_device), bsz,
                                 u.to(cuda_device), temp, None, ws)
    assert relmax(logits.cpu(), lo) <= LOGIT_TOL_F32
    if attn == "gumbel_max":
        a = alphas.cpu()
        valid = torch.zeros(cfg["B"], len(bsz), dtype=torch.bool)
        for t, n in enumerate(bsz):
            valid[:n, t] = True
        assert torch.equal(a.sum(-1)[valid], torch.ones(int(valid.sum())))
        assert ((a == 0) | (a == 1)).all()


@pytest.mark.parametrize("case", ["small_ragged", "ref_dims"])
def test_greedy_vs_oracle(case, cuda_device):
    cfg = dict(CASES[case])
    w, F_rgb, F_dep, _ = make_case(**cfg)
    V, T = cfg["V"], 8
    toks_o, alphas_o, logits_o = O.greedy_decode(w, F_rgb, F_dep, V - 4, T, hoist=True)
    # margins: the reference's argmax(softmax) can tie where logits do not (SURVEY hard part 8)
    top2 = torch.stack(logits_o).topk(2, dim=-1).values
    assert float((top2[..., 0] - top2[..., 1]).min()) > 1e-5, "pick another seed: near-tie in the oracle"
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).eval()
    eng = m._engine(cfg["L"], cuda_device)
    eng.ensure_packed(m._param_list())
    tokens, alphas, logits = eng.greedy(_lib.ATTN_SOFT, F_rgb.to(cuda_device), F_dep.to(cuda_device), V - 4, T,
                                        want_alphas=True, want_logits=True)
    np.testing.assert_array_equal(tokens.cpu().numpy(), toks_o.numpy())
    assert relmax(logits.cpu(), torch.stack(logits_o)) <= LOGIT_TOL_F32
    assert np.abs(alphas.cpu().numpy() - torch.stack(alphas_o).numpy()).max() <= ALPHA_TOL_F32
    # bf16 mode: logits within 2e-2 while the token history is the same
    mb = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device, "bf16").eval()
    engb = mb._engine(cfg["L"], cuda_device)
    engb.ensure_packed(mb._param_list())
    tb, _, lb = engb.greedy(_lib.ATTN_SOFT, F_rgb.to(cuda_device), F_dep.to(cuda_device), V - 4, T, want_logits=True)
    same = (tb.cpu() == toks_o)
    first_div = torch.where(same.all(dim=1), torch.full((cfg["B"],), T), (~same).to(torch.int64).argmax(dim=1))
    lo = torch.stack(logits_o)
    for b in range(cfg["B"]):
        n = int(first_div[b]) + 1 if int(first_div[b]) < T else T
        assert relmax(lb[:n, b].cpu(), lo[:n, b]) <= LOGIT_TOL_BF16


# ------------------------------------------------------------------------------------------
# beam search (the build's own spec; integer parts bit-exact vs the oracle)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,K,V", [(3, 5, 101), (2, 1, 17), (4, 8, 10000), (5, 3, 257)])
def test_beam_select_bit_exact(B, K, V, cuda_device):
    g = torch.Generator().manual_seed(B * 1000 + K)
    logits = torch.randn(B, K, V, generator=g)
    # force exact ties inside and across rows
    logits[:, :, 5] = logits[:, :, 3]
    if K > 1:
        logits[:, 1] = logits[:, 0]
    scores = torch.randn(B, K, generator=g)
    if K > 1:
        scores[:, 1] = scores[:, 0]
    finished = torch.zeros(B, K, dtype=torch.bool)
    finished[0, K - 1] = True
    if B > 1:
        finished[1] = True              # every row finished
    end_id = V - 3
    lib = _lib.load()
    dev = cuda_device
    lg = logits.reshape(B * K, V).to(dev).contiguous()
    lse = torch.empty(B * K, device=dev)
    _lib.check(lib.dic_row_lse(lg.data_ptr(), B * K, V, lse.data_ptr(), _lib.stream_ptr(dev)))
    lse_ref = torch.logsumexp(logits.reshape(B * K, V), dim=1)
    assert np.abs(lse.cpu().numpy() - lse_ref.numpy()).max() <= 2e-6 * max(1.0, float(lse_ref.abs().max()))
    for step0 in (False, True):
        sc = scores.clone()
        if step0:
            sc[:] = float("-inf")
            sc[:, 0] = 0.0
        ns, back, tok, nf = O.beam_select(sc, finished if not step0 else torch.zeros_like(finished),
                                          logits, lse.cpu().reshape(B, K), end_id)
        fin_in = (finished if not step0 else torch.zeros_like(finished)).to(torch.uint8).to(dev)
        o_s = torch.empty(B, K, device=dev)
        o_b = torch.empty(B, K, dtype=torch.int32, device=dev)
        o_t = torch.empty(B, K, dtype=torch.int32, device=dev)
        o_f = torch.empty(B, K, dtype=torch.uint8, device=dev)
        wsz = lib.dic_beam_select_workspace_bytes(B, K)
        wsb = torch.empty(wsz, dtype=torch.uint8, device=dev)
        _lib.check(lib.dic_beam_select(sc.to(dev).data_ptr(), fin_in.data_ptr(), lg.data_ptr(), lse.data_ptr(), B, K,
                                       V, end_id, o_s.data_ptr(), o_b.data_ptr(), o_t.data_ptr(), o_f.data_ptr(),
                                       wsb.data_ptr(), wsz, _lib.stream_ptr(dev)))
        np.testing.assert_array_equal(o_b.cpu().numpy(), back.numpy().astype(np.int32))
        np.testing.assert_array_equal(o_t.cpu().numpy(), tok.numpy().astype(np.int32))
        np.testing.assert_array_equal(o_f.cpu().numpy().astype(bool), nf.numpy())
        np.testing.assert_array_equal(o_s.cpu().numpy().view(np.uint32), ns.numpy().view(np.uint32))


@pytest.mark.parametrize("case,beam", [("small_ragged", 5), ("small_ragged", 1), ("ref_dims", 5), ("mid_uniform", 3)])
def test_beam_search_vs_oracle(case, beam, cuda_device):
    cfg = dict(CASES[case])
    w, F_rgb, F_dep, _ = make_case(**cfg)
    V, T = cfg["V"], 7
    m = build_module(P.CD_RNNDecoderWithSoftAttention, w, cuda_device).eval()
    eng = m._engine(cfg["L"], cuda_device)
    eng.ensure_packed(m._param_list())
    got = eng.beam(F_rgb.to(cuda_device), F_dep.to(cuda_device), V - 4, V - 3, beam, T, trace=True, want_logits=True)
    # oracle driven with the GPU's own lse so that the selection arithmetic sees identical inputs
    lse_gpu = got["lse"].cpu()
    ref = O.beam_search(w, F_rgb, F_dep, V - 4, V - 3, beam, T)
    # numerics of the float part
    B = cfg["B"]
    ref_logits = torch.stack([l.reshape(B * beam, V) for l in ref["logits"]])
    # step 0 is path independent (all rows of an image start from the same state)
    assert relmax(got["logits"][0].cpu(), ref_logits[0]) <= LOGIT_TOL_F32
    # end-to-end: identical tokens, lengths, backpointers; scores within fp32 noise
    np.testing.assert_array_equal(got["tokens"].cpu().numpy(), ref["tokens"].numpy())
    np.testing.assert_array_equal(got["lengths"].cpu().numpy(), ref["lengths"].numpy().astype(np.int32))
    np.testing.assert_array_equal(got["back"].cpu().numpy(), ref["back"].numpy())
    assert np.abs(got["scores"].cpu().numpy() - ref["scores"].numpy()).max() <= 1e-4
    assert np.abs(lse_gpu.numpy() - ref["lse"].numpy()).max() <= 1e-4
    # module-level API
    res = m.beam_search(F_rgb.to(cuda_device), F_dep.to(cuda_device), O.synthetic_vocab(V), beam=beam, max_length=T)
    np.testing.assert_array_equal(res["tokens"].cpu().numpy(), ref["tokens"].numpy())


def test_standalone_attention_modules(cuda_device):
    B, L, D, A, H = 3, 196, 64, 32, 32
    w = O.make_weights(A, 16, D, H, 50, seed=21)
    att = P.Soft_Attention(D, H, A).to(cuda_device)
    with torch.no_grad():
        att.encoder_att.weight.copy_(w["attention.encoder_att.weight"]); att.encoder_att.bias.copy_(w["attention.encoder_att.bias"])
        att.decoder_att.weight.copy_(w["attention.decoder_att.weight"]); att.decoder_att.bias.copy_(w["attention.decoder_att.bias"])
        att.full_att.weight.copy_(w["attention.full_att.weight"]); att.full_att.bias.copy_(w["attention.full_att.bias"])
    g = torch.Generator().manual_seed(22)
    F = torch.rand(B, L, D, generator=g)
    h = torch.randn(B, H, generator=g)
    ctx_o, al_o = O.soft_attention(w, F, h)
    with torch.no_grad():
        ctx, al = att(F.to(cuda_device), h.to(cuda_device))
    assert np.abs(al.cpu().numpy() - al_o.numpy()).max() <= ALPHA_TOL_F32
    assert relmax(ctx.cpu(), ctx_o) <= 1e-5
    hard = P.Hard_Attention(D, H, A).to(cuda_device)
    hard.load_state_dict(att.state_dict())
    torch.manual_seed(5)
    u = torch.rand(B, 196)
    ctx_o, al_o = O.gumbel_max_attention(w, F, h, u)
    torch.manual_seed(5)
    ctx, al = hard.Hard_sample(F.to(cuda_device), h.to(cuda_device), "cuda:0")
    assert al.dtype == torch.int64
    np.testing.assert_array_equal(al.cpu().numpy(), al_o.numpy())
    assert relmax(ctx.cpu(), ctx_o) <= 1e-6
    ctx_o, al_o = O.gumbel_softmax_attention(w, F, h, u, torch.tensor(0.6))
    torch.manual_seed(5)
    with torch.no_grad():
        ctx, al = hard(F.to(cuda_device), h.to(cuda_device), "cuda:0", torch.tensor(0.6))
    assert np.abs(al.cpu().numpy() - al_o.numpy()).max() <= ALPHA_TOL_F32


# ------------------------------------------------------------------------------------------
# GEMM engines
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 53, 29), (256, 512, 2304), (300, 128, 64), (1024, 256, 128)])
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_generic_gemm(M, N, K, dt, cuda_device):
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    tdt = torch.float32 if dt == "f32" else torch.bfloat16
    Ad, Bd = A.to(cuda_device, tdt), Bm.to(cuda_device, tdt)
    Cd = torch.empty(M, N, device=cuda_device)
    code = _lib.DIC_F32 if dt == "f32" else _lib.DIC_BF16
    _lib.check(lib.dic_gemm_nt(0, M, N, K, Ad.data_ptr(), code, Bd.data_ptr(), code, bias.to(cuda_device).data_ptr(),
                               Cd.data_ptr(), _lib.stream_ptr(cuda_device)))
    ref = Ad.double().cpu() @ Bd.double().cpu().t() + bias.double()
    assert relmax(Cd.cpu(), ref) <= 2e-6 * max(1, K) ** 0.5


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (300, 200, 2304), (6272, 128, 2048),
                                   (512, 10000, 128), (129, 2176, 128), (64, 264, 72)])
def test_tcgen05_gemm(M, N, K, cuda_device):
    """tcgen05/TMA engine against an fp64 product of the same bf16 operands."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(cuda_device, torch.bfloat16)
    Bm = torch.randn(N, K, generator=g).to(cuda_device, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    Cd = torch.full((M, N), float("nan"), device=cuda_device)
    _lib.check(lib.dic_gemm_nt(1, M, N, K, A.data_ptr(), _lib.DIC_BF16, Bm.data_ptr(), _lib.DIC_BF16,
                               bias.data_ptr(), Cd.data_ptr(), _lib.stream_ptr(cuda_device)))
    torch.cuda.synchronize()
    ref = A.double().cpu() @ Bm.double().cpu().t() + bias.double().cpu()
    assert torch.isfinite(Cd).all()
    assert relmax(Cd.cpu(), ref) <= 1e-5
