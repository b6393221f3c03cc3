"""Pin the CPU oracle to vectors computed by the UNMODIFIED reference modules
(tests/golden/*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, split_steps
from oracle import decoder_oracle as O

SOFT = ["depth_soft", "base_soft", "depth_soft_peaked"]
HARD = ["depth_hard", "base_hard"]


def _inputs(rec, grad=False):
    F_rgb = torch.from_numpy(rec["F_rgb"]).clone().requires_grad_(grad)
    F_dep = torch.from_numpy(rec["F_dep"]).clone().requires_grad_(grad) if int(rec["depth"]) else None
    return F_rgb, F_dep, torch.from_numpy(rec["captions"]), rec["lengths"].tolist()


@pytest.mark.parametrize("name", SOFT)
@pytest.mark.parametrize("hoist", [False, True])
def test_soft_forward_backward(name, hoist):
    rec, w, g = load_golden(name)
    w = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    F_rgb, F_dep, caps, lengths = _inputs(rec, grad=True)
    logits, bsz, alphas = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, hoist=hoist)
    assert bsz == rec["batch_sizes"].tolist()
    np.testing.assert_allclose(logits.detach().numpy(), rec["logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(alphas.detach().numpy(), rec["alphas"], rtol=0, atol=2e-6)
    V = logits.shape[1]
    loss = O.caption_loss(logits, O.pack_targets(caps, lengths), V - 1, alphas)
    np.testing.assert_allclose(float(loss.detach()), float(rec["loss"]), rtol=1e-6)
    loss.backward()
    for k in O.KEYS:
        ref = g[k].numpy()
        tol = 1e-6 * max(1.0, np.abs(ref).max()) + 1e-7
        np.testing.assert_allclose(w[k].grad.numpy(), ref, rtol=0, atol=tol, err_msg=k)
    np.testing.assert_allclose(F_rgb.grad.numpy(), rec["g_F_rgb"], rtol=0, atol=1e-6)
    if F_dep is not None:
        np.testing.assert_allclose(F_dep.grad.numpy(), rec["g_F_dep"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", SOFT)
def test_soft_greedy(name):
    rec, w, _ = load_golden(name)
    F_rgb, F_dep, _, _ = _inputs(rec)
    V = w["linear.weight"].shape[0]
    T = rec["greedy"].shape[1]
    toks, alphas, _ = O.greedy_decode(w, F_rgb, F_dep, V - 4, T)
    np.testing.assert_array_equal(toks.numpy(), rec["greedy"])
    toks_nosm, _, _ = O.greedy_decode(w, F_rgb, F_dep, V - 4, T, use_softmax=False, hoist=True)
    np.testing.assert_array_equal(toks_nosm.numpy(), rec["greedy"])
    t1, a1, _ = O.greedy_decode(w, F_rgb[:1], None if F_dep is None else F_dep[:1], V - 4, T)
    np.testing.assert_array_equal(t1[0].numpy(), rec["sample_tokens"])
    np.testing.assert_allclose(torch.cat(a1).numpy(), rec["sample_alphas"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", HARD)
def test_hard_paths(name):
    rec, w, g = load_golden(name)
    w = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    F_rgb, F_dep, caps, lengths = _inputs(rec, grad=True)
    sizes = rec["batch_sizes"].tolist()
    u = split_steps(torch.from_numpy(rec["u_fwd"]), sizes)
    logits, bsz, _ = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, attn="gumbel_softmax",
                                       u_steps=u, temp=torch.tensor(float(rec["temp"])))
    np.testing.assert_allclose(logits.detach().numpy(), rec["logits"], rtol=0, atol=2e-6)
    V = logits.shape[1]
    loss = O.caption_loss(logits, O.pack_targets(caps, lengths), V - 1)
    loss.backward()
    for k in O.KEYS:
        ref = g[k].numpy()
        tol = 1e-6 * max(1.0, np.abs(ref).max()) + 1e-7
        np.testing.assert_allclose(w[k].grad.numpy(), ref, rtol=0, atol=tol, err_msg=k)
    np.testing.assert_allclose(F_rgb.grad.numpy(), rec["g_F_rgb"], rtol=0, atol=1e-6)
    with torch.no_grad():
        wd = {k: v.detach() for k, v in w.items()}
        Fr, Fd = F_rgb.detach(), None if F_dep is None else F_dep.detach()
        ue = split_steps(torch.from_numpy(rec["u_eval"]), sizes)
        ev, _, _ = O.decoder_forward(wd, Fr, Fd, caps, lengths, attn="gumbel_max", u_steps=ue)
        np.testing.assert_allclose(ev.numpy(), rec["eval_logits"], rtol=0, atol=2e-6)
        T = rec["greedy"].shape[1]
        B = Fr.shape[0]
        ug = split_steps(torch.from_numpy(rec["u_greedy"]), [B] * T)
        toks, _, _ = O.greedy_decode(wd, Fr, Fd, V - 4, T, attn="gumbel_max", u_steps=ug)
        np.testing.assert_array_equal(toks.numpy(), rec["greedy"])
        us = split_steps(torch.from_numpy(rec["u_sample"]), [1] * T)
        t1, a1, _ = O.greedy_decode(wd, Fr[:1], None if Fd is None else Fd[:1], V - 4, T,
                                    attn="gumbel_max", u_steps=us)
        np.testing.assert_array_equal(t1[0].numpy(), rec["sample_tokens"])
        np.testing.assert_array_equal(torch.cat(a1).numpy(), rec["sample_alphas"])


def test_beam_spec_properties():
    """Beam search is the build's own spec (no reference).  Sanity properties:
    beam=1 equals greedy; best score is non-increasing in t; backpointers in range."""
    w = O.make_weights(16, 8, 24, 16, 37, seed=5)
    g = torch.Generator().manual_seed(6)
    F = torch.rand(4, 20, 24, generator=g)
    V = 37
    T = 7
    greedy, _, _ = O.greedy_decode(w, F, None, V - 4, T, hoist=True, use_softmax=False)
    b1 = O.beam_search(w, F, None, V - 4, V - 3, 1, T)
    # beam=1 follows greedy until the first <end>, then stays frozen on <end>
    for b in range(4):
        n = int(b1["lengths"][b])
        np.testing.assert_array_equal(b1["tokens"][b, :n].numpy(), greedy[b, :n].numpy())
        assert (b1["tokens"][b, n:] == V - 3).all()
    b5 = O.beam_search(w, F, None, V - 4, V - 3, 5, T)
    assert (b5["scores"] >= b1["scores"] - 1e-5).all()
    assert int(b5["back"].min()) >= 0 and int(b5["back"].max()) < 5
    s = b5["all_scores"][:, :, 0]
    assert (s[1:] <= s[:-1] + 1e-6).all()


def test_lookahead_beam_order_is_the_same_search():
    """The kernel order of the CUDA beam path (attention of step t+1 from the un-reordered h_t, gate GEMM in parent
    order, LSTM following the backpointers, the token's share of the gates from a table) restated in the oracle:
    in float64 it must pick the same tokens / backpointers and reach the same scores as the serial specification."""
    A, E, D, H, V, L, B, K, T = 16, 12, 24, 20, 60, 9, 5, 3, 7
    w = {k: v.double() for k, v in O.make_weights(A, E, D, H, V, seed=5).items()}
    g = torch.Generator().manual_seed(6)
    F_rgb = torch.rand(B, L, D, generator=g, dtype=torch.float64)
    F_dep = torch.rand(B, L, D, generator=g, dtype=torch.float64)
    voc = O.synthetic_vocab(V)
    ref = O.beam_search(w, F_rgb, F_dep, voc["<start>"], voc["<end>"], K, T)
    look = O.beam_search_lookahead(w, F_rgb, F_dep, voc["<start>"], voc["<end>"], K, T)
    assert torch.equal(look["tokens"], ref["tokens"])
    assert torch.equal(look["back"], ref["back"])
    assert torch.equal(look["toks"], ref["toks"])
    assert float((look["scores"] - ref["scores"]).abs().max()) <= 1e-10
