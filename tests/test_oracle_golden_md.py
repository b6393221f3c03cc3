"""Oracle vs the reference's concat-fusion (MD_) decoders: cat(features, depth) then the base path
(depth_models.py:376)."""
import numpy as np
import torch

from conftest import load_golden, split_steps
from oracle import decoder_oracle as O


def test_md_soft_oracle():
    rec, w, g = load_golden("md_soft")
    w = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    F = torch.cat((torch.from_numpy(rec["F_rgb"]), torch.from_numpy(rec["F_dep"])), dim=2).requires_grad_(True)
    caps, lengths = torch.from_numpy(rec["captions"]), rec["lengths"].tolist()
    logits, bsz, alphas = O.decoder_forward(w, F, None, caps, lengths, hoist=True)
    assert bsz == rec["batch_sizes"].tolist()
    np.testing.assert_allclose(logits.detach().numpy(), rec["logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(alphas.detach().numpy(), rec["alphas"], rtol=0, atol=2e-6)
    V = logits.shape[1]
    O.caption_loss(logits, O.pack_targets(caps, lengths), V - 1, alphas).backward()
    for k in O.KEYS:
        ref = g[k].numpy()
        np.testing.assert_allclose(w[k].grad.numpy(), ref, rtol=0, atol=1e-6 * max(1.0, np.abs(ref).max()) + 1e-7)
    d_rgb = rec["F_rgb"].shape[2]
    np.testing.assert_allclose(F.grad[:, :, :d_rgb].numpy(), rec["g_F_rgb"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(F.grad[:, :, d_rgb:].numpy(), rec["g_F_dep"], rtol=0, atol=1e-6)
    toks, _, _ = O.greedy_decode({k: v.detach() for k, v in w.items()}, F.detach(), None, V - 4, rec["greedy"].shape[1])
    np.testing.assert_array_equal(toks.numpy(), rec["greedy"])


def test_md_hard_oracle():
    rec, w, _ = load_golden("md_hard")
    F = torch.cat((torch.from_numpy(rec["F_rgb"]), torch.from_numpy(rec["F_dep"])), dim=2)
    caps, lengths = torch.from_numpy(rec["captions"]), rec["lengths"].tolist()
    sizes = rec["batch_sizes"].tolist()
    u = split_steps(torch.from_numpy(rec["u_eval"]), sizes)
    ev, _, _ = O.decoder_forward(w, F, None, caps, lengths, attn="gumbel_max", u_steps=u)
    np.testing.assert_allclose(ev.numpy(), rec["eval_logits"], rtol=0, atol=2e-6)
