"""Parity AT THE BENCHMARKED SIZE (BASELINE.json configs[1] and [2]): the CUDA path against the CPU oracle
(fp64, hoisted form; the oracle itself is pinned to the unmodified reference modules by
tests/test_oracle_golden*.py) with B >= 130 images at the reference dims, so that every per-step GEMM of
the time loops has M*N >= 16384 and runs on the tcgen05 engine (`tc_gemm_kernel`) exactly as in bench.py --
the small-batch parity cases in test_gpu_parity.py exercise the FMA engine for those launches.

  * bf16, B=256, T=20, uniform lengths (the bench configuration): logits / alphas of `forward`, loss of
    `forward_loss`, all 17 parameter gradients and dL/dF of the fused step (with an explicit dropout mask)
  * bf16 and fp32, B=144, ragged and uniform, fewer steps (bounded CPU time)
  * beam search 128 images x 5 beams x 20 steps, fp32 mode: tokens / lengths / backpointers identical to
    `oracle.beam_search` wherever the oracle's own candidate scores are not within 1e-4 of a tie
The engine class counters (`dic_profile_read`) assert which GEMM engine actually ran.

Reference lines matched: depth_models.py:153-207 (forward), depth_train.py:210-216 (loss).
"""
import numpy as np
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from depth_image_captioning_pub_b200.engine import batch_sizes_from_lengths
from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu

L, D, A, E, H, V = 196, 2048, 128, 128, 128, 10000
ATT_KEYS = ("attention.encoder_att.weight", "attention.encoder_att.bias",
            "attention.decoder_att.weight", "attention.decoder_att.bias")


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def relfro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def make_batch(B, T, seed, ragged, feat_dtype):
    g = torch.Generator().manual_seed(seed)
    F_rgb = torch.rand(B, L, D, generator=g).to(feat_dtype)
    F_dep = torch.rand(B, L, D, generator=g).to(feat_dtype)
    if ragged:
        # sorted descending; at least 130 captions stay active for the first steps
        lengths = sorted(torch.randint(3, T + 2, (B,), generator=g).tolist(), reverse=True)
        lengths[0] = T + 1
    else:
        lengths = [T + 1] * B
    caps = torch.full((B, T + 1), V - 1, dtype=torch.int64)
    for b, n in enumerate(lengths):
        caps[b, 0] = V - 4
        caps[b, 1:n - 1] = torch.randint(0, V - 4, (n - 2,), generator=g)
        caps[b, n - 1] = V - 3
    return F_rgb, F_dep, caps, lengths


def oracle_step(w, F_rgb, F_dep, caps, lengths, mask_steps):
    """fp64 oracle of forward + loss + backward on the SAME input values (bf16 annotations are exact in fp64)."""
    torch.set_num_threads(max(1, torch.get_num_threads()))
    wo = {k: v.clone().double().requires_grad_(True) for k, v in w.items()}
    Fr = F_rgb.double().requires_grad_(True)
    Fd = F_dep.double().requires_grad_(True)
    lo, bsz, ao = O.decoder_forward(wo, Fr, Fd, caps, lengths, dropout_masks=mask_steps, hoist=True)
    loss = O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao, 0.7)
    loss.backward()
    return lo.detach(), ao.detach(), float(loss.detach()), {k: v.grad for k, v in wo.items()}, Fr.grad, Fd.grad


def engine_counts(fn):
    """Run fn() with the per-class event profile on -> {class: launches}."""
    lib = _lib.load()
    lib.dic_profile_enable(1)
    try:
        fn()
        torch.cuda.synchronize()
        prof = _lib.profile_read()
    finally:
        lib.dic_profile_enable(0)
    return {k: v[1] for k, v in prof.items()}


def run_case(dev, precision, B, T, ragged, seed, fused, report):
    feat_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    F_rgb, F_dep, caps, lengths = make_batch(B, T, seed, ragged, feat_dtype)
    bsz = batch_sizes_from_lengths(lengths)
    assert bsz[0] >= 130
    w = O.make_weights(A, E, D, H, V, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    mask = (torch.rand(sum(bsz), H, generator=g) >= 0.5).float() * 2.0            # nn.Dropout(0.5) (depth_models.py:197)
    steps, o = [], 0
    for n in bsz:
        steps.append(mask[o:o + n].double())
        o += n
    lo, ao, loss_o, g_o, gFr_o, gFd_o = oracle_step(w, F_rgb, F_dep, caps, lengths, steps)

    m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
    m.load_state_dict(w)
    m.precision = precision
    m = m.to(dev).train()
    mask_d = mask.to(dev)
    m._dropout_mask = lambda total, device: mask_d               # the oracle's mask instead of a device draw
    fr = F_rgb.to(dev).requires_grad_(True)
    fd = F_dep.to(dev).requires_grad_(True)
    cp = caps.to(dev)

    # forward: logits + alphas
    out, alphas = m(fr, fd, cp, lengths)
    assert out.batch_sizes.tolist() == bsz
    el = relmax(out.data.detach().float().cpu(), lo)
    ea = float(np.abs(alphas.detach().cpu().numpy() - ao.numpy()).max())
    if fused:
        loss = m.forward_loss(fr, fd, cp, lengths, ignore_index=V - 1, lam=0.7)
    else:
        tg = O.pack_targets(caps, lengths).to(dev)
        loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1)
        loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    eloss = abs(float(loss.detach()) - loss_o) / max(1.0, abs(loss_o))

    if precision == "fp32":
        # The attention-projection gradients are sums over B*L rows whose mean-annotation part cancels exactly
        # (sum_l de_l = 0 per caption-step): fp32 accumulation noise grows with the row count, and at this batch
        # the CPU fp32 autograd of the reference itself is 2.5e-3 away from fp64 (scripts/grad_noise.py).
        ltol, atol, losstol, gtol, gtol_att, ftol_att = 1e-4, 1e-5, 1e-5, 1e-4, 5e-3, 5e-3
    else:
        # bf16 storage: the spec bounds the logits (2e-2).  The attention-projection gradients pass through
        # softmax-backward and a ReLU mask whose pre-activations are perturbed ~1e-3 by ANY bf16 quantity
        # upstream (annotations, W_enc, att1, h): measured bounds, in max-abs and in Frobenius norm.
        ltol, atol, losstol, gtol, gtol_att, ftol_att = 2e-2, 2e-3, 2e-3, 5e-2, 0.25, 0.12
    bad = []
    if el > ltol: bad.append(("logits", el, ltol))
    if ea > atol: bad.append(("alphas", ea, atol))
    if eloss > losstol: bad.append(("loss", eloss, losstol))
    rows = []
    for k, p in m.named_parameters():
        ref = g_o[k].numpy()
        got = p.grad.double().cpu().numpy()
        assert np.isfinite(got).all(), k
        if k == "attention.full_att.bias":          # exactly zero by softmax shift invariance
            e = float(np.abs(got).max())
            rows.append((k, e, None))
            if e > 1e-5 * max(1.0, float(np.abs(g_o["attention.full_att.weight"].numpy()).max())): bad.append((k, e, "abs"))
            continue
        em, ef = relmax(got, ref), relfro(got, ref)
        rows.append((k, em, ef))
        att = k in ATT_KEYS
        if em > (gtol_att if att else gtol): bad.append((k, "max", em))
        if att and ef > ftol_att: bad.append((k, "fro", ef))
    for name, got, ref in (("dF_rgb", fr.grad, gFr_o), ("dF_depth", fd.grad, gFd_o)):
        em = relmax(got.double().cpu().numpy(), ref.numpy())
        ef = relfro(got.double().cpu().numpy(), ref.numpy())
        rows.append((name, em, ef))
        # Frobenius bound at the gradient tolerance; the max-norm gets slack for isolated ReLU-mask flips: with
        # ~2e7 pre-activations per case a few sit within one fp32 ulp of zero, and one flipped mask element moves
        # its annotation row of dL/dF by ~1e-3 of the tensor's maximum (measured: 1.2e-3 at one row, fp32 mode)
        if ef > gtol: bad.append((name, "fro", ef))
        if em > max(gtol, 5e-3): bad.append((name, "max", em))
    report.append(f"{precision} B={B} T={T} ragged={ragged} fused={fused}: logits {el:.2e} alphas {ea:.2e} loss {eloss:.2e} | "
                  + " ".join(f"{k.replace('attention.', 'att.').replace('.weight', '.w').replace('.bias', '.b')}:{a:.1e}"
                             + (f"/{b:.1e}" if b is not None else "") for k, a, b in rows))
    print(report[-1])
    assert not bad, bad
    return m, (fr, fd, cp, lengths)


def test_bench_config_bf16_fused_step_vs_oracle(cuda_device):
    """B=256, T=20, bf16, forward_loss + backward: the configuration bench.py times."""
    report = []
    m, (fr, fd, cp, lengths) = run_case(cuda_device, "bf16", 256, 20, False, 101, True, report)

    # which engines ran: every GEMM of the two time loops on tcgen05, none on the FMA engine
    def step():
        m.zero_grad(set_to_none=True)
        fd.grad = None
        m.forward_loss(fr, fd, cp, lengths, ignore_index=V - 1, lam=0.7).backward()
    c = engine_counts(step)
    T = 20
    assert c.get("gemm_tcgen05", 0) >= 3 * T + 10, c          # gates, dzg, dh per step + out-of-loop GEMMs
    assert c.get("attn_alpha_fwd", 0) == T, c                 # fused head kernel (h-projection + energies + softmax)
    assert c.get("gemm_fma", 0) <= 4, c                       # (tiny out-of-loop products only)
    print("engine launches per step:", c)
    assert c.get("attn_context_fwd", 0) == T and c.get("attn_stream_bwd", 0) == T, c


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("ragged", [False, True])
def test_large_batch_vs_oracle(precision, ragged, cuda_device):
    """B=144 at the reference dims, ragged and uniform, unfused loss (the reference loop's own expression)."""
    report = []
    m, (fr, fd, cp, lengths) = run_case(cuda_device, precision, 144, 6, ragged, 202 + int(ragged), False, report)
    if precision == "bf16":
        c = engine_counts(lambda: m(fr, fd, cp, lengths))
        bsz = batch_sizes_from_lengths(lengths)
        big = sum(1 for n in bsz if n >= 128)
        assert c.get("gemm_tcgen05", 0) >= big, (c, bsz)         # gate GEMM of every step with >= 128 rows


def test_beam_128x5_vs_oracle(cuda_device):
    """128 images x 5 beams x 20 steps (BASELINE.json configs[2] per GPU), fp32 mode."""
    dev = cuda_device
    B, K, T = 128, 5, 20
    F_rgb, F_dep, _, _ = make_batch(B, T, 303, False, torch.float32)
    w = O.make_weights(A, E, D, H, V, seed=304)
    # a random-init vocabulary projection gives near-uniform word distributions: cumulative scores near -180 then
    # tie within a few fp32 ulps (7.6e-6) for 2/3 of the images.  A 40x sharper projection keeps 118 of the 128
    # images clear of near ties over all 20 steps (measured with the oracle alone).
    w["linear.weight"] = w["linear.weight"] * 40.0
    # record, per image, the smallest gap between neighbouring candidates among the oracle's top K+1 at any
    # step: where it is < 1e-4 an fp32 implementation may legitimately order two hypotheses the other way
    gaps = torch.full((B,), float("inf"))
    orig = O.beam_select

    def recording_select(scores, finished, logits, lse, end_id):
        cand = scores.unsqueeze(2) + (logits - lse.unsqueeze(2))
        frozen = torch.full_like(cand, float("-inf"))
        frozen[:, :, end_id] = scores
        cand = torch.where(finished.unsqueeze(2), frozen, cand)
        top = cand.reshape(B, -1).topk(K + 1, dim=1).values
        d = (top[:, :-1] - top[:, 1:])
        d = torch.where(torch.isfinite(d), d, torch.full_like(d, float("inf")))
        gaps.copy_(torch.minimum(gaps, d.min(dim=1).values))
        return orig(scores, finished, logits, lse, end_id)
    O.beam_select = recording_select
    try:
        ref = O.beam_search(w, F_rgb, F_dep, V - 4, V - 3, K, T)
    finally:
        O.beam_select = orig
    m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
    m.load_state_dict(w)
    m.precision = "fp32"
    m = m.to(dev).eval()
    got = m.beam_search(F_rgb.to(dev), F_dep.to(dev), O.synthetic_vocab(V), beam=K, max_length=T, trace=True)
    clear = gaps > 1e-4
    assert int(clear.sum()) >= int(0.85 * B), f"only {int(clear.sum())} of {B} images free of near ties"
    tok_eq = (got["tokens"].cpu() == ref["tokens"]).all(dim=1)
    back_eq = (got["back"].cpu() == ref["back"]).all(dim=2).all(dim=0)
    len_eq = got["lengths"].cpu().to(torch.int64) == ref["lengths"]
    assert bool(tok_eq[clear].all()) and bool(back_eq[clear].all()) and bool(len_eq[clear].all()), \
        (int((~tok_eq & clear).sum()), int((~back_eq & clear).sum()))
    assert float((got["scores"].cpu() - ref["scores"])[clear].abs().max()) <= 1e-4
    print(f"beam 128x5x20: {int(clear.sum())} images without near ties, all identical; "
          f"{int((~tok_eq).sum())} of the {int((~clear).sum())} near-tie images differ")
