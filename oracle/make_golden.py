"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference ships no golden vectors, so the pin for the
oracle and for the CUDA path is manufactured here: seeded inputs go through the
reference's own nn.Modules on CPU (fp32) and the inputs, weights and outputs
are stored as small fixtures.  Nothing from the reference's sources is copied;
only numbers it computes.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch.nn.utils.rnn import pack_padded_sequence

REF = os.environ.get("DIC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from Captioning_models.Depth_caption_model.depth_models import (  # noqa: E402
    CD_RNNDecoderWithSoftAttention, CD_RNNDecoderWithHardAttention)
from Captioning_models.Base_caption_model.base_caption_models import (  # noqa: E402
    RNNDecoderWithSoftAttention, RNNDecoderWithHardAttention)

# small dims (the reference takes them as ctor args); L must be 196 because
# Hard_Attention fixes k=196 (attention.py:108,124)
A, E, D, H, V, L = 32, 16, 32, 32, 53, 196
B = 3
LENGTHS = [7, 5, 4]            # incl. <start>, sorted descending (util.py:95)
MAXLEN = 6
W2I = {"<start>": V - 4, "<end>": V - 3, "<unk>": V - 2, "<null>": V - 1}


def make_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    F_rgb = torch.rand(B, L, D, generator=g)
    F_dep = torch.rand(B, L, D, generator=g)
    caps = torch.full((B, max(LENGTHS)), W2I["<null>"], dtype=torch.int64)
    for b, n in enumerate(LENGTHS):
        caps[b, 0] = W2I["<start>"]
        caps[b, 1:n - 1] = torch.randint(0, V - 4, (n - 2,), generator=g)
        caps[b, n - 1] = W2I["<end>"]
    return F_rgb, F_dep, caps


def npy(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def draw_u(seed, sizes):
    """Replays the draws Gumbel_softmax makes on the global CPU generator
    (attention.py:17,40): one torch.rand(bs_valid, 196) per step."""
    torch.manual_seed(seed)
    return [torch.rand(n, L) for n in sizes]


def soft_case(name, depth: bool, seed: int, peak: float = 1.0):
    torch.manual_seed(seed)
    cls = CD_RNNDecoderWithSoftAttention if depth else RNNDecoderWithSoftAttention
    m = cls(A, E, D, H, V)
    with torch.no_grad():
        m.attention.full_att.weight.mul_(peak)
    m.eval()
    F_rgb, F_dep, caps = make_inputs(seed + 1)
    F_rgb.requires_grad_(True)
    F_dep.requires_grad_(True)
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    out, alphas = m(*feats, caps, LENGTHS)
    dec = [l - 1 for l in LENGTHS]
    tg = pack_padded_sequence(caps[:, 1:], dec, batch_first=True)
    loss = torch.nn.functional.cross_entropy(out.data, tg.data, ignore_index=W2I["<null>"])
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    rec = {"w." + k: v for k, v in m.state_dict().items()}
    rec.update({"g." + k: p.grad for k, p in m.named_parameters()})
    rec.update(F_rgb=F_rgb, F_dep=F_dep, captions=caps, lengths=np.array(LENGTHS),
               logits=out.data, batch_sizes=out.batch_sizes, alphas=alphas, loss=loss,
               g_F_rgb=F_rgb.grad, depth=np.array(int(depth)), peak=np.array(peak))
    if depth:
        rec["g_F_dep"] = F_dep.grad
    with torch.no_grad():
        fs = tuple(f.detach() for f in feats)
        rec["greedy"] = m.batch_sample(*fs, W2I, max_length=MAXLEN)
        p1, a1 = m.sample(*(f[:1] for f in fs), W2I, max_length=MAXLEN)
        rec["sample_tokens"] = np.array(p1, dtype=np.int64)
        rec["sample_alphas"] = torch.cat(a1, dim=0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npy(rec))
    print(name, "loss", float(loss.detach()))


def hard_case(name, depth: bool, seed: int, temp: float):
    torch.manual_seed(seed)
    cls = CD_RNNDecoderWithHardAttention if depth else RNNDecoderWithHardAttention
    m = cls(A, E, D, H, V, "cpu")
    m.eval()
    F_rgb, F_dep, caps = make_inputs(seed + 1)
    F_rgb.requires_grad_(True)
    F_dep.requires_grad_(True)
    feats = (F_rgb, F_dep) if depth else (F_rgb,)
    dec = [l - 1 for l in LENGTHS]
    sizes = [sum(l > t for l in dec) for t in range(max(dec))]
    tt = torch.tensor(temp)
    # Gumbel-softmax teacher-forced forward + backward (depth_models.py:580-634)
    torch.manual_seed(seed + 2)
    out = m(*feats, caps, LENGTHS, tt)
    tg = pack_padded_sequence(caps[:, 1:], dec, batch_first=True)
    loss = torch.nn.functional.cross_entropy(out.data, tg.data, ignore_index=W2I["<null>"])
    loss.backward()
    u_fwd = torch.cat(draw_u(seed + 2, sizes), dim=0)
    rec = {"w." + k: v for k, v in m.state_dict().items()}
    rec.update({"g." + k: p.grad for k, p in m.named_parameters()})
    rec.update(F_rgb=F_rgb, F_dep=F_dep, captions=caps, lengths=np.array(LENGTHS),
               logits=out.data, batch_sizes=out.batch_sizes, loss=loss, temp=np.array(temp, dtype=np.float32),
               u_fwd=u_fwd, g_F_rgb=F_rgb.grad, depth=np.array(int(depth)))
    if depth:
        rec["g_F_dep"] = F_dep.grad
    with torch.no_grad():
        fs = tuple(f.detach() for f in feats)
        # Gumbel-max eval_forward (depth_models.py:637-689)
        torch.manual_seed(seed + 3)
        ev = m.eval_forward(*fs, caps, LENGTHS)
        rec["eval_logits"] = ev.data
        rec["u_eval"] = torch.cat(draw_u(seed + 3, sizes), dim=0)
        # Gumbel-max greedy decode (depth_models.py:742-789)
        torch.manual_seed(seed + 4)
        rec["greedy"] = m.batch_sample(*fs, W2I, max_length=MAXLEN)
        rec["u_greedy"] = torch.cat(draw_u(seed + 4, [B] * MAXLEN), dim=0)
        torch.manual_seed(seed + 5)
        p1, a1 = m.sample(*(f[:1] for f in fs), W2I, max_length=MAXLEN)
        rec["sample_tokens"] = np.array(p1, dtype=np.int64)
        rec["sample_alphas"] = torch.cat(a1, dim=0)          # int64 one-hots
        rec["u_sample"] = torch.cat(draw_u(seed + 5, [1] * MAXLEN), dim=0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npy(rec))
    print(name, "loss", float(loss.detach()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    soft_case("depth_soft", True, 100)
    soft_case("base_soft", False, 200)
    soft_case("depth_soft_peaked", True, 300, peak=50.0)
    hard_case("depth_hard", True, 400, temp=0.8)
    hard_case("base_hard", False, 500, temp=1.0)
