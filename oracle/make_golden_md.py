"""Golden vectors for the concat-fusion (MD_) decoders, from the UNMODIFIED reference modules.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_md.py

Same recipe as oracle/make_golden.py (kept separate so the earlier fixtures stay byte-stable).
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
    MD_RNNDecoderWithSoftAttention, MD_RNNDecoderWithHardAttention)

A, E, D_RGB, D_DEP, H, V, L = 32, 16, 24, 8, 32, 53, 196     # mlp_dim_encoder = 24 + 8 = 32
B = 3
LENGTHS = [6, 5, 3]
MAXLEN = 5
W2I = {"<start>": V - 4, "<end>": V - 3, "<unk>": V - 2, "<null>": V - 1}


def npy(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def inputs(seed):
    g = torch.Generator().manual_seed(seed)
    F_rgb = torch.rand(B, L, D_RGB, generator=g)
    F_dep = torch.rand(B, L, D_DEP, generator=g)
    caps = torch.full((B, max(LENGTHS)), W2I["<null>"], dtype=torch.int64)
    for b, n in enumerate(LENGTHS):
        caps[b, 0] = W2I["<start>"]
        caps[b, 1:n - 1] = torch.randint(0, V - 4, (n - 2,), generator=g)
        caps[b, n - 1] = W2I["<end>"]
    return F_rgb, F_dep, caps


def main():
    os.makedirs(OUT, exist_ok=True)
    # soft
    torch.manual_seed(600)
    m = MD_RNNDecoderWithSoftAttention(A, E, D_RGB + D_DEP, H, V).eval()
    F_rgb, F_dep, caps = inputs(601)
    F_rgb.requires_grad_(True)
    F_dep.requires_grad_(True)
    out, alphas = m(F_rgb, F_dep, caps, LENGTHS)
    dec = [l - 1 for l in LENGTHS]
    tg = pack_padded_sequence(caps[:, 1:], dec, batch_first=True)
    loss = torch.nn.functional.cross_entropy(out.data, tg.data, ignore_index=W2I["<null>"])
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    rec = {"w." + k: v for k, v in m.state_dict().items()}
    rec.update({"g." + k: p.grad for k, p in m.named_parameters()})
    rec.update(F_rgb=F_rgb, F_dep=F_dep, captions=caps, lengths=np.array(LENGTHS), logits=out.data,
               batch_sizes=out.batch_sizes, alphas=alphas, loss=loss, g_F_rgb=F_rgb.grad, g_F_dep=F_dep.grad)
    with torch.no_grad():
        rec["greedy"] = m.batch_sample(F_rgb.detach(), F_dep.detach(), W2I, max_length=MAXLEN)
    np.savez_compressed(os.path.join(OUT, "md_soft.npz"), **npy(rec))
    print("md_soft loss", float(loss.detach()))
    # hard (Gumbel-max eval paths only: cheap to pin, same kernels as the soft training path otherwise)
    torch.manual_seed(700)
    mh = MD_RNNDecoderWithHardAttention(A, E, D_RGB + D_DEP, H, V, "cpu").eval()
    F_rgb, F_dep, caps = inputs(701)
    sizes = [sum(l > t for l in dec) for t in range(max(dec))]
    rec = {"w." + k: v for k, v in mh.state_dict().items()}
    with torch.no_grad():
        torch.manual_seed(703)
        ev = mh.eval_forward(F_rgb, F_dep, caps, LENGTHS)
        torch.manual_seed(703)
        u_eval = torch.cat([torch.rand(n, L) for n in sizes])
        torch.manual_seed(704)
        greedy = mh.batch_sample(F_rgb, F_dep, W2I, max_length=MAXLEN)
    rec.update(F_rgb=F_rgb, F_dep=F_dep, captions=caps, lengths=np.array(LENGTHS), eval_logits=ev.data,
               batch_sizes=ev.batch_sizes, u_eval=u_eval, greedy=greedy)
    np.savez_compressed(os.path.join(OUT, "md_hard.npz"), **npy(rec))
    print("md_hard done")


if __name__ == "__main__":
    main()
