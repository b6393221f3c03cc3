"""CPU study (fp64 oracle): which bf16 rounding upstream of the attention ReLU drives the error of the
encoder_att / decoder_att gradients in bf16 mode.  A straight-through perturbation is added to the ReLU
pre-activation att1 + att2 for each candidate source (bf16 storage of att1, bf16 W_enc, bf16 rounding of the
RGB+depth sum, bf16 h / W_dec in att2); the gradients are compared with the unperturbed fp64 ones
(max-abs / Frobenius, relative).  Output of a run: profiles/r02_mask_flip_study.txt.

    python scripts/mask_flip_study.py [B]
"""
import sys, itertools
import numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import decoder_oracle as O
torch.set_num_threads(8)
L, D, A, E, H, V = 196, 2048, 128, 128, 128, 2000
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 6
g = torch.Generator().manual_seed(5)
F_rgb = torch.rand(B, L, D, generator=g).bfloat16().double()
F_dep = torch.rand(B, L, D, generator=g).bfloat16().double()
lengths = [T + 1] * B
caps = torch.randint(0, V - 4, (B, T + 1), generator=g); caps[:, 0] = V - 4; caps[:, T] = V - 3
w0 = O.make_weights(A, E, D, H, V, seed=6)
bf = lambda x: x.float().bfloat16().double()
def run(modes):
    w = {k: v.clone().double().requires_grad_(True) for k, v in w0.items()}
    Fsum = (F_rgb + F_dep)
    Wenc = w["attention.encoder_att.weight"]
    pert = torch.zeros(B, L, A, dtype=torch.float64)
    with torch.no_grad():
        att1_exact = Fsum @ Wenc.t() + w["attention.encoder_att.bias"]
        if 'wenc' in modes: pert += Fsum @ (bf(Wenc) - Wenc).t()
        if 'fsum' in modes: pert += (bf(Fsum) - Fsum) @ Wenc.t()
        if 'att1' in modes: pert += bf(att1_exact + pert) - (att1_exact + pert)
        if 'att1_split' in modes:   # hi+lo bf16 storage
            x = att1_exact + pert; hi = bf(x); lo = bf(x - hi); pert += (hi + lo) - x
    orig = O.attention_energy
    def energy(w_, feats, h, att1=None):
        n = feats.shape[0]
        a1 = feats @ w_["attention.encoder_att.weight"].t() + w_["attention.encoder_att.bias"] + pert[:n]
        hh = h
        Wd = w_["attention.decoder_att.weight"]
        att2 = hh @ Wd.t() + w_["attention.decoder_att.bias"]
        if 'att2' in modes:
            with torch.no_grad():
                d2 = bf(h) @ bf(Wd).t() - h @ Wd.t()
            att2 = att2 + d2
        s = torch.relu(a1 + att2.unsqueeze(1))
        return (s @ w_["attention.full_att.weight"].t() + w_["attention.full_att.bias"]).squeeze(2)
    O.attention_energy = energy
    try:
        lo, bsz, ao = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, hoist=False)
        loss = O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao, 0.7)
        loss.backward()
    finally:
        O.attention_energy = orig
    return {k: v.grad.numpy() for k, v in w.items()}, float(pert.abs().mean())
ref, _ = run(())
keys = ["attention.encoder_att.weight", "attention.decoder_att.weight", "attention.decoder_att.bias", "attention.full_att.weight", "f_beta.weight"]
for modes in [('att1',), ('wenc',), ('fsum',), ('att2',), ('att1','wenc','fsum','att2'), ('wenc','fsum','att2'), ('att2',), ('att1_split','att2'), ('att1','att2')]:
    gr, pm = run(modes)
    s = " ".join(f"{k.replace('attention.','')[:14]}:{np.abs(gr[k]-ref[k]).max()/np.abs(ref[k]).max():.1e}/{np.linalg.norm(gr[k]-ref[k])/np.linalg.norm(ref[k]):.1e}" for k in keys)
    print(f"{'+'.join(modes):28s} |pert| {pm:.1e}  {s}", flush=True)
