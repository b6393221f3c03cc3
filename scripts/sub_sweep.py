"""Sweep the number of sub-batch streams (dic_set_substreams) for the training step and decode.

    python scripts/sub_sweep.py [--batch 256] [--decode-batch 128]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import depth_image_captioning_pub_b200 as P  # noqa: E402
from depth_image_captioning_pub_b200 import _lib  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402  (weights only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--decode-batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--subs", default="1,2,3,4,6,8")
ap.add_argument("--no-decode", action="store_true")
ap.add_argument("--no-train", action="store_true")
ap.add_argument("--only-decode-batch", action="store_true")
ap.add_argument("--tag", default="")
args = ap.parse_args()
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
dev = torch.device("cuda", 0)
lib = _lib.load()
B = args.batch
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234))
m.precision = "bf16"
m = m.to(dev).train()
opt = torch.optim.AdamW(list(m.parameters()), lr=1e-3, fused=True)
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(B, 1235, torch.bfloat16)
targets = O.pack_targets(caps, lengths).to(dev)
F_rgb, caps = F_rgb.to(dev), caps.to(dev)
F_dep = F_dep.to(dev).requires_grad_(True)


def train_step():
    loss = m.forward_loss(F_rgb, F_dep, caps, lengths, ignore_index=V - 1, lam=bench.LAM)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    F_dep.grad = None
    return loss


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


subs = [int(x) for x in args.subs.split(",")]
for s in ([] if args.no_train else subs):
    lib.dic_set_substreams(s)
    ms = timed(train_step, args.steps)
    print(f"{args.tag}train  B={B} substreams={s}: {ms:.3f} ms/step  {B * T / ms * 1e3:.0f} tokens/s", flush=True)

if args.no_decode:
    lib.dic_set_substreams(0)
    sys.exit(0)
m.eval()
m.cache_packed_weights = True
voc = O.synthetic_vocab(V)
for Bd in ([args.decode_batch] if args.only_decode_batch else sorted({args.decode_batch, 256})):
    fr, fd = F_rgb[:Bd].contiguous(), F_dep[:Bd].detach().contiguous()
    for s in subs:
        lib.dic_set_substreams(s)
        ms = timed(lambda: m.beam_search(fr, fd, voc, beam=5, max_length=T), args.steps)
        msg = timed(lambda: m.batch_sample(fr, fd, voc, max_length=T), args.steps)
        print(f"decode B={Bd} substreams={s}: beam5 {ms:.3f} ms {Bd / ms * 1e3:.0f} cap/s | greedy {msg:.3f} ms "
              f"{Bd / msg * 1e3:.0f} cap/s", flush=True)
lib.dic_set_substreams(0)
