"""A few training steps of the bench configuration (B=256, T=20, bf16) and nothing else: the command profiled by
ncu for the launch list and the --set full captures under profiles/.   python scripts/train_run.py [steps] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import depth_image_captioning_pub_b200 as P  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402  (weights only)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
dev = torch.device("cuda", 0)
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234))
m.precision = "bf16"
m = m.to(dev).train()
opt = P.FusedAdamW(list(m.parameters()), lr=1e-3)
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(B, 1235, torch.bfloat16)
F_rgb, caps = F_rgb.to(dev), caps.to(dev)
F_dep = F_dep.to(dev).requires_grad_(True)
for _ in range(steps):
    loss = m.forward_loss(F_rgb, F_dep, caps, lengths, ignore_index=V - 1, lam=bench.LAM)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    F_dep.grad = None
torch.cuda.synchronize()
print("loss", float(loss))
