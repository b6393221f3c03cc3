"""Run beam / greedy decoding a few times (for ncu launch lists and quick timing).
Usage: python scripts/decode_run.py [beam|greedy|hard] [batch] [reps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import depth_image_captioning_pub_b200 as P  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "beam"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, D, A, E, H, V, T = 196, 2048, 128, 128, 128, 10000, 20
dev = torch.device("cuda:0")
w = O.make_weights(A, E, D, H, V, seed=1234)
cls = P.CD_RNNDecoderWithHardAttention if mode == "hard" else P.CD_RNNDecoderWithSoftAttention
m = cls(A, E, D, H, V, *(("cuda:0",) if mode == "hard" else ()))
m.load_state_dict(w)
m.precision = "bf16"
m.cache_packed_weights = True
m = m.to(dev).eval()
g = torch.Generator().manual_seed(1)
F_rgb = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
F_dep = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
voc = O.synthetic_vocab(V)


def run():
    if mode == "beam":
        return m.beam_search(F_rgb, F_dep, voc, beam=5, max_length=T)["tokens"]
    return m.batch_sample(F_rgb, F_dep, voc, max_length=T)


run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    run()
t_cpu = (time.perf_counter() - t0) / reps          # host time to enqueue (no sync inside beam/greedy for beam)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"{mode} B={B}: {dt*1e3:.3f} ms per call ({t_cpu*1e3:.3f} ms host enqueue), {B/dt:.0f} captions/s")

if os.environ.get("DIC_DECODE_PROFILE"):
    from depth_image_captioning_pub_b200 import _lib
    lib = _lib.load()
    lib.dic_profile_enable(1)
    run()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    lib.dic_profile_enable(0)
    tot = sum(v[0] for v in prof.values())
    print(f"per-class CUDA-event time of one call (no PDL while profiling), total {tot:.3f} ms:")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        if v[1]:
            print(f"  {k:20s} {v[0]*1e3:9.1f} us  {v[1]:4d} launches  avg {v[0]*1e3/v[1]:7.1f} us")
