"""Host-side enqueue time vs device time of the library calls (is the path launch-bound on the CPU?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
dev = torch.device("cuda", 0)
lib = _lib.load()
B = 256
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234))
m.precision = "bf16"
m = m.to(dev).train()
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(B, 1235, torch.bfloat16)
targets = O.pack_targets(caps, lengths).to(dev)
F_rgb, caps = F_rgb.to(dev), caps.to(dev)
F_dep = F_dep.to(dev).requires_grad_(True)

def host_dev(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    th = td = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        th += t1 - t0; td += e0.elapsed_time(e1) * 1e-3
    return th / n * 1e3, td / n * 1e3

opt = torch.optim.AdamW(list(m.parameters()), lr=1e-3, fused=True)
for s in (1,):
    lib.dic_set_substreams(s)
    def fwd():
        out, alphas = m(F_rgb, F_dep, caps, lengths)
        return out, alphas
    print(f"S={s} train forward: host %.3f ms, device %.3f ms" % host_dev(fwd))
    def fb():
        loss = m.forward_loss(F_rgb, F_dep, caps, lengths, ignore_index=V - 1, lam=0.7)
        loss.backward()
        F_dep.grad = None
        m.zero_grad(set_to_none=True)
    print(f"S={s} train fwd+loss+bwd: host %.3f ms, device %.3f ms" % host_dev(fb))
    def full():
        loss = m.forward_loss(F_rgb, F_dep, caps, lengths, ignore_index=V - 1, lam=0.7)
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True); F_dep.grad = None
    print(f"S={s} full step: host %.3f ms, device %.3f ms" % host_dev(full))
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): full()
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    m.eval(); m.cache_packed_weights = True
    voc = O.synthetic_vocab(V)
    fr, fd = F_rgb[:128].contiguous(), F_dep[:128].detach().contiguous()
    print(f"S={s} beam5 B=128: host %.3f ms, device %.3f ms" % host_dev(lambda: m.beam_search(fr, fd, voc, beam=5, max_length=T)))
    print(f"S={s} greedy B=128: host %.3f ms, device %.3f ms" % host_dev(lambda: m.batch_sample(fr, fd, voc, max_length=T)))
    m.train(); m.cache_packed_weights = False
