import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
import depth_image_captioning_pub_b200 as P
from oracle import decoder_oracle as O
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
dev = torch.device("cuda", 0)
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V); m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234)); m.precision = "bf16"
m = m.to(dev).eval(); m.cache_packed_weights = True
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(128, 1235, torch.bfloat16)
fr, fd = F_rgb.to(dev), F_dep.to(dev)
voc = O.synthetic_vocab(V)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m.batch_sample(fr, fd, voc, max_length=T)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"greedy rep {rep}: device {e0.elapsed_time(e1)/20:.3f} ms/call, wall {(t1-t0)/20*1e3:.3f} ms/call", flush=True)
    if rep == 2:
        for _ in range(20): m.beam_search(fr, fd, voc, beam=5, max_length=T)
