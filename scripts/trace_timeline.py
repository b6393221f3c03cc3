"""Real GPU timeline of the PDL-chained step kernels from the library's in-kernel trace
(dic_trace_start/stop: thread 0 of every CTA stamps %globaltimer at entry / after its dependency
wait / at exit).   python scripts/trace_timeline.py [train|beam|greedy] [--sub S] [--batch B]"""
import argparse, ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="train")
ap.add_argument("--sub", type=int, default=1)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rows", type=int, default=70)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--skip", type=int, default=0, help="launches to skip before printing")
ap.add_argument("--detail", type=int, default=-1, help="kernel id: per-CTA go/exit distribution of one launch")
ap.add_argument("--concurrent", action="store_true", help="kernels of two streams overlap: split launches per kernel id by repeated block index")
args = ap.parse_args()
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
dev = torch.device("cuda", 0)
lib = _lib.load()
B = args.batch
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234))
m.precision = "bf16"
m = m.to(dev).train()
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(B, 1235, torch.bfloat16)
targets = O.pack_targets(caps, lengths).to(dev)
F_rgb, caps = F_rgb.to(dev), caps.to(dev)
F_dep = F_dep.to(dev).requires_grad_(True)
voc = O.synthetic_vocab(V)
lib.dic_set_substreams(args.sub)

def train():
    out, alphas = m(F_rgb, F_dep, caps, lengths)
    loss = torch.nn.functional.cross_entropy(out.data, targets, ignore_index=V - 1)
    loss = loss + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    F_dep.grad = None
    m.zero_grad(set_to_none=True)

if args.what == "train":
    fn = train
else:
    m.eval(); m.cache_packed_weights = True
    Bd = min(B, 128) if args.batch == 256 else B
    fr, fd = F_rgb[:Bd].contiguous(), F_dep[:Bd].detach().contiguous()
    fn = (lambda: m.beam_search(fr, fd, voc, beam=args.beam, max_length=T)) if args.what == "beam" else \
         (lambda: m.batch_sample(fr, fd, voc, max_length=T))
for _ in range(3):
    fn()
torch.cuda.synchronize()
cap = 4_000_000
buf = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
_lib.check(lib.dic_trace_start(buf.data_ptr(), cap))
fn()
cnt = C.c_uint(0)
_lib.check(lib.dic_trace_stop(C.byref(cnt)))
n = min(cnt.value, cap)
rec = np.frombuffer(buf[: n * 32].cpu().numpy().tobytes(),
                    dtype=np.dtype([("t0", "<u8"), ("t1", "<u8"), ("t2", "<u8"), ("kid", "<i4"), ("blk", "<i4")]))
names = {1: "alpha", 2: "ctx", 3: "lstm_fwd", 4: "lstm_bwd", 5: "bwd_stream", 6: "bwd_small", 7: "argmax_embed",
         8: "beam_topk", 9: "beam_merge", 10: "beam_reorder", 0: "gemm", 100: "gemm:hproj", 200: "gemm:gates",
         300: "gemm:dzg", 400: "gemm:dh", 500: "gemm:logits"}
# group into launches: sort by t1 (after-wait), consecutive same kid = one launch
order = np.argsort(rec["t1"], kind="stable")
rec = rec[order]
if args.concurrent:
    # launches of different streams interleave in time; launches of ONE kernel id do not overlap each other and
    # every launch holds each block index once: a repeated index starts the next launch
    rows_ = []
    for kid in np.unique(rec["kid"]):
        r = rec[rec["kid"] == kid]
        seen, start = set(), 0
        for i in range(len(r) + 1):
            if i == len(r) or int(r["blk"][i]) in seen:
                q = r[start:i]
                rows_.append((int(q["t1"].min()), int(kid), len(q), int(q["t0"].min()), int(q["t2"].max())))
                if args.detail == int(kid):
                    dcount = dcount + 1 if "dcount" in globals() else 1
                    if dcount == 8:
                        b0 = int(q["t1"].min())
                        pc = lambda a: " ".join(f"{np.percentile(a, p_):7.2f}" for p_ in (0, 10, 25, 50, 75, 90, 100))
                        print(f"detail kid={kid}: {len(q)} CTAs; percentiles 0/10/25/50/75/90/100 (us rel. first go)")
                        print("  entry:", pc((q["t0"].astype(np.int64) - b0) / 1e3))
                        print("  exit :", pc((q["t2"].astype(np.int64) - b0) / 1e3))
                        print("  entry->exit per CTA:", pc((q["t2"].astype(np.int64) - q["t0"].astype(np.int64)) / 1e3))
                seen, start = set(), i
            if i < len(r):
                seen.add(int(r["blk"][i]))
    rows_.sort()
    T0 = rows_[0][0]
    print(f"{n} records, {len(rows_)} launches, span {(max(x[4] for x in rows_) - T0) / 1e3:.1f} us")
    print(f"{'kernel':14s} {'ctas':>5s} {'entry':>9s} {'go':>9s} {'end':>9s} | {'run':>7s}")
    for k, (t1, kid, nb, t0, t2) in enumerate(rows_):
        if args.skip <= k < args.skip + args.rows:
            print(f"{names.get(kid, str(kid)):14s} {nb:5d} {(t0 - T0) / 1e3:9.1f} {(t1 - T0) / 1e3:9.1f} {(t2 - T0) / 1e3:9.1f} | {(t2 - t1) / 1e3:7.1f}")
    sys.exit(0)
launches = []
i = 0
detail_seen = 0
while i < n:
    j = i
    while j < n and rec["kid"][j] == rec["kid"][i]:
        j += 1
    r = rec[i:j]
    if args.detail >= 0 and int(r["kid"][0]) == args.detail:
        detail_seen += 1
        if detail_seen == 8:      # a launch from the middle of the loop
            base_t = int(r["t1"].min())
            go = np.sort((r["t1"].astype(np.int64) - base_t) / 1e3)
            ex = np.sort((r["t2"].astype(np.int64) - base_t) / 1e3)
            dur = np.sort((r["t2"].astype(np.int64) - r["t1"].astype(np.int64)) / 1e3)
            ent = np.sort((r["t0"].astype(np.int64) - base_t) / 1e3)
            q = lambda a: " ".join(f"{np.percentile(a, p):7.2f}" for p in (0, 10, 50, 90, 100))
            print(f"detail kid={args.detail}: {len(r)} CTAs; percentiles 0/10/50/90/100 (us rel. first go)")
            print("  entry:", q(ent)); print("  go   :", q(go)); print("  exit :", q(ex)); print("  go->exit per CTA:", q(dur))
    launches.append((int(r["kid"][0]), j - i, int(r["t0"].min()), int(r["t1"].min()), int(r["t1"].max()),
                     int(r["t2"].min()), int(r["t2"].max())))
    i = j
T0 = launches[0][2]
print(f"{n} records, {len(launches)} launches, span {(max(l[6] for l in launches) - T0) / 1e3:.1f} us")
print(f"{'kernel':14s} {'ctas':>5s} {'entry':>9s} {'go':>9s} {'end':>9s} | {'run':>7s} {'gap':>6s} {'early':>6s}   (us; run = last exit - first go, gap = first go - prev end, early = go - entry)")
prev_end = None
agg = {}
for k, (kid, nb, t0, t1a, t1b, t2a, t2b) in enumerate(launches):
    run = (t2b - t1a) / 1e3
    gap = (t1a - prev_end) / 1e3 if prev_end is not None else 0.0
    a = agg.setdefault(kid, [0, 0.0, 0.0])
    a[0] += 1; a[1] += run; a[2] += gap
    if args.skip <= k < args.skip + args.rows:
        print(f"{names.get(kid, str(kid)):14s} {nb:5d} {(t0 - T0) / 1e3:9.1f} {(t1a - T0) / 1e3:9.1f} {(t2b - T0) / 1e3:9.1f} | "
              f"{run:7.1f} {gap:6.1f} {(t1a - t0) / 1e3:6.1f}")
    prev_end = t2b
print("\nper kernel: launches, mean run us, mean gap-before us, total us")
for kid, (c, r, g) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print(f"{names.get(kid, str(kid)):14s} {c:4d} {r / c:8.1f} {g / c:8.1f} {r + g:9.1f}")
