"""Kernel timeline of one data-parallel training step (torch.profiler / CUPTI on rank 0): where the three
bucketed all-reduces run relative to the backward kernels.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_timeline.py [--batch 256]
A profiler perturbs the step time; this is for the ORDER and overlap of kernels, not for bench numbers."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import depth_image_captioning_pub_b200 as P  # noqa: E402
from depth_image_captioning_pub_b200.distributed import FlatGradAllReduce  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402  (weights only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--min-us", type=float, default=15.0)
ap.add_argument("--time-only", action="store_true", help="CUDA-event time of 30 steps instead of the profile")
ap.add_argument("--tag", default="")
args = ap.parse_args()
L, D, A, E, H, V, T = bench.L, bench.D, bench.A, bench.E, bench.H, bench.V, bench.T
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
B = args.batch
m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
m.load_state_dict(O.make_weights(A, E, D, H, V, seed=1234))
m.precision = "bf16"
m = m.to(dev).train()
params = list(m.parameters())
opt = P.FusedAdamW(params, lr=1e-3)
F_rgb, F_dep, caps, lengths = bench.synthetic_batch(B, 1235 + rank, torch.bfloat16)
F_rgb, caps = F_rgb.to(dev), caps.to(dev)
F_dep = F_dep.to(dev).requires_grad_(True)
ar = FlatGradAllReduce(params, module=m) if world > 1 else None


def step():
    loss = m.forward_loss(F_rgb, F_dep, caps, lengths, ignore_index=V - 1, lam=bench.LAM)
    if ar is not None:
        ar.arm()
    loss.backward()
    if ar is not None:
        ar(average=True)
    opt.step()
    opt.zero_grad(set_to_none=True)
    F_dep.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if args.time_only:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{args.tag} world={world} B={B}: {float(ms):.3f} ms/step", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last step = after the last adamw kernel but one
    adam = [i for i, e in enumerate(evs) if "adamw" in e.name.lower()]
    lo = adam[-2] + 1 if len(adam) >= 2 else 0
    evs = evs[lo:adam[-1] + 1]
    t0 = evs[0].time_range.start
    print(f"{len(evs)} kernels in the step, span {evs[-1].time_range.end - t0:.1f} us")
    print("   start_us     dur_us  stream  kernel")
    main_stream = None
    for e in evs:
        dur = e.time_range.end - e.time_range.start
        st = getattr(e, "device_resource_id", -1)
        if main_stream is None:
            main_stream = st
        name = e.name[:70]
        if dur >= args.min_us or st != main_stream or "nccl" in name.lower():
            print(f"{e.time_range.start - t0:10.1f} {dur:10.1f}  {st:6}  {name}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
