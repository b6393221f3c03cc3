"""Back-to-back calls of the library's NVLink all-reduce (dic_dp_allreduce) on a gradient-sized buffer.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_allreduce_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depth_image_captioning_pub_b200.distributed import PeerBuffer  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
n = 4818321          # the decoder's 17 tensors at V = 10k
for mm in ("0", "1"):
    os.environ["DIC_DP_MULTIMEM"] = mm
    pb = PeerBuffer(n, dev)
    if mm == "1" and not pb.multicast:
        if rank == 0:
            print("no multicast mapping on this box")
        continue
    for blocks in (8, 16, 32, 64, 148):
        pb.blocks = blocks
        pb.flat.fill_(float(rank + 1))
        pb.all_reduce(average=False)
        torch.cuda.synchronize()
        ok = bool((pb.flat == world * (world + 1) / 2).all())
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            pb.all_reduce(average=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        if rank == 0:
            print(f"world={world} multimem={mm} blocks={blocks:3d}: {us:7.1f} us per all-reduce of {4 * n / 1e6:.1f} MB  correct={ok}", flush=True)
# raw peer copy for calibration: pull the peer's buffer with a copy kernel / copy engine
pb = PeerBuffer(n, dev)
peer = pb.handle.get_buffer((rank + 1) % world, (pb.n_pad,), torch.float32)
local = torch.empty(pb.n_pad, device=dev)
for _ in range(3):
    local.copy_(peer)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    local.copy_(peer)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"peer -> local copy of {4 * pb.n_pad / 1e6:.1f} MB: {us:.1f} us = {4 * pb.n_pad / us / 1e3:.0f} GB/s")
big = torch.empty(64 << 20, device=dev)   # 256 MB
if rank == 0:
    pass
# NCCL for comparison
t = torch.ones(n, device=dev)
for _ in range(5):
    dist.all_reduce(t, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_reduce(t, op=dist.ReduceOp.AVG)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world={world} ncclAllReduce: {e0.elapsed_time(e1) / 50 * 1e3:7.1f} us")
dist.barrier()
dist.destroy_process_group()
