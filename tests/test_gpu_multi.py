"""Two-GPU data-parallel step (skipped on single-GPU boxes): the in-place all-reduce of the flat gradient buffer
(distributed.FlatGradAllReduce(module=...)) -- the library's own NVLink peer-memory kernel (dic_dp_allreduce), the same
kernel overlapped with the dL/dF GEMM, and the bucketed NCCL path -- leaves every parameter gradient equal to the mean
of the two ranks' own gradients AND to the gradient of the whole batch computed on one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, mode):
    import torch.distributed as dist
    import depth_image_captioning_pub_b200 as P
    from depth_image_captioning_pub_b200 import distributed as D
    from oracle import decoder_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    A, E, Dd, H, V, L, B, T = 128, 128, 2048, 128, 1000, 196, 16, 5
    m = P.CD_RNNDecoderWithSoftAttention(A, E, Dd, H, V)
    m.load_state_dict(O.make_weights(A, E, Dd, H, V, seed=3))
    m.precision = "bf16"
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(10)                   # the whole batch, the same on every rank ...
    Fr_all, Fd_all = torch.rand(world * B, L, Dd, generator=g), torch.rand(world * B, L, Dd, generator=g)
    caps_all = torch.randint(0, V - 4, (world * B, T + 1), generator=g)
    caps_all[:, 0] = V - 4
    mine = D.shard_sorted_batch([T + 1] * (world * B), rank, world)     # ... and this rank's round-robin shard of it
    Fr, Fd, caps = Fr_all[mine].to(dev), Fd_all[mine].to(dev), caps_all[mine].to(dev)
    params = list(m.parameters())
    # the whole batch on one GPU, before the module is switched to the flat gradient buffer
    m.forward_loss(Fr_all.to(dev), Fd_all.to(dev), caps_all.to(dev), [T + 1] * (world * B), ignore_index=V - 1).backward()
    whole = torch.cat([p.grad.flatten() for p in params]).clone()
    m.zero_grad(set_to_none=True)
    ar = D.FlatGradAllReduce(params, module=m, mode=mode)
    # reference: own gradients without the reduction, gathered from both ranks
    m.forward_loss(Fr, Fd, caps, [T + 1] * B, ignore_index=V - 1).backward()
    own = torch.cat([p.grad.flatten() for p in params]).clone()
    both = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(both, own)
    mean = sum(both) / world
    m.zero_grad(set_to_none=True)
    # overlapped in-place path
    loss = m.forward_loss(Fr, Fd, caps, [T + 1] * B, ignore_index=V - 1)
    ar.arm()
    loss.backward()
    ar(average=True)
    torch.cuda.synchronize()
    got = torch.cat([p.grad.flatten() for p in params])
    err = float((got - mean).abs().max() / mean.abs().max())
    err_whole = float((got - whole).abs().max() / whole.abs().max())
    # a second step through the same buffers (epoch 2 of the flag protocol), same data: same result
    m.zero_grad(set_to_none=True)
    loss = m.forward_loss(Fr, Fd, caps, [T + 1] * B, ignore_index=V - 1)
    ar.arm()
    loss.backward()
    ar(average=True)
    torch.cuda.synchronize()
    got2 = torch.cat([p.grad.flatten() for p in params])
    err = max(err, float((got2 - mean).abs().max() / mean.abs().max()))
    # every rank holds the same reduced values
    allg = [torch.empty_like(got2) for _ in range(world)]
    dist.all_gather(allg, got2.contiguous())
    same = all(torch.equal(allg[0], a) for a in allg[1:])
    eng = next(iter(m._engines.values()))
    aliased = all(eng.grad_flat.data_ptr() <= p.grad.data_ptr() < eng.grad_flat.data_ptr() + eng.grad_flat.numel() * 4
                  for p in params)
    if rank == 0:
        out.put((err, err_whole, aliased, same, ar.mode))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["p2p", "p2p_overlap", "nccl"])
def test_two_gpu_overlapped_allreduce(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    err, err_whole, aliased, same, used = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert used == mode, f"requested {mode}, ran {used} (peer-memory set-up failed?)"
    assert aliased and same
    assert err <= 5e-3, err        # atomic split-K weight gradients: equal up to fp32 summation order
    # data parallel == one GPU on the whole batch (the per-rank att1 / GEMM tiles differ: bf16 storage noise)
    assert err_whole <= 2e-2, err_whole
