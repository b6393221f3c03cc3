"""Two-GPU data-parallel step through the real NCCL path (skipped on single-GPU boxes): the in-place,
overlapped all-reduce of the flat gradient buffer (distributed.FlatGradAllReduce(module=...).arm()) leaves every
parameter gradient equal to the mean of the two ranks' own gradients."""
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


def _worker(rank, world, port, out):
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
    g = torch.Generator().manual_seed(10 + rank)            # different data per rank
    Fr, Fd = torch.rand(B, L, Dd, generator=g).to(dev), torch.rand(B, L, Dd, generator=g).to(dev)
    caps = torch.randint(0, V - 4, (B, T + 1), generator=g).to(dev)
    caps[:, 0] = V - 4
    params = list(m.parameters())
    ar = D.FlatGradAllReduce(params, module=m)
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
    eng = next(iter(m._engines.values()))
    aliased = all(eng.grad_flat.data_ptr() <= p.grad.data_ptr() < eng.grad_flat.data_ptr() + eng.grad_flat.numel() * 4
                  for p in params)
    if rank == 0:
        out.put((err, aliased))
    dist.destroy_process_group()


def test_two_gpu_overlapped_allreduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, aliased = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert aliased
    assert err <= 2e-3, err        # atomic split-K weight gradients: equal up to fp32 summation order
