"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, the flat-buffer gradient
all-reduce and the data-parallel loss weighting, checked against the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depth_image_captioning_pub_b200 import distributed as D
from oracle import decoder_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges():
    for n in (1, 7, 128, 1024):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    lengths = [9, 9, 8, 7, 7, 5, 3]
    for world in (2, 3):
        idx = [D.shard_sorted_batch(lengths, r, world) for r in range(world)]
        assert sorted(sum(idx, [])) == list(range(len(lengths)))
        for sub in idx:      # every sub-batch stays sorted descending
            sl = [lengths[i] for i in sub]
            assert sl == sorted(sl, reverse=True)


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    A, E, Dd, H, V, L = 8, 4, 8, 8, 19, 12
    w = {k: v.clone().requires_grad_(True) for k, v in O.make_weights(A, E, Dd, H, V, seed=3).items()}
    g = torch.Generator().manual_seed(4)
    lengths = [7, 6, 6, 5, 4, 3]
    Bn = len(lengths)
    F_rgb, F_dep = torch.rand(Bn, L, Dd, generator=g), torch.rand(Bn, L, Dd, generator=g)
    caps = torch.randint(0, V - 4, (Bn, max(lengths)), generator=g)
    caps[:, 0] = V - 4
    idx = D.shard_sorted_batch(lengths, rank, world)
    ll = [lengths[i] for i in idx]
    logits, _, alphas = O.decoder_forward(w, F_rgb[idx], F_dep[idx], caps[idx], ll, hoist=True)
    tg = O.pack_targets(caps[idx], ll)
    ce = torch.nn.functional.cross_entropy(logits, tg)            # token mean on this rank
    reg = ((1.0 - alphas.sum(dim=1)) ** 2).mean()                 # mean over local B*L
    loss = D.dp_loss_weight(int(tg.numel())) * ce + D.dp_loss_weight(len(idx)) * 0.7 * reg
    loss.backward()
    ar = D.FlatGradAllReduce(list(w.values()))
    ar(average=True)
    toks = torch.full((len(idx), 3), rank, dtype=torch.int64)
    gathered = D.gather_tokens(toks) if len(lengths) % world == 0 else None
    if rank == 0:
        # numpy, not tensors: tensor pickling shares file descriptors with a process that is about to exit
        out.put(({k: v.grad.numpy().copy() for k, v in w.items()},
                 None if gathered is None else gathered.numpy().copy()))
    dist.destroy_process_group()


def test_data_parallel_gradients_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, gathered = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference on the whole batch
    A, E, Dd, H, V, L = 8, 4, 8, 8, 19, 12
    w = {k: v.clone().requires_grad_(True) for k, v in O.make_weights(A, E, Dd, H, V, seed=3).items()}
    g = torch.Generator().manual_seed(4)
    lengths = [7, 6, 6, 5, 4, 3]
    Bn = len(lengths)
    F_rgb, F_dep = torch.rand(Bn, L, Dd, generator=g), torch.rand(Bn, L, Dd, generator=g)
    caps = torch.randint(0, V - 4, (Bn, max(lengths)), generator=g)
    caps[:, 0] = V - 4
    logits, _, alphas = O.decoder_forward(w, F_rgb, F_dep, caps, lengths, hoist=True)
    # NB: the per-rank regulariser means are over [B_r, L]; alphas zero-padding differs between the
    # sharded and the full batch (Tmax per rank), which the mean over (B, L) of (1 - sum_t alpha)^2
    # does not see because padded steps contribute alpha = 0 either way.
    loss = torch.nn.functional.cross_entropy(logits, O.pack_targets(caps, lengths)) \
        + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    for k in w:
        ref = w[k].grad
        assert torch.allclose(torch.from_numpy(grads[k]), ref, rtol=1e-4, atol=1e-6), k
    assert gathered is not None and gathered.shape == (6, 3)
    assert (gathered[:3] == 0).all() and (gathered[3:] == 1).all()
