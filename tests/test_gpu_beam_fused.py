"""Beam search, bf16 mode: the fused step (vocabulary projection + log-sum-exp + per-row top-K in one tcgen05 kernel,
merge + reorder in one more; csrc/beam_fused.cuh) against the unfused kernels of the same library, which the
fp32-mode tests pin bit-exactly to oracle.beam_select / oracle.beam_search (tests/test_gpu_parity.py,
tests/test_gpu_bench_size.py).  A call that asks for the lse / backpointer traces takes the unfused path."""
import pytest
import torch

import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O

pytestmark = pytest.mark.gpu

L, D, A, E, H = 196, 2048, 128, 128, 128


def _module(V, dev, seed=1234):
    m = P.CD_RNNDecoderWithSoftAttention(A, E, D, H, V)
    m.load_state_dict(O.make_weights(A, E, D, H, V, seed=seed))
    m.precision = "bf16"
    return m.to(dev).eval()


@pytest.mark.parametrize("B,K,V,T", [(128, 5, 10000, 20), (5, 3, 1000, 7), (130, 5, 2000, 6), (2, 5, 10000, 4)])
def test_fused_beam_step_equals_unfused(B, K, V, T, cuda_device):
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator().manual_seed(77 + B)
    F_rgb = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    F_dep = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    m = _module(V, dev)
    voc = O.synthetic_vocab(V)
    m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T)            # warm-up (packs weights, allocates)
    n0 = lib.dic_launch_count()
    fused = m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T)
    n1 = lib.dic_launch_count()
    ref = m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T, trace=True)
    n2 = lib.dic_launch_count()
    torch.cuda.synchronize()
    # the fused step saves two launches per step; the look-ahead order adds the context gather back
    assert n1 - n0 <= (n2 - n1) - T, (n1 - n0, n2 - n1)
    ft, rt = fused["tokens"].cpu(), ref["tokens"].cpu()
    same = (ft == rt).all(dim=1)
    # The two paths differ in summation order (row log-sum-exp; in the look-ahead order also the token's share of the
    # gates, added outside the gate GEMM).  In bf16 storage a last-bit difference can flip the rounding of an h element
    # (2^-9 relative), and with random weights the vocabulary is full of near ties: a few rows take another, equally
    # good, branch.  So: most rows identical, identical rows agree to 1e-4, and EVERY row's final score within 1 % (a
    # wrong parent / token / context would cost whole units of log-probability per step).
    assert same.float().mean() >= 0.90, float(same.float().mean())
    assert torch.equal(fused["lengths"].cpu()[same], ref["lengths"].cpu()[same])
    ds = (fused["scores"].cpu() - ref["scores"].cpu()).abs()
    smax = max(1.0, float(ref["scores"].abs().max()))
    assert float(ds[same].max()) <= 1e-4 * smax, float(ds[same].max())
    assert float(ds.max()) <= 1e-2 * smax, float(ds.max())
    assert ((ft >= 0) & (ft < V)).all()


@pytest.mark.parametrize("B,K,V,T", [(128, 5, 10000, 20), (64, 3, 1000, 9), (130, 5, 2000, 6), (66, 5, 1000, 1), (7, 5, 1000, 9)])
def test_lookahead_attention_equals_serial_order(B, K, V, T, cuda_device, monkeypatch):
    """Look-ahead attention (dic_api.cu decode_impl): the head + context kernels of step t+1 run on the un-reordered
    h_t, the context pass concurrently with the selection of step t; nothing is reordered, the gate GEMM runs in parent
    order and the LSTM kernel follows the backpointers (DIC_BEAM_LOOKAHEAD is read at every call)."""
    dev = cuda_device
    g = torch.Generator().manual_seed(911 + B)
    F_rgb = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    F_dep = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    m = _module(V, dev, seed=4321)
    voc = O.synthetic_vocab(V)
    lib = _lib.load()
    monkeypatch.setenv("DIC_BEAM_LOOKAHEAD", "0")
    m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T)
    n0 = lib.dic_launch_count()
    serial = m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T)
    n1 = lib.dic_launch_count()
    monkeypatch.setenv("DIC_BEAM_LOOKAHEAD", "1")
    # The initial-state GEMM accumulates its K splits with red.add (the last bits of h0 / c0 change from run to run in
    # EITHER order) and the look-ahead order adds the token's share of the gates outside the gate GEMM: same criteria
    # as test_fused_beam_step_equals_unfused.
    for rep in range(3):          # repeated: a missing dependency would show as run-to-run garbage, not as last bits
        look = m.beam_search(F_rgb, F_dep, voc, beam=K, max_length=T)
        torch.cuda.synchronize()
        same = (look["tokens"] == serial["tokens"]).all(dim=1)
        assert float(same.float().mean()) >= 0.90, float(same.float().mean())
        assert torch.equal(look["lengths"][same], serial["lengths"][same])
        ds = (look["scores"] - serial["scores"]).abs()
        smax = max(1.0, float(serial["scores"].abs().max()))
        assert float(ds[same].max()) <= 1e-4 * smax, float(ds[same].max())
        assert float(ds.max()) <= 1e-2 * smax, float(ds.max())           # rows that took another near-tie branch
    n2 = lib.dic_launch_count()
    assert (n2 - n1) == 3 * ((n1 - n0) + 1), (n1 - n0, n2 - n1)      # the token-table GEMM once per call: it ran


@pytest.mark.parametrize("B,V,T", [(128, 10000, 20), (130, 2000, 6), (5, 1000, 9)])
def test_greedy_lookahead_equals_serial_order(B, V, T, cuda_device, monkeypatch):
    """Greedy decoding with the same look-ahead order (head of step t+1 after the logits GEMM of step t, context pass
    next to the arg-max + embedding kernel): tokens as in the serial order."""
    dev = cuda_device
    g = torch.Generator().manual_seed(1913 + B)
    F_rgb = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    F_dep = torch.rand(B, L, D, generator=g).to(torch.bfloat16).to(dev)
    m = _module(V, dev, seed=99)
    voc = O.synthetic_vocab(V)
    monkeypatch.setenv("DIC_BEAM_LOOKAHEAD", "0")
    serial = m.batch_sample(F_rgb, F_dep, voc, max_length=T)
    monkeypatch.setenv("DIC_BEAM_LOOKAHEAD", "1")
    for rep in range(3):
        look = m.batch_sample(F_rgb, F_dep, voc, max_length=T)
        torch.cuda.synchronize()
        same = (look == serial).all(axis=1)          # numpy [B, T]
        assert float(same.mean()) >= 0.90, float(same.mean())
