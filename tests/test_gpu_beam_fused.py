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
    assert n1 - n0 <= (n2 - n1) - 2 * T, (n1 - n0, n2 - n1)           # two launches fewer per step: the fused path ran
    ft, rt = fused["tokens"].cpu(), ref["tokens"].cpu()
    same = (ft == rt).all(dim=1)
    # the two paths sum the row log-sum-exp in a different order: a last-ulp difference can flip a near tie
    assert same.float().mean() >= 0.98, float(same.float().mean())
    assert torch.equal(fused["lengths"].cpu()[same], ref["lengths"].cpu()[same])
    ds = (fused["scores"].cpu() - ref["scores"].cpu()).abs()
    assert float(ds[same].max()) <= 1e-4 * max(1.0, float(ref["scores"].abs().max())), float(ds[same].max())
    assert ((ft >= 0) & (ft < V)).all()
