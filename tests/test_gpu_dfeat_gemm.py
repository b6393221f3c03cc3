"""Fused dL/dF tensor-core GEMM (csrc/dfeat_tc.cuh) against a float64 restatement of
dF[b] = datt1[b] W_enc + alpha[b]^T dz[b] + dmeanF[b]/L on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

from depth_image_captioning_pub_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,D,A,T", [(3, 50, 64, 32, 8), (5, 196, 256, 128, 20), (2, 196, 2048, 128, 20),
                                       (4, 100, 128, 64, 70), (1, 300, 136, 16, 3)])
def test_dfeat_gemm(B, L, D, A, T, cuda_device):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + L)
    Lp = (L + 7) // 8 * 8
    datt1 = (torch.randn(B * L, A, generator=g) * 0.1).to(torch.bfloat16)
    wenc = (torch.randn(A, D, generator=g) * 0.1).to(torch.bfloat16)
    alpha = torch.softmax(torch.randn(B, T, L, generator=g), dim=2)
    alpha[0, T - 1] = 0                      # an inactive (b, t) row
    a16 = torch.zeros(B, T, Lp, dtype=torch.bfloat16)
    a16[:, :, :L] = alpha.to(torch.bfloat16)
    dz = (torch.randn(T, B, D, generator=g) * 0.1).to(torch.bfloat16)
    dmean = torch.randn(B, D, generator=g)
    ref = (datt1.double().view(B, L, A) @ wenc.double()
           + torch.einsum("btl,tbd->bld", a16[:, :, :L].double(), dz.double())
           + dmean.double()[:, None, :] / L)
    dev = cuda_device
    out = torch.full((B * L, D), float("nan"), dtype=torch.bfloat16, device=dev)
    args = [t.to(dev).contiguous() for t in (datt1, wenc, a16.view(B * T, Lp), dz, dmean)]
    _lib.check(lib.dic_dfeat_gemm(args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), Lp, args[3].data_ptr(),
                                  args[4].data_ptr(), out.data_ptr(), B, L, D, A, T, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    got = out.double().cpu().view(B, L, D)
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-6, err     # bf16 output rounding (2^-9 relative)
