"""tcgen05 / FMA GEMM engines with every operand layout the decoder path uses:
K-major ("TN"), MN-major B ("NN": activations x row-major weights), MN-major A and B
("TN over rows": weight gradients), strided sub-views and atomic split-K."""
import numpy as np
import pytest
import torch

from depth_image_captioning_pub_b200 import _lib

pytestmark = pytest.mark.gpu


def run(engine, A_store, a_m, a_k, B_store, b_n, b_k, M, N, K, bias, ldc, splits, dev):
    lib = _lib.load()
    C = torch.zeros(M, ldc, device=dev)
    code = lambda t: _lib.DIC_BF16 if t.dtype == torch.bfloat16 else _lib.DIC_F32
    _lib.check(lib.dic_gemm_ex(engine, M, N, K, A_store.data_ptr(), code(A_store), a_m, a_k, B_store.data_ptr(),
                               code(B_store), b_n, b_k, None if bias is None else bias.data_ptr(), C.data_ptr(), ldc,
                               splits, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    return C[:, :N]


LAYOUTS = ["kk", "km", "mk", "mm"]     # A layout, B layout: k = K-major, m = MN-major
SHAPES = [(128, 128, 64), (256, 2048, 512), (512, 2304, 5120), (128, 2048, 1568), (200, 136, 100),
          (10000, 128, 1024), (1024, 128, 10000)]


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts(engine, layout, M, N, K, cuda_device):
    if engine == 1 and layout[0] == "m" and M % 8:
        pytest.skip("MN-major needs 16-byte row strides")
    dev = cuda_device
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.double() @ B.double().t() + bias.double().cpu()
    pad = 8      # exercise ld != extent
    Kp = (K + 7) // 8 * 8 + pad     # TMA needs 16-byte global row strides
    if layout[0] == "k":
        As = torch.zeros(M, Kp, dtype=torch.bfloat16); As[:, :K] = A
        a_m, a_k = Kp, 1
    else:
        As = torch.zeros(K, M + pad, dtype=torch.bfloat16); As[:, :M] = A.t()
        a_m, a_k = 1, M + pad
    if layout[1] == "k":
        Bs = torch.zeros(N, Kp, dtype=torch.bfloat16); Bs[:, :K] = B
        b_n, b_k = Kp, 1
    else:
        Bs = torch.zeros(K, N + pad, dtype=torch.bfloat16); Bs[:, :N] = B.t()
        b_n, b_k = 1, N + pad
    ldc = N + 4
    out = run(engine, As.to(dev), a_m, a_k, Bs.to(dev), b_n, b_k, M, N, K, bias, ldc, 1, dev)
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5 * max(1.0, (K / 1000.0) ** 0.5), err


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("layout", ["kk", "mm"])
def test_gemm_split_k_atomic(engine, layout, cuda_device):
    dev = cuda_device
    M, N, K = 512, 256, 4096
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = A.double() @ B.double().t()
    if layout == "kk":
        out = run(engine, A.to(dev), K, 1, B.to(dev), K, 1, M, N, K, None, N, 8, dev)
    else:
        out = run(engine, A.t().contiguous().to(dev), 1, M, B.t().contiguous().to(dev), 1, N, M, N, K, None, N, 8, dev)
    err = float((out.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err


# bf16 C: the tcgen05 engine's shared-memory + bulk-tensor-store epilogue (ldc % 8 == 0, N tile 128) against the FMA
# engine and a float64 product.  Shapes cover rows / columns that end inside a 32 x 64 store box, a padded row
# stride (the canary columns behind N must stay untouched) and the logits shape of the benchmark (5120 x 10000).
@pytest.mark.parametrize("M,N,K,ldc,with_bias", [
    (128, 128, 64, 128, True), (200, 136, 128, 136, True), (333, 1000, 192, 1008, False),
    (5120, 10000, 128, 10000, True), (1568, 128, 2048, 128, True), (130, 264, 64, 272, True)])
def test_gemm_bf16_out_tma_store(M, N, K, ldc, with_bias, cuda_device):
    dev = cuda_device
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev) if with_bias else None
    ref = A.double() @ B.double().t()
    if with_bias:
        ref = ref + bias.double()
    outs = []
    for engine in (1, 0):
        C = torch.full((M, ldc), 7.0, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.dic_gemm_nt_bf16(engine, M, N, K, A.data_ptr(), B.data_ptr(),
                                        None if bias is None else bias.data_ptr(), C.data_ptr(), ldc,
                                        _lib.stream_ptr(dev)))
        torch.cuda.synchronize()
        assert bool((C[:, N:] == 7.0).all()), "columns behind N were written"
        outs.append(C[:, :N].double())
        # one bf16 rounding of an fp32-accumulated product
        err = float((outs[-1] - ref).abs().max() / ref.abs().max())
        assert err <= 2.0 ** -8, (engine, err)
    # both engines accumulate in fp32 and round once: they may differ by one bf16 ulp where the sums differ in
    # the last fp32 bits, never by more
    d = (outs[0] - outs[1]).abs()
    assert bool((d <= 2.0 ** -7 * (outs[1].abs() + 0.02 * ref.abs().max())).all())
