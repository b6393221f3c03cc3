"""Micro-benchmark of the GEMM shapes of one training / decode step (CUDA events, warm, back to
back launches).  Usage: python scripts/gemm_bench.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depth_image_captioning_pub_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
only = sys.argv[2] if len(sys.argv) > 2 else None

# name, M, N, K, A layout, B layout, splits, ideal bytes (for GB/s)
SHAPES = [
    ("att1       F.Wenc^T", 50176, 128, 2048, "k", "k", 1),
    ("logits     H.Wout^T", 5120, 10000, 128, "k", "k", 1),
    ("hproj      h.Wdb^T", 256, 2176, 128, "k", "k", 1),
    ("gates      x.Wg^T (split 16)", 256, 512, 2304, "k", "k", 16),
    ("dzg        G.Wih", 256, 2048, 512, "k", "m", 2),
    ("dh         G.Whdb", 256, 128, 2688, "k", "m", 10),
    ("dF         datt1.Wenc", 50176, 2048, 128, "k", "m", 1),
    ("dHout      dlog.Wout", 5120, 128, 10000, "k", "m", 7),
    ("dW_out     dlog^T.H", 10000, 128, 5120, "m", "m", 3),
    ("dW_ih      G^T.X", 512, 2176, 5120, "m", "m", 4),
    ("dW_enc     datt1^T.F", 128, 2048, 50176, "m", "m", 18),
    ("dW_beta    G^T.h", 2048, 128, 5120, "m", "m", 18),
    ("beam logits (640 rows)", 640, 10000, 128, "k", "k", 1),
]

print(f"{'gemm':34s} {'M':>6s} {'N':>6s} {'K':>6s}   tc us   fma us   tc TFLOP/s")
for name, M, N, K, la, lb, splits in SHAPES:
    if only and not name.startswith(only):
        continue
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    if la == "k":
        As, a_m, a_k = A, K, 1
    else:
        As, a_m, a_k = A.t().contiguous(), 1, M
    if lb == "k":
        Bs, b_n, b_k = B, K, 1
    else:
        Bs, b_n, b_k = B.t().contiguous(), 1, N
    C = torch.zeros(M, N, device=dev)
    res = []
    for engine in (1, 0):
        if engine == 0 and M * N * K > 3e10:
            res.append(float("nan"))
            continue
        def run():
            _lib.check(lib.dic_gemm_ex(engine, M, N, K, As.data_ptr(), _lib.DIC_BF16, a_m, a_k, Bs.data_ptr(),
                                       _lib.DIC_BF16, b_n, b_k, None, C.data_ptr(), N, splits, _lib.stream_ptr(dev)))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps * 1e3)
    tf = 2.0 * M * N * K / (res[0] * 1e-6) / 1e12
    print(f"{name:34s} {M:6d} {N:6d} {K:6d} {res[0]:8.1f} {res[1]:8.1f} {tf:10.1f}")
