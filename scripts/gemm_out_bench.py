"""bf16-output GEMMs of the training step, timed alone (CUDA events, back to back): the logits GEMM
[5120 x 128].[128 x 10000] and the att1 GEMM [50176 x 2048].[2048 x 128].  DIC_TMA_STORE=0 selects the
per-thread store epilogue.   python scripts/gemm_out_bench.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depth_image_captioning_pub_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K in (("logits", 5120, 10000, 128), ("att1", 50176, 128, 2048)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)

    def run():
        _lib.check(lib.dic_gemm_nt_bf16(1, M, N, K, A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), N,
                                        _lib.stream_ptr(dev)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    by = (M * K + N * K + M * N) * 2
    print(f"{name:8s} M={M} N={N} K={K}: {us:7.1f} us  {2.0 * M * N * K / us * 1e-6:7.1f} TFLOP/s  {by / us * 1e-3:7.1f} GB/s "
          f"(TMA store {'off' if os.environ.get('DIC_TMA_STORE') == '0' else 'on'})", flush=True)
