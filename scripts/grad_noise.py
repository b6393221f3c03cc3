"""Diagnostic: per-parameter gradient error of the CUDA path (fp32 / bf16) against an fp64
oracle, next to the CPU fp32 oracle's own error (the achievable floor)."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import depth_image_captioning_pub_b200 as P
from depth_image_captioning_pub_b200 import _lib
from oracle import decoder_oracle as O
from test_gpu_parity import make_case, CASES, build_module

dev = torch.device("cuda:0")
for case in sys.argv[1:] or ["ref_dims"]:
    cfg = dict(CASES[case]); lengths = cfg["lengths"]; V = cfg["V"]
    w, F_rgb, F_dep, caps = make_case(**cfg)
    def oracle(dt):
        wo = {k: v.clone().to(dt).requires_grad_(True) for k, v in w.items()}
        a = F_rgb.clone().to(dt).requires_grad_(True); b = F_dep.clone().to(dt).requires_grad_(True)
        lo, _, ao = O.decoder_forward(wo, a, b, caps, lengths, hoist=True)
        O.caption_loss(lo, O.pack_targets(caps, lengths), V - 1, ao).backward()
        return {k: wo[k].grad.double().numpy() for k in wo}, b.grad.double().numpy(), lo.detach().double().numpy(), ao.detach().double().numpy()
    g64, f64, l64, a64 = oracle(torch.float64)
    g32, f32, l32, a32 = oracle(torch.float32)
    res = {}
    for prec in ("fp32", "bf16"):
        m = build_module(P.CD_RNNDecoderWithSoftAttention, w, dev, prec).eval()
        Fr = F_rgb.to(dev).requires_grad_(True); Fd = F_dep.to(dev).requires_grad_(True)
        out, alphas = m(Fr, Fd, caps.to(dev), lengths)
        tg = O.pack_targets(caps, lengths).to(dev)
        loss = torch.nn.functional.cross_entropy(out.data, tg, ignore_index=V - 1) + 0.7 * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
        loss.backward()
        gp = dict(m.named_parameters())
        res[prec] = ({k: gp[k].grad.double().cpu().numpy() for k in _lib.PARAM_KEYS}, Fd.grad.double().cpu().numpy(),
                     out.data.detach().double().cpu().numpy(), alphas.detach().double().cpu().numpy())
    print(f"== {case}:  relative-to-max errors vs fp64 oracle: cpu_fp32 | gpu_fp32 | gpu_bf16")
    def rel(x, r): return np.abs(x - r).max() / max(np.abs(r).max(), 1e-30)
    print(f"{'logits':34s} {rel(l32,l64):.1e} {rel(res['fp32'][2],l64):.1e} {rel(res['bf16'][2],l64):.1e}")
    print(f"{'alphas(abs)':34s} {np.abs(a32-a64).max():.1e} {np.abs(res['fp32'][3]-a64).max():.1e} {np.abs(res['bf16'][3]-a64).max():.1e}")
    for k in _lib.PARAM_KEYS:
        print(f"{k:34s} {rel(g32[k],g64[k]):.1e} {rel(res['fp32'][0][k],g64[k]):.1e} {rel(res['bf16'][0][k],g64[k]):.1e}   max|ref|={np.abs(g64[k]).max():.1e}")
    print(f"{'d depth_features':34s} {rel(f32,f64):.1e} {rel(res['fp32'][1],f64):.1e} {rel(res['bf16'][1],f64):.1e}")
