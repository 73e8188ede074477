#!/usr/bin/env python
"""Decoder-step GEMM shapes as dependent chains inside one CUDA graph (the way the decoder runs them): us per GEMM.
usage: [EC_GEMM_BN=64|128|256] python scripts/dec_gemm_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
T = 768; M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024; REP = 24
g = lambda n, k: (torch.randn(n, k, device=dev) * k ** -0.5).to(torch.bfloat16)
shapes = {"qkv": (3 * T, T, ops.EPI_BF16), "out": (T, T, ops.EPI_F32_RESID), "fc1": (4 * T, T, ops.EPI_BF16_GELU),
          "fc2": (T, 4 * T, ops.EPI_F32_RESID), "q": (T, T, ops.EPI_BF16)}
res = {}
for name, (N, K, epi) in shapes.items():
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); ws = [g(N, K) for _ in range(REP)]      # distinct weights: cold like the real step
    x = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
    def chain():
        r = x
        for w in ws:
            o = ops.gemm(a, w, None, epi, residual=r)
            if epi == ops.EPI_F32_RESID: r = o
    chain(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): chain()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name] = sorted(ts)[2] * 1e3 / REP
print(f"M={M} BN={os.environ.get('EC_GEMM_BN','auto')}: " + " | ".join(f"{k} {v:.2f}us" for k, v in res.items()), flush=True)
