#!/usr/bin/env python
"""Same-box A/B of GEMM variants: interleaves the shapes of the ViT-L/14 tower, median of 7, L2 flushed.
usage: EC_LIB_PATH=... python scripts/gemm_ab.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
B = 1024; Mv = B * 257
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [("qkv", Mv, 3072, 1024, ops.EPI_BF16), ("out_proj", Mv, 1024, 1024, ops.EPI_F32_RESID),
          ("fc1_gelu", Mv, 4096, 1024, ops.EPI_BF16_GELU), ("fc2", Mv, 1024, 4096, ops.EPI_F32_RESID),
          ("patch", B * 256, 1024, 592, ops.EPI_F32)]
res = {}
for name, M, N, K, epi in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
    ts = []
    for i in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, bias, epi, residual=r); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    res[name] = (t, 2.0 * M * N * K / t / 1e9)
    del a, w, r
print(os.environ.get("EC_LIB_PATH", "default"), os.environ.get("EC_GEMM_PAIR", ""), " | ".join(f"{k} {v[0]:.3f}ms {v[1]:.0f}TF" for k, v in res.items()), flush=True)
