#!/usr/bin/env python
"""One GEMM shape of the ViT-L/14 tower, a few launches (ncu target).  usage: gemm_one.py qkv|out_proj|fc1_gelu|fc2|patch [batch]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
name = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024; Mv = B * 257
S = {"qkv": (Mv, 3072, 1024, ops.EPI_BF16), "out_proj": (Mv, 1024, 1024, ops.EPI_F32_RESID),
     "fc1_gelu": (Mv, 4096, 1024, ops.EPI_BF16_GELU), "fc2": (Mv, 1024, 4096, ops.EPI_F32_RESID),
     "patch": (B * 256, 1024, 592, ops.EPI_F32),
     # decoder-step shapes (B rows)
     "dec_qkv": (B, 2304, 768, ops.EPI_BF16), "dec_out": (B, 768, 768, ops.EPI_F32_RESID),
     "dec_fc1": (B, 3072, 768, ops.EPI_BF16_GELU), "dec_fc2": (B, 768, 3072, ops.EPI_F32_RESID)}
M, N, K, epi = S[name]
a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, w, bias, epi, residual=r); e1.record(); torch.cuda.synchronize()
print(name, M, N, K, "ms", e0.elapsed_time(e1), "TF", 2.0 * M * N * K / e0.elapsed_time(e1) / 1e9)
ops.check_watchdog(dev)
