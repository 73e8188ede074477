#!/usr/bin/env python
"""Per-SM operand ingest rate of the 1-CTA GEMM (TMA L2 -> SMEM): long-K GEMMs with one / many CTAs, L2-warm."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
for (M, N, K) in ((128, 64, 16384), (128, 128, 16384), (128, 256, 8192), (1024, 768, 16384), (1024, 768, 3072), (2048, 1152, 4096)):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, None, ops.EPI_BF16)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, None, ops.EPI_BF16); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2] * 1e-3
    bn = 64 if N <= 768 and M * N < 148 * 128 * 128 else 128
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    per_cta = (128 + bn) * K * 2
    print(f"M={M} N={N} K={K}: {t*1e6:.1f} us; ~{tiles} CTAs (BN~{bn}) x {per_cta/1e6:.2f} MB -> {per_cta/t/1e9:.0f} GB/s per CTA, "
          f"{tiles*per_cta/t/1e12:.2f} TB/s aggregate, {2.0*M*N*K/t/1e12:.0f} TFLOP/s", flush=True)
