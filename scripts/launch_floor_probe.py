#!/usr/bin/env python
"""Fixed cost of one dependent kernel inside a linear CUDA graph: a tiny torch kernel vs the 1-CTA GEMM kernel at
K = 64 (one k-block) for 1 / 96 / 144 CTAs.  usage: python scripts/launch_floor_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0"); REP = 48
def timed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[3] * 1e3 / REP
t = torch.zeros(32, device=dev)
def tiny():
    for _ in range(REP): t.add_(1.0)
print(f"tiny torch kernel: {timed(tiny):.2f} us per node", flush=True)
for (M, N, K, epi, name) in [(128, 64, 64, ops.EPI_BF16, "1 CTA"), (1024, 768, 64, ops.EPI_BF16, "96 CTAs bf16"),
                             (1024, 768, 64, ops.EPI_F32_RESID, "96 CTAs resid"), (1024, 2304, 64, ops.EPI_BF16, "144 CTAs bf16"),
                             (1024, 3072, 64, ops.EPI_BF16_GELU, "192 tiles gelu"), (1024, 768, 768, ops.EPI_BF16, "96 CTAs K=768"),
                             (1024, 768, 3072, ops.EPI_F32_RESID, "96 CTAs K=3072 resid")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    x = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
    def chain():
        r = x
        for _ in range(REP):
            o = ops.gemm(a, w, None, epi, residual=r)
            if epi == ops.EPI_F32_RESID: r = o
    print(f"gemm {M}x{N}x{K} ({name}): {timed(chain):.2f} us per node", flush=True)

# Mixed chains: does alternating the big-smem GEMM with small kernels cost more than the sum of its parts?
M, N, K = 1024, 768, 768
a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
x = torch.randn(M, N, device=dev); gam = torch.ones(N, device=dev); bet = torch.zeros(N, device=dev)
def ln_chain():
    for _ in range(REP): ops.layernorm(x, gam, bet)
def gemm_chain():
    r = x
    for _ in range(REP): r = ops.gemm(a, w, None, ops.EPI_F32_RESID, residual=r)
def gemm_tiny():
    r = x
    for _ in range(REP): r = ops.gemm(a, w, None, ops.EPI_F32_RESID, residual=r); t.add_(1.0)
def gemm_ln():
    r = x
    for _ in range(REP): r = ops.gemm(a, w, None, ops.EPI_F32_RESID, residual=r); ops.layernorm(r, gam, bet)
for name, fn in [("layernorm 1024x768", ln_chain), ("gemm resid 1024x768x768", gemm_chain), ("gemm + tiny", gemm_tiny), ("gemm + layernorm", gemm_ln)]:
    print(f"{name}: {timed(fn):.2f} us per iteration", flush=True)
