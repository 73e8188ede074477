#!/usr/bin/env python
"""Do chains of decoder-shaped GEMMs (qkv -> out -> fc1 -> fc2, width 768) overlap across separate linear graphs?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
T = 768; BLOCKS = 12

def mk(M):
    g = lambda n, k: (torch.randn(n, k, device=dev) * k ** -0.5).to(torch.bfloat16)
    return dict(x=torch.randn(M, T, device=dev), a=torch.randn(M, T, device=dev).to(torch.bfloat16),
                wqkv=g(3 * T, T), wo=g(T, T), w1=g(4 * T, T), w2=g(T, 4 * T))

def chain(d):
    a = d["a"]; x = d["x"]
    for _ in range(BLOCKS):
        q = ops.gemm(a, d["wqkv"], None, ops.EPI_BF16)
        x = ops.gemm(q[:, :T], d["wo"], None, ops.EPI_F32_RESID, residual=x)
        h = ops.gemm(a, d["w1"], None, ops.EPI_BF16_GELU)
        x = ops.gemm(h, d["w2"], None, ops.EPI_F32_RESID, residual=x)
    return x

def linear_graph(M):
    d = mk(M); chain(d); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = chain(d)
    return g, out, d

for M, nch in ((1024, 1), (512, 1), (512, 2), (256, 1), (256, 4), (128, 1), (128, 8)):
    gs = [linear_graph(M) for _ in range(nch)]
    ss = [torch.cuda.Stream() for _ in range(nch)]
    torch.cuda.synchronize()
    def run():
        for (g, _, _), s in zip(gs, ss):
            with torch.cuda.stream(s):
                g.replay()
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    a.record()
    for s in ss: s.wait_stream(cur)
    for _ in range(5): run()
    for s in ss: cur.wait_stream(s)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5
    print(f"M={M} x {nch} graph(s): {t*1e3:.1f} us per round = {t*1e3/(BLOCKS*4):.2f} us per GEMM level ({nch*M} rows)", flush=True)
    del gs
