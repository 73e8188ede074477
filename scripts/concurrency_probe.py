#!/usr/bin/env python
"""Do independent chains of small tcgen05 GEMMs overlap when captured as parallel graph branches?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = K = 768
DEPTH = 64

def mk():
    return (torch.randn(M, K, device=dev).to(torch.bfloat16), torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.03)

def chain(a, w):
    x = a
    for _ in range(DEPTH):
        x = ops.gemm(x, w, None, ops.EPI_BF16)
    return x

def graph_of(nchains):
    data = [mk() for _ in range(nchains)]
    streams = [torch.cuda.Stream() for _ in range(nchains - 1)]
    for a, w in data: chain(a, w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        outs = []
        for i, (a, w) in enumerate(data):
            if i == 0:
                outs.append(chain(a, w))
            else:
                s = streams[i - 1]
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    outs.append(chain(a, w))
        for s in streams:
            cur.wait_stream(s)
    return g, outs

for nch in (1, 2, 4, 8):
    g, outs = graph_of(nch)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5
    print(f"M={M}: {nch} chain(s) x {DEPTH} dependent GEMMs: {t*1e3:.1f} us per replay, {t*1e3/DEPTH:.2f} us per GEMM level", flush=True)

# ---- separate linear graphs replayed on separate streams
def linear_graph():
    a, w = mk()
    chain(a, w); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = chain(a, w)
    return g, out

for nch in (1, 2, 4, 8):
    gs = [linear_graph() for _ in range(nch)]
    ss = [torch.cuda.Stream() for _ in range(nch)]
    torch.cuda.synchronize()
    def run():
        for (g, _), s in zip(gs, ss):
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
    print(f"M={M}: {nch} separate linear graph(s) on own streams: {t*1e3:.1f} us per round, {t*1e3/DEPTH:.2f} us per GEMM level", flush=True)
