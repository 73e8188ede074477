#!/usr/bin/env python
"""Microbenchmark of the absorbed cross-attention pieces at the bench shape (1024 crops, ViT-L/14 decoder):
the streaming attention kernel and the two grouped GEMMs on either side of it.
usage: python scripts/xattn_probe.py [--batch 1024] [--iters 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops  # noqa: E402


def bench(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1)                       # 512 MB write: evicts L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, T, NK, hd = args.batch, 12, 768, 255, 64
    g = torch.Generator().manual_seed(0)
    flush = torch.zeros(128 * 1024 * 1024, device=dev)
    qp = (torch.randn((B, H, T), generator=g) * 0.1).to(torch.bfloat16).to(dev)
    xn = torch.randn((B, NK, T), generator=g).to(torch.bfloat16).to(dev)
    q = torch.randn((B, H * hd), generator=g).to(torch.bfloat16).to(dev)
    wk = (torch.randn((H * T, hd), generator=g) * 0.1).to(torch.bfloat16).to(dev)
    wv = (torch.randn((H * hd, T), generator=g) * 0.03).to(torch.bfloat16).to(dev)
    ctx = torch.randn((B, H * T), generator=g).to(torch.bfloat16).to(dev)
    w_sq = (torch.randn((T, T), generator=g) * 0.03).to(torch.bfloat16).to(dev)
    x = torch.randn((B, T), generator=g).to(torch.bfloat16).to(dev)
    packed = ops.pack_image_tokens(xn.float())
    out = torch.zeros_like(qp)
    med, best = bench(lambda: ops.cross_attention_absorbed(qp, packed, NK, out=out), args.iters, flush)
    gb = B * (NK * T * 2 + 2 * H * T * 2) / 1e9
    print(f"cross_attention_absorbed B={B}: median {med:.1f} us, best {best:.1f} us  ({gb / (med * 1e-6):.0f} GB/s of {gb:.3f} GB)")
    med, best = bench(lambda: ops.gemm_grouped(q, wk, group_cols=T), args.iters, flush)
    print(f"grouped GEMM q' (M={B}, N={H * T}, K={hd}): median {med:.1f} us, best {best:.1f} us")
    med, best = bench(lambda: ops.gemm_grouped(ctx, wv, group_cols=hd), args.iters, flush)
    print(f"grouped GEMM v (M={B}, N={T}, K={T}, groups of {hd}): median {med:.1f} us, best {best:.1f} us")
    med, best = bench(lambda: ops.gemm(x, w_sq), args.iters, flush)
    print(f"plain GEMM (M={B}, N={T}, K={T}): median {med:.1f} us, best {best:.1f} us")
    ops.check_watchdog(dev)


if __name__ == "__main__":
    main()
