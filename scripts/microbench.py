#!/usr/bin/env python
"""Per-kernel timings (CUDA events, warm, inputs >> L2 or L2 flushed) for the shapes of the ViT-L/14 path.
usage: python scripts/microbench.py [--batch 1024] > gpurun_out/microbench.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops  # noqa: E402


def timeit(fn, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"batch": B, "gemm": [], "attention": [], "layernorm": [], "decode_gemm": []}
    Mv = B * 257
    shapes = [("qkv", Mv, 3072, 1024, ops.EPI_BF16), ("out_proj", Mv, 1024, 1024, ops.EPI_F32_RESID),
              ("fc1_gelu", Mv, 4096, 1024, ops.EPI_BF16_GELU), ("fc2", Mv, 1024, 4096, ops.EPI_F32_RESID),
              ("pool_kv", Mv, 1536, 1024, ops.EPI_BF16), ("cross_kv", B * 255, 1536, 768, ops.EPI_BF16),
              ("patch", B * 256, 1024, 592, ops.EPI_F32)]
    for name, M, N, K, epi in shapes:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
        ms = timeit(lambda: ops.gemm(a, w, bias, epi, residual=res))
        out["gemm"].append({"name": name, "M": M, "N": N, "K": K, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9})
        del a, w, res
    for name, M, N, K, epi in [("dec_qkv", B, 2304, 768, ops.EPI_BF16), ("dec_out", B, 768, 768, ops.EPI_F32_RESID),
                               ("dec_fc1", B, 3072, 768, ops.EPI_BF16_GELU), ("dec_fc2", B, 768, 3072, ops.EPI_F32_RESID)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if epi == ops.EPI_F32_RESID else None
        ms = timeit(lambda: ops.gemm(a, w, bias, epi, residual=res), iters=20, flush=flush)
        out["decode_gemm"].append({"name": name, "M": M, "N": N, "K": K, "us": ms * 1e3,
                                   "tflops": 2.0 * M * N * K / ms / 1e9})
    a = torch.randn(B, 768, device=dev).to(torch.bfloat16)
    w = (torch.randn(49408, 768, device=dev) * 768 ** -0.5).to(torch.bfloat16)
    ms = timeit(lambda: ops.gemm_argmax(a, w, want_logits=False), iters=20, flush=flush)
    out["decode_gemm"].append({"name": "vocab_argmax", "M": B, "N": 49408, "K": 768, "us": ms * 1e3,
                               "tflops": 2.0 * B * 49408 * 768 / ms / 1e9})
    qkv = torch.randn(B, 257, 3072, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv[:, :, :1024], qkv[:, :, 1024:2048], qkv[:, :, 2048:], 16))
    out["attention"].append({"name": "vit_l14_self", "ms": ms, "tflops": 4.0 * B * 16 * 257 * 257 * 64 / ms / 1e9})
    ms = timeit(lambda: ops.attention_fused_qkv(qkv, 16, impl=1))
    out["attention"].append({"name": "vit_l14_self_tcgen05", "ms": ms, "tflops": 4.0 * B * 16 * 257 * 257 * 64 / ms / 1e9})
    kv = torch.randn(B, 257, 1536, device=dev).to(torch.bfloat16)
    q = torch.randn(256, 768, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(q, kv[:, :, :768], kv[:, :, 768:], 8))
    out["attention"].append({"name": "pooler", "ms": ms, "tflops": 4.0 * B * 8 * 256 * 257 * 96 / ms / 1e9})
    x = torch.randn(Mv, 1024, device=dev)
    g = torch.ones(1024, device=dev); b = torch.zeros(1024, device=dev)
    ms = timeit(lambda: ops.layernorm(x, g, b, out_bf16=True))
    out["layernorm"].append({"name": "ln_1024_bf16", "rows": Mv, "ms": ms, "gbs": Mv * 1024 * 6 / ms / 1e6})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
