#!/usr/bin/env python
"""ViT-block attention at the image-tower shape (B crops x 16 heads x 257 tokens x 64): block-wise kernel (impl 1) vs the
whole-row kernel (impl 2), accuracy against fp32 and time per launch (cold: distinct inputs; hot loop: 20 launches).
usage: python scripts/attn_probe.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H, S, D = 16, 257, 64
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(5)
qkv = (torch.randn(B, S, 3 * H * D, generator=g) * 1.5).to(torch.bfloat16).to(dev)
# accuracy on the first 4 crops (fp32 reference on the GPU via torch)
sub = qkv[:4]
C = H * D
qf = sub[:, :, :C].float().view(4, S, H, D).transpose(1, 2)
kf = sub[:, :, C:2 * C].float().view(4, S, H, D).transpose(1, 2)
vf = sub[:, :, 2 * C:].float().view(4, S, H, D).transpose(1, 2)
ref = (torch.softmax(qf @ kf.transpose(-1, -2) / D ** 0.5, dim=-1) @ vf).transpose(1, 2).reshape(4, S, C)
for impl in (1, 2):
    out = ops.attention_fused_qkv(qkv, H, impl=impl)
    ops.check_watchdog(dev)
    o = out[:4].float()
    err = (o - ref).abs().max().item()
    rel = ((o - ref).norm() / ref.norm()).item()
    row_err = (o - ref).abs().amax(dim=(0, 2))
    print(f"impl {impl}: max err {err:.4f} rel {rel:.2e}; worst rows {row_err.topk(3).indices.tolist()}; row 256 err {row_err[256]:.4f}; "
          f"finite {bool(torch.isfinite(out.float()).all())}", flush=True)
    for _ in range(3):
        ops.attention_fused_qkv(qkv, H, impl=impl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.attention_fused_qkv(qkv, H, impl=impl)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    tf = 4 * S * S * D * H * B / (ms * 1e-3) / 1e12
    print(f"impl {impl}: B {B} {ms:.3f} ms per launch ({tf:.0f} TFLOP/s)", flush=True)
o1 = ops.attention_fused_qkv(qkv, H, impl=1).float(); o2 = ops.attention_fused_qkv(qkv, H, impl=2).float()
print("v2 vs v1 max diff", (o1 - o2).abs().max().item())
