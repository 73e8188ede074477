import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qkv = torch.randn(B, 257, 3072, device="cuda").to(torch.bfloat16)
for _ in range(2):
    ops.attention_fused_qkv(qkv, 16, impl=1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ops.attention_fused_qkv(qkv, 16, impl=1); b.record(); torch.cuda.synchronize()
print("attention_sm100 B", B, "ms", a.elapsed_time(b))
