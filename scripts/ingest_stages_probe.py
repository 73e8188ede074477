import os, sys, torch
sys.path.insert(0, "/root/repo")
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
for (M, N, K) in ((128, 64, 16384), (1024, 768, 3072)):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, None, ops.EPI_BF16)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, None, ops.EPI_BF16); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(os.environ.get("EC_LIB_PATH", "default")[-12:], M, N, K, f"{sorted(ts)[2]*1e3:.1f} us", flush=True)
