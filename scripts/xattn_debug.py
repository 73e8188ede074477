import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops
dev = torch.device("cuda:0")
def rb(shape, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)
for B in (149, 300, 1024):
    H, T, NK = 12, 768, 255
    qp = rb((B, H, T), 3.0 * T ** -0.5, 21); xn = rb((B, NK, T), 1.0, 22)
    packed = ops.pack_image_tokens(xn.float())
    s = torch.einsum("bht,bjt->bhj", qp.float(), xn.float())
    ref = torch.einsum("bhj,bjt->bht", torch.softmax(s, dim=-1), xn.float())
    for rep in range(8):
        ctx = ops.cross_attention_absorbed(qp, packed, NK)
        torch.cuda.synchronize()
        err = (ctx.float() - ref).abs().amax(dim=(1, 2))
        bad = (err > 0.03).nonzero().flatten().tolist()
        print(f"B={B} rep={rep}: max err {err.max().item():.4f}; bad crops {len(bad)}: {bad[:24]}")
        if bad:
            b = bad[0]
            e = (ctx[b].float() - ref[b]).abs()
            print("   crop", b, "per-head max err", [round(x, 3) for x in e.amax(dim=1).tolist()])
            print("   per-96-col-slice max err", [round(e[:, i * 96:(i + 1) * 96].max().item(), 3) for i in range(8)])
            # does it match another crop's reference?
            d = (ctx[b].float()[None] - ref).abs().amax(dim=(1, 2))
            print("   closest reference crop:", int(d.argmin()), float(d.min()))
ops.check_watchdog(dev)
