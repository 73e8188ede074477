#!/usr/bin/env python
"""Where does a small decoder GEMM launch spend its time?  Needs a library built with EC_NVCC_EXTRA=-DEC_GEMM_TIMING.
Runs greedy decodes at the bench shape and prints, per GEMM shape, the median of CTA 0's stage clocks (us since kernel
entry).  usage: python scripts/gemm_timing_probe.py [--batch 1024] [--mhz 1965]"""
import argparse
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import _native as nat, synth  # noqa: E402
from embodied_captioning_b200.captioner.models.coca.coca_model import CoCa  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402

EV = ["skip word", "prologue", "operands in", "last MMA", "acc ready", "tile stored", "LN row0 done", "LN rows done"]
EPI = {0: "bf16", 1: "gelu", 2: "f32", 3: "resid", 4: "argmax"}


def read_records(batch):
    n = 8192 * 12
    buf = (C.c_int64 * n)()
    nat.check(nat.lib().ec_gemm_timing_read(buf, n), "ec_gemm_timing_read")
    groups = {}
    for i in range(8192):
        r = buf[i * 12:(i + 1) * 12]
        if r[0] == 0 or r[3] != batch or r[9] == 0:
            continue
        groups.setdefault((r[0], r[1], r[2] // 8, r[2] % 8), []).append(r[4:12])
    return groups


def micro(args, dev):
    from embodied_captioning_b200 import ops
    g = torch.Generator().manual_seed(0)
    flush = torch.zeros(128 * 1024 * 1024, device=dev)
    for M, N, K in ((args.batch + 1, 768, 768), (args.batch + 2, 768, 3072)):       # M tags the records
        x = torch.randn((M, K), generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn((N, K), generator=g) * 0.03).to(torch.bfloat16).to(dev)
        for mode in ("hot", "flushed"):
            for _ in range(300):
                if mode == "flushed":
                    flush.add_(1)
                ops.gemm(x, w)
            torch.cuda.synchronize()
            rows = [r for k, v in read_records(M).items() for r in v][-250:]
            med = [statistics.median(x_[e] for x_ in rows) / args.mhz for e in range(8)]
            print(f"micro M={M} N={N} K={K} {mode:8s}: " + ", ".join(f"{EV[e]} {med[e]:.2f}" for e in range(8)))
            # overwrite the ring so that the next mode reads its own records
            y = torch.randn((7, 64), generator=g).to(torch.bfloat16).to(dev)
            wy = torch.randn((64, 64), generator=g).to(torch.bfloat16).to(dev)
            for _ in range(8192):
                ops.gemm(y, wy)
            torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--mhz", type=float, default=1965.0)
    ap.add_argument("--micro", action="store_true", help="one GEMM shape launched back to back: L2-hot, then with an L2 flush between launches")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.micro:
        return micro(args, dev)
    cfg = get_config("coca_ViT-L-14")
    model = CoCa(cfg, device=dev)
    model.load_state_dict(synth.make_state_dict(cfg, seed=1234))
    model.set_decode_graph(True)
    g = torch.Generator().manual_seed(3)
    embs = torch.randn((args.batch, cfg.n_image_embs, cfg.embed_dim), generator=g).to(dev)
    for _ in range(3):
        model.generate(None, image_embs=embs, generation_type="top_k", seq_len=30, fixed_output_length=True, return_logits=False)
    torch.cuda.synchronize()
    n = 8192 * 12
    buf = (C.c_int64 * n)()
    nat.check(nat.lib().ec_gemm_timing_read(buf, n), "ec_gemm_timing_read")
    groups = {}
    for i in range(8192):
        r = buf[i * 12:(i + 1) * 12]
        if r[0] == 0 or r[3] != args.batch or r[9] == 0:
            continue
        groups.setdefault((r[0], r[1], r[2] // 8, r[2] % 8), []).append(r[4:12])
    for (N, K, bn, epi), rows in sorted(groups.items()):
        med = [statistics.median(x[e] for x in rows) / args.mhz for e in range(8)]
        print(f"N={N:5d} K={K:5d} BN={bn:3d} {EPI.get(epi, epi):6s} n={len(rows):4d}: " +
              ", ".join(f"{EV[e]} {med[e]:.2f}" for e in range(8)))


if __name__ == "__main__":
    main()
