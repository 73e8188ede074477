#!/usr/bin/env python
"""Stage clocks of the decoder's GEMM-sequence kernel (needs a library built with EC_NVCC_EXTRA=-DEC_SEQ_TIMING).
Runs one greedy decode at the bench shape and prints, for a few launches of the LAST step, CTA 0's event times in us
(SM cycles / clock) relative to kernel entry.  usage: python scripts/seq_timing_probe.py [--batch 1024] [--tags 0,1,22,23]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import _native as nat, synth  # noqa: E402
from embodied_captioning_b200.captioner.models.coca.coca_model import CoCa  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402

EV = ["prologue", "dep met", "loads issued", "operands in", "acc committed", "acc ready", "tile published",
      "block complete", "rows published"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--tags", default="0,1,12,13,34,35")
    ap.add_argument("--mhz", type=float, default=1965.0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = get_config("coca_ViT-L-14")
    model = CoCa(cfg, device=dev)
    model.load_state_dict(synth.make_state_dict(cfg, seed=1234))
    model.set_decode_graph(True)
    g = torch.Generator().manual_seed(3)
    embs = torch.randn((args.batch, cfg.n_image_embs, cfg.embed_dim), generator=g).to(dev)
    for _ in range(3):
        model.generate(None, image_embs=embs, generation_type="top_k", seq_len=30, fixed_output_length=True, return_logits=False)
    torch.cuda.synchronize()
    n = 64 * 6 * 16
    buf = (C.c_int64 * n)()
    nat.check(nat.lib().ec_seq_timing_read(buf, n), "ec_seq_timing_read")
    for tag in [int(t) for t in args.tags.split(",")]:
        print(f"== launch {tag} of the step")
        for s in range(6):
            row = [buf[(tag * 6 + s) * 16 + e] for e in range(len(EV))]
            if not any(row):
                continue
            print(f"  stage {s}: " + ", ".join(f"{EV[e]} {row[e] / args.mhz:.2f}" for e in range(len(EV)) if row[e]))


if __name__ == "__main__":
    main()
