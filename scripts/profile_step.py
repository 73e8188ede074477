#!/usr/bin/env python
"""One hot-path step inside a cudaProfilerStart/Stop window, for `ncu --profile-from-start off`.
usage: python scripts/profile_step.py [--batch 256] [--stage all|encode|decode]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops, synth  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402
from embodied_captioning_b200.engine import CaptionEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--model", default="coca_ViT-L-14")
    ap.add_argument("--stage", default="all")
    ap.add_argument("--seq-len", type=int, default=30)
    ap.add_argument("--graph", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = get_config(args.model)
    eng = CaptionEngine(args.model, None, device=dev, decode_graph=bool(args.graph),
                        state_dict=synth.make_state_dict(cfg, seed=1234))
    crops = torch.from_numpy(synth.make_crops(args.batch, seed=5678)).to(dev)
    keys = torch.from_numpy(np.arange(args.batch) // 8).to(dev)
    out = eng.caption_crops(crops, keys, args.seq_len)          # warm-up (graph capture, attribute setup)
    torch.cuda.synchronize()
    x, _, _ = ops.preprocess_crops(crops, torch.tensor([0, 0, 224, 224], dtype=torch.int32, device=dev).repeat(args.batch, 1),
                                   torch.arange(args.batch, dtype=torch.int32, device=dev), None)
    lat, embs = eng.model._encode_image_dev(x)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    ev0.record()
    if args.stage in ("all",):
        eng.caption_crops(crops, keys, args.seq_len)
    elif args.stage == "encode":
        eng.model._encode_image_dev(x)
    elif args.stage == "decode":
        eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=args.seq_len,
                           fixed_output_length=True, return_logits=False)
    ev1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled stage", args.stage, "batch", args.batch, "graph", args.graph, "PDL", os.environ.get("EC_PDL", "0"),
          "ms", round(ev0.elapsed_time(ev1), 3))


if __name__ == "__main__":
    main()
