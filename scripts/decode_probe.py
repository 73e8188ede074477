#!/usr/bin/env python
"""Decoder-only timing at the bench shape for several row-chain counts.
usage: python scripts/decode_probe.py [--batch 1024] [--chains 1,2,4,8] [--seq-len 30]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import synth  # noqa: E402
from embodied_captioning_b200.captioner.models.coca.coca_model import CoCa  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--model", default="coca_ViT-L-14")
    ap.add_argument("--chains", default="1,2,4,8")
    ap.add_argument("--seq-len", type=int, default=30)
    ap.add_argument("--graph", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = get_config(args.model)
    model = CoCa(cfg, device=dev)
    model.load_state_dict(synth.make_state_dict(cfg, seed=1234))
    model.set_decode_graph(bool(args.graph))
    g = torch.Generator().manual_seed(3)
    embs = torch.randn((args.batch, cfg.n_image_embs, cfg.embed_dim), generator=g).to(dev)
    ref = None
    for ch in [int(c) for c in args.chains.split(",")]:
        model.set_decode_chains(ch)
        for _ in range(2):
            tok = model.generate(None, image_embs=embs, generation_type="top_k", seq_len=args.seq_len,
                                 fixed_output_length=True, return_logits=False)["text"]
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model.generate(None, image_embs=embs, generation_type="top_k", seq_len=args.seq_len,
                           fixed_output_length=True, return_logits=False)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        same = True if ref is None else bool(torch.equal(ref, tok))
        ref = tok if ref is None else ref
        print(f"decode B={args.batch} chains={ch} graph={args.graph} PDL={os.environ.get('EC_PDL', '0')}: "
              f"{min(ts):.2f} ms (median {sorted(ts)[1]:.2f}); tokens identical to first config: {same}", flush=True)


if __name__ == "__main__":
    main()
