#!/usr/bin/env python
"""Order of work over a stream of 1024-crop batches: (tower, decoder) per batch vs G towers then G decoders.  The kernels and
their shapes are the same either way; only the number of tower -> decoder transitions changes (the decoder loses ~20 ms right
after the power-capped tower, profiles/r02_decode_after_tower.txt).  usage: python scripts/pairing_probe.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops, synth  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402
from embodied_captioning_b200.engine import CaptionEngine  # noqa: E402

dev = torch.device("cuda:0")
cfg = get_config("coca_ViT-L-14")
eng = CaptionEngine("coca_ViT-L-14", None, device=dev, decode_graph=True, state_dict=synth.make_state_dict(cfg, seed=1234))
B = 1024
crops = torch.from_numpy(synth.make_crops(B, seed=5678)).to(dev)
rects = torch.tensor([0, 0, 224, 224], dtype=torch.int32, device=dev).repeat(B, 1)
x, _, _ = ops.preprocess_crops(crops, rects, torch.arange(B, dtype=torch.int32, device=dev), None)


def dec(embs):
    return eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=30, fixed_output_length=True, return_logits=False)["text"]


def run(G, total=8):
    toks = []
    for _ in range(total // G):
        encs = [eng.model._encode_image_dev(x) for _ in range(G)]
        toks += [dec(e[1]) for e in encs]
    return toks


ref = run(1, 2)[0]
for G in (1, 2, 4, 1, 2, 4):
    run(G, 4); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); toks = run(G, 8); b.record(); torch.cuda.synchronize()
    same = all(bool(torch.equal(t, ref)) for t in toks)
    print(f"groups of {G}: {a.elapsed_time(b) / 8:.1f} ms per 1024-crop batch ({1024 * 8 / a.elapsed_time(b) * 1e3:.0f} crops/s); tokens identical: {same}", flush=True)
