#!/usr/bin/env python
"""Why is the decoder slower right after the image tower than on its own?  Times generate() (29 steps, 1024 crops) after:
another generate, the tower, the tower + an idle gap, and 1 / 4 of the tower.
usage: python scripts/decode_after_tower_probe.py"""
import os, sys, time
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
lat, embs = eng.model._encode_image_dev(x)
embs = embs.clone()
xq = x[:B // 4].contiguous()


def dec():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=30, fixed_output_length=True, return_logits=False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


def enc(inp):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.model._encode_image_dev(inp); b.record()
    return a, b


for _ in range(2):
    dec()
res = {}
res["after another decode"] = [dec() for _ in range(4)]
for name, inp, gap in (("after the tower (1024 crops)", x, 0.0), ("after the tower + 20 ms idle", x, 0.02), ("after the tower + 100 ms idle", x, 0.1),
                       ("after a quarter tower (256 crops)", xq, 0.0)):
    ts = []
    for _ in range(4):
        a, b = enc(inp)
        if gap > 0:
            torch.cuda.synchronize(); time.sleep(gap)
        ts.append(dec())
    res[name] = ts
# two decodes after one tower: does the second one recover?
a, b = enc(x); d1 = dec(); d2 = dec()
res["first / second decode after one tower"] = [d1, d2]
for k, v in res.items():
    print(f"{k:40s} " + "  ".join(f"{t:6.1f}" for t in v) + " ms", flush=True)
