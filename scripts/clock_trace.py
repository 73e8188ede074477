#!/usr/bin/env python
"""SM clock and board power at 5 ms resolution while the K1 step loop runs (what clock does the decoder see after the
power-capped image tower?).  usage: python scripts/clock_trace.py > trace.csv"""
import os, subprocess, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops, synth  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402
from embodied_captioning_b200.engine import CaptionEngine  # noqa: E402

dev = torch.device("cuda:0")
cfg = get_config("coca_ViT-L-14")
eng = CaptionEngine("coca_ViT-L-14", None, device=dev, decode_graph=True, state_dict=synth.make_state_dict(cfg, seed=1234))
crops = torch.from_numpy(synth.make_crops(1024, seed=5678)).to(dev)
keys = torch.from_numpy(np.arange(1024) // 8).to(dev)
for _ in range(3):
    eng.caption_crops(crops, keys, 30)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "5", "-i", "0"],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
marks = []
t0 = time.time()
for i in range(8):
    x, _, _ = ops.preprocess_crops(crops, torch.tensor([0, 0, 224, 224], dtype=torch.int32, device=dev).repeat(1024, 1),
                                   torch.arange(1024, dtype=torch.int32, device=dev), None)
    a = time.time(); lat, embs = eng.model._encode_image_dev(x); torch.cuda.synchronize(); b = time.time()
    eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=30, fixed_output_length=True, return_logits=False)
    torch.cuda.synchronize(); c = time.time()
    marks.append((a - t0, b - t0, c - t0))
time.sleep(0.2)
p.terminate()
out = p.stdout.read()
print("# step marks (s since loop start): tower start, tower end = decoder start, decoder end")
for m in marks:
    print("# %.4f %.4f %.4f  (tower %.1f ms, decoder %.1f ms)" % (m[0], m[1], m[2], (m[1] - m[0]) * 1e3, (m[2] - m[1]) * 1e3))
print("# loop started at unix time %.3f" % t0)
print(out)
