#!/usr/bin/env python
"""Timings of the non-headline BASELINE.json configs on one B200 (ViT-L/14, synthetic weights):
K2 episode replay (500 frames 640x480, 8 boxes/frame, real Pillow-exact resize, cumulative per-frame disagreement with
duplicate-view removal + map + reward), K4 long captions (seq_len 60, 5 views/instance).
usage: python scripts/config_probe.py [--frames 500]"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops, synth
from embodied_captioning_b200.captioner import transform as tr
from embodied_captioning_b200.config import get_config
from embodied_captioning_b200.engine import CaptionEngine
from embodied_captioning_b200.episode import EpisodeReplay

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=500); ap.add_argument("--model", default="coca_ViT-L-14")
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = get_config(args.model)
eng = CaptionEngine(args.model, None, device=dev, decode_graph=True, state_dict=synth.make_state_dict(cfg, seed=1234))

def timed(fn):
    torch.cuda.synchronize(); t0 = time.time(); r = fn(); torch.cuda.synchronize(); return r, time.time() - t0

# ---- K2
F, K = args.frames, 8
frames, boxes, fidx, ids = synth.make_frames_and_boxes(F, K, 480, 640, seed=31)
rng = np.random.default_rng(5); ids = ids.copy(); fresh = rng.random(len(ids)) < 0.1
ids[fresh] = 5_000_000 + np.arange(int(fresh.sum()))
P = 200_000
pts = np.stack([rng.uniform(-9.5, 9.5, P), rng.uniform(0, 3, P), rng.uniform(-9.5, 9.5, P)], 1).astype(np.float32)
pobj = rng.choice(np.unique(ids), size=P)
bounds = ((-10.0, 0.0, -10.0), (10.0, 3.0, 10.0))
rep = EpisodeReplay(eng, dedup=True)
frames_pinned = torch.from_numpy(frames).pin_memory()
rep.run(frames_pinned[:16], boxes[:128], fidx[:128], ids[:128], seq_len=30, every_frame=False)      # warm-up
res, t = timed(lambda: rep.run(frames_pinned, boxes, fidx, ids, policy=tr.POLICY_AGENT, seq_len=30, chunk_crops=1024,
                               every_frame=True, points=pts, point_ids=pobj, bounds=bounds))
n = len(ids)
_, t_cap = timed(lambda: rep.run(frames_pinned, boxes, fidx, ids, policy=tr.POLICY_AGENT, seq_len=30, chunk_crops=1024, every_frame=False))
print(f"K2 episode replay: {F} frames 640x480 (host, pinned), {n} crops, agent box policy, Pillow-exact resize: "
      f"{t:.3f} s end to end = {n / t:.0f} crops/s; captioning alone {t_cap:.3f} s = {n / t_cap:.0f} crops/s; "
      f"{F} cumulative per-frame reductions (dedup) + {F} maps of {res.frames[-1].dmap.shape[0]}x{res.frames[-1].dmap.shape[1]} "
      f"over {P} points + rewards: {(t - t_cap) / F * 1e3:.2f} ms per frame; last reward {float(res.frames[-1].reward):.4f}", flush=True)

# ---- K4
crops, kid = synth.make_instance_crops(205, 5, seed=8)            # 1025 -> 1024 crops, 5 views per instance
crops, kid = crops[:1024], kid[:1024]
c_dev, k_dev = torch.from_numpy(crops).to(dev), torch.from_numpy(kid).to(dev)
for sl in (30, 60):
    eng.caption_crops(c_dev, k_dev, seq_len=sl)
    _, t = timed(lambda: eng.caption_crops(c_dev, k_dev, seq_len=sl))
    print(f"K4 shape: 1024 crops, 5 views/instance, seq_len {sl}: {t * 1e3:.1f} ms = {1024 / t:.0f} crops/s", flush=True)
ops.check_watchdog(dev)
