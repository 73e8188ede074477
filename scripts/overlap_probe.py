#!/usr/bin/env python
"""Cross-batch overlap probe: the decoder of batch i on one stream next to the image tower of batch
i + 1 on another, against the same two calls back to back.  No kernel is changed: the tower's persistent GEMMs take every
SM for ~1.2 ms at a time, every decoder GEMM CTA needs an SM to itself.
usage: python scripts/overlap_probe.py [--batch 1024]"""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import ops, synth  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402
from embodied_captioning_b200.engine import CaptionEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--seq-len", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = get_config("coca_ViT-L-14")
eng = CaptionEngine("coca_ViT-L-14", None, device=dev, decode_graph=True, state_dict=synth.make_state_dict(cfg, seed=1234))
crops = torch.from_numpy(synth.make_crops(args.batch, seed=5678)).to(dev)
rects = torch.tensor([0, 0, 224, 224], dtype=torch.int32, device=dev).repeat(args.batch, 1)
x, _, _ = ops.preprocess_crops(crops, rects, torch.arange(args.batch, dtype=torch.int32, device=dev), None)
lat, embs = eng.model._encode_image_dev(x)
embs = embs.clone()


def enc():
    return eng.model._encode_image_dev(x)


def dec():
    return eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=args.seq_len, fixed_output_length=True,
                              return_logits=False)["text"]


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


ref = dec()
t_enc, t_dec = timed(enc), timed(dec)
t_seq = timed(lambda: (enc(), dec()))
for prio in (0, -1):
    s_dec = torch.cuda.Stream(device=dev, priority=prio)
    s_enc = torch.cuda.Stream(device=dev)
    out = {}

    def both():
        cur = torch.cuda.current_stream()
        s_dec.wait_stream(cur); s_enc.wait_stream(cur)
        with torch.cuda.stream(s_dec):
            out["tok"] = dec()
        with torch.cuda.stream(s_enc):
            enc()
        cur.wait_stream(s_dec); cur.wait_stream(s_enc)
    t_both = timed(both)
    same = bool(torch.equal(out["tok"], ref))
    print(f"batch {args.batch}: tower alone {t_enc:.1f} ms, decoder alone {t_dec:.1f} ms, back to back {t_seq:.1f} ms, "
          f"two streams (decoder priority {prio}) {t_both:.1f} ms; decoder tokens unchanged: {same}", flush=True)
