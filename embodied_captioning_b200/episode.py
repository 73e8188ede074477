"""Episode replay (BASELINE.json configs[2]): RGB frames + detector boxes -> captions -> per-instance multi-view
disagreement after every frame -> disagreement map + exploration reward.

B200-first form of the reference's online loop: per box `rgb[y1:y2, x1:x2]` -> `captioner(PIL)` -> SBERT embedding
(agents/goal_exploration/goal_exploration.py:95-105), embeddings appended to the per-object sets of the world model
(utils/projection_utils.py:388-430), `unique(dim=0)` per object (:574-578), `_cosine_distance` per object
(:743-754), rasterisation (:760-785) and reward (envs/env_habitat3.py:75-84).  Here all crops of a chunk of frames go
through ONE gather launch, one encoder pass and one decode loop; the cumulative per-frame reductions run on the
device over a growing prefix of the (frame-ordered) embedding table.

The caption-text encoder of the reference (SBERT) is not part of the CUDA path: `embed_fn` maps the
token matrix [N, T] (device) to embeddings [N, D] (device); the default uses the CoCa pooled image latent, which is
what BASELINE.json's north_star substitutes for the similarity path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from . import ops
from .captioner.transform import POLICY_AGENT, box_rects
from .disagreement import disagreement_map


@dataclass
class FrameResult:
    frame: int
    inst_keys: torch.Tensor      # int64 [M] ascending
    inst_scores: torch.Tensor    # f32 [M]
    inst_counts: torch.Tensor    # int32 [M] distinct views
    reward: Optional[torch.Tensor] = None
    dmap: Optional[torch.Tensor] = None


@dataclass
class EpisodeResult:
    tokens: torch.Tensor         # int64 [N, seq_len]
    latent: torch.Tensor         # f32 [N, E]
    embeddings: torch.Tensor     # f32 [N, D] (what the disagreement ran on)
    status: torch.Tensor         # int32 [N] (1 = empty / unsupported crop)
    order: np.ndarray            # permutation that put the detections into frame order
    frames: List[FrameResult] = field(default_factory=list)


class EpisodeReplay:
    def __init__(self, engine, embed_fn: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None,
                 dedup: bool = True, mode: int = ops.MODE_FULL_MEAN_1MINUS):
        """engine: CaptionEngine.  embed_fn(tokens[N,T], latent[N,E]) -> f32 [N,D] on the device."""
        self.engine = engine
        self.embed_fn = embed_fn or (lambda tokens, latent: latent)
        self.dedup = dedup
        self.mode = mode

    @torch.no_grad()
    def run(self, frames, boxes, frame_idx, instance_ids, policy: int = POLICY_AGENT, seq_len: int = 30,
            chunk_crops: int = 1024, every_frame: bool = True, points=None, point_ids=None, bounds=None) -> EpisodeResult:
        """frames uint8 [F,H,W,3] (host or device); boxes float [N,4] xyxy; frame_idx int [N]; instance_ids int [N].
        Optional world points (f32 [P,3] + object id per point) and map bounds ((lx,ly,lz),(ux,uy,uz)) add the
        disagreement map and the reward to every frame result."""
        eng = self.engine
        dev = eng.device
        frames_t = torch.as_tensor(frames)
        F_, H, W = int(frames_t.shape[0]), int(frames_t.shape[1]), int(frames_t.shape[2])
        fidx = np.asarray(frame_idx, dtype=np.int64)
        order = np.argsort(fidx, kind="stable")                       # frame order = the order the agent saw them
        boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)[order]
        fidx = fidx[order]
        ids = np.asarray(instance_ids, dtype=np.int64)[order]
        rects, flags = box_rects(boxes, H, W, policy)
        n = len(fidx)
        toks, lats, stats = [], [], []
        for s in range(0, n, chunk_crops):
            e = min(n, s + chunk_crops)
            f0, f1 = int(fidx[s]), int(fidx[e - 1]) + 1                # only the frames this chunk touches travel
            fr = frames_t[f0:f1].to(dev, non_blocking=True)
            out = eng.caption_frames(fr, torch.from_numpy(rects[s:e]).to(dev),
                                     torch.from_numpy((fidx[s:e] - f0).astype(np.int32)).to(dev),
                                     torch.from_numpy(flags[s:e]).to(dev), None, seq_len)
            toks.append(out["tokens"]); lats.append(out["latent"]); stats.append(out["status"])
        tokens = torch.cat(toks) if toks else torch.zeros((0, seq_len), dtype=torch.int64, device=dev)
        latent = torch.cat(lats) if lats else torch.zeros((0, eng.cfg.embed_dim), device=dev)
        status = torch.cat(stats) if stats else torch.zeros((0,), dtype=torch.int32, device=dev)
        emb = self.embed_fn(tokens, latent).to(torch.float32).contiguous()
        keys = torch.from_numpy(ids).to(dev)
        res = EpisodeResult(tokens=tokens, latent=latent, embeddings=emb, status=status, order=order)
        if n == 0:
            return res
        # rows [0, upto) are the detections of frames <= f
        upto = np.searchsorted(fidx, np.arange(F_), side="right")
        todo = range(F_) if every_frame else [F_ - 1]
        pts = None if points is None else torch.as_tensor(points, dtype=torch.float32).to(dev)
        pids = None if point_ids is None else torch.as_tensor(np.asarray(point_ids, dtype=np.int64)).to(dev)
        for f in todo:
            u = int(upto[f])
            if u == 0:
                continue
            k, sc, ct = ops.instance_disagreement(emb[:u], keys[:u], self.mode, dedup=self.dedup)
            fr_res = FrameResult(frame=f, inst_keys=k, inst_scores=sc, inst_counts=ct)
            if pts is not None and bounds is not None:
                seen = torch.isin(pids, k)                             # points of objects detected so far
                fr_res.dmap, fr_res.reward = disagreement_map(pts[seen], pids[seen], k, sc, bounds[0], bounds[1])
            res.frames.append(fr_res)
        return res
