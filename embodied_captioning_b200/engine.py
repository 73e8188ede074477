"""Batched captioning engine: crops -> preprocess -> image tower -> greedy decode -> instance disagreement.

This is the B200-first form of the reference's hot loops (goal_exploration.py:95-105 batch-1 loop over
boxes; projection_utils.py:743-754 per-object loop): one gather launch for all crops, one encoder pass,
one KV-cached decode loop without per-token host syncs, one segmented reduction.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _native as nat
from . import ops
from .captioner.models.coca.factory import create_model
from .captioner.transform import POLICY_AGENT, box_rects


class CaptionEngine:
    def __init__(self, model_name: str = "coca_ViT-L-14", pretrained: Optional[str] = "synthetic",
                 device="cuda:0", decode_graph: bool = True, state_dict=None):
        self.device = nat.cuda_device(device)
        self.model = create_model(model_name, pretrained if state_dict is None else None, device=self.device)
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.model.set_decode_graph(decode_graph)
        self.cfg = self.model.cfg
        self._pinned: Dict[str, torch.Tensor] = {}

    # -------------------------------------------------------------------------------------------- device path
    @torch.no_grad()
    def caption_preprocessed(self, x_bf16: torch.Tensor, keys: Optional[torch.Tensor], seq_len: int = 30,
                             min_seq_len: int = 5, mode: int = ops.MODE_FULL_MEAN_1MINUS,
                             stage_mark=None) -> Dict[str, torch.Tensor]:
        """`stage_mark` (optional callable) is invoked after the image tower and after the decoder: bench.py records a
        CUDA event there to time the stages inside its step loop."""
        latent, embs = self.model._encode_image_dev(x_bf16)
        if stage_mark is not None:
            stage_mark()
        gen = self.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=seq_len,
                                  min_seq_len=min_seq_len, fixed_output_length=True, return_logits=False)
        if stage_mark is not None:
            stage_mark()
        out = {"tokens": gen["text"], "latent": latent}
        if keys is not None:
            k, s, c = ops.instance_disagreement(latent, keys, mode)
            out.update(inst_keys=k, inst_scores=s, inst_counts=c)
        return out

    @torch.no_grad()
    def caption_frames(self, frames: torch.Tensor, rects: torch.Tensor, frame_idx: torch.Tensor,
                       flags: Optional[torch.Tensor], keys: Optional[torch.Tensor], seq_len: int = 30,
                       min_seq_len: int = 5, mode: int = ops.MODE_FULL_MEAN_1MINUS):
        """frames uint8 [F,H,W,3], rects int32 [N,4], frame_idx int32 [N] (all on device).  N = 0 (a frame without
        detections: the reference's per-box loops simply do not run, goal_exploration.py:95-105) returns empty tensors
        without touching the GPU path."""
        if rects.shape[0] == 0:
            dev = self.device
            out = {"tokens": torch.zeros((0, seq_len), dtype=torch.int64, device=dev),
                   "latent": torch.zeros((0, self.cfg.embed_dim), dtype=torch.float32, device=dev),
                   "status": torch.zeros((0,), dtype=torch.int32, device=dev)}
            if keys is not None:
                out.update(inst_keys=torch.zeros((0,), dtype=torch.int64, device=dev),
                           inst_scores=torch.zeros((0,), dtype=torch.float32, device=dev),
                           inst_counts=torch.zeros((0,), dtype=torch.int32, device=dev))
            return out
        x, _, status = ops.preprocess_crops(frames, rects, frame_idx, flags, size=self.cfg.image_size)
        out = self.caption_preprocessed(x, keys, seq_len, min_seq_len, mode)
        out["status"] = status
        return out

    @torch.no_grad()
    def caption_boxes(self, frames: torch.Tensor, boxes, frame_idx, keys=None, policy: int = POLICY_AGENT,
                      seq_len: int = 30):
        rects, flags = box_rects(boxes, frames.shape[1], frames.shape[2], policy)
        dev = self.device
        k = None if keys is None else torch.as_tensor(np.asarray(keys, dtype=np.int64)).to(dev)
        return self.caption_frames(frames.to(dev), torch.from_numpy(rects).to(dev),
                                   torch.as_tensor(np.asarray(frame_idx, dtype=np.int32)).to(dev),
                                   torch.from_numpy(flags).to(dev), k, seq_len)

    @torch.no_grad()
    def caption_crops(self, crops: torch.Tensor, keys: Optional[torch.Tensor], seq_len: int = 30):
        """crops uint8 [N,h,w,3] on device (each crop is its own frame)."""
        n, h, w, _ = crops.shape
        rects = torch.tensor([0, 0, w, h], dtype=torch.int32, device=crops.device).repeat(n, 1)
        fi = torch.arange(n, dtype=torch.int32, device=crops.device)
        return self.caption_frames(crops, rects, fi, None, keys, seq_len)

    # -------------------------------------------------------------------------------------------- host path
    def _pin(self, tag: str, like: torch.Tensor) -> torch.Tensor:
        cur = self._pinned.get(tag)
        if cur is None or cur.shape != like.shape or cur.dtype != like.dtype:
            cur = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            self._pinned[tag] = cur
        return cur

    @torch.no_grad()
    def caption_crops_host(self, crops_pinned: torch.Tensor, keys_pinned: torch.Tensor, seq_len: int = 30):
        """End-to-end call with HOST buffers: H2D of the uint8 crops + keys, the full device path, D2H of the
        token matrix and the per-instance scores.  Returns host tensors and the byte counts moved."""
        dev = self.device
        crops = crops_pinned.to(dev, non_blocking=True)
        keys = keys_pinned.to(dev, non_blocking=True)
        out = self.caption_crops(crops, keys, seq_len)
        tok = self._pin("tok", out["tokens"])
        tok.copy_(out["tokens"], non_blocking=True)
        sc = self._pin("sc", out["inst_scores"])
        sc.copy_(out["inst_scores"], non_blocking=True)
        ik = self._pin("ik", out["inst_keys"])
        ik.copy_(out["inst_keys"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h2d = crops_pinned.numel() * crops_pinned.element_size() + keys_pinned.numel() * 8
        d2h = tok.numel() * 8 + sc.numel() * 4 + ik.numel() * 8
        return {"tokens": tok, "inst_scores": sc, "inst_keys": ik, "h2d_bytes": h2d, "d2h_bytes": d2h}
