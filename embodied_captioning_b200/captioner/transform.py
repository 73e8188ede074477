"""Image transform of the captioner as a GPU gather kernel.

Replaces upstream `image_transform_v2(PreprocessCfg, is_train=False)` built at
`experimenting_env/captioner/models/coca/factory.py:420-430` and called at `coca.py:28`
(Resize(224, BICUBIC, shortest edge) -> CenterCrop(224) -> ToTensor -> Normalize).  The fp32 output is
bit-identical to Pillow + torchvision (tests/test_preprocess_gpu.py); box policies are
integer host arithmetic restated from the reference call sites.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import _native as nat
from .. import ops

POLICY_AGENT = 0            # goal_exploration.py:95-99, frontier_exploration.py:133-137
POLICY_PSEUDOLABELER = 1    # detector/pseudolabeler.py:629-643,670-675
POLICY_EVAL = 2             # scripts/evaluate_finetuned_model.py:33-37
POLICY_PSEUDOCAPTIONER = 3  # captioner/pseudocaptioner.py:179-191,345-350


def box_rects(boxes, frame_h: int, frame_w: int, policy: int):
    """float xyxy boxes [N,4] -> (int32 rects [N,4], int32 flags [N]) following the reference call sites.

    numpy-slice policies (agent, pseudocaptioner) clamp the rectangle to the frame; PIL-crop policies
    (pseudolabeler, eval) keep it and the kernel reads zeros outside the frame.  flags bit0 = BGR->RGB.
    """
    b = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    n = b.shape[0]
    rects = np.zeros((n, 4), dtype=np.int64)
    flags = np.zeros((n,), dtype=np.int32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]

    def clamp(r):
        r[:, 0] = np.clip(r[:, 0], 0, frame_w); r[:, 2] = np.clip(r[:, 2], 0, frame_w)
        r[:, 1] = np.clip(r[:, 1], 0, frame_h); r[:, 3] = np.clip(r[:, 3], 0, frame_h)
        return r

    if policy == POLICY_AGENT:
        rects[:] = np.trunc(b).astype(np.int64)                       # int(x.item())
        rects = clamp(rects)
    elif policy == POLICY_PSEUDOLABELER:
        ef = np.float32(0.2)                                          # torch fp32 arithmetic
        width, height = frame_h, frame_w                              # image.shape[0], image.shape[1] (sic)
        bw, bh = (x2 - x1).astype(np.float32), (y2 - y1).astype(np.float32)
        nx1 = (x1 - ef * bw).astype(np.float32); ny1 = (y1 - ef * bh).astype(np.float32)
        nx2 = (x2 + ef * bw).astype(np.float32); ny2 = (y2 + ef * bh).astype(np.float32)
        rects[:, 0] = np.trunc(np.maximum(nx1, 0)); rects[:, 1] = np.trunc(np.maximum(ny1, 0))
        rects[:, 2] = np.trunc(np.minimum(nx2, width)); rects[:, 3] = np.trunc(np.minimum(ny2, height))
        flags[:] = 1
    elif policy == POLICY_PSEUDOCAPTIONER:
        d = b.astype(np.float64)                                      # python float arithmetic
        bw, bh = d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]
        rects[:, 0] = np.trunc(np.maximum(d[:, 0] - 0.1 * bw, 0)); rects[:, 1] = np.trunc(np.maximum(d[:, 1] - 0.1 * bh, 0))
        rects[:, 2] = np.trunc(np.minimum(d[:, 2] + 0.1 * bw, 1280)); rects[:, 3] = np.trunc(np.minimum(d[:, 3] + 0.1 * bh, 1280))
        rects = clamp(rects)
        flags[:] = 1
    elif policy == POLICY_EVAL:
        # The CSV boxes may be floats: the +-10 expansion and the boundary tests run on the raw values (python float
        # arithmetic), and PIL's `Image.crop` rounds each coordinate half-to-even (`int(round(x))`), not truncates.
        d = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
        e = np.empty_like(d)
        e[:, 0] = np.where(d[:, 0] - 10 >= 0, d[:, 0] - 10, 0)
        e[:, 1] = np.where(d[:, 1] - 10 >= 0, d[:, 1] - 10, d[:, 1])
        e[:, 2] = np.where(d[:, 2] + 10 <= frame_w, d[:, 2] + 10, d[:, 2])
        e[:, 3] = np.where(d[:, 3] + 10 <= frame_h, d[:, 3] + 10, d[:, 3])
        rects[:] = np.rint(e).astype(np.int64)                        # np.rint = round-half-even = python round()
    else:
        raise ValueError(f"unknown crop policy {policy}")
    return rects.astype(np.int32), flags


def preprocess_boxes(frames: torch.Tensor, boxes, frame_idx, policy: int = POLICY_AGENT, size: int = 224,
                     want_f32: bool = False):
    """frames uint8 [F,H,W,3] (device), boxes float [N,4], frame_idx int [N] -> bf16 [N,3,S,S] (and f32), status."""
    dev = frames.device
    rects, flags = box_rects(boxes, frames.shape[1], frames.shape[2], policy)
    r = torch.from_numpy(rects).to(dev)
    f = torch.from_numpy(flags).to(dev)
    fi = torch.as_tensor(np.asarray(frame_idx, dtype=np.int32)).to(dev)
    return ops.preprocess_crops(frames, r, fi, f, size=size, want_f32=want_f32)


class ImageTransform:
    """Callable drop-in for `preprocess_val`: PIL.Image / HWC uint8 array -> f32 tensor [3, S, S]."""

    def __init__(self, size: int = 224, device="cuda:0"):
        self.size = size
        self.device = nat.cuda_device(device)

    def _to_u8(self, img) -> np.ndarray:
        if hasattr(img, "convert"):                 # PIL.Image
            img = np.asarray(img.convert("RGB"))
        arr = np.ascontiguousarray(np.asarray(img))
        if arr.ndim != 3 or arr.shape[2] not in (3, 4) or arr.dtype != np.uint8:
            raise TypeError("expected a PIL image or an HxWx3 uint8 array")
        return arr[:, :, :3] if arr.shape[2] == 4 else arr

    def batch(self, images: Sequence, want_f32: bool = True) -> torch.Tensor:
        """Images of arbitrary sizes -> [N,3,S,S]; same-sized images share one launch."""
        arrs = [self._to_u8(im) for im in images]
        out = torch.empty((len(arrs), 3, self.size, self.size), device=self.device,
                          dtype=torch.float32 if want_f32 else torch.bfloat16)
        by_shape = {}
        for i, a in enumerate(arrs):
            by_shape.setdefault(a.shape, []).append(i)
        for shp, idxs in by_shape.items():
            if shp[0] == 0 or shp[1] == 0:
                raise ValueError("empty crop (zero height or width)")
            fr = torch.from_numpy(np.stack([arrs[i] for i in idxs])).to(self.device)
            n = len(idxs)
            rects = torch.tensor([[0, 0, shp[1], shp[0]]] * n, dtype=torch.int32, device=self.device)
            fi = torch.arange(n, dtype=torch.int32, device=self.device)
            b16, f32, status = ops.preprocess_crops(fr, rects, fi, None, size=self.size, want_f32=want_f32,
                                                    want_bf16=not want_f32)
            if int(status.max().item()) != 0:
                raise ValueError("crop outside the supported down-scale range (short edge > 11.5 x output size)")
            out[torch.as_tensor(idxs, device=self.device)] = f32 if want_f32 else b16
        return out

    def __call__(self, img) -> torch.Tensor:
        return self.batch([img])[0]
