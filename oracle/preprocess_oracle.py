"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the crop + image-transform path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.

Follows, for crop indexing (integer, bit-exact):
  * agent policy          experimenting_env/agents/goal_exploration/goal_exploration.py:95-99
                          (= agents/frontier_exploration/frontier_exploration.py:133-137)
  * pseudolabeler policy  experimenting_env/detector/pseudolabeler.py:629-643,670-675
  * evaluation policy     scripts/evaluate_finetuned_model.py:33-37
  * pseudocaptioner       experimenting_env/captioner/pseudocaptioner.py:179-191,345-350
and for the transform (built at captioner/models/coca/factory.py:420-430, called coca.py:28) the
upstream open_clip `image_transform_v2(is_train=False)` = torchvision Resize(224, BICUBIC) on the
shortest edge -> CenterCrop(224) -> ToTensor -> Normalize(OPENAI mean/std), whose resize is Pillow's
two-pass antialiased bicubic with 8-bit intermediates and 22-bit fixed-point coefficients
(SURVEY.md Appendix B).  PIN: `tests/test_preprocess_oracle.py` checks this file bit-for-bit against
the Pillow / torchvision installed in the build container on random shapes, and the outputs are
committed under tests/golden/ so the GPU box (which may lack nothing but needs no Pillow) can use them.
"""
from __future__ import annotations

import math

import numpy as np

POLICY_AGENT = 0
POLICY_PSEUDOLABELER = 1
POLICY_EVAL = 2
POLICY_PSEUDOCAPTIONER = 3

MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
PRECISION_BITS = 22


# --------------------------------------------------------------------------- A1 crop boxes
def crop_rect(policy: int, box, frame_h: int, frame_w: int):
    """Integer source rectangle (x1, y1, x2, y2) + (pad_outside, bgr_to_rgb) for one detection box.

    `pad_outside` True means PIL `Image.crop` semantics (pixels outside the frame read as 0);
    False means numpy slicing semantics (the rectangle is clamped to the frame).
    """
    f32 = np.float32
    if policy != POLICY_EVAL:
        x1, y1, x2, y2 = (f32(v) for v in box)
    if policy == POLICY_AGENT:
        r = [int(x1), int(y1), int(x2), int(y2)]          # int() truncation, then rgb[y1:y2, x1:x2]
        return _clamp(r, frame_h, frame_w), False, False
    if policy in (POLICY_PSEUDOLABELER, POLICY_PSEUDOCAPTIONER):
        if policy == POLICY_PSEUDOLABELER:
            ef = f32(0.2)                                  # torch fp32 tensor arithmetic
            width, height = frame_h, frame_w               # image.shape[0], image.shape[1] (sic)
            bw, bh = f32(x2 - x1), f32(y2 - y1)
            nx1 = f32(x1 - f32(ef * bw)); ny1 = f32(y1 - f32(ef * bh))
            nx2 = f32(x2 + f32(ef * bw)); ny2 = f32(y2 + f32(ef * bh))
        else:
            ef = 0.1                                       # python float arithmetic on the stored box
            width, height = 1280, 1280
            x1, y1, x2, y2 = (float(v) for v in (x1, y1, x2, y2))
            bw, bh = x2 - x1, y2 - y1
            nx1 = x1 - ef * bw; ny1 = y1 - ef * bh
            nx2 = x2 + ef * bw; ny2 = y2 + ef * bh
        r = [int(max(nx1, 0)), int(max(ny1, 0)), int(min(nx2, width)), int(min(ny2, height))]
        if policy == POLICY_PSEUDOLABELER:
            return r, True, True                           # PIL crop of the BGR->RGB frame
        return _clamp(r, frame_h, frame_w), False, True    # numpy slice, then BGR->RGB
    if policy == POLICY_EVAL:
        # evaluate_finetuned_model.py:33-37 on the raw (possibly float) CSV values, then `Image.crop`, which rounds
        # each coordinate with `int(round(x))` (Pillow Image.crop: `map(int, map(round, box))`, half-to-even)
        b = [float(v) if not isinstance(v, (int, np.integer)) else int(v) for v in box]
        e = [b[0] - 10 if (b[0] - 10) >= 0 else 0,
             b[1] - 10 if (b[1] - 10) >= 0 else b[1],
             b[2] + 10 if (b[2] + 10) <= frame_w else b[2],
             b[3] + 10 if (b[3] + 10) <= frame_h else b[3]]
        return [int(round(v)) for v in e], True, False
    raise ValueError(f"unknown crop policy {policy}")


def _clamp(r, h, w):
    """numpy basic-slice semantics for non-negative bounds."""
    x1, y1, x2, y2 = r
    x1 = min(max(x1, 0), w); x2 = min(max(x2, 0), w)
    y1 = min(max(y1, 0), h); y2 = min(max(y2, 0), h)
    return [x1, y1, x2, y2]


def extract_crop(frame: np.ndarray, rect, pad_outside: bool, bgr_to_rgb: bool) -> np.ndarray:
    x1, y1, x2, y2 = rect
    H, W = frame.shape[:2]
    out = np.zeros((max(y2 - y1, 0), max(x2 - x1, 0), 3), dtype=np.uint8)
    sx1, sy1, sx2, sy2 = max(x1, 0), max(y1, 0), min(x2, W), min(y2, H)
    if sx2 > sx1 and sy2 > sy1:
        out[sy1 - y1:sy2 - y1, sx1 - x1:sx2 - x1] = frame[sy1:sy2, sx1:sx2, :3]
    if bgr_to_rgb:
        out = out[:, :, ::-1]
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------- A2 transform
def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc`: per output index (xmin, int32 taps)."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    bounds, taps = [], []
    ss = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in k:
            ww += v
        if ww != 0.0:
            k = [v / ww for v in k]
        ki = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k]
        bounds.append(xmin)
        taps.append(np.array(ki, dtype=np.int64))
    return bounds, taps


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    in_size = img.shape[axis]
    bounds, taps = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for o in range(out_size):
        k = taps[o]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(k, src[bounds[o]:bounds[o] + len(k)], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resized_size(w: int, h: int, size: int = 224):
    """torchvision `_compute_resized_output_size` for an int size: shortest edge -> size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_short, new_long) if w <= h else (new_long, new_short)   # (new_w, new_h)


def resize_pil_exact(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    h, w = img.shape[:2]
    if w != new_w:                       # horizontal pass first, 8-bit intermediate
        img = _resample_axis(img, new_w, 1)
    if h != new_h:
        img = _resample_axis(img, new_h, 0)
    return img


def center_crop_offsets(w: int, h: int, size: int = 224):
    """torchvision center_crop: int(round((dim - size) / 2.0)) with Python's round-half-even."""
    return int(round((w - size) / 2.0)), int(round((h - size) / 2.0))    # (left, top)


def transform_u8(crop: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 [h,w,3] -> uint8 [size,size,3] after Resize(shortest edge) + CenterCrop."""
    h, w = crop.shape[:2]
    nw, nh = resized_size(w, h, size)
    img = resize_pil_exact(crop, nw, nh)
    left, top = center_crop_offsets(nw, nh, size)
    return img[top:top + size, left:left + size]


def normalize_f32(img_u8: np.ndarray) -> np.ndarray:
    """ToTensor + Normalize in fp32: ((u8/255) - mean) / std, CHW."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)


def transform(crop: np.ndarray, size: int = 224) -> np.ndarray:
    return normalize_f32(transform_u8(crop, size))


def preprocess_boxes(frames, boxes, frame_idx, policy: int, size: int = 224):
    """frames uint8 [F,H,W,3] (or list), boxes float [N,4] -> (f32 [N,3,size,size], int32 rects [N,4])."""
    out = np.empty((len(boxes), 3, size, size), dtype=np.float32)
    rects = np.empty((len(boxes), 4), dtype=np.int32)
    for i, (b, fi) in enumerate(zip(boxes, frame_idx)):
        fr = frames[int(fi)]
        rect, pad, swap = crop_rect(policy, b, fr.shape[0], fr.shape[1])
        rects[i] = rect
        out[i] = transform(extract_crop(fr, rect, pad, swap), size)
    return out, rects
