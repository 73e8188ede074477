"""Deterministic synthetic weights and crops (there are no checkpoints or datasets offline).

The state dict uses upstream open_clip's CoCa tensor names and shapes (the
layout `factory.py:131-180 load_checkpoint` expects), so real checkpoints load through the same path.
Init scales follow upstream's `init_parameters`; every value is rounded to a
bf16-representable number so that the fp32 oracle and the bf16-operand CUDA path see *identical*
weights and only activation rounding differs.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import CoCaConfig


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def state_dict_spec(cfg: CoCaConfig) -> Dict[str, tuple]:
    """name -> shape of every tensor of the open_clip CoCa layout for `cfg`."""
    W, E, T, P = cfg.vision_width, cfg.embed_dim, cfg.text_width, cfg.patch_size
    spec: Dict[str, tuple] = {"logit_scale": ()}

    def block(prefix, width, cross=False):
        spec[prefix + "ln_1.weight"] = (width,)
        spec[prefix + "ln_1.bias"] = (width,)
        if cross:
            spec[prefix + "ln_1_kv.weight"] = (width,)
            spec[prefix + "ln_1_kv.bias"] = (width,)
        spec[prefix + "attn.in_proj_weight"] = (3 * width, width)
        spec[prefix + "attn.in_proj_bias"] = (3 * width,)
        spec[prefix + "attn.out_proj.weight"] = (width, width)
        spec[prefix + "attn.out_proj.bias"] = (width,)
        spec[prefix + "ln_2.weight"] = (width,)
        spec[prefix + "ln_2.bias"] = (width,)
        spec[prefix + "mlp.c_fc.weight"] = (4 * width, width)
        spec[prefix + "mlp.c_fc.bias"] = (4 * width,)
        spec[prefix + "mlp.c_proj.weight"] = (width, 4 * width)
        spec[prefix + "mlp.c_proj.bias"] = (width,)

    spec["visual.conv1.weight"] = (W, 3, P, P)
    spec["visual.class_embedding"] = (W,)
    spec["visual.positional_embedding"] = (cfg.vision_tokens, W)
    spec["visual.ln_pre.weight"] = (W,)
    spec["visual.ln_pre.bias"] = (W,)
    for i in range(cfg.vision_layers):
        block(f"visual.transformer.resblocks.{i}.", W)
    spec["visual.attn_pool.query"] = (cfg.n_queries, E)
    spec["visual.attn_pool.attn.q_proj_weight"] = (E, E)
    spec["visual.attn_pool.attn.k_proj_weight"] = (E, W)
    spec["visual.attn_pool.attn.v_proj_weight"] = (E, W)
    spec["visual.attn_pool.attn.in_proj_bias"] = (3 * E,)
    spec["visual.attn_pool.attn.out_proj.weight"] = (E, E)
    spec["visual.attn_pool.attn.out_proj.bias"] = (E,)
    spec["visual.attn_pool.ln_q.weight"] = (E,)
    spec["visual.attn_pool.ln_q.bias"] = (E,)
    spec["visual.attn_pool.ln_k.weight"] = (W,)
    spec["visual.attn_pool.ln_k.bias"] = (W,)
    spec["visual.ln_post.weight"] = (E,)
    spec["visual.ln_post.bias"] = (E,)
    spec["visual.proj"] = (E, E)

    spec["text.token_embedding.weight"] = (cfg.vocab_size, T)
    spec["text.cls_emb"] = (T,)
    spec["text.positional_embedding"] = (cfg.context_length + 1, T)
    for i in range(cfg.text_layers):
        block(f"text.transformer.resblocks.{i}.", T)
    spec["text.ln_final.weight"] = (T,)
    spec["text.ln_final.bias"] = (T,)
    spec["text.text_projection"] = (T, E)

    for i in range(cfg.mm_layers):
        block(f"text_decoder.resblocks.{i}.", T)
        block(f"text_decoder.cross_attn.{i}.", T, cross=True)
    spec["text_decoder.ln_final.weight"] = (T,)
    spec["text_decoder.ln_final.bias"] = (T,)
    spec["text_decoder.text_projection"] = (T, cfg.vocab_size)
    return spec


def count_params(cfg: CoCaConfig) -> int:
    return sum(int(np.prod(s)) if len(s) else 1 for s in state_dict_spec(cfg).values())


def make_state_dict(cfg: CoCaConfig, seed: int = 1234, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights in the open_clip layout, bf16-representable fp32 tensors.

    `gain` scales the linear-layer std; LayerNorm affine parameters are jittered around (1, 0) and
    biases get small noise so that no code path is trivially zero.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    spec = state_dict_spec(cfg)
    W, T = cfg.vision_width, cfg.text_width
    sd: Dict[str, torch.Tensor] = {}

    def randn(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    for name, shape in spec.items():
        width = W if name.startswith("visual.transformer") else T
        layers = cfg.vision_layers if name.startswith("visual.transformer") else cfg.text_layers
        leaf = name.split(".")[-1]
        if name == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        elif ".ln_" in name or name.split(".")[-2].startswith("ln_"):
            t = (1.0 + randn(shape, 0.05)) if leaf == "weight" else randn(shape, 0.02)
        elif name == "visual.conv1.weight":
            t = randn(shape, gain * (cfg.patch_dim ** -0.5))
        elif name in ("visual.class_embedding", "visual.positional_embedding"):
            t = randn(shape, W ** -0.5)
        elif name == "visual.proj":
            t = randn(shape, gain * cfg.embed_dim ** -0.5)
        elif name == "visual.attn_pool.query":
            t = randn(shape, 1.0)
        elif name.startswith("visual.attn_pool.attn."):
            if leaf.endswith("bias"):
                t = randn(shape, 0.02)
            else:
                t = randn(shape, gain * shape[-1] ** -0.5)
        elif name == "text.token_embedding.weight":
            t = randn(shape, 0.02)
        elif name == "text.positional_embedding":
            t = randn(shape, 0.01)
        elif name == "text.cls_emb":
            t = randn(shape, 0.01)
        elif name in ("text.text_projection", "text_decoder.text_projection"):
            t = randn(shape, gain * T ** -0.5)
        elif name.endswith("attn.in_proj_weight"):
            t = randn(shape, gain * width ** -0.5)
        elif name.endswith("attn.out_proj.weight") or name.endswith("mlp.c_proj.weight"):
            t = randn(shape, gain * (width ** -0.5) * ((2 * layers) ** -0.5))
        elif name.endswith("mlp.c_fc.weight"):
            t = randn(shape, gain * (2 * width) ** -0.5)
        elif leaf == "bias" or leaf.endswith("_bias"):
            t = randn(shape, 0.02)
        else:
            raise KeyError(f"no init rule for {name}")
        sd[name] = _bf16_round(t.reshape(shape))
    return sd


def make_crops(n: int, size: int = 224, seed: int = 5678) -> np.ndarray:
    """`uint8[n,size,size,3]` crops: low-frequency colour structure plus uniform noise
    (SURVEY.md section 8(d) "Synthetic inputs")."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, size, dtype=np.float32), np.linspace(0, 1, size, dtype=np.float32),
                         indexing="ij")
    out = np.empty((n, size, size, 3), dtype=np.uint8)
    for i in range(n):
        base = rng.uniform(40, 215, size=3).astype(np.float32)
        fx, fy = rng.uniform(0.5, 3.0, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        amp = rng.uniform(10, 60, size=3)
        img = np.stack([base[c] + amp[c] * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph[c]) for c in range(3)], -1)
        img += rng.uniform(-24, 24, size=img.shape)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def make_frames_and_boxes(n_frames: int, boxes_per_frame: int, height: int = 480, width: int = 640,
                          seed: int = 5678):
    """Episode-replay shaped input (BASELINE.json configs[2]): `uint8[F,H,W,3]` frames and float xyxy boxes
    with w,h in U[32,320] clipped to the frame; returns (frames, boxes[F*k,4] float32, frame_idx[F*k] int32,
    instance_key[F*k] int64) with instance id = box slot."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(n_frames, height, width, 3), dtype=np.uint8)
    # add a per-frame smooth gradient so resampling sees structure
    ramp = (np.linspace(0, 64, width, dtype=np.float32)[None, :, None] +
            np.linspace(0, 64, height, dtype=np.float32)[:, None, None])
    frames = np.clip(frames.astype(np.float32) * 0.5 + ramp, 0, 255).astype(np.uint8)
    n = n_frames * boxes_per_frame
    w = rng.uniform(32, 320, size=n)
    h = rng.uniform(32, 320, size=n)
    x1 = rng.uniform(0, width - 32, size=n)
    y1 = rng.uniform(0, height - 32, size=n)
    x2 = np.minimum(x1 + w, width)
    y2 = np.minimum(y1 + h, height)
    boxes = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
    frame_idx = np.repeat(np.arange(n_frames, dtype=np.int32), boxes_per_frame)
    keys = np.tile(np.arange(boxes_per_frame, dtype=np.int64), n_frames)
    return frames, boxes, frame_idx, keys


def make_instance_crops(n_instances: int, views: int, size: int = 224, seed: int = 5678, view_noise: float = 6.0):
    """Multi-view shaped input: `views` noisy, slightly shifted views of each of `n_instances` base crops.
    Returns (uint8 [n_instances*views, size, size, 3], instance ids int64 [n_instances*views])."""
    base = make_crops(n_instances, size=size + 16, seed=seed)
    rng = np.random.default_rng(seed + 1)
    out = np.empty((n_instances * views, size, size, 3), dtype=np.uint8)
    ids = np.repeat(np.arange(n_instances, dtype=np.int64), views)
    for i in range(n_instances):
        for v in range(views):
            dx, dy = rng.integers(0, 17, size=2)
            img = base[i, dy:dy + size, dx:dx + size].astype(np.float32)
            img += rng.normal(0, view_noise, size=img.shape)
            out[i * views + v] = np.clip(img, 0, 255).astype(np.uint8)
    return out, ids


def fit_confident_head(hidden: torch.Tensor, targets: torch.Tensor, vocab_size: int, margin: float = 12.0,
                       ridge: float = 1e-3) -> torch.Tensor:
    """Closed-form "training" of the vocabulary head on given decoder states.

    Random-init weights give almost flat 49k-way logits, so greedy arg-max is decided by numerical noise.  A
    trained captioner is confident; this emulates that cheaply: given final-LayerNorm states `hidden` [S, T]
    (one per (crop, position)) and the token each state should emit, `targets` [S], it solves the ridge
    regression  P = argmin |hidden P - margin * onehot(targets)|^2 + ridge |P|^2  and returns P [T, vocab]
    (a drop-in `text_decoder.text_projection`), rounded to bf16-representable values.
    """
    H = hidden.to(torch.float64)
    S, T = H.shape
    A = H.t() @ H
    A += ridge * float(A.diagonal().mean()) * torch.eye(T, dtype=torch.float64)
    # H^T Y with Y = margin * onehot: column v collects the states whose target is v
    HtY = torch.zeros((T, vocab_size), dtype=torch.float64)
    HtY.index_add_(1, targets.to(torch.long), (H * margin).t().contiguous())
    P = torch.linalg.solve(A, HtY)
    return _bf16_round(P.to(torch.float32))
