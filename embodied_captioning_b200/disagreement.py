"""Per-instance multi-view caption disagreement with the reference's function names.

  * `_cosine_distance(mat[D,V]) -> 0-d tensor`        experimenting_env/utils/projection_utils.py:928-942
    (pluggable as `SemanticPointCloud.distance_f`, projection_utils.py:330-332)
  * `mean_cosine_similarity(captions | embeddings)`   scripts/compute_cosine_sim.py:11-22
  * `instance_disagreement(emb[N,D], keys[N])`        batched form of the dict comprehension at
    projection_utils.py:743-754 / the pandas groupby at compute_cosine_sim.py:159-179

All arithmetic runs in libec_b200 (`csrc/disagreement.cu`); there is no CPU path.  The caption-text
encoder (SBERT all-MiniLM-L6-v2 in the reference: goal_exploration.py:57,102-103) is host-side and
injectable through `set_encoder` -- it is not one of the CUDA subsystems.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import ops
from .ops import MODE_FULL_MEAN_1MINUS, MODE_PAIRS_MEAN_COS

model: Optional[Callable] = None      # module-global encoder, same name as compute_cosine_sim.py:156


def set_encoder(encoder) -> None:
    """`encoder.encode(list[str], convert_to_numpy=True) -> [V, D]` (sentence-transformers protocol)."""
    global model
    model = encoder


def _device(device=None) -> torch.device:
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _cosine_distance(mat: torch.Tensor) -> torch.Tensor:
    """mean over all V x V entries of 1 - cos(view_i, view_j); `mat` is [D, V] (columns are views)."""
    if len(mat) == 1:                       # quirk kept: tests D, not V (projection_utils.py:939-940)
        return torch.tensor(0)
    dev = mat.device if mat.is_cuda else _device()
    emb = mat.detach().to(dev, torch.float32).t().contiguous()
    keys = torch.zeros((emb.shape[0],), dtype=torch.int64, device=dev)
    _, score, _ = ops.instance_disagreement(emb, keys, MODE_FULL_MEAN_1MINUS)
    out = score[0]
    return out if mat.is_cuda else out.cpu()


def mean_cosine_similarity(captions_list, device=None) -> float:
    """Mean over i<j of cos(e_i, e_j); NaN when fewer than two captions.  Accepts a list of strings
    (encoded with the injected `model`) or an array/tensor of embeddings [V, D]."""
    if len(captions_list) < 2:
        return float("nan")
    if isinstance(captions_list, (np.ndarray, torch.Tensor)):
        emb = torch.as_tensor(captions_list)
    else:
        if model is None:
            raise RuntimeError("no caption encoder set: call disagreement.set_encoder(SentenceTransformer(...))")
        emb = torch.as_tensor(model.encode(list(captions_list), convert_to_numpy=True))
    dev = emb.device if emb.is_cuda else _device(device)
    emb = emb.to(dev, torch.float32).contiguous()
    keys = torch.zeros((emb.shape[0],), dtype=torch.int64, device=dev)
    _, score, _ = ops.instance_disagreement(emb, keys, MODE_PAIRS_MEAN_COS)
    return float(score[0].item())


def make_keys(episode_ids: Sequence[int], object_ids: Sequence[int]) -> np.ndarray:
    """(episode_id, object_id) -> one int64 key (pseudocaptioner.py:137); both ids must be in [0, 2^31)."""
    e = np.asarray(episode_ids, dtype=np.int64)
    o = np.asarray(object_ids, dtype=np.int64)
    if (e < 0).any() or (o < 0).any() or (e >= 2 ** 31).any() or (o >= 2 ** 31).any():
        raise ValueError("episode/object ids must fit in 31 bits")
    return (e << 32) | o


def split_keys(keys: np.ndarray):
    k = np.asarray(keys, dtype=np.int64)
    return k >> 32, k & 0xFFFFFFFF


def pandas_group_order(keys: np.ndarray) -> np.ndarray:
    """Permutation putting key-sorted groups into the order of
    `df.groupby(["episode_id","object_id"])` on string-cast ids (compute_cosine_sim.py:125-138,159)."""
    e, o = split_keys(keys)
    labels = [(str(int(a)), str(int(b))) for a, b in zip(e, o)]
    return np.array(sorted(range(len(labels)), key=lambda i: labels[i]), dtype=np.int64)


def instance_disagreement(emb: torch.Tensor, keys: torch.Tensor, mode: int = MODE_FULL_MEAN_1MINUS,
                          dedup: bool = False):
    """emb [N,D] f32, keys [N] int64 (device tensors) -> (keys[M], score[M], count[M]) sorted by key.
    dedup=True applies the world model's `unique(dim=0)` per instance first (projection_utils.py:574-578)."""
    return ops.instance_disagreement(emb.contiguous(), keys.contiguous(), mode, dedup)


def map_geometry(lower_bound, upper_bound, map_scale: float = 0.025):
    """Grid size and integer origin of the top-down map (projection_utils.py:683-699) -> (H, W, off_x, off_y)."""
    H = int(abs(upper_bound[2] - lower_bound[2]) / map_scale)
    W = int(abs(upper_bound[0] - lower_bound[0]) / map_scale)
    return H, W, int(np.floor(lower_bound[0] / map_scale)), int(np.ceil(lower_bound[2] / map_scale))


def disagreement_map(points: torch.Tensor, point_keys: torch.Tensor, inst_keys: torch.Tensor,
                     inst_scores: torch.Tensor, lower_bound, upper_bound, map_scale: float = 0.025):
    """Top-down disagreement map + exploration reward on the device.

    Replaces the score gather / centre write / six inflate writes of `get_topdown_semantic`
    (projection_utils.py:760-785) and `get_reward` (env_habitat3.py:75-84).  points f32 [P,3], point_keys int64 [P]
    (object id of every 3-D point), (inst_keys, inst_scores) the key-sorted result of `instance_disagreement`.
    Returns (map f32 [H,W], reward 0-d f32 tensor), both on the device."""
    from . import _native as nat
    H, W, off_x, off_y = map_geometry(lower_bound, upper_bound, map_scale)
    if H <= 0 or W <= 0:
        raise ValueError("empty map bounds")
    dev = points.device
    if not points.is_cuda:
        raise nat.NativeError("expected CUDA tensors (there is no CPU path)")
    pts = points.to(torch.float32).contiguous()
    pk = point_keys.to(torch.int64).contiguous()
    ik = inst_keys.to(torch.int64).contiguous()
    isc = inst_scores.to(torch.float32).contiguous()
    P, M = pts.shape[0], ik.shape[0]
    out = torch.empty((H, W), device=dev, dtype=torch.float32)
    reward = torch.zeros((), device=dev, dtype=torch.float32)
    lib = nat.lib()
    ws = torch.empty((lib.ec_disagreement_map_workspace_bytes(P, H, W),), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        nat.check(lib.ec_disagreement_map(nat.ptr(pts) if P else None, P, nat.ptr(pk) if P else None,
                                          nat.ptr(ik) if M else None, nat.ptr(isc) if M else None, M,
                                          float(np.float32(map_scale)), off_x, off_y, H, W, out.data_ptr(),
                                          reward.data_ptr(), ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream(dev).cuda_stream), "ec_disagreement_map")
    return out, reward


def group_mean(scores: torch.Tensor) -> float:
    """NaN-skipping mean over groups (pandas `.mean()` at compute_cosine_sim.py:189-195)."""
    s = scores[~torch.isnan(scores)]
    return float(s.double().mean().item()) if s.numel() else float("nan")
