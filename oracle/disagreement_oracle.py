"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-instance multi-view disagreement.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.

Follows:
  * `_cosine_distance(mat[D,V])`        experimenting_env/utils/projection_utils.py:928-942
  * dict-comprehension over objects     experimenting_env/utils/projection_utils.py:743-754
  * duplicate-view removal              experimenting_env/utils/projection_utils.py:574-578 (`unique(dim=0)`)
  * `mean_cosine_similarity`            scripts/compute_cosine_sim.py:11-22 (pairs i<j, scipy cosine, NaN if <2)
  * grouping key (episode_id, object_id) experimenting_env/captioner/pseudocaptioner.py:137,
                                        scripts/compute_cosine_sim.py:125-138,159 (string-cast, lexicographic)

PARITY UNPINNED by reference tests (none record expected values: test_cosine_similarity.py:12-19,86-93
only defines the pairs-only convention).  Pinned here by the algebraic identity
A11 = (V-1)/V * (1 - A12) and by scipy's `cosine` in the test-suite.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def cosine_distance(mat: torch.Tensor) -> torch.Tensor:
    """projection_utils.py:928-942 verbatim semantics; `mat` is [D, V] (columns are views)."""
    if len(mat) == 1:
        return torch.tensor(0)
    norm = (mat * mat).sum(0, keepdims=True) ** 0.5
    return (1 - (mat.T @ mat) / norm / norm.T).mean()


def mean_cosine_similarity_embeddings(emb: np.ndarray) -> float:
    """compute_cosine_sim.py:11-22 on already-encoded embeddings [V, D] (float64 accumulation like scipy)."""
    n = len(emb)
    if n < 2:
        return float("nan")
    e = np.asarray(emb, dtype=np.float64)
    sims = []
    for i in range(n):
        for j in range(i + 1, n):
            uv = np.dot(e[i], e[j])
            sims.append(uv / np.sqrt(np.dot(e[i], e[i]) * np.dot(e[j], e[j])))
    return float(np.mean(sims)) if sims else float("nan")


def group_indices(keys: np.ndarray) -> "OrderedDict[int, np.ndarray]":
    """instance key -> ascending crop indices, groups in ascending integer key order."""
    keys = np.asarray(keys, dtype=np.int64)
    order = np.argsort(keys, kind="stable")
    out: "OrderedDict[int, np.ndarray]" = OrderedDict()
    sk = keys[order]
    starts = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1]])
    ends = np.r_[starts[1:], len(sk)]
    for s, e in zip(starts, ends):
        out[int(sk[s])] = order[s:e]
    return out


def instance_disagreement(emb: np.ndarray, keys: np.ndarray, mode: str = "full", dedup: bool = False):
    """Per-instance scores in ascending key order -> (keys[M] int64, score[M] float32/64, count[M] int32).

    mode="full": A11 (`_cosine_distance`, mean over all VxV of 1 - cos, float32 torch arithmetic).
    mode="pairs": A12 (mean over i<j of cos; NaN when fewer than two views).
    dedup: apply `unique(dim=0)` (exact duplicate rows removed, rows sorted) first, as
           `SemanticPointCloud.preprocess` does (projection_utils.py:574-578).
    """
    groups = group_indices(keys)
    ks, sc, ct = [], [], []
    for k, idx in groups.items():
        e = torch.from_numpy(np.asarray(emb[idx], dtype=np.float32))
        if dedup:
            e = e.unique(dim=0)
        ks.append(k)
        ct.append(e.shape[0])
        if mode == "full":
            sc.append(float(cosine_distance(e.T.contiguous())))
        else:
            sc.append(mean_cosine_similarity_embeddings(e.numpy()))
    return np.array(ks, dtype=np.int64), np.array(sc, dtype=np.float64), np.array(ct, dtype=np.int32)


def string_key_group_order(episode_ids, object_ids):
    """pandas `groupby(["episode_id","object_id"])` on string-cast ids (compute_cosine_sim.py:125-138,159):
    returns the unique (episode, object) string pairs in lexicographic order."""
    pairs = sorted({(str(e), str(o)) for e, o in zip(episode_ids, object_ids)})
    return pairs
