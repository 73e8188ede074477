"""Consensus pseudo-label preparation (captioner/pseudocaptioner.py of the reference): caption filter, grouping by
(episode_id, object_id) and the per-instance caption frequency table that feeds the LLM prompt.

  * `filter_captions`             pseudocaptioner.py:96-123   (case-insensitive SUBSTRING test against a word list)
  * `group key`                   pseudocaptioner.py:137      ((id_episode, id_object))
  * `compute_captions_frequency`  pseudocaptioner.py:156-177  ({key: [[freq, caption], ...]}, first-seen order)

The histogram runs on the device over the decoder's token matrix (`ops.caption_frequency`): two captions are the same
string iff their token rows are the same (greedy decode pads after EOS), so the string table is obtained by decoding
one representative row per distinct caption instead of every crop.  The LLM call itself stays out of scope.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy as np
import torch

from .. import ops
from ..disagreement import make_keys, split_keys

# word list of pseudocaptioner.py:98-119 (data, reproduced so that the filter decides identically)
BANNED_WORDS = (
    "person man woman boy girl child children adult kid baby human people group crowd "
    "dog cat bird fish horse animal pet elephant lion tiger monkey mouse rabbit cow "
    "pig sheep deer bear chicken duck goat camel snake frog turtle whale dolphin insect bug spider "
    "blurry picture image photo portrait painting drawing sketch screenshot artwork filter 3d rendering "
    "thing stuff object item something stuff device equipment material machine gadget "
    "unknown unidentified indistinguishable"
).split() + ["living room", "kitchen", "bedroom", "bathroom", "dining room", "living room", "room"] + (
    "car vehicle bike truck street road tree forest mountain park outdoor sky landscape scenery "
    "running jumping walking talking playing sitting standing moving holding eating drinking flying swimming driving"
).split()


def filter_captions(caption: str) -> bool:
    """True when the caption contains none of the banned words as a substring (so "cat" also bans "location")."""
    low = caption.lower()
    return not any(w.lower() in low for w in BANNED_WORDS)


def compute_captions_frequency(tokens: torch.Tensor, episode_ids, object_ids, decode: Callable[[np.ndarray], str],
                               apply_filter: bool = True) -> Dict[Tuple[int, int], List[list]]:
    """tokens int [N,T] on the device (one decoded caption per detection, in detection order); ids [N].
    -> {(episode_id, object_id): [[freq, caption], ...]}: keys in first-seen order, captions in first-seen order,
    exactly like grouping the kept detections and counting strings (pseudocaptioner.py:125-177)."""
    keys = make_keys(episode_ids, object_ids)
    k, first, count = ops.caption_frequency(tokens, torch.from_numpy(keys).to(tokens.device))
    k, first, count = k.cpu().numpy(), first.cpu().numpy(), count.cpu().numpy()
    rep = tokens[torch.from_numpy(first.astype(np.int64)).to(tokens.device)].cpu().numpy() if len(first) else []
    # distinct token rows may still decode to the same string (e.g. different ids rendering alike): merge on strings
    per_key: Dict[int, Dict[str, list]] = {}
    key_first: Dict[int, int] = {}
    for kk, f, c, row in zip(k, first, count, rep):
        text = decode(row)
        if apply_filter and not filter_captions(text):
            continue
        d = per_key.setdefault(int(kk), {})
        if text in d:
            d[text][0] += int(c); d[text][2] = min(d[text][2], int(f))
        else:
            d[text] = [int(c), text, int(f)]
        key_first[int(kk)] = min(key_first.get(int(kk), 1 << 62), int(f))
    out: Dict[Tuple[int, int], List[list]] = {}
    for kk in sorted(per_key, key=lambda q: key_first[q]):             # dict insertion order of grouped_captions
        e, o = split_keys(np.array([kk]))
        ent = sorted(per_key[kk].values(), key=lambda v: v[2])         # first-seen order inside the instance
        out[(int(e[0]), int(o[0]))] = [[v[0], v[1]] for v in ent]
    return out
