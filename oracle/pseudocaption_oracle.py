"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the caption grouping / frequency table of the pseudo-captioner.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.

Follows experimenting_env/captioner/pseudocaptioner.py:
  * filter_captions              :96-123   (substring test, case-insensitive)
  * group_captions               :125-150  (key = (id_episode, id_object); kept detections only; dict insertion order)
  * compute_captions_frequency   :156-177  ([[freq, caption], ...] in first-seen order)
PARITY UNPINNED by reference tests (there are none for these functions).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple


def filter_captions(caption: str, banned_words: Sequence[str]) -> bool:
    if not any(banned_word.lower() in caption.lower() for banned_word in banned_words):
        return True
    return False


def captions_frequency(captions: Sequence[str], episode_ids, object_ids, banned_words: Sequence[str] = (),
                       apply_filter: bool = True) -> Dict[Tuple[int, int], List[list]]:
    grouped: Dict[Tuple[int, int], List[str]] = {}
    for cap, e, o in zip(captions, episode_ids, object_ids):
        key = (int(e), int(o))
        if (not apply_filter) or filter_captions(cap, banned_words):
            if key not in grouped:
                grouped[key] = [cap]
            else:
                grouped[key].append(cap)
    out = {}
    for key in grouped:
        freq: Dict[str, int] = {}
        for cap in grouped[key]:
            freq[cap] = freq.get(cap, 0) + 1
        out[key] = [[f, c] for c, f in freq.items()]
    return out
