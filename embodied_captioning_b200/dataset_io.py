"""On-disk formats either side of the captioner hot path and the batched feeders built on them.

  * observation files  `episode_{:06d}_step_{:05d}_modality_{mod}_id_{cam}.npz`
        naming / parsing: experimenting_env/utils/sensors_utils.py:13-38, writer :41-65
  * detection files    `np.savez_compressed(path, {"instances": ..., "image": rgb})`, read back as
        `np.load(path, allow_pickle=True)["arr_0"].item()`   (detector/pseudolabeler.py:835-843,
        captioner/pseudocaptioner.py:132-144, scripts/evaluate_finetuned_model.py:29)
  * evaluation CSV     columns filename, episode_id, object_id, bounding_box, proposed_caption,
        reference_caption                                     (scripts/evaluate_finetuned_model.py:154-160)

The reference stores `instances` as a pickled detectron2 `Instances`; detectron2 is not a dependency here, so files
written by this module carry the same fields as a plain dict (`pred_boxes` float32 [n,4], `infos` list of dicts with
`id_episode` / `id_object`, `captions` list of str).  `load_detections` accepts both forms (the pickled one needs
detectron2 importable, exactly as in the reference).
"""
from __future__ import annotations

import ast
import csv
import os
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np

CSV_HEADER = ["filename", "episode_id", "object_id", "bounding_box", "proposed_caption", "reference_caption"]


# ------------------------------------------------------------------------------------------ observation files
@dataclass
class SenseInfo:
    base_path: str
    mod: str
    episode: int = 0
    camera_id: int = 0
    step: int = 0

    def get_path(self) -> str:
        return os.path.join(self.base_path, f"episode_{self.episode:06d}_step_{self.step:05d}_modality_{self.mod}"
                                            f"_id_{self.camera_id}.npz")


def _field(path: str, name: str) -> str:
    stem = os.path.split(os.path.splitext(path)[0])[1]
    return stem[stem.find(name):].split("_")[1]


def get_sense_info(path: str) -> SenseInfo:
    return SenseInfo(os.path.dirname(path), _field(path, "modality"), int(_field(path, "episode")),
                     int(_field(path, "id")), int(_field(path, "step")))


def save_obs(exp_path: str, episode_id: int, observations: Dict[str, np.ndarray], timestamp: int) -> List[str]:
    paths = []
    for camera_id, (modality, obs) in enumerate(observations.items()):
        p = SenseInfo(exp_path, modality, int(episode_id), int(camera_id), int(timestamp)).get_path()
        np.savez_compressed(p, obs)
        paths.append(p)
    return paths


def load_obs(path: str) -> np.ndarray:
    return np.load(path, allow_pickle=True)["arr_0"]


# ------------------------------------------------------------------------------------------ detection files
@dataclass
class Detections:
    filename: str
    image: np.ndarray            # uint8 [H,W,3] (BGR when written by the pseudolabeler, pseudolabeler.py:670-675)
    boxes: np.ndarray            # float32 [n,4] xyxy
    episode_ids: np.ndarray      # int64 [n]
    object_ids: np.ndarray       # int64 [n]
    captions: List[str]


def save_detections(path: str, image: np.ndarray, boxes, infos: Sequence[dict], captions: Sequence[str]) -> str:
    inst = {"pred_boxes": np.asarray(boxes, dtype=np.float32).reshape(-1, 4), "infos": list(infos),
            "captions": list(captions)}
    np.savez_compressed(path, {"instances": inst, "image": np.asarray(image)})
    return path if path.endswith(".npz") else path + ".npz"


def load_detections(path: str) -> Detections:
    d = np.load(path, allow_pickle=True)["arr_0"].item()
    inst, image = d["instances"], d["image"]
    if isinstance(inst, dict):
        boxes, infos, caps = inst["pred_boxes"], inst["infos"], inst.get("captions", [])
    else:                                     # detectron2 Instances (pseudocaptioner.py:136-143)
        boxes = inst.pred_boxes.tensor.numpy() if len(inst) else np.zeros((0, 4), np.float32)
        infos, caps = inst.infos, getattr(inst, "captions", [])
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    return Detections(path, np.asarray(image), boxes,
                      np.array([int(i["id_episode"]) for i in infos], dtype=np.int64),
                      np.array([int(i["id_object"]) for i in infos], dtype=np.int64), [str(c) for c in caps])


# ------------------------------------------------------------------------------------------ batched feeders
def iter_frame_batches(paths: Iterable[str], max_crops: int = 1024):
    """Group detection files into batches of same-sized frames holding at most `max_crops` boxes:
    yields (frames uint8 [F,H,W,3], boxes [N,4], frame_idx [N], episode_ids [N], object_ids [N], files [F])."""
    cur: Dict[tuple, list] = {}
    counts: Dict[tuple, int] = {}

    def flush(shape):
        dets = cur.pop(shape)
        counts.pop(shape)
        frames = np.stack([d.image[:, :, :3] for d in dets])
        boxes = np.concatenate([d.boxes for d in dets]) if dets else np.zeros((0, 4), np.float32)
        fidx = np.concatenate([np.full(len(d.boxes), i, np.int32) for i, d in enumerate(dets)])
        return (frames, boxes, fidx, np.concatenate([d.episode_ids for d in dets]),
                np.concatenate([d.object_ids for d in dets]), [d.filename for d in dets])

    for p in paths:
        d = load_detections(p)
        shape = d.image.shape[:2]
        if shape in cur and counts[shape] + len(d.boxes) > max_crops:
            yield flush(shape)
        cur.setdefault(shape, []).append(d)
        counts[shape] = counts.get(shape, 0) + len(d.boxes)
    for shape in list(cur.keys()):
        yield flush(shape)


def evaluate_csv(engine, test_csv: str, dest_csv: str, decode: Callable[[np.ndarray], str], seq_len: int = 30,
                 max_crops: int = 1024) -> int:
    """Batched `scripts/evaluate_finetuned_model.py:150-160`: for every row of `test_csv` (filename, episode_id,
    object_id, bounding_box, caption) crop the frame with the +-10 px policy (:33-37), caption it, and write
    `dest_csv` with the reference's header.  Rows keep the input order.  Returns the number of rows written."""
    import torch
    from .captioner.transform import POLICY_EVAL
    rows = list(csv.DictReader(open(test_csv, newline="")))
    captions: List[Optional[str]] = [None] * len(rows)
    by_file: Dict[str, List[int]] = {}
    for i, r in enumerate(rows):
        by_file.setdefault(r["filename"], []).append(i)
    pending: Dict[tuple, list] = {}      # frame shape -> [(image, row indices)]

    def run(shape):
        items = pending.pop(shape)
        frames = torch.from_numpy(np.stack([im for im, _ in items]))
        boxes, fidx, ridx = [], [], []
        for f, (_, idxs) in enumerate(items):
            for i in idxs:
                boxes.append([float(v) for v in ast.literal_eval(rows[i]["bounding_box"])])
                fidx.append(f); ridx.append(i)
        # float64: the +-10 px policy and PIL's rounding act on the raw CSV values (transform.box_rects, POLICY_EVAL)
        out = engine.caption_boxes(frames, np.asarray(boxes, np.float64), fidx, None, POLICY_EVAL, seq_len)
        toks = out["tokens"].cpu().numpy()
        for j, i in enumerate(ridx):
            captions[i] = decode(toks[j])

    for fn, idxs in by_file.items():
        img = np.load(fn, allow_pickle=True)["arr_0"].item()["image"][:, :, :3]
        shape = img.shape[:2]
        pending.setdefault(shape, []).append((np.ascontiguousarray(img), idxs))
        if sum(len(ix) for _, ix in pending[shape]) >= max_crops:
            run(shape)
    for shape in list(pending.keys()):
        run(shape)
    with open(dest_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(CSV_HEADER)
        for r, c in zip(rows, captions):
            w.writerow([r["filename"], r["episode_id"], r["object_id"], r["bounding_box"], c, r["caption"]])
    return len(rows)
