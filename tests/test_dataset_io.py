"""CPU: on-disk formats either side of the hot path and the caption filter (N4 host part)."""
import csv
import os

import numpy as np
import pytest

from embodied_captioning_b200 import dataset_io as dio
from embodied_captioning_b200.captioner import pseudocaptioner as pc
from oracle import pseudocaption_oracle as pco


def test_observation_file_naming_roundtrip(tmp_path):
    info = dio.SenseInfo(str(tmp_path), "rgb", episode=12, camera_id=3, step=45)
    p = info.get_path()
    assert os.path.basename(p) == "episode_000012_step_00045_modality_rgb_id_3.npz"      # sensors_utils.py:23-27
    back = dio.get_sense_info(p)
    assert (back.episode, back.step, back.mod, back.camera_id) == (12, 45, "rgb", 3)
    rgb = np.random.default_rng(0).integers(0, 256, size=(6, 8, 3), dtype=np.uint8)
    depth = np.random.default_rng(1).random((6, 8)).astype(np.float32)
    paths = dio.save_obs(str(tmp_path), 7, {"rgb": rgb, "depth": depth}, 9)
    assert [os.path.basename(q) for q in paths] == ["episode_000007_step_00009_modality_rgb_id_0.npz",
                                                    "episode_000007_step_00009_modality_depth_id_1.npz"]
    assert np.array_equal(dio.load_obs(paths[0]), rgb) and np.array_equal(dio.load_obs(paths[1]), depth)


def test_detection_file_roundtrip_and_batches(tmp_path):
    rng = np.random.default_rng(3)
    files = []
    for f in range(5):
        h, w = (48, 64) if f != 3 else (32, 32)
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        n = int(rng.integers(1, 4))
        boxes = np.stack([rng.uniform(0, 10, n), rng.uniform(0, 10, n), rng.uniform(12, 30, n), rng.uniform(12, 30, n)], 1)
        infos = [{"id_episode": 2, "id_object": int(rng.integers(0, 6))} for _ in range(n)]
        p = dio.save_detections(str(tmp_path / f"det_{f}.npz"), img, boxes, infos, [f"cap {f} {i}" for i in range(n)])
        files.append(p)
        raw = np.load(p, allow_pickle=True)["arr_0"].item()              # the reference's read (pseudocaptioner.py:132)
        assert set(raw.keys()) == {"instances", "image"} and np.array_equal(raw["image"], img)
        d = dio.load_detections(p)
        assert d.boxes.dtype == np.float32 and np.allclose(d.boxes, boxes.astype(np.float32))
        assert d.object_ids.tolist() == [i["id_object"] for i in infos] and d.captions[0] == f"cap {f} 0"
    batches = list(dio.iter_frame_batches(files, max_crops=4))
    assert sum(len(b[1]) for b in batches) == sum(len(dio.load_detections(p).boxes) for p in files)
    for frames, boxes, fidx, eids, oids, names in batches:
        assert frames.ndim == 4 and len(names) == frames.shape[0] and fidx.max() == frames.shape[0] - 1
        assert len(boxes) == len(fidx) == len(eids) == len(oids)
    assert {b[0].shape[1:3] for b in batches} == {(48, 64), (32, 32)}


def test_caption_filter_matches_restated_reference():
    caps = ["a wooden chair next to a table", "a nice location", "A DOG on a sofa", "a lamp", "the bedroom wall",
            "a carpet", "carton box", "some 3D shape", ""]
    for c in caps:
        assert pc.filter_captions(c) == pco.filter_captions(c, pc.BANNED_WORDS)
    assert pc.filter_captions("a wooden chair next to a table") and pc.filter_captions("a lamp")
    assert not pc.filter_captions("a nice location")        # substring semantics: "cat" inside "location"
    assert not pc.filter_captions("a carpet")               # "car" and "pet"
    ref_path = "/root/reference/experimenting_env/captioner/pseudocaptioner.py"
    if os.path.isfile(ref_path):                            # build container only: the word list is the reference's
        import ast
        src = open(ref_path).read()
        a = src.index("banned_words = [")
        ref = ast.literal_eval(src[a + len("banned_words = "):src.index("]", a) + 1])
        assert ref == list(pc.BANNED_WORDS)


def test_frequency_oracle_known_answer():
    caps = ["a red chair", "a chair", "a red chair", "a dog", "a table", "a chair"]
    out = pco.captions_frequency(caps, [1, 1, 1, 1, 1, 1], [5, 5, 5, 5, 9, 5], pc.BANNED_WORDS)
    assert out == {(1, 5): [[2, "a red chair"], [2, "a chair"]], (1, 9): [[1, "a table"]]}


def test_eval_csv_header():
    assert dio.CSV_HEADER == ["filename", "episode_id", "object_id", "bounding_box", "proposed_caption",
                              "reference_caption"]                   # evaluate_finetuned_model.py:156
