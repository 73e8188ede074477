"""GPU: crop/resize/normalise gather kernel is bit-exact (fp32 output) against the oracle and the committed
Pillow/torchvision golden hashes; integer crop indexing is bit-exact for the four reference box policies."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from embodied_captioning_b200 import ops, synth
from embodied_captioning_b200.captioner import transform as tr
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.json")))


def golden_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("e", GOLD["entries"], ids=lambda e: f"{e['h']}x{e['w']}")
def test_kernel_matches_pillow_golden(device, e):
    img = golden_image(e["seed"], e["h"], e["w"])
    t = tr.ImageTransform(224, device)
    out = t(img)
    assert out.dtype == torch.float32 and tuple(out.shape) == (3, 224, 224)
    got = out.cpu().numpy()
    assert hashlib.sha256(got.tobytes()).hexdigest() == e["sha256"], \
        f"max abs diff vs oracle {np.abs(got - po.transform(img)).max()}"


@pytest.mark.parametrize("policy", [tr.POLICY_AGENT, tr.POLICY_PSEUDOLABELER, tr.POLICY_EVAL, tr.POLICY_PSEUDOCAPTIONER])
def test_boxes_on_frames_bit_exact(device, policy):
    frames, boxes, fidx, _ = synth.make_frames_and_boxes(3, 8, 480, 640, seed=11 + policy)
    if policy == tr.POLICY_EVAL:
        boxes = np.trunc(boxes)
    # touches the frame border: the expansion policies leave the frame (PIL zero padding).  The pseudolabeler
    # clamps x against the frame HEIGHT (sic, pseudolabeler.py:631), so its border box sits at the bottom edge.
    boxes[0] = [100, 400, 200, 480] if policy == tr.POLICY_PSEUDOLABELER else [600, 440, 640, 480]
    rects, _ = tr.box_rects(boxes, 480, 640, policy)
    valid = (rects[:, 2] > rects[:, 0]) & (rects[:, 3] > rects[:, 1])   # the reference's PIL crop raises otherwise
    assert valid.sum() >= 8
    b16, f32, status = tr.preprocess_boxes(torch.from_numpy(frames).to(device), boxes, fidx, policy, want_f32=True)
    assert status.cpu().numpy().astype(bool).tolist() == (~valid).tolist()   # empty crops are flagged, not faulted
    ref, ref_rects = po.preprocess_boxes(frames, boxes[valid], fidx[valid], policy)
    assert np.array_equal(rects[valid], ref_rects)                        # crop indexing: bit-exact
    got = f32.cpu().numpy()[valid]
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"
    assert torch.equal(b16.cpu()[torch.from_numpy(valid)], torch.from_numpy(ref).to(torch.bfloat16))


def test_identity_path_for_224_crops(device):
    crops = synth.make_crops(5, seed=3)
    t = tr.ImageTransform(224, device)
    out = t.batch(list(crops)).cpu().numpy()
    ref = np.stack([po.normalize_f32(c) for c in crops])
    assert np.array_equal(out, ref)


def test_empty_and_oversized_crops_are_flagged_not_faulted(device):
    frames = torch.from_numpy(golden_image(1, 64, 64)[None]).to(device)
    rects = torch.tensor([[10, 10, 10, 30], [0, 0, 64, 64], [5, 5, 60, 60]], dtype=torch.int32, device=device)
    fi = torch.tensor([0, 0, 7], dtype=torch.int32, device=device)       # last one: bad frame index
    b16, _, status = ops.preprocess_crops(frames, rects, fi, None)
    assert status.cpu().tolist() == [1, 0, 1]
    assert float(b16[0].float().abs().max()) == 0.0
