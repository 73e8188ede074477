"""CPU: the numpy preprocess oracle is bit-exact against the committed Pillow/torchvision golden hashes
(and against the live libraries when they are installed)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.json")))


def golden_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("e", GOLD["entries"], ids=lambda e: f"{e['h']}x{e['w']}")
def test_oracle_matches_pillow_golden(e):
    out = po.transform(golden_image(e["seed"], e["h"], e["w"]))
    assert out.dtype == np.float32 and out.shape == (3, 224, 224)
    assert hashlib.sha256(out.tobytes()).hexdigest() == e["sha256"]


def test_oracle_matches_live_pillow_when_available():
    PIL = pytest.importorskip("PIL")
    tv = pytest.importorskip("torchvision")
    from PIL import Image
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode
    tf = T.Compose([T.Resize(224, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(mean=po.MEAN.tolist(), std=po.STD.tolist())])
    rng = np.random.default_rng(7)
    for _ in range(6):
        h, w = int(rng.integers(20, 500)), int(rng.integers(20, 500))
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(tf(Image.fromarray(img, "RGB")).numpy(), po.transform(img))


def test_identity_for_224_crops():
    img = golden_image(3, 224, 224)
    assert np.array_equal(po.transform_u8(img), img)


def test_center_crop_round_half_even():
    assert po.center_crop_offsets(417, 224) == (96, 0)      # (417-224)/2 = 96.5 -> 96
    assert po.center_crop_offsets(419, 224) == (98, 0)      # 97.5 -> 98
    assert po.resized_size(640, 480) == (298, 224)


def test_crop_policies():
    fr_h, fr_w = 480, 640
    box = [100.7, 50.2, 300.9, 200.5]
    assert po.crop_rect(po.POLICY_AGENT, box, fr_h, fr_w) == ([100, 50, 300, 200], False, False)
    r, pad, swap = po.crop_rect(po.POLICY_PSEUDOLABELER, box, fr_h, fr_w)
    assert pad and swap and r == [60, 20, 340, 230]
    # the pseudolabeler clamps x to image.shape[0] (= height) and y to image.shape[1] (= width): sic
    r2, _, _ = po.crop_rect(po.POLICY_PSEUDOLABELER, [400, 300, 630, 470], fr_h, fr_w)
    assert r2[2] == 480 and r2[3] == 504
    r3, pad3, _ = po.crop_rect(po.POLICY_EVAL, [5, 5, 635, 475], fr_h, fr_w)
    assert pad3 and r3 == [0, 5, 635, 475]
    r4, pad4, swap4 = po.crop_rect(po.POLICY_PSEUDOCAPTIONER, box, 1280, 1280)
    assert (not pad4) and swap4 and r4 == [80, 35, 320, 215]


def test_extract_crop_pads_outside_frame():
    fr = golden_image(5, 10, 12)
    c = po.extract_crop(fr, [-2, -1, 5, 4], True, False)
    assert c.shape == (5, 7, 3) and (c[0] == 0).all() and (c[:, :2] == 0).all()
    assert np.array_equal(c[1:, 2:], fr[0:4, 0:5])
