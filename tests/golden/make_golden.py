"""Generates the committed golden fixtures (run in the build container, where Pillow/torchvision exist):

  preprocess_golden.json : sha256 of the fp32 [3,224,224] output of the REAL torchvision/Pillow transform
                           (Resize(224,BICUBIC) -> CenterCrop -> ToTensor -> Normalize) on seeded random
                           images of several shapes -- pins oracle/preprocess_oracle.py and the CUDA kernel.
  coca_tiny_golden.npz   : fp32-oracle outputs (image latent, image_embs checksum, greedy tokens, teacher-forced
                           logits sample, per-instance disagreement) for the seeded `coca_tiny` weights/crops.

usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from embodied_captioning_b200 import synth  # noqa: E402
from embodied_captioning_b200.config import get_config  # noqa: E402
from oracle import disagreement_oracle as do  # noqa: E402
from oracle import preprocess_oracle as po  # noqa: E402
from oracle.coca_oracle import OracleCoCa  # noqa: E402

SHAPES = [(224, 224), (300, 224), (224, 300), (100, 57), (57, 100), (480, 640), (33, 320), (225, 226), (97, 97),
          (417, 193), (640, 480), (1000, 700)]


def golden_image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def main():
    from PIL import Image
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode
    tf = T.Compose([T.Resize(224, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(mean=po.MEAN.tolist(), std=po.STD.tolist())])
    entries = []
    for i, (h, w) in enumerate(SHAPES):
        img = golden_image(1000 + i, h, w)
        ref = tf(Image.fromarray(img, "RGB")).numpy().astype(np.float32)
        entries.append({"seed": 1000 + i, "h": h, "w": w, "sha256": hashlib.sha256(ref.tobytes()).hexdigest(),
                        "sum": float(ref.astype(np.float64).sum())})
    import PIL
    import torchvision
    json.dump({"pillow": PIL.__version__, "torchvision": torchvision.__version__, "entries": entries},
              open(os.path.join(HERE, "preprocess_golden.json"), "w"), indent=1)

    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=1234)
    crops = synth.make_crops(16, seed=5678)
    x = torch.from_numpy(np.stack([po.transform(c) for c in crops]))
    m = OracleCoCa(cfg, sd)
    lat, embs = m.encode_image_tokens(x)
    gen = m.generate(image_embs=embs, seq_len=12, mode="kvcache")
    text = gen["text"]
    tf_logits = m.forward_logits(embs[:4], text[:4, :6])
    keys = np.array([3, 1, 3, 2, 1, 3, 2, 2, 1, 3, 5, 5, 1, 2, 3, 8], dtype=np.int64)
    k, s_full, c = do.instance_disagreement(lat.numpy(), keys, mode="full")
    _, s_pairs, _ = do.instance_disagreement(lat.numpy(), keys, mode="pairs")
    np.savez_compressed(os.path.join(HERE, "coca_tiny_golden.npz"),
                        latent=lat.numpy(), embs_mean=embs.mean(dim=1).numpy(), embs_l2=embs.norm(dim=(1, 2)).numpy(),
                        tokens=text.numpy(), tf_logits_top8=tf_logits.topk(8, dim=-1).values.numpy(),
                        tf_logits_top8_idx=tf_logits.topk(8, dim=-1).indices.numpy(),
                        keys=keys, inst_keys=k, inst_full=s_full, inst_pairs=s_pairs, inst_count=c)
    print("wrote goldens:", len(entries), "preprocess entries; tokens", tuple(text.shape))


if __name__ == "__main__":
    main()
