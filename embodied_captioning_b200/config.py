"""Architecture hyper-parameters of the CoCa captioner towers.

Mirrors the JSON registry of the reference
(`experimenting_env/captioner/models/coca/model_configs/coca_ViT-L-14.json`, loaded by
`experimenting_env/captioner/models/coca/factory.py:27-57`).  `coca_ViT-B-32` is not shipped by the
reference; its values are upstream open_clip's, confirmed by the
253.56 M parameter count.  The `coca_tiny*` entries exist only for fast parity tests.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, asdict

_CFG_DIR = os.path.join(os.path.dirname(__file__), "captioner", "models", "coca", "model_configs")

SOT_TOKEN_ID = 49406  # coca_model.py:232
EOS_TOKEN_ID = 49407  # coca_model.py:233
PAD_TOKEN_ID = 0      # coca_model.py:103,234

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass(frozen=True)
class CoCaConfig:
    name: str
    embed_dim: int            # E: pooler / latent / decoder width
    image_size: int
    patch_size: int
    vision_width: int         # W
    vision_layers: int
    vision_head_dim: int      # open_clip: head_width = 64
    pooler_heads: int         # attn_pooler_heads (vision_cfg)
    n_queries: int            # 256 learned queries (coca_model.py:54)
    text_width: int           # T
    text_heads: int
    text_layers: int
    mm_layers: int
    mm_heads: int
    context_length: int       # 76 (+1 CLS slot in the positional table)
    vocab_size: int
    sot_token_id: int = SOT_TOKEN_ID
    eos_token_id: int = EOS_TOKEN_ID
    pad_token_id: int = PAD_TOKEN_ID

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def n_patches(self) -> int:
        return self.grid * self.grid

    @property
    def vision_tokens(self) -> int:
        return self.n_patches + 1

    @property
    def vision_heads(self) -> int:
        return self.vision_width // self.vision_head_dim

    @property
    def patch_dim(self) -> int:
        return 3 * self.patch_size * self.patch_size

    @property
    def n_image_embs(self) -> int:      # rows of image_embs handed to the decoder (x[:,1:])
        return self.n_queries - 1

    def to_dict(self):
        return asdict(self)


def _from_open_clip_json(name: str, d: dict, **extra) -> CoCaConfig:
    v, t, m = d["vision_cfg"], d["text_cfg"], d["multimodal_cfg"]
    assert t["width"] == m["width"] == d["embed_dim"], "CoCa decoder width must equal embed_dim"
    return CoCaConfig(
        name=name,
        embed_dim=d["embed_dim"],
        image_size=v["image_size"],
        patch_size=v["patch_size"],
        vision_width=v["width"],
        vision_layers=v["layers"],
        vision_head_dim=v.get("head_width", 64),
        pooler_heads=v.get("attn_pooler_heads", 8),
        n_queries=v.get("attn_pooler_queries", 256),
        text_width=t["width"],
        text_heads=t["heads"],
        text_layers=t["layers"],
        mm_layers=m["layers"],
        mm_heads=m["heads"],
        context_length=t["context_length"],
        vocab_size=t["vocab_size"],
        **extra,
    )


def list_models():
    return sorted(f[:-5] for f in os.listdir(_CFG_DIR) if f.endswith(".json"))


def get_config(name: str) -> CoCaConfig:
    """`factory.py:92-97 get_model_config` equivalent; raises RuntimeError like `factory.py:232-233`."""
    path = os.path.join(_CFG_DIR, name + ".json")
    if not os.path.isfile(path):
        raise RuntimeError(f"Model config for {name} not found; available models {list_models()}.")
    with open(path) as f:
        d = json.load(f)
    extra = {}
    for k in ("sot_token_id", "eos_token_id", "pad_token_id"):
        if k in d:
            extra[k] = d[k]
    return _from_open_clip_json(name, d, **extra)
