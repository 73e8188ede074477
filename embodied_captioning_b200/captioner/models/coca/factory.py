"""Model / transform construction with the reference's entry points
(experimenting_env/captioner/models/coca/factory.py:378-432 `create_model_and_transforms`,
:183-338 `create_model`, :131-180 `load_checkpoint`).

`pretrained` may be
  * a path to a checkpoint file (`torch.load`; keys `state_dict` / `model` unwrapped, `module.` stripped --
    evaluate_finetuned_model.py:139-146, predictor_utils.py:182-185),
  * an open_clip pretrained tag (e.g. 'mscoco_finetuned_laion2B-s13B-b90k'): resolved only against
    `cache_dir` (there is no network); missing -> RuntimeError like factory.py:310-314,
  * 'synthetic' or 'synthetic:<seed>': seeded random-init weights in the same layout (synth.py),
  * None: the model is returned without weights (load_state_dict later).
"""
import glob
import os
from typing import Optional, Union

import torch

from ....config import get_config, list_models
from ....synth import make_state_dict
from ...tokenizer import get_tokenizer as _get_tokenizer, decode  # noqa: F401  (re-exported)
from ...transform import ImageTransform
from .coca_model import CoCa


def _unwrap(ckpt):
    if isinstance(ckpt, dict):
        for key in ("state_dict", "model"):
            if key in ckpt and isinstance(ckpt[key], dict):
                ckpt = ckpt[key]
                break
    return {(k[7:] if k.startswith("module.") else k): v for k, v in ckpt.items()}


def _resolve_pretrained(model_name: str, pretrained: str, cache_dir: Optional[str]):
    if pretrained.startswith("synthetic"):
        seed = int(pretrained.split(":")[1]) if ":" in pretrained else 1234
        return make_state_dict(get_config(model_name), seed=seed)
    if os.path.isfile(pretrained):
        return _unwrap(torch.load(pretrained, map_location="cpu", weights_only=False))
    if cache_dir and os.path.isdir(cache_dir):
        for pat in (f"*{pretrained}*", f"*{model_name}*{pretrained}*"):
            for hit in sorted(glob.glob(os.path.join(cache_dir, "**", pat), recursive=True)):
                if os.path.isfile(hit) and hit.endswith((".pt", ".bin", ".pth")):
                    return _unwrap(torch.load(hit, map_location="cpu", weights_only=False))
    raise RuntimeError(f"Pretrained weights ({pretrained}) not found for model {model_name} "
                       f"(no network: place the checkpoint under cache_dir={cache_dir!r} or pass a file path).")


# open_clip's precision strings (factory.py:189,254-268).  The native path has ONE arithmetic -- bf16 GEMM / attention
# operands, fp32 accumulation, fp32 residual stream / LayerNorm / softmax -- whatever is asked for; the reference's
# `torch.cuda.amp.autocast()` around generate (evaluate_finetuned_model.py:42) is likewise a no-op here.  Unknown
# strings raise instead of being silently accepted.
PRECISIONS = ("fp32", "fp16", "bf16", "amp", "amp_bf16", "amp_bfloat16", "pure_fp16", "pure_bf16")


def create_model(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                 device: Union[str, torch.device] = "cpu", cache_dir: Optional[str] = None, **_ignored):
    if precision not in PRECISIONS:
        raise ValueError(f"precision={precision!r}: expected one of {PRECISIONS} (all run the native bf16-operand / "
                         "fp32-accumulate arithmetic)")
    cfg = get_config(model_name.replace("/", "-"))
    want = torch.device(device)
    cuda = want if want.type == "cuda" else torch.device("cuda", torch.cuda.current_device()
                                                         if torch.cuda.is_available() else 0)
    model = CoCa(cfg, device=cuda)
    if want.type != "cuda":
        model.to(want)        # compute stays on the GPU; outputs are returned on `want`
    if pretrained:
        model.load_state_dict(_resolve_pretrained(cfg.name, pretrained, cache_dir))
    return model


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                device: Union[str, torch.device] = "cpu", cache_dir: Optional[str] = None, **kwargs):
    """-> (model, preprocess_train, preprocess_val); training augmentation is outside the hot path, so both
    transforms are the inference transform."""
    model = create_model(model_name, pretrained, precision=precision, device=device, cache_dir=cache_dir, **kwargs)
    tf = ImageTransform(size=model.cfg.image_size, device=model.device)
    return model, tf, tf


def get_tokenizer(model_name: str = "", **kw):
    try:
        vocab = get_config(model_name.replace("/", "-")).vocab_size
    except RuntimeError:
        vocab = 49408
    return _get_tokenizer(model_name, vocab_size=vocab, **kw)
