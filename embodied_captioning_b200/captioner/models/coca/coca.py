"""Drop-in captioner object (experimenting_env/captioner/models/coca/coca.py:19-33).

`CoCa(cfg)(PIL.Image) -> {"text": str, "logits": list}`; the same `self.outputs` dict is reused and mutated
across calls like the reference (captioning_predictor.py:12, coca.py:30-33).  `caption_batch` is the batched
extension used by the B200 engine (one preprocess launch, one encoder pass, one decode loop for N crops).
"""
import torch

from ...captioning_predictor import CaptioningPredictor
from .factory import create_model_and_transforms, get_tokenizer


class CoCa(CaptioningPredictor):
    def __init__(self, cfg=None, cache_dir="/mnt/storage/tgalliena/model_cache"):
        super().__init__(cfg)
        self.model, _, self.preprocess = create_model_and_transforms(cfg.model_name, pretrained=cfg.checkpoint_name,
                                                                     cache_dir=cache_dir, device="cuda")
        self.tokenizer = get_tokenizer(cfg.model_name)
        self._device = self.model.device

    def to(self, *args, **kwargs):
        self.model.to(*args, **kwargs)
        self._device = self.model.output_device or self.model.device
        return self

    def _text(self, tokens):
        return self.tokenizer.decode(tokens).split("<end_of_text>")[0].replace("<start_of_text>", "")

    def forward(self, inputs):
        inputs = self.preprocess(inputs).unsqueeze(0).to(self.model.device)
        outputs = self.model.generate(inputs, generation_type='top_k')
        self.outputs["text"] = self._text(outputs["text"][0])
        self.outputs["logits"] = outputs["logits"]
        return self.outputs

    @torch.no_grad()
    def caption_batch(self, images, seq_len=30, return_logits=False):
        """list of PIL / HWC uint8 crops -> list of caption strings (+ token matrix)."""
        x = self.preprocess.batch(images, want_f32=False)
        out = self.model.generate(x, generation_type='top_k', seq_len=seq_len, return_logits=return_logits)
        toks = out["text"].cpu()
        return [self._text(t) for t in toks], out
