"""Base captioner object (experimenting_env/captioner/captioning_predictor.py:8-53).

The reference derives from `pytorch_lightning.LightningModule`; only `nn.Module` behaviour (`eval`, `to`,
`device`, call) is used by its callers (goal_exploration.py:55-56, predictor_utils.py:166-208), so this base
is a plain `nn.Module` and does not require Lightning.
"""
import torch
import torch.nn.functional as F
from torch import nn


class CaptioningPredictor(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.perplexity = 0.0
        self.outputs = {}
        if cfg is not None:
            self.input_height, self.input_width = cfg.height, cfg.width

    @property
    def device(self):
        return getattr(self, "_device", torch.device("cpu"))

    def compute_perplexity(self, logits=None):
        """exp(-sum(log max softmax) / T) as a double (captioning_predictor.py:34-47)."""
        if logits is None:
            steps = self.outputs["logits"]
            logits = torch.stack([s.to(torch.float32).cpu() for s in steps], dim=1)   # (n_samples, seq_len, vocab)
        probs = F.softmax(logits, dim=-1)
        probs = torch.max(probs, dim=-1).values
        sum_log_probs = -probs.log().sum()
        num_tokens = probs.shape[1]
        self.perplexity = torch.exp(sum_log_probs / num_tokens).double()
        return self.perplexity

    def print_caption(self):
        print(self.outputs["text"])
