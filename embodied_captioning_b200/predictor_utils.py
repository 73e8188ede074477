"""`Captioner` wrapper returning only the caption string (experimenting_env/utils/predictor_utils.py:166-208)."""
import logging

import torch
from torch import nn

from .captioner.models.coca.coca import CoCa
from .captioner.utils.utils import Configuration

logger = logging.getLogger(__name__)


class Captioner(nn.Module):
    def __init__(self, cfg=None, input_format=None, load_checkpoint=False, checkpoint_path=None, metadata=None,
                 model=None):
        super().__init__()
        self.model = self.get_captioner(cfg.captioner) if model is None else model
        self.model.eval()
        assert self.model is not None, "No model provided"
        logger.info("Captioner model loaded successfully")
        if load_checkpoint and checkpoint_path is not None:
            checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
            # the reference loads into the WRAPPER (predictor_utils.py:184), so its checkpoints carry `model.`-prefixed
            # keys; the inner load strips `model.` / `module.` prefixes
            self.model.model.load_state_dict(checkpoint['model'])
            logger.info(f"Captioner model checkpoint loaded successfully from {checkpoint_path}")
        # the reference pins the captioner to the CPU here (predictor_utils.py:187); the B200 build has no
        # CPU path, so compute stays on the GPU and only the returned tensors follow the caller's device.
        self.model.to('cpu')

    def get_captioner(self, cfg):
        if cfg.arch_name == "coca":
            captioner_cfg = Configuration(arch_name=cfg.arch_name, model_name=cfg.model_name,
                                          checkpoint_name=cfg.checkpoint_name, height=cfg.height, width=cfg.width)
            return CoCa(captioner_cfg.captioner).eval()
        raise NotImplementedError(f"arch_name={cfg.arch_name!r} is outside the B200-native hot path")

    def forward(self, x):
        out = self.model(x)
        return out['text']
