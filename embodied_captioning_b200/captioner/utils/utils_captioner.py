"""`select_captioner(cfg)` (experimenting_env/captioner/utils/utils_captioner.py:4-11)."""
from ..models.coca.coca import CoCa


def select_captioner(cfg):
    arch_name = cfg.arch_name
    assert arch_name.lower() in ["coca", "blip2"], "Currently, only 'coca' and 'blip2' architectures are supported."
    if arch_name.lower() == "coca":
        return CoCa(cfg)
    # BLIP-2 (OPT-2.7B int8 through bitsandbytes) is a second captioner family outside the B200 hot path
    # (SURVEY.md C8): it is not re-implemented here.
    raise NotImplementedError("arch_name='blip2' is outside the B200-native hot path; use the reference's BLIP2 wrapper")
