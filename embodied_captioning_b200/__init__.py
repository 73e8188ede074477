"""embodied_captioning_b200 -- B200-native (sm_100a) captioner forward + per-instance multi-view
caption disagreement, drop-in for the `experimenting_env/captioner` Python API of
hsp-iit/embodied-captioning.  See DESIGN.md / INTEGRATION.md.

Importing this package never touches the GPU; every compute call goes through `libec_b200.so`
(hand-written CUDA behind the C ABI of `include/ec_b200.h`) and raises if it is missing.
"""
from .config import CoCaConfig, get_config, list_models  # noqa: F401

__version__ = "0.1.0"
