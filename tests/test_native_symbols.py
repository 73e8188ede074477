"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/ec_b200.h declares
(no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ec_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from embodied_captioning_b200 import _native as nat
    assert _declared() == sorted(nat.EXPORTS)


def test_library_loads_and_exports_every_symbol():
    from embodied_captioning_b200 import _native as nat
    if not os.path.isfile(nat.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    L = nat.lib()
    assert L.ec_abi_version() == 1
    for name in _declared():
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_gpu():
    import torch
    from embodied_captioning_b200 import _native as nat
    from embodied_captioning_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nat.NativeError):
        ops.layernorm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))
    from embodied_captioning_b200.captioner.utils.utils import Configuration
    from embodied_captioning_b200.captioner.utils.utils_captioner import select_captioner
    cfg = Configuration(arch_name="coca", model_name="coca_tiny", checkpoint_name="synthetic", height=224, width=224)
    with pytest.raises(nat.NativeError):
        select_captioner(cfg.captioner)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "embodied_captioning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
