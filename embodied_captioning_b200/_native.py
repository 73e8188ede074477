"""ctypes binding of `libec_b200.so` (C ABI declared in `include/ec_b200.h`).

There is no CPU or eager-PyTorch fallback: importing works anywhere (so CPU-only tooling can
inspect symbols), but every compute entry point raises `NativeError` unless the shared library
is present and the device is an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EC_LIB_PATH") or os.path.join(_HERE, "libec_b200.so")   # override: A/B builds only

EC_OK = 0
ERROR_NAMES = {-1: "EC_ERR_INVALID", -2: "EC_ERR_CUDA", -3: "EC_ERR_UNSUPPORTED", -4: "EC_ERR_WORKSPACE",
               -5: "EC_ERR_STATE", -6: "EC_ERR_DEVICE", -7: "EC_ERR_TIMEOUT"}

EXPORTS = [
    "ec_abi_version", "ec_last_error", "ec_model_create", "ec_model_destroy", "ec_model_set_tensor",
    "ec_model_finalize", "ec_model_set_option", "ec_preprocess_crops", "ec_encode_workspace_bytes",
    "ec_encode_image", "ec_decode_workspace_bytes", "ec_decode_greedy", "ec_forward_logits",
    "ec_disagreement_workspace_bytes", "ec_instance_disagreement", "ec_instance_disagreement_peers",
    "ec_gemm_bf16", "ec_gemm_bf16_ln", "ec_gemm_argmax", "ec_layernorm", "ec_attention", "ec_check_watchdog",
    "ec_launch_count", "ec_profile_gemm", "ec_profile_gemm_read", "ec_attention_fused_qkv",
    "ec_disagreement_map_workspace_bytes", "ec_disagreement_map",
    "ec_caption_frequency_workspace_bytes", "ec_caption_frequency",
    "ec_textenc_create", "ec_textenc_destroy", "ec_textenc_set_tensor", "ec_textenc_finalize",
    "ec_textenc_workspace_bytes", "ec_textenc_encode",
    "ec_gemm_bf16_grouped", "ec_gemm_bf16_splitk", "ec_cross_attention_absorbed", "ec_pack_image_tokens", "ec_packed_image_tokens_bytes",
    "ec_seq_timing_read", "ec_gemm_timing_read",
]
DEBUG_EXPORTS = ("ec_seq_timing_read", "ec_gemm_timing_read")      # absent from A/B builds of older sources


class NativeError(RuntimeError):
    pass


class EcConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "embed_dim", "image_size", "patch_size", "vision_width", "vision_layers", "vision_head_dim",
        "pooler_heads", "n_queries", "text_width", "text_heads", "text_layers", "mm_layers", "mm_heads",
        "context_length", "vocab_size", "sot_token_id", "eos_token_id", "pad_token_id")]


class EcTextEncConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab_size", "hidden", "layers", "heads", "intermediate", "max_positions",
                                         "type_vocab")] + [("layer_norm_eps", C.c_float)]


_lib: Optional[C.CDLL] = None


def build_hint() -> str:
    return ("libec_b200.so is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `bash embodied_captioning_b200/csrc/build.sh` (nvcc, sm_100a)")


def lib() -> C.CDLL:
    """Load (once) and return the shared library with argument types declared."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeError(build_hint())
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    L.ec_abi_version.restype = C.c_int
    L.ec_last_error.restype = C.c_char_p
    L.ec_model_create.argtypes = [C.POINTER(EcConfig), C.c_int, C.POINTER(vp)]
    L.ec_model_destroy.argtypes = [vp]
    L.ec_model_set_tensor.argtypes = [vp, C.c_char_p, vp, i64]
    L.ec_model_finalize.argtypes = [vp, vp]
    L.ec_model_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.ec_preprocess_crops.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.ec_encode_workspace_bytes.argtypes = [vp, i32]
    L.ec_encode_workspace_bytes.restype = sz
    L.ec_encode_image.argtypes = [vp, vp, i32, vp, vp, vp, sz, vp]
    L.ec_decode_workspace_bytes.argtypes = [vp, i32, i32]
    L.ec_decode_workspace_bytes.restype = sz
    L.ec_decode_greedy.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    L.ec_forward_logits.argtypes = [vp, vp, vp, i32, i32, vp, vp, sz, vp]
    L.ec_disagreement_workspace_bytes.argtypes = [i32]
    L.ec_disagreement_workspace_bytes.restype = sz
    L.ec_instance_disagreement.argtypes = [vp, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    L.ec_instance_disagreement_peers.argtypes = [vp, i32, i32, i64, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                                 sz, vp]
    L.ec_gemm_bf16.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, i32, vp]
    L.ec_gemm_bf16_ln.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, i32, vp, vp]
    L.ec_gemm_argmax.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.ec_layernorm.argtypes = [vp, i64, vp, vp, vp, i64, i32, i32, i32, vp]
    L.ec_attention.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, i32, i32, i32, i32, i32, vp]
    L.ec_check_watchdog.argtypes = [vp]
    L.ec_attention_fused_qkv.argtypes = [vp, i64, i32, i32, i32, vp, i64, i64, i32, i32, i32, i32, i32, vp]
    L.ec_disagreement_map_workspace_bytes.argtypes = [i32, i32, i32]
    L.ec_disagreement_map_workspace_bytes.restype = sz
    L.ec_disagreement_map.argtypes = [vp, i32, vp, vp, vp, i32, C.c_float, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.ec_caption_frequency_workspace_bytes.argtypes = [i32]
    L.ec_caption_frequency_workspace_bytes.restype = sz
    L.ec_caption_frequency.argtypes = [vp, i32, i32, vp, i32, vp, vp, vp, vp, vp, sz, vp]
    L.ec_textenc_create.argtypes = [C.POINTER(EcTextEncConfig), C.c_int, C.POINTER(vp)]
    L.ec_textenc_destroy.argtypes = [vp]
    L.ec_textenc_set_tensor.argtypes = [vp, C.c_char_p, vp, i64]
    L.ec_textenc_finalize.argtypes = [vp]
    L.ec_textenc_workspace_bytes.argtypes = [vp, i32, i32]
    L.ec_textenc_workspace_bytes.restype = sz
    L.ec_textenc_encode.argtypes = [vp, vp, vp, i32, i32, vp, vp, sz, vp]
    L.ec_gemm_bf16_splitk.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp, vp]
    L.ec_gemm_bf16_grouped.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp]
    L.ec_packed_image_tokens_bytes.argtypes = [i32, i32, i32]
    L.ec_packed_image_tokens_bytes.restype = sz
    L.ec_pack_image_tokens.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.ec_cross_attention_absorbed.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]
    L.ec_launch_count.argtypes = []
    L.ec_launch_count.restype = C.c_int64
    L.ec_profile_gemm.argtypes = [i32]
    L.ec_profile_gemm_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    ab_build = bool(os.environ.get("EC_LIB_PATH"))
    for name in DEBUG_EXPORTS:
        if hasattr(L, name):
            getattr(L, name).argtypes = [C.POINTER(C.c_int64), i32]
    for name in EXPORTS:
        if ab_build and name in DEBUG_EXPORTS:
            continue
        getattr(L, name)          # AttributeError here means the .so is older than the header: rebuild
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != EC_OK:
        msg = lib().ec_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'libec_b200'} failed: {ERROR_NAMES.get(rc, rc)}: {msg}")


def ptr(t) -> Optional[int]:
    """Device (or host) pointer of a torch tensor, None for None."""
    return None if t is None else t.data_ptr()


def current_stream(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def cuda_device(device=None):
    """Normalise `device` to an indexed CUDA device: an index-less 'cuda' means the CURRENT device (what every torch
    allocation resolves it to), never device 0 -- the native handle and torch's buffers must agree under torchrun."""
    import torch
    d = torch.device("cuda" if device is None else device)
    if d.type != "cuda":
        raise NativeError(f"device {d} is not a CUDA device; there is no CPU execution path")
    if d.index is not None:
        return d
    return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("embodied_captioning_b200 needs an sm_100 (B200) GPU; no CUDA device is visible and there "
                          "is no CPU fallback")
