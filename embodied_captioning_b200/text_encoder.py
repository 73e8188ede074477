"""Caption-text encoder on the device.

Drop-in for the reference's `SentenceTransformer("all-MiniLM-L6-v2")` object
(agents/goal_exploration/goal_exploration.py:57 `self.encoder = ...`, :102-103 `self.encoder.encode(caption)`;
detector/pseudolabeler.py:568,651,677 `encode(..., convert_to_tensor=True)`): same `encode` protocol, so it can be
handed to `disagreement.set_encoder`.  WordPiece ids -> 6-layer BERT (hidden 384, 12 heads x 32, GELU, post-LN,
eps 1e-12) -> masked mean pooling -> L2 normalise, all in `libec_b200` (`csrc/textenc.cu`); no CPU path.

Neither the MiniLM weights nor its WordPiece vocabulary exist offline: weights are loaded from an HF `BertModel`
state dict (`load_state_dict`), the tokenizer is injected (`tokenizer(list[str]) -> {"input_ids", "attention_mask"}`,
e.g. `transformers.BertTokenizerFast(vocab_file=...)`).  `encode_ids` takes ids that are already on the device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as nat


@dataclass(frozen=True)
class TextEncoderConfig:
    vocab_size: int = 30522
    hidden: int = 384
    layers: int = 6
    heads: int = 12
    intermediate: int = 1536
    max_positions: int = 512
    type_vocab: int = 2
    layer_norm_eps: float = 1e-12


MINILM_L6 = TextEncoderConfig()                                   # sentence-transformers/all-MiniLM-L6-v2
TINY = TextEncoderConfig(vocab_size=512, hidden=128, layers=2, heads=4, intermediate=256, max_positions=64)


def state_dict_spec(cfg: TextEncoderConfig) -> Dict[str, tuple]:
    H, I = cfg.hidden, cfg.intermediate
    spec = {"embeddings.word_embeddings.weight": (cfg.vocab_size, H),
            "embeddings.position_embeddings.weight": (cfg.max_positions, H),
            "embeddings.token_type_embeddings.weight": (cfg.type_vocab, H),
            "embeddings.LayerNorm.weight": (H,), "embeddings.LayerNorm.bias": (H,)}
    for l in range(cfg.layers):
        p = f"encoder.layer.{l}."
        for n in ("query", "key", "value"):
            spec[p + f"attention.self.{n}.weight"] = (H, H)
            spec[p + f"attention.self.{n}.bias"] = (H,)
        spec[p + "attention.output.dense.weight"] = (H, H); spec[p + "attention.output.dense.bias"] = (H,)
        spec[p + "attention.output.LayerNorm.weight"] = (H,); spec[p + "attention.output.LayerNorm.bias"] = (H,)
        spec[p + "intermediate.dense.weight"] = (I, H); spec[p + "intermediate.dense.bias"] = (I,)
        spec[p + "output.dense.weight"] = (H, I); spec[p + "output.dense.bias"] = (H,)
        spec[p + "output.LayerNorm.weight"] = (H,); spec[p + "output.LayerNorm.bias"] = (H,)
    return spec


def synthetic_state_dict(cfg: TextEncoderConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Seeded BERT-style init (N(0, 0.02), LayerNorm gains near 1), values representable in bf16."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_spec(cfg).items():
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif "embeddings" in name:
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) * (shape[1] ** -0.5)
        sd[name] = t.to(torch.bfloat16).to(torch.float32)
    return sd


class CaptionTextEncoder:
    def __init__(self, cfg: TextEncoderConfig = MINILM_L6, device="cuda:0",
                 tokenizer: Optional[Callable] = None, max_seq_length: int = 256):
        self.cfg = cfg
        self.device = nat.cuda_device(device)
        self.tokenizer = tokenizer
        self.max_seq_length = min(max_seq_length, cfg.max_positions)
        self._handle = C.c_void_p()
        self._loaded = False
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if self._handle:
                nat.lib().ec_textenc_destroy(self._handle)
        except Exception:
            pass

    def get_sentence_embedding_dimension(self) -> int:
        return self.cfg.hidden

    def load_state_dict(self, state_dict, strict: bool = True):
        nat.require_cuda()
        lib = nat.lib()
        c = self.cfg
        if not self._handle:
            ec = nat.EcTextEncConfig(c.vocab_size, c.hidden, c.layers, c.heads, c.intermediate, c.max_positions,
                                     c.type_vocab, c.layer_norm_eps)
            nat.check(lib.ec_textenc_create(C.byref(ec), self.device.index, C.byref(self._handle)), "ec_textenc_create")
        strip = lambda k: k[5:] if k.startswith("bert.") else (k[13:] if k.startswith("0.auto_model.") else k)
        sd = {strip(k): v for k, v in state_dict.items()}
        spec = state_dict_spec(c)
        missing = [k for k in spec if k not in sd]
        if strict and missing:
            raise RuntimeError(f"text encoder state_dict: missing keys {missing[:6]}{'...' if len(missing) > 6 else ''}")
        with torch.cuda.device(self.device):
            for name, shape in spec.items():
                if name not in sd:
                    continue
                t = sd[name].detach().to("cpu", torch.float32).contiguous()
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {tuple(shape)}")
                nat.check(lib.ec_textenc_set_tensor(self._handle, name.encode(), t.data_ptr(), t.numel()),
                          f"ec_textenc_set_tensor({name})")
            nat.check(lib.ec_textenc_finalize(self._handle), "ec_textenc_finalize")
        self._loaded = True
        return self

    @torch.no_grad()
    def encode_ids(self, input_ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        """input_ids int [N, L] and lengths int [N] (real tokens per row) on the device -> f32 [N, hidden]."""
        if not self._loaded:
            raise RuntimeError("text encoder weights are not loaded: call load_state_dict()")
        if not input_ids.is_cuda:
            raise nat.NativeError("expected CUDA tensors (there is no CPU path)")
        ids = input_ids.to(torch.int32).contiguous()
        lens = lengths.to(device=ids.device, dtype=torch.int32).contiguous()
        N, L = ids.shape
        out = torch.empty((N, self.cfg.hidden), device=ids.device, dtype=torch.float32)
        if N == 0:
            return out
        lib = nat.lib()
        need = lib.ec_textenc_workspace_bytes(self._handle, N, L)
        if self._ws is None or self._ws.numel() < need or self._ws.device != ids.device:
            self._ws = torch.empty((need,), device=ids.device, dtype=torch.uint8)
        with torch.cuda.device(ids.device):
            nat.check(lib.ec_textenc_encode(self._handle, ids.data_ptr(), lens.data_ptr(), N, L, out.data_ptr(),
                                            self._ws.data_ptr(), self._ws.numel(),
                                            torch.cuda.current_stream(ids.device).cuda_stream), "ec_textenc_encode")
        return out

    def encode(self, sentences, convert_to_tensor: bool = False, convert_to_numpy: bool = True, batch_size: int = 4096,
               **_):
        """sentence-transformers `encode` protocol (str or list[str]); needs the injected WordPiece tokenizer."""
        if self.tokenizer is None:
            raise RuntimeError("no WordPiece tokenizer: pass tokenizer= (e.g. transformers.BertTokenizerFast(vocab_file))")
        single = isinstance(sentences, str)
        sents: Sequence[str] = [sentences] if single else list(sentences)
        outs = []
        for s in range(0, len(sents), batch_size):
            enc = self.tokenizer(list(sents[s:s + batch_size]))
            ids_l = [list(x)[:self.max_seq_length] for x in enc["input_ids"]]
            L = max(len(x) for x in ids_l)
            ids = np.zeros((len(ids_l), L), dtype=np.int32)
            for i, x in enumerate(ids_l):
                ids[i, :len(x)] = x
            lens = np.array([len(x) for x in ids_l], dtype=np.int32)
            outs.append(self.encode_ids(torch.from_numpy(ids).to(self.device), torch.from_numpy(lens).to(self.device)))
        emb = torch.cat(outs) if outs else torch.zeros((0, self.cfg.hidden), device=self.device)
        if single:
            emb = emb[0]
        if convert_to_tensor:
            return emb
        return emb.cpu().numpy() if convert_to_numpy else emb
