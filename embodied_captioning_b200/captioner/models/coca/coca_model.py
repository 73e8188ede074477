"""`CoCa` model object with the reference's public surface, backed by libec_b200 (sm_100a CUDA).

Mirrors `experimenting_env/captioner/models/coca/coca_model.py`:
  * `CoCa.encode_image`            coca_model.py:162-164
  * `CoCa.forward(image, text, image_latent=, image_embs=, output_labels=)`   coca_model.py:170-203
  * `CoCa.generate(image, ..., generation_type='top_k')`                      coca_model.py:205-333
  * `load_state_dict` of an open_clip-layout checkpoint (factory.py:131-180)

Differences that are deliberate: beam search is not implemented (no call
site of the reference uses it; transformers >= 5 removed `BeamSearchScorer`), and `top_k` must be 1
(the only value the reference ever passes: greedy).  `temperature`, `top_p`, `repetition_penalty != 1`
raise instead of silently changing semantics.  There is no CPU execution path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from .... import _native as nat
from ....config import CoCaConfig, OPENAI_MEAN, OPENAI_STD
from ....synth import state_dict_spec

GENERATION_TYPES = {"top_k": "top_k", "top_p": None, "beam_search": None}


def strip_checkpoint_prefixes(state_dict):
    """Keys as the reference's checkpoints carry them: `module.` from DDP (evaluate_finetuned_model.py:141-146) and
    `model.` from the `CoCa` wrapper whose state dict `Captioner` saves and loads (predictor_utils.py:182-185); the
    wrapper prefix is only dropped when every tensor key carries it (no open_clip CoCa tensor starts with `model.`)."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    if sd and all(k.startswith("model.") for k in sd):
        sd = {k[6:]: v for k, v in sd.items()}
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    return sd


class _Visual:
    """Stand-in for `model.visual`: callers only read `.preprocess_cfg` / `.image_size` (factory.py:418)."""

    def __init__(self, cfg: CoCaConfig):
        self.image_size = (cfg.image_size, cfg.image_size)
        self.preprocess_cfg = {"size": self.image_size, "mode": "RGB", "mean": OPENAI_MEAN, "std": OPENAI_STD,
                               "interpolation": "bicubic", "resize_mode": "shortest", "fill_color": 0}


class CoCa(nn.Module):
    MAX_BATCH = 1024      # crops per native call; larger batches are chunked

    def __init__(self, cfg: CoCaConfig, device="cuda:0", pad_id: int = 0):
        super().__init__()
        self.cfg = cfg
        self.pad_id = cfg.pad_token_id if pad_id == 0 else pad_id
        self.context_length = cfg.context_length
        self.visual = _Visual(cfg)
        self._device = nat.cuda_device(device)     # always indexed: 'cuda' = the current device, not device 0
        self.output_device: Optional[torch.device] = None     # set by .to('cpu') for callers that asked for CPU
        self._handle = C.c_void_p()
        self._loaded = False
        self._ws: Dict[str, torch.Tensor] = {}
        self._state_dict_ref: Optional[Dict[str, torch.Tensor]] = None
        self._graph = False
        self._tc_attention = True
        self._chains = 0
        self._absorbed = os.environ.get("EC_ABSORB", "1") != "0"     # A/B switch for probes
        self._seq = None                                             # None: the library default (EC_SEQ, off)
        self._dummy = nn.Parameter(torch.zeros(1), requires_grad=False)   # lets `.parameters()` work

    # ------------------------------------------------------------------ nn.Module protocol
    @property
    def device(self) -> torch.device:
        return self._device

    def to(self, *args, **kwargs):          # `.to(device_str)` at goal_exploration.py:55-56, predictor_utils.py:187
        dev = None
        for a in args:
            if isinstance(a, (str, torch.device)):
                dev = torch.device(a)
        dev = torch.device(kwargs["device"]) if "device" in kwargs else dev
        if dev is None:
            return self
        if dev.type == "cuda":
            new = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
            self.output_device = None
            if new != self._device:
                self._device = new
                self._release()
                if self._state_dict_ref is not None:
                    self.load_state_dict(self._state_dict_ref)
        else:
            # No CPU execution path exists: compute stays on the B200, results are handed back on `dev`.
            self.output_device = dev
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else 0))

    def cpu(self):
        return self.to("cpu")

    def _release(self):
        if self._handle:
            nat.lib().ec_model_destroy(self._handle)
            self._handle = C.c_void_p()
        self._loaded = False
        self._ws.clear()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _ensure_handle(self):
        if self._handle:
            return
        nat.require_cuda()
        c = self.cfg
        ec = nat.EcConfig(c.embed_dim, c.image_size, c.patch_size, c.vision_width, c.vision_layers, c.vision_head_dim,
                          c.pooler_heads, c.n_queries, c.text_width, c.text_heads, c.text_layers, c.mm_layers,
                          c.mm_heads, c.context_length, c.vocab_size, c.sot_token_id, c.eos_token_id, c.pad_token_id)
        nat.check(nat.lib().ec_model_create(C.byref(ec), self._device.index, C.byref(self._handle)),
                  "ec_model_create")
        nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_graph", int(self._graph)), "ec_model_set_option")
        nat.check(nat.lib().ec_model_set_option(self._handle, b"tc_attention", int(self._tc_attention)))
        nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_chains", self._chains))
        nat.check(nat.lib().ec_model_set_option(self._handle, b"absorbed_cross", int(self._absorbed)))
        if self._seq is not None:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_seq", int(self._seq)))

    def set_absorbed_cross(self, enabled: bool):
        """Decoder cross-attention in the absorbed-projection form (default) or over per-layer K/V caches; the
        absorbed weights are built when the weights are loaded, so switching it ON reloads them."""
        enabled = bool(enabled)
        if enabled == self._absorbed:
            return
        self._absorbed = enabled
        if self._handle:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"absorbed_cross", int(enabled)))
            if enabled and self._state_dict_ref is not None:
                self.load_state_dict(self._state_dict_ref)
            self._ws.pop("decode", None)

    def set_decode_seq(self, enabled: bool):
        """Decoder steps with one persistent GEMM-sequence launch per block, or (default: measured faster) one launch
        per GEMM / LayerNorm; the results are bit-identical."""
        self._seq = bool(enabled)
        if self._handle:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_seq", int(self._seq)))

    def set_decode_graph(self, enabled: bool):
        self._graph = bool(enabled)
        if self._handle:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_graph", int(self._graph)))

    def set_decode_chains(self, chains: int):
        """Independent row chains of the decoder step (0 = choose from the batch size)."""
        self._chains = int(chains)
        if self._handle:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"decode_chains", self._chains))

    def set_tc_attention(self, enabled: bool):
        """ViT-block attention on tcgen05/TMEM (default) or on the mma.sync flash kernel."""
        self._tc_attention = bool(enabled)
        if self._handle:
            nat.check(nat.lib().ec_model_set_option(self._handle, b"tc_attention", int(self._tc_attention)))

    def load_state_dict(self, state_dict, strict: bool = True):
        """open_clip-layout tensors (names/shapes: SURVEY.md Appendix A); `module.` prefixes are stripped like
        scripts/evaluate_finetuned_model.py:141-146 does."""
        sd = strip_checkpoint_prefixes(state_dict)
        spec = state_dict_spec(self.cfg)
        missing = [k for k in spec if k not in sd]
        unexpected = [k for k in sd if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for CoCa: missing keys {missing[:8]}"
                               f"{'...' if len(missing) > 8 else ''}, unexpected keys {unexpected[:8]}")
        if self._loaded:
            self._release()       # a finalized native model may be sealed (LayerNorm folded into its weights): start over
        self._ensure_handle()
        lib = nat.lib()
        with torch.cuda.device(self._device):
            for name, shape in spec.items():
                if name not in sd:
                    continue
                t = sd[name].detach().to(device="cpu", dtype=torch.float32).contiguous()
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {name}: checkpoint {tuple(t.shape)} vs model {tuple(shape)}")
                nat.check(lib.ec_model_set_tensor(self._handle, name.encode(), t.data_ptr(), t.numel()),
                          f"ec_model_set_tensor({name})")
            nat.check(lib.ec_model_finalize(self._handle, torch.cuda.current_stream(self._device).cuda_stream),
                      "ec_model_finalize")
        self._loaded = True
        self._state_dict_ref = sd
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def state_dict(self, *a, **k):
        return dict(self._state_dict_ref or {})

    # ------------------------------------------------------------------ workspaces (torch owns device memory)
    def _workspace(self, tag: str, nbytes: int) -> torch.Tensor:
        cur = self._ws.get(tag)
        if cur is None or cur.numel() < nbytes:
            self._ws.pop(tag, None)
            cur = torch.empty((nbytes,), device=self._device, dtype=torch.uint8)
            self._ws[tag] = cur
        return cur

    def _out(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.output_device is None else t.to(self.output_device)

    def _require_loaded(self):
        if not self._loaded:
            raise RuntimeError("CoCa weights are not loaded: call load_state_dict() (or create the model through "
                               "create_model_and_transforms with a checkpoint / 'synthetic')")

    # ------------------------------------------------------------------ image tower
    def _encode_image_dev(self, images: torch.Tensor, want_latent=True, want_embs=True):
        self._require_loaded()
        c = self.cfg
        x = images.to(device=self._device)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        assert x.shape[1:] == (3, c.image_size, c.image_size), f"expected [B,3,{c.image_size},{c.image_size}]"
        x = x.to(torch.bfloat16).contiguous()
        N = x.shape[0]
        lib = nat.lib()
        latent = torch.empty((N, c.embed_dim), device=self._device, dtype=torch.float32) if want_latent else None
        embs = torch.empty((N, c.n_image_embs, c.embed_dim), device=self._device, dtype=torch.float32) if want_embs else None
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            for s in range(0, N, self.MAX_BATCH):
                e = min(N, s + self.MAX_BATCH)
                ws = self._workspace("encode", lib.ec_encode_workspace_bytes(self._handle, e - s))
                nat.check(lib.ec_encode_image(self._handle, x[s:e].data_ptr(), e - s,
                                              None if latent is None else latent[s:e].data_ptr(),
                                              None if embs is None else embs[s:e].data_ptr(), ws.data_ptr(), ws.numel(),
                                              stream), "ec_encode_image")
        return latent, embs

    def _encode_image(self, images, normalize: bool = True):
        if not normalize:
            raise NotImplementedError("normalize=False is not used on the captioning path (coca_model.py:154)")
        latent, embs = self._encode_image_dev(images)
        return self._out(latent), self._out(embs)

    def encode_image(self, images, normalize: bool = True):
        latent, _ = self._encode_image_dev(images, want_embs=False) if normalize else self._encode_image(images, False)
        return self._out(latent)

    def encode_text(self, text, normalize: bool = True):
        raise NotImplementedError("the pooled text latent is not on the captioning hot path (SURVEY.md A4)")

    # ------------------------------------------------------------------ teacher-forced forward
    def forward(self, image, text: Optional[torch.Tensor] = None, image_latent: Optional[torch.Tensor] = None,
                image_embs: Optional[torch.Tensor] = None, output_labels: bool = True):
        if image_latent is None or image_embs is None:
            image_latent, image_embs = self._encode_image_dev(image)
        if text is None:
            return {"image_features": self._out(image_latent), "image_embs": self._out(image_embs)}
        text = text.to(self._device)
        labels = text[:, 1:] if output_labels else None
        tin = text[:, :-1] if output_labels else text           # coca_model.py:187-190
        logits = self._forward_logits_dev(image_embs.to(self._device, torch.float32), tin)
        out = {"image_features": self._out(image_latent), "logits": self._out(logits)}
        if labels is not None:
            out["labels"] = self._out(labels)
        return out

    def _forward_logits_dev(self, embs: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        self._require_loaded()
        c = self.cfg
        embs = embs.contiguous()
        N, t = text.shape
        if N * t > 0:
            lo, hi = int(text.min().item()), int(text.max().item())
            if lo < 0 or hi >= c.vocab_size:      # the embedding gather has no bounds check on the device
                raise nat.NativeError(f"ec_forward_logits failed: EC_ERR_INVALID: token ids must lie in [0, "
                                      f"{c.vocab_size}), got [{lo}, {hi}] (ignore_index / foreign-tokenizer ids?)")
        tx = text.to(torch.int32).contiguous()
        out = torch.empty((t, N, c.vocab_size), device=self._device, dtype=torch.float32)
        lib = nat.lib()
        with torch.cuda.device(self._device):
            ws = self._workspace("decode", lib.ec_decode_workspace_bytes(self._handle, N, t + 1))
            nat.check(lib.ec_forward_logits(self._handle, embs.data_ptr(), tx.data_ptr(), N, t, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream(self._device).cuda_stream), "ec_forward_logits")
        return out.permute(1, 0, 2)

    # ------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(self, image, text=None, seq_len=30, max_seq_len=77, temperature=1., generation_type="beam_search",
                 top_p=0.1, top_k=1, pad_token_id=None, eos_token_id=None, sot_token_id=None, num_beams=6,
                 num_beam_groups=3, min_seq_len=5, stopping_criteria=None, repetition_penalty=1.0,
                 fixed_output_length=False, image_embs: Optional[torch.Tensor] = None, return_logits: bool = True):
        """`generate(image, generation_type='top_k')` of coca_model.py:205-333 (greedy: top_k = 1).

        Returns {"text": int64 [B, T<=seq_len] incl. SOT, "logits": list[T-1] of f32 [B_active, vocab]} -- the
        ragged list of coca_model.py:311-313.  `return_logits=False` (extension) skips materialising logits.
        """
        assert seq_len > min_seq_len, "seq_len must be larger than min_seq_len"          # coca_model.py:228
        if generation_type != "top_k":
            if generation_type in GENERATION_TYPES:
                raise NotImplementedError(f"generation_type={generation_type!r}: only 'top_k' (the path every caller of "
                                          "the reference uses) is implemented")
            raise ValueError("generation_type has to be one of | top_k | top_p | beam_search |.")  # coca_model.py:273-276
        if top_k != 1 or temperature != 1.0 or repetition_penalty != 1.0 or text is not None or stopping_criteria:
            raise NotImplementedError("only greedy decoding (top_k=1, temperature=1, no prompt) is implemented")
        c = self.cfg
        for given, own, nm in ((pad_token_id, c.pad_token_id, "pad"), (eos_token_id, c.eos_token_id, "eos"),
                               (sot_token_id, c.sot_token_id, "sot")):
            if given is not None and int(given) != own:
                raise NotImplementedError(f"{nm}_token_id override differs from the model config")
        self._require_loaded()
        if image_embs is None:
            _, image_embs = self._encode_image_dev(image, want_latent=False)
        embs = image_embs.to(self._device, torch.float32).contiguous()
        N = embs.shape[0]
        tokens = torch.empty((N, seq_len), device=self._device, dtype=torch.int32)
        logits_all: List[Optional[torch.Tensor]] = []
        lib = nat.lib()
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            for s in range(0, N, self.MAX_BATCH):
                e = min(N, s + self.MAX_BATCH)
                lg = (torch.empty((seq_len - 1, e - s, c.vocab_size), device=self._device, dtype=torch.float32)
                      if return_logits else None)
                ws = self._workspace("decode", lib.ec_decode_workspace_bytes(self._handle, e - s, seq_len))
                nat.check(lib.ec_decode_greedy(self._handle, embs[s:e].data_ptr(), e - s, seq_len, min_seq_len,
                                               tokens[s:e].data_ptr(), nat.ptr(lg), ws.data_ptr(), ws.numel(), stream),
                          "ec_decode_greedy")
                logits_all.append(lg)
        tok = tokens.to(torch.int64)
        if fixed_output_length and not return_logits:      # the engine path: nothing below depends on where rows ended
            return {"text": self._out(tok), "logits": []}
        # The reference loop stops as soon as every row has emitted EOS/PAD: T = 1 + last column any row was active.
        ended = (tok[:, 1:] == c.eos_token_id) | (tok[:, 1:] == c.pad_token_id)
        first_end = torch.where(ended.any(dim=1), ended.float().argmax(dim=1) + 1,
                                torch.full((N,), seq_len - 1, device=tok.device))
        T = seq_len if fixed_output_length else int(first_end.max().item()) + 1
        out = {"text": self._out(tok[:, :T]), "logits": []}
        if return_logits:
            lg = logits_all[0] if len(logits_all) == 1 else torch.cat(logits_all, dim=1)
            for step in range(T - 1):
                active = first_end >= (step + 1)           # rows still generating at this step
                if not bool(active.any()):
                    break
                out["logits"].append(self._out(lg[step] if bool(active.all()) else lg[step][active]))
        return out
