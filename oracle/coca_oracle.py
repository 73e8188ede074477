"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch fp32) of the reference captioner math.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path never does.

PARITY STATUS: the arithmetic of the reference lives in `open_clip_torch` (un-pinned in the reference's
`requirements.txt:15`, not installed here, not vendored), so `coca_model.py` cannot be imported and there are no
reference outputs to compare with.  What CAN be pinned offline is pinned (tests/test_oracle_pins.py): every block of
this file is compared to <= 1e-5 with the real `torch.nn.MultiheadAttention` (self, cross, and the pooler's kdim/vdim
form), `nn.LayerNorm`, `nn.GELU`, `nn.Conv2d`, `nn.Embedding` assembled under upstream's parameter names (strict
`load_state_dict` of the open_clip-layout state dict), and `generate` is compared token for token (ragged logits
included) with the reference's own sampling statements executed on the real transformers `MinLengthLogitsProcessor` /
`RepetitionPenaltyLogitsProcessor` / `TopKLogitsWarper` / `MaxLengthCriteria` and `torch.multinomial`.
PARITY UNPINNED remains true only for open_clip's glue code (the order in which it calls those objects), restated here
from upstream's published modules as wired by the reference:

  * wiring, `forward`, `generate(generation_type='top_k')`:
        experimenting_env/captioner/models/coca/coca_model.py:92-333
  * tower sizes: experimenting_env/captioner/models/coca/model_configs/coca_ViT-L-14.json
  * upstream semantics restated: VisionTransformer with
    AttentionalPooler (k = v = ln_k(x), q = ln_q(query)), TextTransformer(embed_cls=True),
    MultimodalTransformer (self block then cross block with ln_1_kv, each with its own MLP),
    nn.MultiheadAttention arithmetic, exact-erf GELU, LayerNorm eps 1e-5.
  * logits processors: transformers `MinLengthLogitsProcessor`, `TopKLogitsWarper(1)`,
    `RepetitionPenaltyLogitsProcessor(1.0)` (coca_model.py:235-240,271,311-320).

The only reference-held known-answer vectors on this path are the three perplexity examples of
`captioning_predictor.py:66-98`; `tests/test_oracle.py` checks them.  Structural pin: the parameter
counts (638.45 M / 253.56 M) equal upstream's published figures.

`precision="fp32"` is the reference arithmetic.  `precision="bf16_operands"` re-runs the same graph
with every GEMM/attention operand rounded to bf16 at the points where the CUDA path rounds (fp32
accumulation, fp32 residual stream / LayerNorm / softmax); it is used to separate logic errors from
rounding effects in the token-parity tests.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _r(x: torch.Tensor, emulate: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


class OracleCoCa:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], precision: str = "fp32"):
        assert precision in ("fp32", "bf16_operands")
        self.cfg = cfg
        self.sd = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}
        self.emu = precision == "bf16_operands"

    # ---------------------------------------------------------------- primitives
    def _ln(self, x, prefix):
        return F.layer_norm(x, (x.shape[-1],), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5)

    def _lin(self, x, w, b=None):
        y = _r(x, self.emu) @ _r(w, self.emu).t()
        return y if b is None else y + b

    def _attn_core(self, q, k, v, heads, mask=None):
        """q [B,Sq,C], k/v [B,Sk,C] already projected; nn.MultiheadAttention arithmetic."""
        B, Sq, C = q.shape
        Sk = k.shape[1]
        d = C // heads
        q, k, v = (_r(t, self.emu) for t in (q, k, v))
        q = q.view(B, Sq, heads, d).transpose(1, 2)
        k = k.view(B, Sk, heads, d).transpose(1, 2)
        v = v.view(B, Sk, heads, d).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
        if mask is not None:
            s = s + mask
        p = torch.softmax(s, dim=-1)
        o = _r(p, self.emu) @ v
        return _r(o.transpose(1, 2).reshape(B, Sq, C), self.emu)

    def _mha_self(self, x, prefix, heads, mask=None):
        C = x.shape[-1]
        qkv = self._lin(x, self.sd[prefix + ".in_proj_weight"], self.sd[prefix + ".in_proj_bias"])
        q, k, v = qkv.split(C, dim=-1)
        o = self._attn_core(q, k, v, heads, mask)
        return self._lin(o, self.sd[prefix + ".out_proj.weight"], self.sd[prefix + ".out_proj.bias"])

    def _mlp(self, x, prefix):
        h = self._lin(x, self.sd[prefix + ".c_fc.weight"], self.sd[prefix + ".c_fc.bias"])
        h = F.gelu(h)  # exact erf form: coca_model.py:65 (nn.GELU)
        return self._lin(h, self.sd[prefix + ".c_proj.weight"], self.sd[prefix + ".c_proj.bias"])

    def _block(self, x, prefix, heads, mask=None):
        x = x + self._mha_self(self._ln(x, prefix + ".ln_1"), prefix + ".attn", heads, mask)
        return x + self._mlp(self._ln(x, prefix + ".ln_2"), prefix + ".mlp")

    def _cross_block(self, x, kv, prefix, heads):
        C = x.shape[-1]
        w, b = self.sd[prefix + ".attn.in_proj_weight"], self.sd[prefix + ".attn.in_proj_bias"]
        kvn = self._ln(kv, prefix + ".ln_1_kv")
        q = self._lin(self._ln(x, prefix + ".ln_1"), w[:C], b[:C])
        k = self._lin(kvn, w[C:2 * C], b[C:2 * C])
        v = self._lin(kvn, w[2 * C:], b[2 * C:])
        o = self._attn_core(q, k, v, heads)
        x = x + self._lin(o, self.sd[prefix + ".attn.out_proj.weight"], self.sd[prefix + ".attn.out_proj.bias"])
        return x + self._mlp(self._ln(x, prefix + ".ln_2"), prefix + ".mlp")

    # ---------------------------------------------------------------- vision tower (A3)
    def patchify(self, images: torch.Tensor) -> torch.Tensor:
        """[B,3,H,W] -> [B, G*G, 3*P*P] in conv-weight (c, py, px) order."""
        P = self.cfg.patch_size
        B, C, H, W = images.shape
        x = images.reshape(B, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5)
        return x.reshape(B, (H // P) * (W // P), C * P * P)

    def vision_trunk(self, images: torch.Tensor) -> torch.Tensor:
        """Patch embedding + class token + positions + ln_pre + the ViT blocks: the tokens [B, 1 + G*G, W] that the
        attentional pooler reads (pinned against transformers' CLIPVisionModel in tests/test_oracle_pins.py)."""
        cfg, sd = self.cfg, self.sd
        images = _r(images.to(torch.float32), self.emu)
        x = self._lin(self.patchify(images), sd["visual.conv1.weight"].reshape(cfg.vision_width, -1))
        cls = sd["visual.class_embedding"].expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]
        x = self._ln(x, "visual.ln_pre")
        for i in range(cfg.vision_layers):
            x = self._block(x, f"visual.transformer.resblocks.{i}", cfg.vision_heads)
        return x

    def encode_image_tokens(self, images: torch.Tensor):
        """coca_model.py:152-155 `_encode_image` (normalize=True) -> (image_latent [B,E], image_embs [B,255,E])."""
        cfg, sd = self.cfg, self.sd
        x = self.vision_trunk(images)
        # AttentionalPooler: q = ln_q(query) (input independent), k = v = ln_k(x)
        E = cfg.embed_dim
        kx = self._ln(x, "visual.attn_pool.ln_k")
        qx = self._ln(sd["visual.attn_pool.query"], "visual.attn_pool.ln_q")
        bias = sd["visual.attn_pool.attn.in_proj_bias"]
        q = self._lin(qx, sd["visual.attn_pool.attn.q_proj_weight"], bias[:E])
        k = self._lin(kx, sd["visual.attn_pool.attn.k_proj_weight"], bias[E:2 * E])
        v = self._lin(kx, sd["visual.attn_pool.attn.v_proj_weight"], bias[2 * E:])
        q = q.unsqueeze(0).expand(x.shape[0], -1, -1)
        o = self._attn_core(q, k, v, cfg.pooler_heads)
        o = self._lin(o, sd["visual.attn_pool.attn.out_proj.weight"], sd["visual.attn_pool.attn.out_proj.bias"])
        o = self._ln(o, "visual.ln_post")
        pooled, tokens = o[:, 0], o[:, 1:]
        pooled = self._lin(pooled, sd["visual.proj"].t())
        return F.normalize(pooled, dim=-1), tokens

    def encode_image(self, images, normalize: bool = True):
        lat, _ = self.encode_image_tokens(images)
        return lat

    # ---------------------------------------------------------------- text tower (A4)
    def _causal(self, n):
        return torch.full((n, n), float("-inf")).triu_(1)

    def text_tokens(self, text: torch.Tensor) -> torch.Tensor:
        """Unimodal text tower, `tokens = x[:, :-1]` output only (the pooled CLS output is unused by
        `generate`).  The CLS pad-mask only alters the CLS row, so rows of real tokens see a pure
        causal mask; the CLS position is therefore omitted here -- `text_tokens_with_cls`
        keeps it and `tests/test_oracle.py` asserts both agree."""
        cfg, sd = self.cfg, self.sd
        t = text.shape[1]
        x = sd["text.token_embedding.weight"][text] + sd["text.positional_embedding"][:t]
        mask = self._causal(t)
        for i in range(cfg.text_layers):
            x = self._block(x, f"text.transformer.resblocks.{i}", cfg.text_heads, mask)
        return x

    def text_tokens_with_cls(self, text: torch.Tensor):
        """Full upstream TextTransformer.forward with the appended CLS token and its pad mask."""
        cfg, sd = self.cfg, self.sd
        B, t = text.shape
        x = sd["text.token_embedding.weight"][text]
        x = torch.cat([x, sd["text.cls_emb"].expand(B, 1, -1)], dim=1)
        x = x + sd["text.positional_embedding"][:t + 1]
        cls_mask = (text != cfg.pad_token_id).unsqueeze(1)                       # [B,1,t]
        cls_mask = F.pad(cls_mask, (1, 0, cls_mask.shape[2], 0), value=True)      # [B,t+1,t+1]
        add = torch.zeros(cls_mask.shape).masked_fill_(~cls_mask, float("-inf"))
        mask = (self._causal(t + 1)[None] + add).unsqueeze(1)                     # broadcast over heads
        for i in range(cfg.text_layers):
            x = self._block(x, f"text.transformer.resblocks.{i}", cfg.text_heads, mask)
        pooled = self._ln(x[:, -1], "text.ln_final")
        pooled = self._lin(pooled, sd["text.text_projection"].t())
        return F.normalize(pooled, dim=-1), x[:, :-1]

    # ---------------------------------------------------------------- multimodal decoder (A5)
    def decoder_logits(self, image_embs: torch.Tensor, token_embs: torch.Tensor) -> torch.Tensor:
        cfg, sd = self.cfg, self.sd
        t = token_embs.shape[1]
        mask = self._causal(t)
        x = token_embs
        for i in range(cfg.mm_layers):
            x = self._block(x, f"text_decoder.resblocks.{i}", cfg.mm_heads, mask)
            x = self._cross_block(x, image_embs, f"text_decoder.cross_attn.{i}", cfg.mm_heads)
        x = self._ln(x, "text_decoder.ln_final")
        return self._lin(x, sd["text_decoder.text_projection"].t())

    def decoder_hidden(self, image_embs: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        """ln_final output of the multimodal decoder for every position of `text` ([B,t,T]); the vocabulary
        projection of these states is `forward_logits`.  Used to fit the trained-like head of the parity tests."""
        cfg = self.cfg
        x = self.text_tokens(text)
        mask = self._causal(x.shape[1])
        for i in range(cfg.mm_layers):
            x = self._block(x, f"text_decoder.resblocks.{i}", cfg.mm_heads, mask)
            x = self._cross_block(x, image_embs, f"text_decoder.cross_attn.{i}", cfg.mm_heads)
        return self._ln(x, "text_decoder.ln_final")

    def forward_logits(self, image_embs, text):
        """coca_model.py:170-203 with output_labels=False -> logits [B,t,V]."""
        return self.decoder_logits(image_embs, self.text_tokens(text))

    # ---------------------------------------------------------------- generate (A7)
    @torch.no_grad()
    def generate(self, images=None, seq_len: int = 30, min_seq_len: int = 5, max_seq_len: int = 77,
                 temperature: float = 1.0, image_embs: Optional[torch.Tensor] = None,
                 mode: str = "reforward", keep_logits: bool = True, fixed_output_length: bool = False):
        """coca_model.py:205-333 for generation_type='top_k', top_k=1.

        mode="reforward": the reference's loop (full forward over the prefix every step).
        mode="kvcache": same arithmetic with cached per-layer K/V (the "fair" CPU variant).
        Returns {"text": int64 [B,T], "logits": list of f32 [B_active,V]} exactly like the reference
        (the logits list is ragged: finished rows drop out, coca_model.py:311-313).
        """
        assert seq_len > min_seq_len, "seq_len must be larger than min_seq_len"
        cfg = self.cfg
        if image_embs is None:
            _, image_embs = self.encode_image_tokens(images)
        B = image_embs.shape[0]
        sot, eos, pad = cfg.sot_token_id, cfg.eos_token_id, cfg.pad_token_id
        text = torch.full((B, 1), sot, dtype=torch.long)
        out_logits: List[torch.Tensor] = []
        cache = _KVCache(self, image_embs) if mode == "kvcache" else None
        while True:
            x = text[:, -max_seq_len:]
            cur_len = x.shape[1]
            if cache is None:
                logits = self.forward_logits(image_embs, x)[:, -1]
            else:
                logits = cache.step(x[:, -1], cur_len - 1)
            mask = (text[:, -1] == eos) | (text[:, -1] == pad)
            sample = torch.full((B, 1), pad, dtype=torch.long)
            if mask.all():
                if not fixed_output_length:
                    break
            else:
                lg = logits[~mask].clone()
                if cur_len < min_seq_len:                 # MinLengthLogitsProcessor(min_seq_len, eos)
                    lg[:, eos] = float("-inf")
                # RepetitionPenaltyLogitsProcessor(1.0) is the identity
                if keep_logits:
                    out_logits.append(lg / temperature)
                # TopKLogitsWarper(1): keep scores >= k-th best (ties kept), softmax, multinomial;
                # with k=1 and no exact tie this is argmax
                if cur_len + 1 == seq_len:
                    sample[~mask, 0] = eos
                else:
                    sample[~mask, 0] = lg.argmax(dim=-1)
            text = torch.cat([text, sample], dim=-1)
            if text.shape[1] >= seq_len:                  # MaxLengthCriteria(seq_len)
                break
        return {"text": text, "logits": out_logits}


class _KVCache:
    """KV-cached single-token step; arithmetic identical to the re-forward path row by row."""

    def __init__(self, m: OracleCoCa, image_embs: torch.Tensor):
        self.m = m
        cfg, sd = m.cfg, m.sd
        C = cfg.text_width
        self.cross = []
        for i in range(cfg.mm_layers):
            p = f"text_decoder.cross_attn.{i}"
            w, b = sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"]
            kvn = m._ln(image_embs, p + ".ln_1_kv")
            self.cross.append((_r(m._lin(kvn, w[C:2 * C], b[C:2 * C]), m.emu),
                               _r(m._lin(kvn, w[2 * C:], b[2 * C:]), m.emu)))
        self.k: Dict[str, List[torch.Tensor]] = {}
        self.v: Dict[str, List[torch.Tensor]] = {}

    def _self_block(self, x, prefix, heads):
        m, sd = self.m, self.m.sd
        C = x.shape[-1]
        h = m._ln(x, prefix + ".ln_1")
        qkv = m._lin(h, sd[prefix + ".attn.in_proj_weight"], sd[prefix + ".attn.in_proj_bias"])
        q, k, v = qkv.split(C, dim=-1)
        self.k.setdefault(prefix, []).append(_r(k, m.emu))
        self.v.setdefault(prefix, []).append(_r(v, m.emu))
        K = torch.cat(self.k[prefix], dim=1)
        V = torch.cat(self.v[prefix], dim=1)
        o = m._attn_core(q, K, V, heads)
        x = x + m._lin(o, sd[prefix + ".attn.out_proj.weight"], sd[prefix + ".attn.out_proj.bias"])
        return x + m._mlp(m._ln(x, prefix + ".ln_2"), prefix + ".mlp")

    def step(self, tok: torch.Tensor, pos: int) -> torch.Tensor:
        m, cfg, sd = self.m, self.m.cfg, self.m.sd
        C = cfg.text_width
        x = (sd["text.token_embedding.weight"][tok] + sd["text.positional_embedding"][pos]).unsqueeze(1)
        for i in range(cfg.text_layers):
            x = self._self_block(x, f"text.transformer.resblocks.{i}", cfg.text_heads)
        for i in range(cfg.mm_layers):
            x = self._self_block(x, f"text_decoder.resblocks.{i}", cfg.mm_heads)
            p = f"text_decoder.cross_attn.{i}"
            w, b = sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"]
            q = m._lin(m._ln(x, p + ".ln_1"), w[:C], b[:C])
            o = m._attn_core(q, self.cross[i][0], self.cross[i][1], cfg.mm_heads)
            x = x + m._lin(o, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
            x = x + m._mlp(m._ln(x, p + ".ln_2"), p + ".mlp")
        x = m._ln(x, "text_decoder.ln_final")
        return m._lin(x[:, 0], sd["text_decoder.text_projection"].t())


def compute_perplexity(logits_list: List[torch.Tensor]) -> torch.Tensor:
    """captioning_predictor.py:34-47: exp(-sum(log max softmax) / T), double."""
    logits = torch.stack([l.to(torch.float32) for l in logits_list], dim=1)  # [n, T, V]
    probs = F.softmax(logits, dim=-1).max(dim=-1).values
    return torch.exp(-probs.log().sum() / probs.shape[1]).double()
