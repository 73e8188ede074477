"""CPU: pins the hand-written model oracle (oracle/coca_oracle.py) to the REAL objects the reference wires.

open_clip is not installed, but every arithmetic object its CoCa is assembled from is:
`torch.nn.MultiheadAttention` (self-attention, cross-attention, and the pooler's `kdim/vdim` form), `nn.LayerNorm`,
`nn.GELU`, `nn.Conv2d`, `nn.Embedding` (coca_model.py:58-81,110-135 build the towers from these through open_clip),
and the sampler objects of `generate` -- transformers `MinLengthLogitsProcessor`, `RepetitionPenaltyLogitsProcessor`,
`TopKLogitsWarper`, `MaxLengthCriteria` and `torch.multinomial` (coca_model.py:235-244,271,294-327).

The modules below are assembled with the parameter NAMES of SURVEY.md Appendix A, so `load_state_dict(strict=True)`
of the synthetic open_clip-layout state dict is itself a check of names and shapes; the forward passes then have to
agree with the oracle to fp32 round-off.  What stays unpinned is only open_clip's own glue code (the order in which it
calls these objects), which cannot be imported offline -- stated in DESIGN.md section 1.
"""
import numpy as np
import pytest
import torch
from torch import nn

from embodied_captioning_b200 import synth
from embodied_captioning_b200.config import get_config
from oracle import preprocess_oracle as po
from oracle.coca_oracle import OracleCoCa

TOL = 1e-5


class ResBlock(nn.Module):
    """open_clip ResidualAttentionBlock as published: pre-LN MHA + pre-LN MLP(c_fc, GELU, c_proj); `is_cross` adds
    ln_1_kv on the key/value input."""

    def __init__(self, width, heads, is_cross=False):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads, batch_first=True)
        if is_cross:
            self.ln_1_kv = nn.LayerNorm(width)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, 4 * width))
        self.mlp.add_module("gelu", nn.GELU())
        self.mlp.add_module("c_proj", nn.Linear(4 * width, width))

    def forward(self, x, kv=None, attn_mask=None):
        q = self.ln_1(x)
        k = v = self.ln_1_kv(kv) if kv is not None else q
        x = x + self.attn(q, k, v, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class Pooler(nn.Module):
    """open_clip AttentionalPooler: learned queries, MHA with kdim = vdim = context width, ln_q / ln_k."""

    def __init__(self, d_model, context_dim, heads, n_queries):
        super().__init__()
        self.query = nn.Parameter(torch.zeros(n_queries, d_model))
        self.attn = nn.MultiheadAttention(d_model, heads, kdim=context_dim, vdim=context_dim, batch_first=True)
        self.ln_q = nn.LayerNorm(d_model)
        self.ln_k = nn.LayerNorm(context_dim)

    def forward(self, x):
        k = self.ln_k(x)
        q = self.ln_q(self.query).unsqueeze(0).expand(x.shape[0], -1, -1)
        return self.attn(q, k, k, need_weights=False)[0]


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.fixture(scope="module")
def tiny():
    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=4321)
    return cfg, sd, OracleCoCa(cfg, sd)


def _causal(n):
    return torch.full((n, n), float("-inf")).triu_(1)


@torch.no_grad()
def test_vit_block_matches_real_multihead_attention(tiny):
    cfg, sd, orc = tiny
    blk = ResBlock(cfg.vision_width, cfg.vision_heads).eval()
    blk.load_state_dict(_sub(sd, "visual.transformer.resblocks.1."), strict=True)
    x = torch.randn(3, cfg.vision_tokens, cfg.vision_width, generator=torch.Generator().manual_seed(1))
    got = orc._block(x, "visual.transformer.resblocks.1", cfg.vision_heads)
    ref = blk(x)
    assert (got - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())


@torch.no_grad()
def test_text_block_causal_mask_matches_real_multihead_attention(tiny):
    cfg, sd, orc = tiny
    blk = ResBlock(cfg.text_width, cfg.text_heads).eval()
    blk.load_state_dict(_sub(sd, "text.transformer.resblocks.0."), strict=True)
    t = 9
    x = torch.randn(4, t, cfg.text_width, generator=torch.Generator().manual_seed(2))
    got = orc._block(x, "text.transformer.resblocks.0", cfg.text_heads, _causal(t))
    ref = blk(x, attn_mask=_causal(t))
    assert (got - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())


@torch.no_grad()
def test_cross_block_matches_real_multihead_attention(tiny):
    cfg, sd, orc = tiny
    blk = ResBlock(cfg.text_width, cfg.mm_heads, is_cross=True).eval()
    blk.load_state_dict(_sub(sd, "text_decoder.cross_attn.1."), strict=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 6, cfg.text_width, generator=g)
    kv = torch.randn(4, cfg.n_image_embs, cfg.embed_dim, generator=g)
    got = orc._cross_block(x, kv, "text_decoder.cross_attn.1", cfg.mm_heads)
    ref = blk(x, kv=kv)
    assert (got - ref).abs().max().item() <= TOL * max(1.0, ref.abs().max().item())


@torch.no_grad()
def test_image_tower_matches_real_torch_modules(tiny):
    """conv1 (nn.Conv2d, stride = patch, no bias) -> cls + pos -> ln_pre -> blocks -> AttentionalPooler (kdim/vdim MHA)
    -> ln_post -> split -> proj -> F.normalize, against OracleCoCa.encode_image_tokens."""
    cfg, sd, orc = tiny
    W, E = cfg.vision_width, cfg.embed_dim
    conv1 = nn.Conv2d(3, W, cfg.patch_size, stride=cfg.patch_size, bias=False)
    conv1.load_state_dict({"weight": sd["visual.conv1.weight"]}, strict=True)
    ln_pre, ln_post = nn.LayerNorm(W), nn.LayerNorm(E)
    ln_pre.load_state_dict(_sub(sd, "visual.ln_pre."), strict=True)
    ln_post.load_state_dict(_sub(sd, "visual.ln_post."), strict=True)
    blocks = [ResBlock(W, cfg.vision_heads).eval() for _ in range(cfg.vision_layers)]
    for i, b in enumerate(blocks):
        b.load_state_dict(_sub(sd, f"visual.transformer.resblocks.{i}."), strict=True)
    pool = Pooler(E, W, cfg.pooler_heads, cfg.n_queries).eval()
    pool.load_state_dict(_sub(sd, "visual.attn_pool."), strict=True)

    crops = synth.make_crops(3, seed=77)
    img = torch.from_numpy(np.stack([po.transform(c) for c in crops]))
    x = conv1(img)                                             # [B, W, G, G]
    x = x.reshape(x.shape[0], W, -1).permute(0, 2, 1)
    x = torch.cat([sd["visual.class_embedding"].expand(x.shape[0], 1, -1), x], dim=1) + sd["visual.positional_embedding"]
    x = ln_pre(x)
    for b in blocks:
        x = b(x)
    x = ln_post(pool(x))
    pooled, tokens = x[:, 0], x[:, 1:]
    latent = torch.nn.functional.normalize(pooled @ sd["visual.proj"], dim=-1)
    got_lat, got_tok = orc.encode_image_tokens(img)
    assert got_tok.shape == (3, cfg.n_image_embs, E)
    assert (got_tok - tokens).abs().max().item() <= 2 * TOL * max(1.0, tokens.abs().max().item())
    assert (got_lat - latent).abs().max().item() <= 2 * TOL


@torch.no_grad()
def test_forward_logits_match_real_modules_end_to_end(tiny):
    """Teacher-forced logits (coca_model.py:170-203, output_labels=False): nn.Embedding + positional table ->
    causal text blocks -> (self, cross) decoder pairs -> ln_final -> text_projection, all on real torch modules."""
    cfg, sd, orc = tiny
    T = cfg.text_width
    emb = nn.Embedding(cfg.vocab_size, T)
    emb.load_state_dict({"weight": sd["text.token_embedding.weight"]}, strict=True)
    tblocks = [ResBlock(T, cfg.text_heads).eval() for _ in range(cfg.text_layers)]
    dself = [ResBlock(T, cfg.mm_heads).eval() for _ in range(cfg.mm_layers)]
    dcross = [ResBlock(T, cfg.mm_heads, is_cross=True).eval() for _ in range(cfg.mm_layers)]
    for i, b in enumerate(tblocks):
        b.load_state_dict(_sub(sd, f"text.transformer.resblocks.{i}."), strict=True)
    for i in range(cfg.mm_layers):
        dself[i].load_state_dict(_sub(sd, f"text_decoder.resblocks.{i}."), strict=True)
        dcross[i].load_state_dict(_sub(sd, f"text_decoder.cross_attn.{i}."), strict=True)
    ln_final = nn.LayerNorm(T)
    ln_final.load_state_dict(_sub(sd, "text_decoder.ln_final."), strict=True)

    g = torch.Generator().manual_seed(5)
    embs = torch.randn(3, cfg.n_image_embs, cfg.embed_dim, generator=g)
    text = torch.randint(1, cfg.vocab_size - 2, (3, 7), generator=g)
    t = text.shape[1]
    x = emb(text) + sd["text.positional_embedding"][:t]
    for b in tblocks:
        x = b(x, attn_mask=_causal(t))
    for s, c in zip(dself, dcross):
        x = c(s(x, attn_mask=_causal(t)), kv=embs)
    ref = ln_final(x) @ sd["text_decoder.text_projection"]
    got = orc.forward_logits(embs, text)
    assert (got - ref).abs().max().item() <= 5 * TOL * max(1.0, ref.abs().max().item())


@torch.no_grad()
@pytest.mark.parametrize("fixed", [False, True])
def test_generate_loop_matches_hf_processors_and_multinomial(tiny, fixed):
    """The reference's sampling statements (coca_model.py:235-244,271,294-327) executed with the REAL transformers
    objects and torch.multinomial on the oracle's logits: with top_k = 1 the warper leaves one finite logit per row, so
    multinomial is deterministic and must reproduce the oracle's arg-max loop token for token, ragged logits included."""
    from transformers import (LogitsProcessorList, MaxLengthCriteria, MinLengthLogitsProcessor,
                              RepetitionPenaltyLogitsProcessor, StoppingCriteriaList, TopKLogitsWarper)
    cfg, sd, orc = tiny
    seq_len, min_seq_len, max_seq_len, temperature = 12, 5, 77, 1.0
    eos, pad, sot = cfg.eos_token_id, cfg.pad_token_id, cfg.sot_token_id
    g = torch.Generator().manual_seed(11)
    embs = torch.randn(6, cfg.n_image_embs, cfg.embed_dim, generator=g)
    # weights under which rows stop at different steps: with plain random init the decoder state hardly depends on the
    # image (all rows decode alike), so the cross-attention branch is amplified and the EOS column becomes a strong
    # random direction -- EOS is then the arg-max whenever a row's state happens to point along it
    sd2 = dict(sd)
    proj = sd["text_decoder.text_projection"].clone()
    proj[:, eos] = 0.25 * torch.randn(proj.shape[0], generator=torch.Generator().manual_seed(1))
    sd2["text_decoder.text_projection"] = proj
    for i in range(cfg.mm_layers):
        sd2[f"text_decoder.cross_attn.{i}.attn.out_proj.weight"] = sd[f"text_decoder.cross_attn.{i}.attn.out_proj.weight"] * 40
        sd2[f"text_decoder.cross_attn.{i}.attn.in_proj_weight"] = sd[f"text_decoder.cross_attn.{i}.attn.in_proj_weight"] * 4
    orc2 = OracleCoCa(cfg, sd2)

    logit_processor = LogitsProcessorList([MinLengthLogitsProcessor(min_seq_len, torch.tensor([eos])),
                                           RepetitionPenaltyLogitsProcessor(1.0)])
    logit_warper = TopKLogitsWarper(1)
    stopping = StoppingCriteriaList([MaxLengthCriteria(max_length=seq_len)])
    text = torch.ones((embs.shape[0], 1), dtype=torch.long) * sot
    out_logits = []
    while True:
        x = text[:, -max_seq_len:]
        cur_len = x.shape[1]
        logits = orc2.forward_logits(embs, x)[:, -1]
        mask = (text[:, -1] == eos) | (text[:, -1] == pad)
        sample = torch.ones((text.shape[0], 1), dtype=torch.long) * pad
        if mask.all():
            if not fixed:
                break
        else:
            logits = logits[~mask, :]
            filtered = logit_processor(x[~mask, :], logits)
            out_logits.append(filtered / temperature)
            filtered = logit_warper(x[~mask, :], filtered)
            probs = torch.softmax(filtered / temperature, dim=-1)
            if cur_len + 1 == seq_len:
                sample[~mask, :] = torch.ones((int((~mask).sum()), 1), dtype=torch.long) * eos
            else:
                sample[~mask, :] = torch.multinomial(probs, 1)
        text = torch.cat((text, sample), dim=-1)
        if all(stopping(text, None)):
            break

    for mode in ("reforward", "kvcache"):
        got = orc2.generate(image_embs=embs, seq_len=seq_len, min_seq_len=min_seq_len, mode=mode,
                            fixed_output_length=fixed)
        assert got["text"].shape == text.shape and torch.equal(got["text"], text), mode
        assert len(got["logits"]) == len(out_logits)
        for a, b in zip(got["logits"], out_logits):
            assert a.shape == b.shape
            fin = torch.isfinite(b)
            assert torch.equal(torch.isfinite(a), fin)
            assert (a[fin] - b[fin]).abs().max().item() <= (1e-4 if mode == "kvcache" else 0.0)
    ends = ((text == eos) | (text == pad)).float().argmax(dim=1)
    assert len(set(ends.tolist())) > 1, "test set-up: rows should stop at different steps"


def test_vision_trunk_matches_transformers_clip_vision_model():
    """The WIRING of the image tower's trunk -- patch conv without bias, class token first, learned positions, ln_pre,
    pre-LN blocks with erf-GELU MLPs -- against an independent third-party implementation of the same published
    architecture: transformers' `CLIPVisionModel` (the OpenAI CLIP ViT that open_clip's `VisionTransformer`
    re-implements; `hidden_act="gelu"` as in open_clip's CoCa configs) loaded with the oracle's weights.  Its
    `last_hidden_state` is the trunk output before any post-norm, i.e. what the attentional pooler reads."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=31)
    W, L = cfg.vision_width, cfg.vision_layers
    hf_cfg = CLIPVisionConfig(hidden_size=W, intermediate_size=4 * W, num_hidden_layers=L,
                              num_attention_heads=cfg.vision_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                              hidden_act="gelu", layer_norm_eps=1e-5, attention_dropout=0.0)
    hf = CLIPVisionModel(hf_cfg).eval()
    m = {"vision_model.embeddings.class_embedding": sd["visual.class_embedding"],
         "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
         "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
         "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"],
         "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"]}
    for i in range(L):
        src, dst = f"visual.transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        w_in, b_in = sd[src + "attn.in_proj_weight"], sd[src + "attn.in_proj_bias"]
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            m[dst + f"self_attn.{nm}.weight"] = w_in[j * W:(j + 1) * W]
            m[dst + f"self_attn.{nm}.bias"] = b_in[j * W:(j + 1) * W]
        m[dst + "self_attn.out_proj.weight"] = sd[src + "attn.out_proj.weight"]
        m[dst + "self_attn.out_proj.bias"] = sd[src + "attn.out_proj.bias"]
        for a, b in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            m[dst + b + ".weight"] = sd[src + a + ".weight"]
            m[dst + b + ".bias"] = sd[src + a + ".bias"]
    own = hf.state_dict()
    missing = [k for k in own if k not in m and "post_layernorm" not in k and "position_ids" not in k]
    assert not missing, missing
    hf.load_state_dict({k: v.clone().float() for k, v in m.items()}, strict=False)
    g = torch.Generator().manual_seed(5)
    images = torch.randn((3, 3, cfg.image_size, cfg.image_size), generator=g)
    with torch.no_grad():
        ref = hf(pixel_values=images).last_hidden_state
    got = OracleCoCa(cfg, sd).vision_trunk(images)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


def test_text_tower_matches_transformers_clip_text_model():
    """The unimodal text tower's trunk (token + position embeddings, causal pre-LN blocks) against transformers'
    `CLIPTextModel` -- the OpenAI CLIP text transformer that open_clip's `TextTransformer` re-implements -- loaded with
    the oracle's weights: the hidden states BEFORE the final LayerNorm are the token rows the multimodal decoder reads
    (`tokens = x[:, :-1]`; the appended CLS token is last, so under the causal mask no token row depends on it)."""
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=32)
    W, L = cfg.text_width, cfg.text_layers
    n_pos = sd["text.positional_embedding"].shape[0]
    hf_cfg = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=W, intermediate_size=4 * W, num_hidden_layers=L,
                            num_attention_heads=cfg.text_heads, max_position_embeddings=n_pos, hidden_act="gelu",
                            layer_norm_eps=1e-5, attention_dropout=0.0, eos_token_id=cfg.eos_token_id,
                            bos_token_id=cfg.sot_token_id, pad_token_id=cfg.pad_token_id)
    hf = CLIPTextModel(hf_cfg).eval()
    m = {"text_model.embeddings.token_embedding.weight": sd["text.token_embedding.weight"],
         "text_model.embeddings.position_embedding.weight": sd["text.positional_embedding"],
         "text_model.final_layer_norm.weight": sd["text.ln_final.weight"],
         "text_model.final_layer_norm.bias": sd["text.ln_final.bias"]}
    for i in range(L):
        src, dst = f"text.transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        w_in, b_in = sd[src + "attn.in_proj_weight"], sd[src + "attn.in_proj_bias"]
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            m[dst + f"self_attn.{nm}.weight"] = w_in[j * W:(j + 1) * W]
            m[dst + f"self_attn.{nm}.bias"] = b_in[j * W:(j + 1) * W]
        m[dst + "self_attn.out_proj.weight"] = sd[src + "attn.out_proj.weight"]
        m[dst + "self_attn.out_proj.bias"] = sd[src + "attn.out_proj.bias"]
        for a, b in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            m[dst + b + ".weight"] = sd[src + a + ".weight"]
            m[dst + b + ".bias"] = sd[src + a + ".bias"]
    missing = [k for k in hf.state_dict() if k not in m and "position_ids" not in k]
    assert not missing, missing
    hf.load_state_dict({k: v.clone().float() for k, v in m.items()}, strict=False)
    g = torch.Generator().manual_seed(6)
    text = torch.randint(1, cfg.vocab_size - 2, (4, 9), generator=g)
    with torch.no_grad():
        ref = hf(input_ids=text, output_hidden_states=True).hidden_states[-1]
    got = OracleCoCa(cfg, sd).text_tokens(text)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
