"""TEST INFRASTRUCTURE ONLY -- CPU fp32 reference of the caption-text encoder.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.

The reference embeds captions with `SentenceTransformer("all-MiniLM-L6-v2")`
(agents/goal_exploration/goal_exploration.py:57,102-103; detector/pseudolabeler.py:568,651,677).  sentence-transformers
is an un-pinned dependency (requirements.txt:16) that is absent here; its pipeline for this model is
  Transformer(BertModel) -> Pooling(mean over attention-masked tokens, sum_mask clamped at 1e-9) -> Normalize(p=2)
[from upstream, unverified offline].  The BERT part IS the real implementation: HF `transformers.BertModel` (installed,
5.5.0), run on the CPU in fp32 with the test's seeded weights.  PARITY: pinned to HF BertModel; the pooling /
normalisation restatement and the absence of the real MiniLM weights remain unpinned.
"""
from __future__ import annotations

import torch


def build_hf_bert(cfg, state_dict):
    from transformers import BertConfig, BertModel
    hc = BertConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                    num_attention_heads=cfg.heads, intermediate_size=cfg.intermediate, hidden_act="gelu",
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                    max_position_embeddings=cfg.max_positions, type_vocab_size=cfg.type_vocab,
                    layer_norm_eps=cfg.layer_norm_eps)
    m = BertModel(hc, add_pooling_layer=False)
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    assert not unexpected, unexpected
    return m.eval().float()


@torch.no_grad()
def sentence_embeddings(model, input_ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """input_ids int64 [N, L], lengths [N] -> f32 [N, hidden]: masked mean pooling + L2 normalisation."""
    L = input_ids.shape[1]
    mask = (torch.arange(L)[None, :] < lengths[:, None]).long()
    tok = model(input_ids=input_ids.long(), attention_mask=mask).last_hidden_state
    m = mask.unsqueeze(-1).to(tok.dtype)
    emb = (tok * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
    return torch.nn.functional.normalize(emb, p=2, dim=1)
