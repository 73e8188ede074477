"""CPU: pins the oracle (test infrastructure) against the only known-answer material the reference holds for
this path and against algebraic self-checks."""
import math

import numpy as np
import pytest
import torch

from embodied_captioning_b200 import synth
from embodied_captioning_b200.config import get_config
from oracle import preprocess_oracle as po
from oracle.coca_oracle import OracleCoCa, compute_perplexity

# the three logits tensors of experimenting_env/captioner/captioning_predictor.py:68-69,79,89 and their targets
PPL_CASES = [
    ([[[0.3659, 0.7025, 0.3104]], [[0.0097, 0.6577, 0.1947]], [[0.5659, 0.0025, 0.0104]], [[0.9097, 0.0577, 0.7947]]],
     [[1], [1], [0], [0]]),
    ([[[0.5659, 0.0025, 0.0104]], [[0.9097, 0.0577, 0.7947]]], [[0], [0]]),
    ([[[0.3659, 0.7025, 0.3104]], [[0.0097, 0.6577, 0.1947]]], [[1], [1]]),
]


def _torcheval_perplexity(inp, target):
    """What torcheval.metrics.text.Perplexity computes: exp(mean negative log-likelihood of the targets)."""
    x = torch.tensor(inp, dtype=torch.float64)
    t = torch.tensor(target)
    logp = torch.log_softmax(x, dim=-1).gather(-1, t.unsqueeze(-1)).squeeze(-1)
    return math.exp(float(-logp.sum() / logp.numel()))


@pytest.mark.parametrize("inp,target", PPL_CASES)
def test_perplexity_known_answers(inp, target):
    expected = _torcheval_perplexity(inp, target)
    x = torch.tensor(inp).permute(1, 0, 2)                 # the reference passes input.permute(1, 0, 2)
    steps = [x[:, i] for i in range(x.shape[1])]
    got = float(compute_perplexity(steps))
    assert got == pytest.approx(expected, rel=1e-3)
    from embodied_captioning_b200.captioner.captioning_predictor import CaptioningPredictor
    p = CaptioningPredictor()
    assert float(p.compute_perplexity(x)) == pytest.approx(expected, rel=1e-3)
    p.outputs["logits"] = steps
    assert float(p.compute_perplexity()) == pytest.approx(expected, rel=1e-3)


def test_parameter_counts_match_upstream():
    assert synth.count_params(get_config("coca_ViT-L-14")) == 638450177      # 638.45 M (open_clip model card)
    assert synth.count_params(get_config("coca_ViT-B-32")) == 253560065      # 253.56 M


@pytest.fixture(scope="module")
def tiny():
    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=1234)
    crops = synth.make_crops(5, seed=5678)
    x = torch.from_numpy(np.stack([po.transform(c) for c in crops]))
    return cfg, sd, x


def test_kvcache_equals_reforward(tiny):
    cfg, sd, x = tiny
    m = OracleCoCa(cfg, sd)
    _, embs = m.encode_image_tokens(x)
    a = m.generate(image_embs=embs, seq_len=12, mode="reforward")
    b = m.generate(image_embs=embs, seq_len=12, mode="kvcache")
    assert torch.equal(a["text"], b["text"])
    for la, lb in zip(a["logits"], b["logits"]):
        fin = torch.isfinite(la)
        assert torch.equal(fin, torch.isfinite(lb))
        assert float((la[fin] - lb[fin]).abs().max()) < 1e-4


def test_cls_token_never_feeds_token_embs(tiny):
    cfg, sd, x = tiny
    m = OracleCoCa(cfg, sd)
    text = torch.tensor([[cfg.sot_token_id, 5, 9, 0, 0], [cfg.sot_token_id, 7, 7, 3, cfg.eos_token_id]])
    _, with_cls = m.text_tokens_with_cls(text)
    assert float((with_cls - m.text_tokens(text)).abs().max()) < 1e-5


def test_generate_invariants(tiny):
    cfg, sd, x = tiny
    m = OracleCoCa(cfg, sd)
    out = m.generate(images=x, seq_len=10, min_seq_len=5, mode="kvcache")
    t = out["text"]
    assert (t[:, 0] == cfg.sot_token_id).all()
    assert (t[:, 1:5] != cfg.eos_token_id).all()            # EOS masked while the prefix is shorter than 5
    assert t.shape[1] <= 10
    if t.shape[1] == 10:
        last = t[:, 9]
        assert ((last == cfg.eos_token_id) | (last == cfg.pad_token_id)).all()
    assert len(out["logits"]) == t.shape[1] - 1


def test_batch_rows_are_independent(tiny):
    cfg, sd, x = tiny
    m = OracleCoCa(cfg, sd)
    full = m.generate(images=x, seq_len=9, mode="kvcache")["text"]
    one = m.generate(images=x[2:3], seq_len=9, mode="kvcache")["text"]
    assert torch.equal(full[2, :one.shape[1]], one[0])


def test_bf16_operand_mode_stays_close(tiny):
    cfg, sd, x = tiny
    lat, embs = OracleCoCa(cfg, sd).encode_image_tokens(x)
    late, embse = OracleCoCa(cfg, sd, precision="bf16_operands").encode_image_tokens(x)
    assert float(torch.nn.functional.cosine_similarity(lat, late).min()) > 0.9999
    assert float((embs - embse).norm() / embs.norm()) < 1e-2
