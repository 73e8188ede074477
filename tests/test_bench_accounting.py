"""CPU: the accounting behind the bench line's roofline fractions (bench.py) against SURVEY.md section 8(a) / 8(d)."""
import importlib.util
import os
import sys

import pytest

from embodied_captioning_b200 import synth
from embodied_captioning_b200.config import get_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_parameter_counts_are_upstreams_published_figures():
    # the restated model has exactly upstream's parameter counts
    assert synth.count_params(get_config("coca_ViT-L-14")) == 638_450_177
    assert synth.count_params(get_config("coca_ViT-B-32")) == 253_560_065


def test_decoder_algorithmic_bytes_is_the_surveys_0_33_tb(bench):
    # K1, 29 steps, all rows active: ~ 29 x (0.557 + 9.63 + 0.0755 t + 0.0755) GB ~ 0.33 TB
    cfg = get_config("coca_ViT-L-14")
    total = bench.decoder_algorithmic_bytes(cfg, 1024, 29)
    assert abs(total - 330.34e9) < 0.2e9, total
    T, L = cfg.text_width, cfg.mm_layers
    per_step_weights = 2 * (cfg.text_layers * 12 * T * T + L * 12 * T * T + L * 10 * T * T + T * cfg.vocab_size)
    assert abs(per_step_weights - 557.2e6) < 1e6                       # "weights 557.2 MB bf16 once per step per batch"
    cross_per_crop = L * cfg.n_image_embs * 2 * T * 2
    assert abs(cross_per_crop - 9.40e6) < 0.01e6                       # "cross-K/V reads 9.40 MB x B"
    # linear in the batch apart from the weights, and additive over steps
    a, b = bench.decoder_algorithmic_bytes(cfg, 512, 29), bench.decoder_algorithmic_bytes(cfg, 1024, 29)
    assert abs((b - 29 * per_step_weights) - 2 * (a - 29 * per_step_weights)) < 1e3


def test_encoder_flops_per_crop(bench):
    # 163.64 GFLOP per crop (L/14), 9.19 (B/32): re-derived from the configuration
    for name, want in (("coca_ViT-L-14", 163.64), ("coca_ViT-B-32", 9.19)):
        c = get_config(name)
        t, W, E, Q = c.vision_tokens, c.vision_width, c.embed_dim, c.n_queries
        blocks = c.vision_layers * (2 * t * W * (3 * W + W + 8 * W) + 4 * t * t * W)
        patch = 2 * c.n_patches * c.patch_dim * W
        pooler = 2 * t * W * 2 * E + 2 * Q * E * E + 4 * Q * t * E + 2 * E * E
        assert abs((blocks + patch + pooler) / 1e9 - want) < 0.02 * want, (name, (blocks + patch + pooler) / 1e9)
        assert bench.ENC_GFLOP[name] == want


def test_l2_note_names_the_streams_of_the_configuration(bench):
    k1 = bench.l2_note(get_config("coca_ViT-L-14"), 1024, 30)
    k0 = bench.l2_note(get_config("coca_ViT-B-32"), 64, 30)
    assert "1.28 GB of weights" in k1 and "0.40 GB of image tokens" in k1 and "348 times" in k1
    assert "0.51 GB of weights" in k0 and k0 != k1
