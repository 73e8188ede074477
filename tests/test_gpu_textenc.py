"""GPU: caption-text encoder (MiniLM-class BERT sentence encoder) against HF `BertModel` in fp32 on
the CPU with the same seeded weights, and the caption -> embedding -> per-instance disagreement chain (A9 -> A11)."""
import numpy as np
import pytest
import torch

from embodied_captioning_b200 import disagreement as dis
from embodied_captioning_b200 import ops
from embodied_captioning_b200 import text_encoder as te
from oracle import disagreement_oracle as do
from oracle import textenc_oracle as teo

pytestmark = pytest.mark.gpu
COS_MIN = 0.999          # same bar as the pooled image embeddings (bf16 GEMM operands vs the fp32 reference)


def _batch(cfg, n, L, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, cfg.vocab_size, (n, L), generator=g)
    lens = torch.randint(2, L + 1, (n,), generator=g)
    lens[0] = L; lens[1] = 1
    ids[:, 0] = 1                                           # [CLS]-like first token
    for i in range(n):
        ids[i, lens[i]:] = 0                                # [PAD]
    return ids, lens


@pytest.mark.parametrize("cfg,n,L", [(te.TINY, 33, 20), (te.MINILM_L6, 64, 32), (te.MINILM_L6, 5, 9)],
                         ids=["tiny", "minilm-l6", "minilm-l6-short"])
def test_sentence_embeddings_match_hf_bert(device, cfg, n, L):
    sd = te.synthetic_state_dict(cfg, seed=7)
    enc = te.CaptionTextEncoder(cfg, device=device).load_state_dict(sd)
    ids, lens = _batch(cfg, n, L, seed=n + L)
    got = enc.encode_ids(ids.to(device), lens.to(device)).cpu()
    ops.check_watchdog(device)
    ref = teo.sentence_embeddings(teo.build_hf_bert(cfg, sd), ids, lens)
    cos = torch.nn.functional.cosine_similarity(got, ref).min().item()
    err = (got - ref).abs().max().item()
    print(f"{n} x {L} tokens, hidden {cfg.hidden}: min cosine {cos:.6f}, max abs err {err:.2e}")
    assert got.shape == (n, cfg.hidden)
    assert torch.allclose(got.norm(dim=1), torch.ones(n), atol=1e-4)
    assert cos >= COS_MIN
    # padding content must not leak into the embedding (prefix attention mask + masked mean)
    ids2 = ids.clone()
    for i in range(n):
        ids2[i, lens[i]:] = 5
    got2 = enc.encode_ids(ids2.to(device), lens.to(device)).cpu()
    assert torch.equal(got, got2)


def test_caption_to_disagreement_chain(device):
    """encode(list[str]) protocol with an injected tokenizer, plugged into `mean_cosine_similarity` /
    `instance_disagreement` (compute_cosine_sim.py:11-22, projection_utils.py:743-754): scores within 1e-3 of the
    fp32 chain (HF BERT -> pooling -> oracle reduction)."""
    cfg = te.TINY
    sd = te.synthetic_state_dict(cfg, seed=11)
    words = {}

    def tokenizer(sents):                                   # stand-in for WordPiece: whitespace words -> ids
        out = []
        for s in sents:
            out.append([1] + [words.setdefault(w, 3 + (len(words) % (cfg.vocab_size - 3))) for w in s.split()] + [2])
        return {"input_ids": out, "attention_mask": [[1] * len(x) for x in out]}

    enc = te.CaptionTextEncoder(cfg, device=device, tokenizer=tokenizer).load_state_dict(sd)
    rng = np.random.default_rng(4)
    vocab = ["a", "red", "wooden", "chair", "table", "lamp", "near", "the", "window", "small", "large", "sofa", "white"]
    caps = [" ".join(rng.choice(vocab, size=int(rng.integers(2, 9)))) for _ in range(60)]
    inst = rng.integers(0, 7, size=60).astype(np.int64)
    emb = enc.encode(caps, convert_to_tensor=True)
    assert emb.shape == (60, cfg.hidden) and emb.is_cuda
    one = enc.encode(caps[3])
    assert isinstance(one, np.ndarray) and one.shape == (cfg.hidden,)
    assert np.allclose(one, emb[3].cpu().numpy(), atol=2e-3)        # batch-size independent up to padding length
    # fp32 reference chain on the same ids
    tk = tokenizer(caps)
    L = max(len(x) for x in tk["input_ids"])
    ids = torch.zeros((60, L), dtype=torch.long)
    for i, x in enumerate(tk["input_ids"]):
        ids[i, :len(x)] = torch.tensor(x)
    lens = torch.tensor([len(x) for x in tk["input_ids"]])
    ref_emb = teo.sentence_embeddings(teo.build_hf_bert(cfg, sd), ids, lens).numpy()
    k, s, c = ops.instance_disagreement(emb, torch.from_numpy(inst).to(device), 0)
    rk, rs, rc = do.instance_disagreement(ref_emb, inst, mode="full")
    assert k.cpu().tolist() == rk.tolist() and c.cpu().tolist() == rc.tolist()
    assert np.abs(s.cpu().numpy() - rs).max() < 1e-3
    dis.set_encoder(enc)
    group = [caps[i] for i in np.flatnonzero(inst == 2)]
    ref_sim = do.mean_cosine_similarity_embeddings(ref_emb[inst == 2])
    assert abs(dis.mean_cosine_similarity(group) - ref_sim) < 1e-3
    dis.set_encoder(None)
