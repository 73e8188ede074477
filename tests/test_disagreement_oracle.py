"""CPU: disagreement oracle vs scipy and the A11/A12 identity."""
import numpy as np
import torch
from scipy.spatial.distance import cosine

from oracle import disagreement_oracle as do


def test_pairs_mean_matches_scipy():
    rng = np.random.default_rng(0)
    e = rng.normal(size=(7, 384)).astype(np.float32)
    ref = np.mean([1 - cosine(e[i], e[j]) for i in range(7) for j in range(i + 1, 7)])
    assert abs(do.mean_cosine_similarity_embeddings(e) - ref) < 1e-6
    assert np.isnan(do.mean_cosine_similarity_embeddings(e[:1]))


def test_full_and_pairs_identity():
    rng = np.random.default_rng(1)
    for V in (2, 3, 8, 31):
        e = rng.normal(size=(V, 96)).astype(np.float32)
        a11 = float(do.cosine_distance(torch.from_numpy(e).T.contiguous()))
        a12 = do.mean_cosine_similarity_embeddings(e)
        assert abs(a11 - (V - 1) / V * (1 - a12)) < 1e-5


def test_single_view_is_zero_up_to_ulp():
    e = torch.randn(1, 384)
    assert abs(float(do.cosine_distance(e.T.contiguous()))) < 1e-6


def test_grouping_and_dedup():
    rng = np.random.default_rng(2)
    emb = rng.normal(size=(12, 32)).astype(np.float32)
    emb[5] = emb[4]                                   # exact duplicate view inside instance 7
    keys = np.array([7, 3, 7, 3, 7, 7, 9, 3, 9, 9, 9, 1], dtype=np.int64)
    k, s, c = do.instance_disagreement(emb, keys, mode="full")
    assert k.tolist() == [1, 3, 7, 9] and c.tolist() == [1, 3, 4, 4]
    k2, s2, c2 = do.instance_disagreement(emb, keys, mode="full", dedup=True)
    assert c2.tolist() == [1, 3, 3, 4]
    perm = rng.permutation(12)
    k3, s3, c3 = do.instance_disagreement(emb[perm], keys[perm], mode="full")
    assert np.allclose(s, s3, atol=1e-6) and k3.tolist() == k.tolist()


def test_string_key_order_is_lexicographic():
    order = do.string_key_group_order([10, 9, 10, 2], [1, 5, 11, 3])
    assert order == [("10", "1"), ("10", "11"), ("2", "3"), ("9", "5")]
