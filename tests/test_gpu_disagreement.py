"""GPU: per-instance disagreement vs the oracle -- grouping bit-exact, scores within 1e-3 (north_star)."""
import numpy as np
import pytest
import torch

from embodied_captioning_b200 import disagreement as dis
from embodied_captioning_b200 import ops
from oracle import disagreement_oracle as do

pytestmark = pytest.mark.gpu
TOL = 1e-3      # tolerance stated by BASELINE.json north_star for per-instance disagreement scores


def _case(n, d, n_inst, seed, key_scale=1):
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(n_inst, d)).astype(np.float32)
    inst = rng.integers(0, n_inst, size=n)
    emb = (centers[inst] + 0.5 * rng.normal(size=(n, d))).astype(np.float32)
    keys = (inst.astype(np.int64) * key_scale + (7 if key_scale > 1 else 0))
    return emb, keys


@pytest.mark.parametrize("n,d,n_inst,ks", [(64, 512, 8, 1), (1, 384, 1, 1), (500, 384, 37, 2 ** 33),
                                           (3000, 768, 50, 1), (2049, 128, 300, 1), (4100, 768, 4100, 3)])
@pytest.mark.parametrize("mode", [0, 1])
def test_matches_oracle(device, n, d, n_inst, ks, mode):
    emb, keys = _case(n, d, n_inst, seed=n + d, key_scale=ks)
    k, s, c = ops.instance_disagreement(torch.from_numpy(emb).to(device), torch.from_numpy(keys).to(device), mode)
    rk, rs, rc = do.instance_disagreement(emb, keys, mode="full" if mode == 0 else "pairs")
    assert k.cpu().tolist() == rk.tolist()                 # instance grouping: bit-exact
    assert c.cpu().tolist() == rc.tolist()
    got = s.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(rs))
    ok = ~np.isnan(rs)
    assert np.abs(got[ok] - rs[ok]).max(initial=0.0) < TOL


def test_permutation_invariance_and_determinism(device):
    emb, keys = _case(5000, 768, 64, seed=5)
    e, k = torch.from_numpy(emb).to(device), torch.from_numpy(keys).to(device)
    _, s1, _ = ops.instance_disagreement(e, k, 0)
    _, s1b, _ = ops.instance_disagreement(e, k, 0)
    assert torch.equal(s1, s1b)                            # run-to-run bit-identical
    perm = torch.randperm(5000, device=device)
    _, s2, _ = ops.instance_disagreement(e[perm].contiguous(), k[perm].contiguous(), 0)
    assert (s1 - s2).abs().max().item() < 1e-5


def test_identity_between_conventions_at_scale(device):
    """A11 = (V-1)/V (1 - A12) on a K3-shaped problem (100k crops, 2k instances, D=768)."""
    rng = np.random.default_rng(9)
    n, d, m = 100_000, 768, 2000
    inst = torch.from_numpy(rng.integers(0, m, size=n)).to(device)
    emb = torch.randn(n, d, device=device) + torch.randn(m, d, device=device)[inst] * 2
    k0, s0, c0 = ops.instance_disagreement(emb, inst, 0)
    k1, s1, c1 = ops.instance_disagreement(emb, inst, 1)
    assert torch.equal(k0, k1) and torch.equal(c0, c1) and int(c0.sum()) == n
    assert torch.equal(k0, torch.unique(inst))
    v = c0.float()
    assert ((s0 - (v - 1) / v * (1 - s1)).abs().max().item()) < 1e-4


def test_reference_named_functions(device):
    mat = torch.randn(384, 6)
    ref = float(do.cosine_distance(mat))
    assert abs(float(dis._cosine_distance(mat)) - ref) < TOL
    assert abs(float(dis._cosine_distance(mat.to(device))) - ref) < TOL
    assert abs(float(dis._cosine_distance(torch.randn(384, 1)))) < 1e-5      # V = 1 -> ~0
    e = np.random.default_rng(3).normal(size=(5, 384)).astype(np.float32)
    assert abs(dis.mean_cosine_similarity(e) - do.mean_cosine_similarity_embeddings(e)) < TOL
    assert np.isnan(dis.mean_cosine_similarity(e[:1]))

    class Enc:
        def encode(self, caps, convert_to_numpy=True):
            return e[:len(caps)]
    dis.set_encoder(Enc())
    assert abs(dis.mean_cosine_similarity(["a", "b", "c"]) - do.mean_cosine_similarity_embeddings(e[:3])) < TOL


@pytest.mark.parametrize("mode", [0, 1])
def test_duplicate_view_removal_matches_unique_dim0(device, mode):
    """`unique(dim=0)` of SemanticPointCloud.preprocess (projection_utils.py:574-578): exact-duplicate views of an
    instance are dropped before the reduction; counts are the numbers of DISTINCT views (bit-exact)."""
    rng = np.random.default_rng(17)
    n, d, m = 3000, 384, 40
    inst = rng.integers(0, m, size=n)
    base = rng.normal(size=(200, d)).astype(np.float32)          # only 200 distinct rows -> many exact duplicates
    emb = base[rng.integers(0, 200, size=n)].copy()
    emb[5] = 0.0; emb[6] = -0.0; inst[5] = inst[6] = 3            # -0.0 == 0.0 for torch.unique
    emb[7, 0] = 1e-30
    keys = inst.astype(np.int64) * 1000003
    k, s, c = ops.instance_disagreement(torch.from_numpy(emb).to(device), torch.from_numpy(keys).to(device), mode,
                                        dedup=True)
    # zero rows have no direction (0/0): keep them out of the score comparison but inside the count comparison
    rk, rs, rc = do.instance_disagreement(emb, keys, mode="full" if mode == 0 else "pairs", dedup=True)
    assert k.cpu().tolist() == rk.tolist()
    assert c.cpu().tolist() == rc.tolist()
    assert int(c.sum()) < n
    got = s.cpu().numpy().astype(np.float64)
    ok = ~np.isnan(rs) & ~np.isnan(got) & (rk != 3 * 1000003)
    assert ok.sum() >= m - 2
    assert np.abs(got[ok] - rs[ok]).max() < TOL
    # without the flag nothing is dropped
    _, _, c_all = ops.instance_disagreement(torch.from_numpy(emb).to(device), torch.from_numpy(keys).to(device), mode)
    assert int(c_all.sum()) == n


def test_disagreement_map_and_reward_match_numpy_semantics(device):
    """N2: score gather + centre write + six inflate writes + reward, against numpy executing the reference's own
    assignments (oracle/raster_oracle.py): last write wins, negative indices wrap, out-of-range assignment dropped."""
    from oracle import raster_oracle as ro
    rng = np.random.default_rng(23)
    lower, upper = (-3.0, 0.0, -2.5), (4.2, 3.0, 3.7)
    for case in range(4):
        P, M = 20000, 25
        pts = np.stack([rng.uniform(lower[0] + 0.2, upper[0] - 0.2, P), rng.uniform(0, 3, P),
                        rng.uniform(lower[2] + 0.2, upper[2] - 0.2, P)], axis=1).astype(np.float32)
        obj = rng.integers(0, M, size=P).astype(np.int64) + 5_000_000          # detection ids (matching.py:156-171)
        if case == 1:        # points just left of / above the origin: negative indices that wrap
            pts[:50, 0] = lower[0] - 0.03 - 0.04 * rng.uniform(0.0, 1.0, 50).astype(np.float32)
            pts[50:90, 2] = lower[2] - 0.06
        if case == 2:        # one point far outside: every assignment touching it is dropped whole
            pts[0] = (upper[0] + 50.0, 1.0, 0.0)
        if case == 3:        # a point on the last row/column: only SOME inflate assignments leave the grid
            H, W, ox, oy = ro.grid_geometry(lower, upper)
            pts[0] = ((W - 1 + ox + 0.5) * 0.025, 1.0, 0.0)
        scores = rng.uniform(0.0, 0.6, M).astype(np.float32)
        scores[::5] = 0.0                                                      # single-view objects: 0 -> not painted
        ik = np.arange(M, dtype=np.int64) + 5_000_000
        ref_map, ref_reward = ro.disagreement_map(pts, obj, {int(k): torch.tensor(v) for k, v in zip(ik, scores)},
                                                  lower, upper)
        got_map, got_reward = dis.disagreement_map(torch.from_numpy(pts).to(device), torch.from_numpy(obj).to(device),
                                                   torch.from_numpy(ik).to(device), torch.from_numpy(scores).to(device),
                                                   lower, upper)
        assert got_map.shape == ref_map.shape
        assert np.array_equal(got_map.cpu().numpy(), ref_map), f"case {case}: map differs"      # bit-exact
        assert abs(float(got_reward) - ref_reward) <= 1e-5 * max(1.0, abs(ref_reward))
        if case == 2:
            assert ref_map.sum() == 0.0
        if case == 0:
            assert ref_map.sum() > 0
    # empty point cloud
    m0, r0 = dis.disagreement_map(torch.zeros((0, 3), device=device), torch.zeros((0,), dtype=torch.int64, device=device),
                                  torch.zeros((0,), dtype=torch.int64, device=device), torch.zeros((0,), device=device),
                                  lower, upper)
    assert float(m0.abs().sum()) == 0.0 and float(r0) == 0.0
