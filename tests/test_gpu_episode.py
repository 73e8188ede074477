"""GPU: episode-replay shape (BASELINE.json configs[2]) at reduced length -- frames + detector boxes -> crops ->
captions -> cumulative per-instance disagreement with duplicate-view removal after every frame -> map + reward,
against the CPU oracles on the same seeded inputs; and the long-caption shape (configs[4], seq_len 60)."""
import numpy as np
import pytest
import torch

from embodied_captioning_b200 import ops, synth
from embodied_captioning_b200.captioner import transform as tr
from embodied_captioning_b200.config import get_config
from embodied_captioning_b200.engine import CaptionEngine
from embodied_captioning_b200.episode import EpisodeReplay
from oracle import disagreement_oracle as do
from oracle import preprocess_oracle as po
from oracle import raster_oracle as ro
from oracle.coca_oracle import OracleCoCa

pytestmark = pytest.mark.gpu


def _caption_embedding(table):
    """Stand-in for the reference's caption-text encoder (SBERT, absent offline): a deterministic bag-of-tokens
    embedding, so that identical captions give bit-identical embeddings (what `unique(dim=0)` relies on)."""
    def f(tokens, latent):
        e = table.to(tokens.device)[tokens.clamp(min=0)]              # [N, T, D]
        return e.sum(dim=1)
    return f


def test_episode_replay_matches_oracles(device):
    cfg = get_config("coca_tiny")
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = CaptionEngine("coca_tiny", None, device=device, decode_graph=True, state_dict=sd)
    F_, K = 10, 8
    frames, boxes, fidx, ids = synth.make_frames_and_boxes(F_, K, 480, 640, seed=31)
    rng = np.random.default_rng(5)
    ids = ids.copy()
    fresh = rng.random(len(ids)) < 0.1                                 # ~10 % fresh detection ids (matching.py:156-171)
    ids[fresh] = 5_000_000 + np.arange(int(fresh.sum()))
    # repeat some detections exactly (same frame content + same box) so duplicate views really occur
    frames[5] = frames[2]
    boxes[5 * K:6 * K] = boxes[2 * K:3 * K]
    ids[5 * K:6 * K] = ids[2 * K:3 * K]
    perm = rng.permutation(len(ids))                                   # detections arrive in arbitrary order
    table = torch.from_numpy(np.random.default_rng(9).normal(size=(cfg.vocab_size, 384)).astype(np.float32))
    lower, upper = (-2.0, 0.0, -2.0), (2.0, 3.0, 2.0)
    P = 4000
    pts = np.stack([rng.uniform(-1.9, 1.9, P), rng.uniform(0, 3, P), rng.uniform(-1.9, 1.9, P)], 1).astype(np.float32)
    pobj = rng.choice(np.unique(ids), size=P)
    rep = EpisodeReplay(eng, embed_fn=_caption_embedding(table), dedup=True)
    res = rep.run(torch.from_numpy(frames), boxes[perm], fidx[perm], ids[perm], policy=tr.POLICY_AGENT, seq_len=12,
                  chunk_crops=32, points=pts, point_ids=pobj, bounds=(lower, upper))
    ops.check_watchdog(device)
    assert int(res.status.max()) == 0 and len(res.frames) == F_
    order = perm[res.order]                                            # row r of the result = original detection order[r]
    assert np.array_equal(fidx[order], np.sort(fidx))

    # --- crops -> model input: bit-exact against the Pillow-exact oracle (numpy slice + transform)
    ref_x, ref_rects = po.preprocess_boxes(frames, boxes[order], fidx[order], tr.POLICY_AGENT)
    rects, _ = tr.box_rects(boxes[order], 480, 640, tr.POLICY_AGENT)
    assert np.array_equal(rects, ref_rects)
    _, f32, _ = tr.preprocess_boxes(torch.from_numpy(frames).to(device), boxes[order], fidx[order], tr.POLICY_AGENT,
                                    want_f32=True)
    assert np.array_equal(f32.cpu().numpy(), ref_x)

    # --- captions: fp32 oracle on the same inputs
    orc = OracleCoCa(cfg, sd)
    lat_ref, embs_ref = orc.encode_image_tokens(torch.from_numpy(ref_x))
    assert torch.nn.functional.cosine_similarity(res.latent.cpu(), lat_ref).min().item() >= 0.999
    tok_ref = orc.generate(image_embs=embs_ref, seq_len=12, mode="kvcache", fixed_output_length=True)["text"]
    same = (res.tokens.cpu()[:, :tok_ref.shape[1]] == tok_ref).all(dim=1)
    print(f"episode replay: {len(order)} crops, identical greedy sequences on {same.float().mean() * 100:.1f}%")
    # the duplicated detections must give bit-identical captions, hence bit-identical embeddings
    inv = np.empty(len(order), dtype=np.int64)
    inv[order] = np.arange(len(order))                                 # original detection index -> result row
    dup_a, dup_b = inv[2 * K:3 * K], inv[5 * K:6 * K]
    assert torch.equal(res.tokens[dup_a], res.tokens[dup_b])
    assert torch.equal(res.embeddings[dup_a], res.embeddings[dup_b])

    # --- cumulative per-frame disagreement with unique(dim=0), then map + reward, from the GPU's own embeddings
    emb = res.embeddings.cpu().numpy()
    k_sorted = ids[order]
    saw_dedup = False
    for fr in res.frames:
        u = int(np.searchsorted(fidx[order], fr.frame, side="right"))
        rk, rs, rc = do.instance_disagreement(emb[:u], k_sorted[:u], mode="full", dedup=True)
        assert fr.inst_keys.cpu().tolist() == rk.tolist()              # grouping: bit-exact
        assert fr.inst_counts.cpu().tolist() == rc.tolist()            # distinct-view counts: bit-exact
        _, _, rc_all = do.instance_disagreement(emb[:u], k_sorted[:u], mode="full", dedup=False)
        saw_dedup |= bool((rc_all != rc).any())
        assert np.abs(fr.inst_scores.cpu().numpy() - rs).max() < 1e-3
        seen = np.isin(pobj, rk)
        # the oracle map is painted from the GPU's scores so that the map comparison is bit-exact on its own
        cd = {int(k): torch.tensor(float(s), dtype=torch.float32) for k, s in zip(rk, fr.inst_scores.cpu().numpy())}
        ref_map, ref_reward = ro.disagreement_map(pts[seen], pobj[seen], cd, lower, upper)
        assert np.array_equal(fr.dmap.cpu().numpy(), ref_map)
        assert abs(float(fr.reward) - ref_reward) <= 1e-5 * max(1.0, abs(ref_reward))
    assert saw_dedup, "the test episode must exercise duplicate-view removal"


def test_long_caption_shape_seq_len_60(device):
    """BASELINE.json configs[4]: max 60 tokens (the reference never passes seq_len, SURVEY.md A7: an extrapolation
    of the same loop).  KV caches, positional table and forced EOS at index 59 against the fp32 oracle."""
    cfg = get_config("coca_small")
    sd = synth.make_state_dict(cfg, seed=77)
    eng = CaptionEngine("coca_small", None, device=device, decode_graph=True, state_dict=sd)
    crops, ids = synth.make_instance_crops(4, 5, seed=8)               # 5 views per instance
    x = torch.from_numpy(np.stack([po.transform(c) for c in crops]))
    out = eng.caption_crops(torch.from_numpy(crops).to(device), torch.from_numpy(ids).to(device), seq_len=60)
    ops.check_watchdog(device)
    tok = out["tokens"].cpu()
    assert tok.shape == (20, 60) and (tok[:, 0] == cfg.sot_token_id).all()
    assert (tok[:, 1:5] != cfg.eos_token_id).all()
    # a row ends on EOS *or* on a sampled PAD id (mask = last in {EOS, PAD}, coca_model.py:304); the others run
    # to the end and get the forced EOS at the last slot (coca_model.py:317-318)
    ended = ((tok[:, 1:59] == cfg.eos_token_id) | (tok[:, 1:59] == cfg.pad_token_id)).any(dim=1)
    assert (~ended).any()
    assert (tok[~ended, 59] == cfg.eos_token_id).all()
    assert (tok[ended, 59] == cfg.pad_token_id).all()
    orc = OracleCoCa(cfg, sd)
    _, embs_ref = orc.encode_image_tokens(x)
    ref = orc.generate(image_embs=embs_ref, seq_len=60, mode="kvcache", fixed_output_length=True)["text"]
    # compare up to the first near-tie split per row: random-init logits are almost flat
    agree = (tok == ref)
    first_split = torch.where(agree.all(dim=1), torch.full((20,), 60), (~agree).float().argmax(dim=1))
    print(f"seq_len 60: identical sequences {agree.all(dim=1).float().mean() * 100:.0f}%, "
          f"median first split at token {int(first_split.median())}")
    assert int(first_split.median()) >= 20
    with pytest.raises(Exception):
        eng.model.generate(None, image_embs=out["latent"][:, None, :].expand(-1, cfg.n_image_embs, -1).contiguous(),
                           generation_type="top_k", seq_len=cfg.context_length + 5)


def test_caption_frequency_table_matches_string_oracle(device):
    """N4: per-instance caption histogram on token rows == counting strings (pseudocaptioner.py:156-177)."""
    from embodied_captioning_b200.captioner import pseudocaptioner as pc
    from oracle import pseudocaption_oracle as pco
    rng = np.random.default_rng(41)
    n, T, n_caps = 5000, 30, 60
    vocab = rng.integers(1, 900, size=(n_caps, T))
    vocab[:, 0] = 1022
    for c in range(n_caps):                                           # EOS then PAD, like greedy decode output
        L = int(rng.integers(6, T - 1))
        vocab[c, L] = 1023; vocab[c, L + 1:] = 0
    vocab[7] = vocab[3]; vocab[7, 4] += 1                              # near-duplicate captions
    which = rng.integers(0, n_caps, size=n)
    tokens = vocab[which]
    eids = rng.integers(0, 3, size=n); oids = rng.integers(0, 40, size=n)
    words = ["chair", "table", "lamp", "dog", "sofa", "room", "shelf", "plant"]
    def decode(row):
        return " ".join(words[int(t) % len(words)] for t in row[1:] if t not in (0, 1023))
    got = pc.compute_captions_frequency(torch.from_numpy(tokens).to(device), eids, oids, decode)
    ref = pco.captions_frequency([decode(r) for r in tokens], eids, oids, pc.BANNED_WORDS)
    assert list(got.keys()) == list(ref.keys())                        # dict insertion order of the instances
    assert got == ref
    got_all = pc.compute_captions_frequency(torch.from_numpy(tokens).to(device), eids, oids, decode, apply_filter=False)
    ref_all = pco.captions_frequency([decode(r) for r in tokens], eids, oids, (), apply_filter=False)
    assert got_all == ref_all and sum(f for v in got_all.values() for f, _ in v) == n
    # raw kernel output: ordered by (key, first occurrence), counts add up, first rows are first occurrences
    keys = torch.from_numpy((eids.astype(np.int64) << 32) | oids.astype(np.int64)).to(device)
    k, first, cnt = ops.caption_frequency(torch.from_numpy(tokens).to(device), keys)
    k, first, cnt = k.cpu().numpy(), first.cpu().numpy(), cnt.cpu().numpy()
    assert int(cnt.sum()) == n
    assert all((k[i], first[i]) < (k[i + 1], first[i + 1]) for i in range(len(k) - 1))
    seen = {}
    for i in range(n):
        seen.setdefault((int(keys[i]), tokens[i].tobytes()), i)
    assert sorted(seen.values()) == sorted(first.tolist())
    e0 = ops.caption_frequency(torch.zeros((0, T), dtype=torch.int32, device=device),
                               torch.zeros((0,), dtype=torch.int64, device=device))
    assert len(e0[0]) == 0


def test_evaluate_csv_end_to_end_vs_reference_loop(device, tmp_path):
    """Drop-in for scripts/evaluate_finetuned_model.py:150-160: detection .npz files + test CSV written with dataset_io,
    `evaluate_csv(engine, ...)`, the destination CSV read back; compared with the reference's per-row loop restated on
    the same files (real PIL crop with the +-10 px policy :33-37, Pillow-exact transform, fp32 oracle generate) -- float
    boxes included.  Weights carry a trained-like head so that token equality is a meaningful bar."""
    import ast
    import csv

    from PIL import Image

    from embodied_captioning_b200 import dataset_io as dio
    from embodied_captioning_b200.captioner.models.coca.factory import create_model
    from embodied_captioning_b200.engine import CaptionEngine

    cfg = get_config("coca_small")
    sd = synth.make_state_dict(cfg, seed=808)
    rng = np.random.default_rng(11)
    rows, files = [], []
    for f in range(6):
        h, w = (120, 160) if f % 3 else (96, 96)
        img = synth.make_crops(1, size=max(h, w), seed=900 + f)[0][:h, :w].copy()
        n = int(rng.integers(1, 4))
        x1 = rng.uniform(0, w - 40, n); y1 = rng.uniform(0, h - 40, n)
        boxes = np.stack([x1, y1, x1 + rng.uniform(24, 60, n), y1 + rng.uniform(24, 60, n)], 1)
        boxes[:, 2] = np.minimum(boxes[:, 2], w); boxes[:, 3] = np.minimum(boxes[:, 3], h)
        if f % 2 == 0:
            boxes = np.round(boxes)                      # integer boxes in half of the files, float ones in the rest
        infos = [{"id_episode": 3, "id_object": int(rng.integers(0, 5))} for _ in range(n)]
        p = dio.save_detections(str(tmp_path / f"det_{f}.npz"), img, boxes, infos, [f"ref caption {f} {i}" for i in range(n)])
        files.append(p)
        for i in range(n):
            bb = [float(v) for v in boxes[i]]
            rows.append({"filename": p, "episode_id": 3, "object_id": infos[i]["id_object"],
                         "bounding_box": str([int(v) for v in bb] if f % 2 == 0 else bb), "caption": f"ref caption {f} {i}"})
    test_csv, dest_csv = str(tmp_path / "test.csv"), str(tmp_path / "dest.csv")
    with open(test_csv, "w", newline="") as fh:
        wr = csv.DictWriter(fh, fieldnames=["filename", "episode_id", "object_id", "bounding_box", "caption"])
        wr.writeheader()
        wr.writerows(rows)

    # reference loop on the same files: PIL crop -> transform -> (teach the head) -> generate
    crops = []
    for r in rows:
        rgb = np.load(r["filename"], allow_pickle=True)["arr_0"].item()["image"]
        im = Image.fromarray(rgb).convert("RGB")
        bb = ast.literal_eval(r["bounding_box"])
        b1 = [bb[0] - 10 if (bb[0] - 10) >= 0 else 0, bb[1] - 10 if (bb[1] - 10) >= 0 else bb[1],
              bb[2] + 10 if (bb[2] + 10) <= im.size[0] else bb[2], bb[3] + 10 if (bb[3] + 10) <= im.size[1] else bb[3]]
        crops.append(po.transform(np.asarray(im.crop(b1))))
    x = torch.from_numpy(np.stack(crops))
    seq_len = 14
    orc0 = OracleCoCa(cfg, sd)
    _, embs = orc0.encode_image_tokens(x)
    g = torch.Generator().manual_seed(5)
    text = torch.randint(1, cfg.vocab_size - 2, (len(rows), seq_len), generator=g)
    text[:, 0] = cfg.sot_token_id
    lens = torch.randint(6, seq_len - 1, (len(rows),), generator=g)
    for i in range(len(rows)):
        text[i, lens[i]] = cfg.eos_token_id
        text[i, lens[i] + 1:] = cfg.pad_token_id
    H = orc0.decoder_hidden(embs, text[:, :-1])
    valid = text[:, 1:] != cfg.pad_token_id
    sd = dict(sd)
    sd["text_decoder.text_projection"] = synth.fit_confident_head(H[valid], text[:, 1:][valid], cfg.vocab_size)
    ref = OracleCoCa(cfg, sd).generate(image_embs=embs, seq_len=seq_len, mode="kvcache", keep_logits=False,
                                       fixed_output_length=True)["text"]

    eng = CaptionEngine("coca_small", None, device=device, decode_graph=False, state_dict=sd)
    render = lambda toks: " ".join(str(int(t)) for t in toks)
    n = dio.evaluate_csv(eng, test_csv, dest_csv, decode=render, seq_len=seq_len, max_crops=5)
    ops.check_watchdog(device)
    assert n == len(rows)
    back = list(csv.DictReader(open(dest_csv, newline="")))
    assert list(back[0].keys()) == dio.CSV_HEADER
    same = 0
    for r, b, t in zip(rows, back, ref):
        assert (b["filename"], b["bounding_box"], b["reference_caption"]) == (r["filename"], r["bounding_box"], r["caption"])
        assert (int(b["episode_id"]), int(b["object_id"])) == (r["episode_id"], r["object_id"])
        same += int(b["proposed_caption"] == render(t))
    assert same == len(rows), f"{same}/{len(rows)} captions identical to the reference loop"
    with pytest.raises(ValueError):
        create_model("coca_small", precision="int8")


@pytest.mark.gpu
def test_frame_without_detections_returns_empty_results(device):
    """A frame with no boxes (the reference's per-box loops do not run, goal_exploration.py:95-105): empty tensors of
    the right shapes and dtypes, no native call; a single crop and a single-view instance still work."""
    sd = synth.make_state_dict(get_config("coca_tiny"), seed=3)
    eng = CaptionEngine("coca_tiny", None, device=device, decode_graph=True, state_dict=sd)
    frames = torch.zeros((1, 48, 64, 3), dtype=torch.uint8, device=device)
    out = eng.caption_boxes(frames, np.zeros((0, 4), np.float32), np.zeros((0,), np.int32), keys=np.zeros((0,), np.int64))
    assert out["tokens"].shape == (0, 30) and out["latent"].shape[0] == 0 and out["status"].numel() == 0
    assert out["inst_keys"].numel() == 0 and out["inst_scores"].numel() == 0 and out["inst_counts"].numel() == 0
    one = eng.caption_boxes(frames, np.array([[4, 4, 40, 40]], np.float32), np.zeros((1,), np.int32), keys=np.array([7], np.int64))
    ops.check_watchdog(device)
    assert one["tokens"].shape == (1, 30) and one["inst_keys"].cpu().tolist() == [7] and one["inst_counts"].cpu().tolist() == [1]
