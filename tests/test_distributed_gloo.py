"""CPU, world_size 2, gloo: host logic of the one exchange step (pad, all-gather, key handling, pad removal,
ordering).  The reduction itself is injected (oracle) because the product reduction only exists as CUDA."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import disagreement_oracle as do


def _oracle_reduce(emb, keys, mode):
    k, s, c = do.instance_disagreement(emb.numpy(), keys.numpy(), mode="full" if mode == 0 else "pairs")
    return torch.from_numpy(k), torch.from_numpy(s), torch.from_numpy(c)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from embodied_captioning_b200 import distributed as D
    rng = np.random.default_rng(0)
    emb = torch.from_numpy(rng.normal(size=(n, d)).astype(np.float32))
    keys = torch.from_numpy(rng.integers(0, 9, size=n).astype(np.int64))
    s, e = D.shard_range(n, rank, world)
    out = {}
    for mode in (0, 1):
        k, sc, ct = D.allgather_disagreement(emb[s:e], keys[s:e], n, mode=mode, reduce_fn=_oracle_reduce)
        out[mode] = (k.numpy(), sc.numpy(), ct.numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_disagreement_world2_matches_single_process():
    n, d, world = 37, 16, 2          # 37 rows: ragged shards (19 + 18) exercise the padding path
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    emb = rng.normal(size=(n, d)).astype(np.float32)
    keys = rng.integers(0, 9, size=n).astype(np.int64)
    for mode, name in ((0, "full"), (1, "pairs")):
        k, s, c = do.instance_disagreement(emb, keys, mode=name)
        for r in range(world):
            kk, ss, cc = results[r][mode]
            assert kk.tolist() == k.tolist() and cc.tolist() == c.tolist()
            assert np.allclose(ss, s, atol=1e-6, equal_nan=True)
