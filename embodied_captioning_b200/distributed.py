"""Multi-GPU plumbing: one process per GPU, crops sharded contiguously by batch, one exchange step.

The captioner forward shards with no communication.  The instance reduction needs all
views of an instance together; two forms are provided:

  * `allgather_disagreement`  -- NCCL all-gather of (embeddings, keys) then the replicated segmented
    reduction on every rank (the baseline named by BASELINE.json's north_star).
  * `peer_disagreement`       -- only the 8-byte keys are all-gathered; every rank reduces the instances it
    owns (m % world == rank) by reading the view rows straight out of its peers' HBM over NVLink
    (symmetric-memory peer pointers inside `segment_reduce_kernel`), so each embedding row crosses
    NVLink at most once and no gathered copy is ever materialised; the M scores are then summed.

Rows are reduced in ascending (key, global crop index) order in both forms, so scores are bit-identical for
any number of ranks.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

PAD_KEY = torch.iinfo(torch.int64).max


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of crops owned by `rank` (block sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rows_per_rank(n: int, world: int) -> int:
    return (n + world - 1) // world


def pad_local(emb: torch.Tensor, keys: torch.Tensor, rows: int):
    """Pad the local shard to `rows` rows; padded rows carry PAD_KEY (dropped after the reduction) and a
    unit embedding so that no NaN is ever produced."""
    n, d = emb.shape
    if n == rows:
        return emb.contiguous(), keys.contiguous()
    pe = torch.zeros((rows, d), dtype=emb.dtype, device=emb.device)
    pe[:, 0] = 1.0
    pe[:n] = emb
    pk = torch.full((rows,), PAD_KEY, dtype=torch.int64, device=keys.device)
    pk[:n] = keys
    return pe, pk


def gather_padded(emb: torch.Tensor, keys: torch.Tensor, rows: int, group=None):
    """all-gather of the padded shards -> (emb_all [world*rows, D], keys_all [world*rows])."""
    world = dist.get_world_size(group)
    pe, pk = pad_local(emb, keys, rows)
    emb_all = torch.empty((world * rows, pe.shape[1]), dtype=pe.dtype, device=pe.device)
    keys_all = torch.empty((world * rows,), dtype=torch.int64, device=pk.device)
    dist.all_gather_into_tensor(emb_all, pe, group=group)
    dist.all_gather_into_tensor(keys_all, pk, group=group)
    return emb_all, keys_all


def drop_pad(keys: torch.Tensor, score: torch.Tensor, count: torch.Tensor):
    keep = keys != PAD_KEY
    return keys[keep], score[keep], count[keep]


def allgather_disagreement(emb: torch.Tensor, keys: torch.Tensor, n_total: int, mode: int = 0, group=None,
                           reduce_fn=None):
    """Baseline exchange: NCCL all-gather, then the replicated reduction (`reduce_fn` defaults to the CUDA
    kernel; the gloo CPU tests inject the oracle to exercise the host logic)."""
    world = dist.get_world_size(group)
    emb_all, keys_all = gather_padded(emb, keys, rows_per_rank(n_total, world), group)
    if reduce_fn is None:
        from . import ops
        reduce_fn = ops.instance_disagreement
    k, s, c = reduce_fn(emb_all, keys_all, mode)
    return drop_pad(k, s, c)


class PeerExchange:
    """Symmetric-memory buffer holding this rank's embedding shard, mapped into every peer."""

    def __init__(self, rows: int, dim: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rows, self.dim = rows, dim
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((rows, dim), dtype=torch.float32, device=self.device)
        self.handle = symm.rendezvous(self.buf, self.group.group_name)
        self.world = self.handle.world_size
        self.rank = self.handle.rank
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def barrier(self):
        self.handle.barrier(channel=0)


def peer_disagreement(xchg: PeerExchange, emb: torch.Tensor, keys: torch.Tensor, n_total: int, mode: int = 0):
    """Fused gather+reduce over NVLink peer memory (see module docstring)."""
    from . import _native as nat
    world, rank, rows = xchg.world, xchg.rank, xchg.rows
    assert rows == rows_per_rank(n_total, world) and emb.shape[1] == xchg.dim
    pe, pk = pad_local(emb, keys, rows)
    xchg.buf.copy_(pe)
    keys_all = torch.empty((world * rows,), dtype=torch.int64, device=xchg.device)
    dist.all_gather_into_tensor(keys_all, pk, group=xchg.group)
    xchg.barrier()                                   # every shard is in place before any peer reads it
    n = world * rows
    dev = xchg.device
    ok = torch.empty((n,), device=dev, dtype=torch.int64)
    sc = torch.zeros((n,), device=dev, dtype=torch.float32)
    ct = torch.empty((n,), device=dev, dtype=torch.int32)
    m = torch.zeros((1,), device=dev, dtype=torch.int32)
    lib = nat.lib()
    ws = torch.empty((max(lib.ec_disagreement_workspace_bytes(n), 16),), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        nat.check(lib.ec_instance_disagreement_peers(
            xchg.peer_table.data_ptr(), world, rows, xchg.dim, keys_all.data_ptr(), n, xchg.dim, mode, rank, world,
            ok.data_ptr(), sc.data_ptr(), ct.data_ptr(), m.data_ptr(), ws.data_ptr(), ws.numel(),
            torch.cuda.current_stream(dev).cuda_stream), "ec_instance_disagreement_peers")
    M = int(m.item())
    sc = sc[:M]
    # non-owned instances hold 0 (NaN stays NaN only on its owner): sum = gather of the owned scores
    owned_nan = torch.isnan(sc)
    sc = torch.where(owned_nan, torch.zeros_like(sc), sc)
    nanflag = owned_nan.to(torch.float32)
    dist.all_reduce(sc, group=xchg.group)
    dist.all_reduce(nanflag, group=xchg.group)
    sc = torch.where(nanflag > 0, torch.full_like(sc, float("nan")), sc)
    xchg.barrier()                                   # peers finished reading before the buffer is reused
    return drop_pad(ok[:M], sc, ct[:M])
