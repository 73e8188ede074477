#!/usr/bin/env python
"""Multi-GPU exchange-step check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multigpu_check.py
Each rank holds a contiguous shard of the view embeddings; the per-instance disagreement is computed by
(a) NCCL all-gather + replicated reduction and (b) the peer-memory fused gather+reduce, and both are compared
bit-for-bit with the single-GPU kernel on the full set and (1e-3) with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embodied_captioning_b200 import distributed as D  # noqa: E402
from embodied_captioning_b200 import ops  # noqa: E402
from oracle import disagreement_oracle as do  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for n, d, m in ((1000, 768, 37), (100_003, 768, 2000), (257, 512, 300)):
        rng = np.random.default_rng(7)
        inst = rng.integers(0, m, size=n).astype(np.int64)
        emb = (rng.normal(size=(m, d))[inst] * 2 + rng.normal(size=(n, d))).astype(np.float32)
        s, e = D.shard_range(n, rank, world)
        emb_l, key_l = torch.from_numpy(emb[s:e]).to(dev), torch.from_numpy(inst[s:e]).to(dev)
        # shards must be the padded contiguous layout: every rank holds rows_per_rank rows except the tail
        rows = D.rows_per_rank(n, world)
        s2, e2 = min(rank * rows, n), min((rank + 1) * rows, n)
        emb_l, key_l = torch.from_numpy(emb[s2:e2]).to(dev), torch.from_numpy(inst[s2:e2]).to(dev)
        full_k, full_s, full_c = ops.instance_disagreement(torch.from_numpy(emb).to(dev), torch.from_numpy(inst).to(dev), 0)
        for mode in (0, 1):
            fk, fs, fc = ops.instance_disagreement(torch.from_numpy(emb).to(dev), torch.from_numpy(inst).to(dev), mode)
            ak, a_s, ac = D.allgather_disagreement(emb_l, key_l, n, mode)
            xchg = D.PeerExchange(rows, d, device=dev)
            pk, ps, pc = D.peer_disagreement(xchg, emb_l, key_l, n, mode)
            same_a = torch.equal(ak, fk) and torch.equal(ac, fc) and torch.equal(a_s.nan_to_num(-7), fs.nan_to_num(-7))
            same_p = torch.equal(pk, fk) and torch.equal(pc, fc) and torch.equal(ps.nan_to_num(-7), fs.nan_to_num(-7))
            rk, rs, rc = do.instance_disagreement(emb, inst, mode="full" if mode == 0 else "pairs") if n <= 2000 else (None, None, None)
            orc = True
            if rk is not None:
                got = fs.cpu().numpy().astype(np.float64)
                orc = (fk.cpu().tolist() == rk.tolist() and np.array_equal(np.isnan(got), np.isnan(rs))
                       and np.nanmax(np.abs(got - rs), initial=0.0) < 1e-3)
            if rank == 0:
                print(f"n={n} d={d} mode={mode}: allgather bit-identical={same_a} peer bit-identical={same_p} oracle={orc}")
            ok = ok and same_a and same_p and orc
            del xchg
    # timing of the two exchange forms on the K3 shape (100k crops, D=768)
    n, d, m = 100_000, 768, 2000
    rows = D.rows_per_rank(n, world)
    emb_l = torch.randn(rows, d, device=dev)
    key_l = torch.randint(0, m, (rows,), device=dev)
    xchg = D.PeerExchange(rows, d, device=dev)
    for name, fn in (("allgather", lambda: D.allgather_disagreement(emb_l, key_l, rows * world, 0)),
                     ("peer", lambda: D.peer_disagreement(xchg, emb_l, key_l, rows * world, 0))):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"K3 exchange+reduce ({world} GPUs, {rows * world} x {d}): {name} {t.item():.3f} ms")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
