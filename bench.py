#!/usr/bin/env python
"""Headline benchmark: object-crop captions/sec (CoCa ViT-L/14, greedy) + % of roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input (BASELINE.json configs[1]):
1024 uint8 224x224 crops per GPU -> preprocess gather kernel -> ViT-L/14 image tower (tcgen05 GEMMs) ->
29 KV-cached greedy decode steps (max 30 tokens) -> per-instance disagreement over the pooled latents
(all-gathered over NCCL when N > 1).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "object-crop captions/sec (CoCa ViT-L/14 greedy)"
UNIT = "crops/s"
MODEL = os.environ.get("EC_BENCH_MODEL", "coca_ViT-L-14")
BATCH = int(os.environ.get("EC_BENCH_BATCH", "1024"))
SEQ_LEN = 30
VIEWS = 8                        # views per object instance in the synthetic batch
ENC_GFLOP = {"coca_ViT-L-14": 163.64, "coca_ViT-B-32": 9.19}     # SURVEY.md section 8(a) A3, per crop
_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.1)
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def decoder_algorithmic_bytes(cfg, batch, steps):
    """weights once per step per batch + cross-K/V per active crop + self-K/V reads/writes (bf16)."""
    T, L = cfg.text_width, cfg.mm_layers
    block = 12 * T * T                                   # qkv + out + fc1 + fc2 weights of one block
    w_params = cfg.text_layers * block + L * block + L * (block - 2 * T * T) + T * cfg.vocab_size
    cross = L * cfg.n_image_embs * 2 * T * 2             # bytes per crop per step
    self_tok = (cfg.text_layers + L) * 2 * T * 2          # bytes per crop per cached token
    total = 0
    for t in range(steps):
        total += 2 * w_params + batch * (cross + self_tok * (t + 1) + self_tok)
    return total


def gemm_algorithmic_bytes(cfg, batch):
    """Algorithmic HBM bytes of the tcgen05 GEMM launches of one step that bench.py times (image tower + cross-K/V
    projection): A (bf16) + W (bf16) + output (+ fp32 residual read) per launch, DESIGN.md section 3."""
    G2 = (cfg.image_size // cfg.patch_size) ** 2
    Mv, Mp = batch * (G2 + 1), batch * G2
    W, E, T, Q = cfg.vision_width, cfg.embed_dim, cfg.text_width, cfg.n_queries
    pk = -(-3 * cfg.patch_size ** 2 // 8) * 8

    def g(M, N, K, out_bytes, resid=False):
        return 2 * M * K + 2 * N * K + out_bytes * M * N + (4 * M * N if resid else 0)
    per_layer = g(Mv, 3 * W, W, 2) + g(Mv, W, W, 4, True) + g(Mv, 4 * W, W, 2) + g(Mv, W, 4 * W, 4, True)
    total = g(Mp, W, pk, 4) + cfg.vision_layers * per_layer + g(Mv, 2 * E, W, 2) + g(batch * Q, E, E, 4) + g(batch, E, E, 4)
    total += cfg.mm_layers * g(batch * (Q - 1), 2 * T, T, 2)
    launches = 1 + 4 * cfg.vision_layers + 3 + cfg.mm_layers
    return total, launches


def measured_gemm_traffic():
    """Per-launch DRAM bytes of the same launches from the committed ncu launch list (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.isfile(p) and MODEL == "coca_ViT-L-14" and BATCH == 1024:
        d = json.load(open(p))
        return d["dram_bytes_per_launch"], d["source"]
    return None, None


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU algorithm for this path (oracle port: Pillow-exact preprocess,
    fp32 CoCa, generate() with batch 1 and no KV cache, python-loop disagreement) on all host threads."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from embodied_captioning_b200 import synth
    from embodied_captioning_b200.config import get_config
    from oracle import disagreement_oracle as do
    from oracle import preprocess_oracle as po
    from oracle.coca_oracle import OracleCoCa
    cfg = get_config(MODEL)
    sd = synth.make_state_dict(cfg, seed=1234)
    orc = OracleCoCa(cfg, sd)
    crops = synth.make_crops(16, seed=5678)
    keys = np.arange(16) // 2

    def one_pass(n):
        lats = []
        for i in range(n):                               # batch 1, like coca.py:28-29
            x = torch.from_numpy(po.transform(crops[i % 16]))[None]
            lat, embs = orc.encode_image_tokens(x)
            orc.generate(image_embs=embs, seq_len=SEQ_LEN, mode="reforward", keep_logits=False)
            lats.append(lat[0].numpy())
        do.instance_disagreement(np.stack(lats), keys[:n], mode="full")

    t0 = time.time(); one_pass(1); t1 = time.time() - t0
    per_step = max(1, min(16, int(8.0 / max(t1, 1e-3))))  # ~8 s of CPU work per step
    for _ in range(max(args.warmup - 1, 0)):
        one_pass(per_step)
    t0 = time.time()
    for _ in range(args.steps):
        one_pass(per_step)
    dt = time.time() - t0
    v = per_step * args.steps / dt
    sample = f"{per_step} of {BATCH} crops per step, batch-1 re-forward generate (no KV cache), fp32"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} greedy captioning, max {SEQ_LEN} tokens + per-instance disagreement "
                               f"(CPU reference path, bounded sample)", "sample_crops_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner there) are sent to
    # stderr for the whole run and the line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from embodied_captioning_b200 import _native as nat
    from embodied_captioning_b200 import distributed as D
    from embodied_captioning_b200 import ops, synth
    from embodied_captioning_b200.config import get_config
    from embodied_captioning_b200.engine import CaptionEngine

    if not os.path.isfile(nat.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = nat.lib()
    peaks = load_peaks()
    cfg = get_config(MODEL)
    sd = synth.make_state_dict(cfg, seed=1234)
    eng = CaptionEngine(MODEL, None, device=dev, decode_graph=True, state_dict=sd)

    crops_host = torch.from_numpy(synth.make_crops(BATCH, seed=5678 + rank)).pin_memory()
    n_inst = max(BATCH * world // VIEWS, 1)
    keys_host = torch.from_numpy(np.random.default_rng(99 + rank).integers(0, n_inst, size=BATCH)).pin_memory()
    crops = crops_host.to(dev)
    keys = keys_host.to(dev)

    def step():
        x, _, _ = ops.preprocess_crops(crops, eng_rects, eng_fi, None, size=cfg.image_size)
        out = eng.caption_preprocessed(x, None if world > 1 else keys, SEQ_LEN)
        if world > 1:
            out["inst"] = D.allgather_disagreement(out["latent"], keys, BATCH * world, 0)
        return out

    eng_rects = torch.tensor([0, 0, cfg.image_size, cfg.image_size], dtype=torch.int32, device=dev).repeat(BATCH, 1)
    eng_fi = torch.arange(BATCH, dtype=torch.int32, device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    ops.check_watchdog(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nat.check(lib.ec_profile_gemm(1))
    launches0 = lib.ec_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    launches = lib.ec_launch_count() - launches0
    import ctypes as C
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_int64()
    nat.check(lib.ec_profile_gemm_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n)))
    nat.check(lib.ec_profile_gemm(0))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    value = BATCH * world / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D + D2H inside the timed region)
    def step_host():
        if world == 1:
            return eng.caption_crops_host(crops_host, keys_host, SEQ_LEN)
        # N > 1: H2D of the local shard, local captioning, NCCL exchange + reduction, D2H of tokens and scores
        c = crops_host.to(dev, non_blocking=True)
        k = keys_host.to(dev, non_blocking=True)
        x, _, _ = ops.preprocess_crops(c, eng_rects, eng_fi, None, size=cfg.image_size)
        o = eng.caption_preprocessed(x, None, SEQ_LEN)
        ik, isc, _ = D.allgather_disagreement(o["latent"], k, BATCH * world, 0)
        tok = o["tokens"].cpu()
        sc = isc.cpu()
        ikc = ik.cpu()
        return {"h2d_bytes": crops_host.numel() + keys_host.numel() * 8,
                "d2h_bytes": tok.numel() * 8 + sc.numel() * 4 + ikc.numel() * 8}
    step_host()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_host()
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = BATCH * world / (float(ems.item()) * 1e-3)

    # ---- per-stage breakdown (one extra untimed-for-value step) for the decoder HBM line
    def timed(fn):
        # second of two calls: the first one may (re)capture the decoder's step graph for this argument set
        fn(); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(dev)
        return r, a.elapsed_time(b)
    (x, _, _), t_pre = timed(lambda: ops.preprocess_crops(crops, eng_rects, eng_fi, None, size=cfg.image_size))
    (lat, embs), t_enc = timed(lambda: eng.model._encode_image_dev(x))
    _, t_dec = timed(lambda: eng.model.generate(None, image_embs=embs, generation_type="top_k", seq_len=SEQ_LEN,
                                                fixed_output_length=True, return_logits=False))
    _, t_dis = timed(lambda: ops.instance_disagreement(lat, keys, 0))
    dec_bytes = decoder_algorithmic_bytes(cfg, BATCH, SEQ_LEN - 1)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    gemm_tflops = (g_fl.value / (g_ms.value * 1e-3) / 1e12) if g_ms.value > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    alg_bytes, alg_launches = gemm_algorithmic_bytes(cfg, BATCH)
    traffic, traffic_src = measured_gemm_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{MODEL} greedy captioning of {BATCH} synthetic 224x224 crops per GPU, max {SEQ_LEN} "
                               f"tokens, + per-instance disagreement ({VIEWS} views/instance)",
                   "crops_per_gpu": BATCH, "seq_len": SEQ_LEN, "weights": "random-init, bf16-representable, seed 1234",
                   "l2_hygiene": "per-step working set (>= 9 GB activations + K/V) is far larger than the 126 MB L2",
                   "parallelism": f"batch-sharded x{world}" + (", NCCL all-gather of latents+keys" if world > 1 else "")},
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": gemm_tflops / peak_tf if peak_tf else None, "traffic": traffic,
                     "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg_bytes / alg_launches,
                     "algorithmic_flops_per_launch": (g_fl.value / g_n.value) if g_n.value else None,
                     "avg_launch_ms": (g_ms.value / g_n.value) if g_n.value else None,
                     "kernel": "gemm_tcgen05_pair_kernel<256,*> (tcgen05 cta_group::2; image tower + cross-K/V "
                               "projection GEMM launches of the step)",
                     "launches_timed": int(g_n.value), "peak_source": peaks["source"] + " bf16_tflops_sustained",
                     "encoder_only": {"ms": t_enc, "tflops": BATCH * ENC_GFLOP.get(MODEL, 0) / t_enc if t_enc else None,
                                      "frac_of_burst_peak": (BATCH * ENC_GFLOP.get(MODEL, 0) / t_enc / peaks["bf16_tflops"])
                                      if t_enc else None}},
        "roofline_decoder": {"bound": "hbm", "achieved": dec_bytes / (t_dec * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": dec_bytes / (t_dec * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "ms": t_dec, "algorithmic_bytes": dec_bytes},
        "stages_ms": {"preprocess": t_pre, "encoder": t_enc, "decoder": t_dec, "disagreement": t_dis},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(res["h2d_bytes"]),
                "d2h_bytes_per_step": int(res["d2h_bytes"])},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, sd)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(cfg, sd):
    """Oracle port timed on the host cores on a bounded sample of the same workload (reference-style loop:
    batch 1, fp32, full re-forward per token -- coca_model.py:294-303)."""
    import numpy as np
    import torch
    from embodied_captioning_b200 import synth
    from oracle import disagreement_oracle as do
    from oracle import preprocess_oracle as po
    from oracle.coca_oracle import OracleCoCa
    orc = OracleCoCa(cfg, sd)
    crops = synth.make_crops(8, seed=5678)
    t0 = time.time()
    lats, n = [], 0
    while n < 8 and (time.time() - t0 < 15.0 or n == 0):
        x = torch.from_numpy(po.transform(crops[n]))[None]
        lat, embs = orc.encode_image_tokens(x)
        orc.generate(image_embs=embs, seq_len=SEQ_LEN, mode="reforward", keep_logits=False)
        lats.append(lat[0].numpy()); n += 1
    do.instance_disagreement(np.stack(lats), np.arange(n) // 2, mode="full")
    dt = time.time() - t0
    t1 = time.time()
    x = torch.from_numpy(np.stack([po.transform(c) for c in crops[:n]]))
    _, embs = orc.encode_image_tokens(x)
    orc.generate(image_embs=embs, seq_len=SEQ_LEN, mode="kvcache", keep_logits=False)
    dt_fair = time.time() - t1
    return {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} of {BATCH} crops, batch-1 re-forward generate (no KV cache), fp32, Pillow-exact preprocess",
            "fair_batched_kvcache_value": n / dt_fair}


if __name__ == "__main__":
    main()
