#!/usr/bin/env python
"""bench.py -- pairs/sec of the contrastive GradCache training step (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one full training step of nomic-bert-base text-text InfoNCE, bf16, seq 512, GLOBAL batch 16384 pairs,
GradCache chunk 64 (reference configs/train/contrastive_pretrain.yaml): chunked no-grad forwards of queries and
documents, the fused InfoNCE loss + its embedding gradients (bf16 all-gather / reduce-scatter across ranks), chunked
re-forward + backward of both towers (shared weights), gradient all-reduce, global-norm clip + AdamW.  The global batch
is fixed, so N GPUs each take 16384/N pairs ("scaling": "strong").  Synthetic token ids, random-init weights.

Printed JSON (rank 0): value = whole-job pairs/s with inputs resident in HBM; e2e = the same step through the public
API from pinned HOST buffers (H2D copies + a D2H read of the loss inside the timed region); roofline = the dominant
kernel (tcgen05 GEMM) timed live with CUDA events on the launching stream, against MEASURED_PEAKS.json;
cpu_baseline = the oracle port of the reference step on this box's host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 16384
SEQ_LEN = 512
CHUNK = 64
LOGIT_SCALE = 50.0
LR, WD, CLIP = 2.0e-4, 0.01, 1.0
METRIC = "pairs/sec at global_bs=16384 (nomic-bert-base text-text InfoNCE, bf16, seq=512, GradCache)"
WORKLOAD = "configs[1]: nomic-bert-base text-text InfoNCE bf16 seq=512 global_bs=16384 GradCache chunk=64"


def ncu_gemm_traffic():
    """Mean DRAM bytes (read + write) per launch of the GEMM kernel, from the committed `ncu --set full` capture of the
    four forward GEMMs of one layer (profiles/r01_gemm_final_ncu_full_raw.csv); None if the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01_gemm_final_ncu_full_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        units = rows[1]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = [float(r[ir]) * mult.get(units[ir], 1.0) + float(r[iw]) * mult.get(units[iw], 1.0) for r in rows[2:]]
        return sum(tot) / len(tot) if tot else None
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(d.get("bf16_tflops", 1590.0)), hbm_gbs=float(d.get("hbm_gbs", 6650.0)), source="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_step_sample(n_pairs, seq, threads):
    """The oracle port of one reference training step on a bounded sample (``n_pairs`` pairs x ``seq`` tokens): both tower
    passes of nomic-bert-base, mean pool, normalize, clip_loss forward + backward (fp32, torch CPU kernels)."""
    import numpy as np
    import torch
    from oracle import infonce as O
    from oracle.encoder import EncoderConfig, biencoder_forward, random_state_dict
    torch.set_num_threads(threads)
    cfg = EncoderConfig()
    sd = {k: v.requires_grad_() for k, v in random_state_dict(cfg, seed=0, ln_jitter=0.0).items()}
    g = torch.Generator().manual_seed(42)
    q_ids = torch.randint(0, 30000, (n_pairs, seq), generator=g)
    d_ids = torch.randint(0, 30000, (n_pairs, seq), generator=g)
    mask = torch.ones(n_pairs, seq, dtype=torch.long)
    t0 = time.perf_counter()
    q = biencoder_forward(sd, cfg, q_ids, mask)
    d = biencoder_forward(sd, cfg, d_ids, mask)
    o = O.clip_loss_fwd_bwd(q.detach().numpy(), d.detach().numpy(), LOGIT_SCALE)
    torch.autograd.backward([q, d], [torch.from_numpy(o["dq"]).float(), torch.from_numpy(o["dd"]).float()])
    dt = time.perf_counter() - t0
    return dt, float(o["loss"])


CPU_SAMPLE_PAIRS = 2


def cpu_threads():
    """Threads for the CPU arm: torch's intra-op pool scales poorly past ~32 threads on these few-thousand-token
    samples (128 threads measured slower than 32), so the arm uses min(cores, 32) and reports that number."""
    return min(os.cpu_count() or 1, 32)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    n_pairs = CPU_SAMPLE_PAIRS
    for _ in range(min(args.warmup, 1)):  # CPU warm-up: one bounded sample is plenty
        cpu_step_sample(1, SEQ_LEN, threads)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_step_sample(n_pairs, SEQ_LEN, threads)
        times.append(dt)
    total = sum(times)
    value = n_pairs * len(times) / total
    sample = f"{n_pairs} pairs x seq {SEQ_LEN} per step (fwd+bwd of both towers + InfoNCE), oracle port, fp32 torch CPU kernels"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "seq_len": SEQ_LEN, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    import contrastors_b200 as cb
    from contrastors_b200 import _lib, ops
    from contrastors_b200.parallel import allreduce_gradients, broadcast_parameters

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n_local = GLOBAL_BATCH // world

    torch.manual_seed(0)
    model = cb.BiEncoder(cb.BiEncoderConfig(encoder=cb.nomic_bert_base())).to(dev)
    model.trunk.reset_parameters(seed=0)
    model.train()
    broadcast_parameters(model)
    logit_scale = cb.LogitScale(logit_scale=LOGIT_SCALE, trainable_logit_scale=False).to(dev)

    g = torch.Generator().manual_seed(42 + rank)
    host = {k: torch.randint(0, 30000, (n_local, SEQ_LEN), generator=g).pin_memory() for k in ("query_input_ids", "document_input_ids")}
    seq_lens = torch.full((n_local,), SEQ_LEN, dtype=torch.int64)  # CPU-side lengths (a loader has them for free)
    ones = torch.ones(n_local, SEQ_LEN, dtype=torch.int64, device=dev)
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    from contrastors_b200.trainer import training_step

    def train_step(batch):
        # the reference's collate schema (text_text_loader.py tokenize_pairs) + CPU-side lengths a loader has for free
        full = {"query_input_ids": batch["query_input_ids"], "query_attention_mask": ones, "query_seq_lens": seq_lens,
                "document_input_ids": batch["document_input_ids"], "document_attention_mask": ones,
                "document_seq_lens": seq_lens, "dataset_name": "synthetic"}
        return training_step(model, full, logit_scale, lr=LR, chunk_size=CHUNK, weight_decay=WD, max_grad_norm=CLIP)

    def timed(fn, steps):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        train_step(resident)
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM (per-step inputs of 2 x 16.8 MB / rank exceed nothing, but every step
    # streams ~10 GB of activations per chunk through HBM, far beyond the 126 MB L2: no extra L2 flush is needed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ops.TIMER = ops.KernelTimer(sample_every=16)
    ms_dev = timed(lambda: train_step(resident), args.steps)
    timer, ops.TIMER = ops.TIMER, None
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    losses = []

    def e2e_step():
        batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        losses.append(train_step(batch).item())

    e2e_steps = max(1, min(args.steps, 2))
    ms_e2e = timed(e2e_step, e2e_steps)

    if rank != 0:
        dist.destroy_process_group()
        return
    pk = peaks()
    summ = timer.summary()
    gm = summ.get("gemm")
    roof = None
    if gm:
        achieved = gm["work_avg"] / (gm["ms_avg"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "cx::gemm_kernel (tcgen05, all encoder linears fwd/dgrad/wgrad)", "achieved": achieved,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": ncu_gemm_traffic(),
                "traffic_note": "mean DRAM read+write bytes per launch over the 4 forward GEMMs of a layer (ncu --set full, profiles/r01_gemm_final_ncu_full_raw.csv); algorithmic operand+result bytes of the same 4 launches average 289 MB",
                "peak_source": pk["source"] + " bf16_tflops_sustained", "launches_per_step": gm["launches"] / args.steps,
                "avg_launch_ms": gm["ms_avg"], "sampled_launches": gm["sampled"],
                "share_of_step": gm["ms_avg"] * gm["launches"] / ms_dev}
        extra = {}
        for kind in ("attn_fwd", "attn_bwd", "infonce_fwd", "infonce_bwd"):
            if kind in summ:
                k = summ[kind]
                tf = k["work_avg"] / (k["ms_avg"] * 1e-3) / 1e12
                extra[kind] = {"achieved_tflops": tf, "frac": tf / pk["tflops"], "avg_launch_ms": k["ms_avg"],
                               "share_of_step": k["ms_avg"] * k["launches"] / ms_dev}
        if "infonce_fwd" in summ and "infonce_bwd" in summ:
            f, b = summ["infonce_fwd"], summ["infonce_bwd"]
            tf = (f["work_avg"] + b["work_avg"]) / ((f["ms_avg"] + b["ms_avg"]) * 1e-3) / 1e12
            extra["infonce"] = {"achieved_tflops": tf, "frac": tf / pk["tflops_burst"], "peak": pk["tflops_burst"],
                                "note": "6*N*M*D algorithmic FLOPs / (fwd+bwd time); burst peak (kernel timed alone)"}
        roof["other_kernels"] = extra
    # CPU baseline: bounded sample of the same step on the host cores (oracle port)
    threads = cpu_threads()
    cpu_pairs = CPU_SAMPLE_PAIRS
    cpu_dt, _ = cpu_step_sample(cpu_pairs, SEQ_LEN, threads)
    out = {
        "metric": METRIC, "value": GLOBAL_BATCH * args.steps / (ms_dev * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "per_gpu_batch": n_local, "seq_len": SEQ_LEN,
                   "parallelism": f"dp{world}", "l2": "working set per step (>10 GB of activations per chunk) >> 126 MB L2"},
        "e2e": {"value": GLOBAL_BATCH * e2e_steps / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes * world,
                "d2h_bytes_per_step": 4 * world, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": {"value": cpu_pairs / cpu_dt, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": f"{cpu_pairs} pairs x seq {SEQ_LEN}: fwd+bwd of both towers + InfoNCE, oracle port (fp32 torch CPU kernels)"},
        "loss": losses[-1] if losses else None,
    }
    print(json.dumps(out))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
