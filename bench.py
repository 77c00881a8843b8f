#!/usr/bin/env python
"""bench.py -- pairs/sec of the contrastive GradCache training step (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one full training step of nomic-bert-base text-text InfoNCE, bf16, seq 512, GLOBAL batch 16384 pairs,
GradCache chunk 64 (reference configs/train/contrastive_pretrain.yaml): chunked no-grad forwards of queries and
documents, the fused InfoNCE loss + its embedding gradients (bf16 all-gather / reduce-scatter across ranks), chunked
re-forward + backward of both towers (shared weights), gradient all-reduce, global-norm clip + AdamW.  The global batch
is fixed, so N GPUs each take 16384/N pairs ("scaling": "strong").  Synthetic token ids, random-init weights.

Printed JSON (rank 0): value = whole-job pairs/s with inputs resident in HBM; e2e = the same K steps through the public
API from pinned HOST buffers (H2D copies + a D2H read of the loss inside the timed region); roofline = the dominant
kernel (tcgen05 GEMM) timed live with CUDA events on the launching stream (1 launch in 17: coprime with the 16 GEMMs of
a layer), against MEASURED_PEAKS.json; comm_ms = device time of every collective of a step; selfcheck = the loss of the
first 64 pairs at step 0 against the CPU oracle; gpu_baseline = the UNMODIFIED reference (its pure-PyTorch NomicBertModel
+ its grad_cache_loss, from baseline/_ref) on the same B200 under bf16 autocast over a bounded sample; cpu_baseline = the
same reference code on this box's host cores (median of 3 after a warm-up).  The three checker legs (selfcheck, gpu_baseline,
cpu_baseline) run on rank 0 at N = 1 only; at N > 1 their keys are null.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore", message=".*Disabling autocast.*")  # the reference's hard-coded cuda autocast on the CPU arm
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 16384
SEQ_LEN = 512
CHUNK = 64
LOGIT_SCALE = 50.0
SELFCHECK_PAIRS = 64
LR, WD, CLIP = 2.0e-4, 0.01, 1.0
METRIC = "pairs/sec at global_bs=16384 (nomic-bert-base text-text InfoNCE, bf16, seq=512, GradCache)"
WORKLOAD = "configs[1]: nomic-bert-base text-text InfoNCE bf16 seq=512 global_bs=16384 GradCache chunk=64"


def ncu_gemm_traffic():
    """Mean DRAM bytes (read + write) per launch of the GEMM kernel, from the committed `ncu --set full` capture of the
    four forward GEMMs of one layer (profiles/r02g_gemm_ncu_full_raw.csv); None if the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", "r02g_gemm_ncu_full_raw.csv")
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


# ------------------------------------------------------------------------------------------------- reference arms
CPU_SAMPLE_PAIRS = 4   # per CPU step: 4 pairs x 512 tokens through the full GradCache step (chunk 2)
CPU_CHUNK = 2
GPU_BASE_CHUNKS = 8    # gpu_baseline: 8 chunks x 64 pairs x 512 tokens per tower per step


def cpu_threads():
    """Threads for the CPU arm: torch's intra-op pool scales poorly past ~32 threads on these few-thousand-token
    samples (128 threads measured slower than 32), so the arm uses min(cores, 32) and reports that number."""
    return min(os.cpu_count() or 1, 32)


def _ensure_group(backend):
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    import tempfile
    # a private 1-rank group through a FileStore: under torchrun a tcp:// rendezvous is redirected to the elastic agent's store
    # (TORCHELASTIC_USE_AGENT_STORE) and a worker that asks for its own port waits for a server that never comes
    store = dist.FileStore(os.path.join(tempfile.mkdtemp(prefix="cxbench_"), "store"), 1)
    dist.init_process_group(backend, store=store, rank=0, world_size=1)
    return True


class ReferenceStep:
    """One training step of the reference's own code on ``device``: its pure-PyTorch NomicBertModel (modeling_hf_nomic_bert.py:
    1650) in a mean-pool + normalize tower, its ``grad_cache_loss`` (loss.py:187-213; bf16 autocast on CUDA, fp32 on the
    host), ``clip_grad_norm_`` + ``torch.optim.AdamW`` + ``zero_grad(set_to_none=True)`` as trainers/base.py:366-393.  Falls back
    to the oracle port of the same step when neither /root/reference nor baseline/_ref is present (``kind`` says which)."""

    def __init__(self, device, n_pairs, chunk):
        import torch
        from oracle import ref_loader
        self.torch, self.device, self.n_pairs, self.chunk = torch, device, n_pairs, chunk
        g = torch.Generator().manual_seed(42)
        self.q = torch.randint(0, 30000, (n_pairs, SEQ_LEN), generator=g).to(device)
        self.d = torch.randint(0, 30000, (n_pairs, SEQ_LEN), generator=g).to(device)
        self.mask = torch.ones(n_pairs, SEQ_LEN, dtype=torch.long, device=device)
        if ref_loader.available():
            from oracle import ref_tower
            self.kind = "reference"
            torch.manual_seed(0)
            self.ref, self.model = ref_tower.build(device)
            self.model.train()
            self.scale = ref_tower.LogitScale(LOGIT_SCALE).to(device)
            self.opt = torch.optim.AdamW(self.model.parameters(), lr=LR, weight_decay=WD)
            self.source = ref_loader.source()
        else:
            from oracle.encoder import EncoderConfig, random_state_dict
            self.kind = "port"
            self.cfg = EncoderConfig()
            self.sd = {k: v.to(device).requires_grad_() for k, v in random_state_dict(self.cfg, seed=0, ln_jitter=0.0).items()}
            self.opt = torch.optim.AdamW(list(self.sd.values()), lr=LR, weight_decay=WD)
            self.source = "oracle port"

    def step(self):
        torch = self.torch
        if self.kind == "reference":
            loss = self.ref.loss.grad_cache_loss(self.model, {"input_ids": self.q, "attention_mask": self.mask}, self.model,
                                                 {"input_ids": self.d, "attention_mask": self.mask}, self.chunk, self.scale)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), CLIP)
        else:
            from oracle import infonce as O
            from oracle.encoder import biencoder_forward
            q = biencoder_forward(self.sd, self.cfg, self.q, self.mask)
            d = biencoder_forward(self.sd, self.cfg, self.d, self.mask)
            o = O.clip_loss_fwd_bwd(q.detach().cpu().numpy(), d.detach().cpu().numpy(), LOGIT_SCALE)
            torch.autograd.backward([q, d], [torch.from_numpy(o["dq"]).float().to(q.device), torch.from_numpy(o["dd"]).float().to(d.device)])
            torch.nn.utils.clip_grad_norm_(list(self.sd.values()), CLIP)
            loss = torch.tensor(o["loss"])
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return float(loss)

    def describe(self, where):
        what = ("the reference's HF NomicBertModel + grad_cache_loss + AdamW (" + self.source + ")") if self.kind == "reference" \
            else "oracle port of the step (reference package not present)"
        return f"{self.n_pairs} pairs x seq {SEQ_LEN} per step, GradCache chunk {self.chunk}, {where}: {what}"


def cpu_baseline(threads, steps=3, warmup=1):
    """(pairs/s median over ``steps`` after ``warmup``, per-step seconds, kind, sample description, last loss)."""
    import torch
    torch.set_num_threads(threads)
    created = _ensure_group("gloo")
    try:
        ref = ReferenceStep(torch.device("cpu"), CPU_SAMPLE_PAIRS, CPU_CHUNK)
        for _ in range(warmup):
            ref.step()
        times, loss = [], None
        for _ in range(steps):
            t0 = time.perf_counter()
            loss = ref.step()
            times.append(time.perf_counter() - t0)
    finally:
        if created:
            import torch.distributed as dist
            dist.destroy_process_group()
    med = sorted(times)[len(times) // 2]
    return CPU_SAMPLE_PAIRS / med, times, ref.kind, ref.describe(f"fp32 on {threads} host threads"), loss


def gpu_baseline(dev, steps=3, warmup=1):
    """The reference code on the same B200: bf16 autocast (loss.py:139,156,175), its own kernels (torch SDPA / cuBLAS / ATen).
    Needs an initialised process group (clip_loss calls dist.get_world_size()); the caller's NCCL group is used."""
    import torch
    n_pairs = GPU_BASE_CHUNKS * CHUNK
    ref = ReferenceStep(dev, n_pairs, CHUNK)
    for _ in range(warmup):
        ref.step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    loss = None
    for _ in range(steps):
        loss = ref.step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    out = {"value": n_pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "kind": ref.kind,
           "dtype": "bf16 autocast, fp32 master weights", "loss": loss,
           "sample": ref.describe("on cuda:%d (1 GPU; data-parallel replicas scale it by N)" % dev.index)}
    del ref
    torch.cuda.empty_cache()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    import torch
    torch.set_num_threads(threads)
    created = _ensure_group("gloo")
    ref = ReferenceStep(torch.device("cpu"), CPU_SAMPLE_PAIRS, CPU_CHUNK)
    for _ in range(max(1, min(args.warmup, 2))):  # CPU warm-up: thread pool + allocator; two bounded samples are plenty
        ref.step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ref.step()
        times.append(time.perf_counter() - t0)
    if created:
        import torch.distributed as dist
        dist.destroy_process_group()
    total = sum(times)
    value = CPU_SAMPLE_PAIRS * len(times) / total
    sample = ref.describe(f"fp32 on {threads} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "seq_len": SEQ_LEN, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": ref.kind, "sample": sample,
                         "median_pairs_per_s": CPU_SAMPLE_PAIRS / sorted(times)[len(times) // 2]},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def oracle_selfcheck(model, logit_scale, q_ids, d_ids):
    """Loss of the first ``n`` pairs through the product path (GradCache forward only: no-grad chunked embeddings + fused
    InfoNCE) against the CPU oracle on the same weights: the bench checks its own arithmetic before it times anything."""
    import torch
    import contrastors_b200 as cb
    from oracle import infonce as O
    from oracle.encoder import EncoderConfig, biencoder_forward
    n = q_ids.shape[0]
    with torch.no_grad():
        eq = model(q_ids, seq_lens=torch.full((n,), SEQ_LEN))["embedding"]
        ed = model(d_ids, seq_lens=torch.full((n,), SEQ_LEN))["embedding"]
        got = float(cb.clip_loss(eq, ed, logit_scale).item())
        sd = {k[len("trunk."):]: v.detach().float().cpu() for k, v in model.state_dict().items()}
        cfg = EncoderConfig()
        t0 = time.perf_counter()
        oq = biencoder_forward(sd, cfg, q_ids.cpu(), None).numpy()
        od = biencoder_forward(sd, cfg, d_ids.cpu(), None).numpy()
        want = float(O.clip_loss_fwd_bwd(oq, od, LOGIT_SCALE)["loss"])
        emb_err = float(max(abs(eq.cpu().numpy() - oq).max(), abs(ed.cpu().numpy() - od).max()))
    rel = abs(got - want) / max(abs(want), 1e-6)
    out = {"pairs": n, "loss": got, "oracle_loss": want, "rel_err": rel, "max_abs_embedding_err": emb_err, "tolerance": 2e-2,
           "ok": bool(rel <= 2e-2), "oracle_seconds": time.perf_counter() - t0,
           "note": "fp32 CPU oracle vs bf16 tower at random init (logits ~ scale * cos ~ 50 * O(1)): bf16-level agreement"}
    if not out["ok"]:
        raise SystemExit(f"bench self-check FAILED: loss {got} vs oracle {want} on the first {n} pairs")
    return out


# ------------------------------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    import contrastors_b200 as cb
    from contrastors_b200 import _lib, ops
    from contrastors_b200.parallel import broadcast_parameters

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

    selfcheck = None
    if rank == 0 and world == 1 and not args.no_selfcheck:  # the checker legs (oracle self-check, both baselines) run at N = 1 only
        selfcheck = oracle_selfcheck(model, logit_scale, resident["query_input_ids"][:SELFCHECK_PAIRS],
                                     resident["document_input_ids"][:SELFCHECK_PAIRS])
    for _ in range(args.warmup):
        train_step(resident)
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM (per-step inputs of 2 x 16.8 MB / rank exceed nothing, but every step
    # streams ~10 GB of activations per chunk through HBM, far beyond the 126 MB L2: no extra L2 flush is needed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    from contrastors_b200 import distributed as cxd
    ops.TIMER = ops.KernelTimer(sample_every=17)  # coprime with the 16 GEMMs of a layer: every GEMM shape gets sampled
    cxd.COMM_EVENTS = []
    ms_dev = timed(lambda: train_step(resident), args.steps)
    timer, ops.TIMER = ops.TIMER, None
    comm_events, cxd.COMM_EVENTS = cxd.COMM_EVENTS, None
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    losses = []

    from contrastors_b200.trainer import BatchPrefetcher
    e2e_steps = args.steps
    state = {}

    def e2e_step():
        # the public input edge: pinned host batches through BatchPrefetcher (batch i+1 is copied on a side stream under step i;
        # the first batch's copy, like every other, happens inside the timed region), loss read back to the host every step
        if "it" not in state:
            state["it"] = BatchPrefetcher((host for _ in range(e2e_steps)), dev)
        losses.append(train_step(next(state["it"])).item())

    ms_e2e = timed(e2e_step, e2e_steps)
    comm = {}
    for kind, a, b in comm_events:
        c = comm.setdefault(kind, [0, 0.0])
        c[0] += 1
        c[1] += a.elapsed_time(b)

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
                "traffic_note": "mean DRAM read+write bytes per launch over the 4 forward GEMMs of a layer (ncu --set full, profiles/r02g_gemm_ncu_full_raw.csv); algorithmic operand+result bytes of the same 4 launches average 289 MB (the fc1 launch also stores the 403 MB of pre-activations in training form)",
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
    # the model, optimizer state and activations of our arm are no longer needed: free them before the baselines run
    del model
    torch.cuda.empty_cache()
    dist.destroy_process_group()
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        created = _ensure_group("gloo")  # the reference's clip_loss asks for a process group; 1 rank: its gather is the identity
        try:
            gpu_base = gpu_baseline(dev)
        except Exception as ex:  # the baseline must never take the bench line down with it
            gpu_base = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
        if created:
            dist.destroy_process_group()
    # CPU baseline: the reference's own step on the host cores, bounded sample, median of 3 after a warm-up
    threads = cpu_threads()
    cpu_base = None
    if world == 1:
        cpu_value, cpu_times, cpu_kind, cpu_sample, _ = cpu_baseline(threads)
        cpu_base = {"value": cpu_value, "unit": "pairs/s", "cores": threads, "kind": cpu_kind, "sample": cpu_sample,
                    "step_seconds": cpu_times, "stat": "median of 3 after 1 warm-up"}
    comm_ms = {k: {"calls_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in comm.items()}
    comm_total = sum(v["ms_per_step"] for v in comm_ms.values())
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
        "comm_ms": {"total_per_step": comm_total, "share_of_step": comm_total / (ms_dev / args.steps), "by_kind": comm_ms,
                    "note": "device time of rank 0's collectives (CUDA events on the stream each is launched on; the chunk gathers and "
                            "the gradient buckets run on side streams under compute, so this is occupancy, not exposed time)"},
        "selfcheck": selfcheck,
        "gpu_baseline": gpu_base,
        "cpu_baseline": cpu_base,
        "loss": losses[-1] if losses else None,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------- image-text arm (configs[2] / [4])
def run_image_text(args):
    """BASELINE configs[2]: ViT-B/16 + nomic-bert-base, 224^2 images, text seq 77, global batch 32768, GradCache chunk 64, both towers
    trainable, trainable logit scale 1 / 0.07; with --config lit it is configs[4]: a FROZEN ViT-L/14 + the trainable text tower at
    global batch 65536.  Informational line for profiles/ (the driver's metric is the text-text line)."""
    import torch
    import torch.distributed as dist
    import contrastors_b200 as cb
    from contrastors_b200 import _lib
    from contrastors_b200.parallel import broadcast_parameters
    from contrastors_b200.trainer import dual_training_step
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lit = args.config == "lit"
    global_batch = args.global_batch or (65536 if lit else 32768)
    n_local, seq = global_batch // world, 77
    torch.manual_seed(0)
    # LiT: the 1024-wide frozen ViT-L/14 meets the 768-wide text tower through a (trainable) projection, as BiEncoder.proj does
    vision = cb.VisionBiEncoder(cb.VisionBiEncoderConfig(encoder=cb.vit_l14() if lit else cb.vit_b16(), freeze=lit,
                                                         projection_dim=768 if lit else None)).to(dev)
    text = cb.BiEncoder(cb.BiEncoderConfig(encoder=cb.nomic_bert_base())).to(dev)
    vision.trunk.reset_parameters(seed=1)
    text.trunk.reset_parameters(seed=0)
    broadcast_parameters(vision, text)
    ls = cb.LogitScale(logit_scale=1.0 / 0.07, trainable_logit_scale=True).to(dev)
    g = torch.Generator().manual_seed(42 + rank)
    px = torch.randn(n_local, 3, 224, 224, generator=g, dtype=torch.bfloat16).to(dev)
    ids = torch.randint(0, 30000, (n_local, seq), generator=g).to(dev)
    ones = torch.ones(n_local, seq, dtype=torch.int64, device=dev)
    lens = torch.full((n_local,), seq, dtype=torch.int64)

    def step():
        return dual_training_step(vision, text, {"input_ids": px}, {"input_ids": ids, "attention_mask": ones, "seq_lens": lens}, ls,
                                  lr=LR, chunk_size=CHUNK, weight_decay=WD, max_grad_norm=CLIP)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    n0 = _lib.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        flops_pair = (4 if not lit else 1) * (161.7e9 if lit else 34.9e9) + 4 * 245.4e6 * seq  # GradCache: 4 forward-equivalents per trainable tower
        v = global_batch * args.steps / (ms.item() * 1e-3)
        print(json.dumps({"metric": f"pairs/sec at global_bs={global_batch} ({'frozen ViT-L/14 (LiT)' if lit else 'ViT-B/16'} + nomic-bert-base image-text, bf16, 224^2, text seq 77, GradCache)",
                          "value": v, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms.item() / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": f"configs[{4 if lit else 2}]", "global_batch": global_batch, "per_gpu_batch": n_local,
                                     "image": 224, "seq_len": seq, "chunk": CHUNK, "parallelism": f"dp{world}",
                                     "loss": "clip_loss bidirectional (symmetric) at world size 1, image->text otherwise"},
                          "gpu_launches": int(_lib.launch_count() - n0), "model_tflops_per_s": v * flops_pair / 1e12,
                          "loss": float(loss)}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="text", choices=["text", "image_text", "lit"],
                    help="text = BASELINE configs[1] (the driver's metric); image_text / lit = configs[2] / configs[4], informational")
    ap.add_argument("--global-batch", type=int, default=0, help="override the global batch of the image-text configs")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the step-0 loss check against the CPU oracle")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-on-the-same-GPU leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config != "text":
        run_image_text(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
