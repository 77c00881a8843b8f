"""Matryoshka loss (dims {768, 512, 256, 128}, weights 1) next to the plain loss at the same shape: forward + backward, CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("nccl", rank=0, world_size=1)
import contrastors_b200 as cb
from contrastors_b200 import loss as L

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
ls = cb.LogitScale(logit_scale=50.0).cuda()
dims, w = [768, 512, 256, 128], [1.0, 1.0, 1.0, 1.0]
for n, m in [(256, 2048), (2048, 16384), (8192, 8192)]:
    g = torch.Generator().manual_seed(1)
    q = torch.randn(n, 768, generator=g).cuda().requires_grad_()
    d = torch.randn(m, 768, generator=g).cuda().requires_grad_()

    def plain():
        q.grad = d.grad = None
        spec = L._NceSpec(label_offset=0, label_stride=m // n, mult=1.0, normalize=True)
        L._fused_infonce(q, d, ls, spec).backward()

    def mat():
        q.grad = d.grad = None
        cb.matryoshka_clip_loss(q, d, ls, dims, w).backward()

    def mat_loop():  # the per-prefix loop the single accumulation replaces (dims not multiples of 64 would take this path)
        q.grad = d.grad = None
        tot = 0
        for k in dims:
            spec = L._NceSpec(label_offset=0, label_stride=m // n, mult=1.0, normalize=True, dims=[k], weights=[1.0])
            tot = tot + L._fused_infonce(q, d, ls, spec)
        tot.backward()

    p, s1, s4 = timeit(plain), timeit(mat), timeit(mat_loop)
    res[f"{n}x{m}x768"] = dict(plain_ms=p, matryoshka_single_accumulation_ms=s1, per_prefix_loop_ms=s4, ratio_single_vs_plain=s1 / p,
                               ratio_loop_vs_plain=s4 / p)
    print(n, m, res[f"{n}x{m}x768"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_matryoshka.json", "w"), indent=1)
dist.destroy_process_group()
