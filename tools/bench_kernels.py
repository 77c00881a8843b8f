"""Micro-timings of the C-ABI kernels (CUDA events, L2 flushed between iterations). Informational, not bench.py."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
SHAPES = [(32768, 2304, 768, 0, 0, False), (32768, 768, 768, 0, 0, False), (32768, 768, 3072, 0, 0, False),
          (32768, 3072, 768, 0, 1, False), (32768, 768, 6144, 0, 1, False), (32768, 768, 2304, 0, 1, False), (32768, 768, 768, 0, 1, False),
          (2304, 768, 32768, 1, 1, True), (6144, 768, 32768, 1, 1, True), (8192, 8192, 8192, 0, 0, False)]
for (M, N, K, am, bm, f32) in SHAPES:
    a = torch.randn((M, K) if am == 0 else (K, M), device="cuda").to(torch.bfloat16)
    b = torch.randn((N, K) if bm == 0 else (K, N), device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    ent = {}
    for cl in (2, 4):  # CTA pairs vs two pairs sharing B by TMA multicast (M % 512 == 0 only; otherwise both are pair mode)
        ops.gemm_select_cluster(cl)
        ms = timeit(lambda: ops.gemm(a, b, a_major=am, b_major=bm, out=out))
        ent[f"cluster{cl}_ms"] = ms
        ent[f"cluster{cl}_tflops"] = 2 * M * N * K / ms / 1e9
    ops.gemm_select_cluster(0)
    A = a if am == 0 else a.t()
    B = b.t() if bm == 0 else b
    ms_t = timeit(lambda: torch.matmul(A, B))
    ent.update(torch_ms=ms_t, torch_tflops=2 * M * N * K / ms_t / 1e9)
    res[f"gemm_{M}x{N}x{K}_a{am}b{bm}{'_f32' if f32 else ''}"] = ent
    print(list(res.items())[-1], flush=True)
# fc1 + SwiGLU (inference form and training form that also stores the pre-activations)
x = torch.randn(32768, 768, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(6144, 768, device="cuda") / 28).to(torch.bfloat16)
for keep in (False, True):
    ent = {}
    for cl in (2, 4):
        ops.gemm_select_cluster(cl)
        ms = timeit(lambda: ops.gemm_swiglu(x, w1, keep_preact=keep))
        ent[f"cluster{cl}_ms"] = ms
        ent[f"cluster{cl}_tflops"] = 2 * 32768 * 6144 * 768 / ms / 1e9
    ops.gemm_select_cluster(0)
    res[f"gemm_swiglu_32768x3072x768_keep{int(keep)}"] = ent
    print(list(res.items())[-1], flush=True)
dm = torch.randn(32768, 768, device="cuda").to(torch.bfloat16)
w2 = (torch.randn(768, 3072, device="cuda") / 28).to(torch.bfloat16)
yg = torch.randn(32768, 6144, device="cuda").to(torch.bfloat16)
ms_f = timeit(lambda: ops.gemm_swiglu_bwd(dm, w2, yg))
ms_2 = timeit(lambda: ops.swiglu_bwd(ops.gemm(dm, w2, b_major=1), yg))
res["gemm_swiglu_bwd_32768x3072x768"] = dict(fused_ms=ms_f, dgrad_then_swiglu_bwd_ms=ms_2, hbm_bytes=2 * 32768 * 6144 * 2 + 32768 * 768 * 2,
                                             fused_gbs=(2 * 32768 * 6144 * 2 + 32768 * 768 * 2) / ms_f / 1e6)
print(list(res.items())[-1], flush=True)
del x, w1, a, b, out, dm, w2, yg

n, m, dim = 2048, 16384, 768
g = torch.Generator().manual_seed(1234)
q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=-1).cuda()
d = torch.nn.functional.normalize(torch.randn(m, dim, generator=g), dim=-1).cuda()
qb, _ = ops.rows_to_bf16(q)
db, _ = ops.rows_to_bf16(d)
ws = ops.infonce_workspace(n, m, dim, "cuda")
lse, argmax, ll, stats = ops.infonce_fwd(qb, db, dim, 50.0, None, None, None, 0, 8, ws)
dq = torch.empty(n, dim, device="cuda")
dd = torch.empty(m, dim, device="cuda")
st = torch.zeros(4, device="cuda")
f = timeit(lambda: ops.infonce_fwd(qb, db, dim, 50.0, None, None, None, 0, 8, ws))
bwd = timeit(lambda: ops.infonce_bwd(qb, db, dim, 50.0, None, None, None, 0, 8, lse, 1.0 / n, None, dq, dd, False, st, ws))
res["infonce_fwd_ms"] = f
res["infonce_bwd_ms"] = bwd
res["infonce_tflops_algorithmic"] = 6 * n * m * dim / (f + bwd) / 1e9
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_kernels.json", "w"), indent=1)
