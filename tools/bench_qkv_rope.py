"""A/B: QKV GEMM + separate RoPE kernel vs the GEMM with RoPE fused into its epilogue (CUDA events, L2 flushed)."""
import os, sys
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
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


T, d, H, Dh, S = 32768, 768, 12, 64, 512
x = torch.randn(T, d, device="cuda").to(torch.bfloat16)
w = (torch.randn(3 * d, d, device="cuda") * 0.02).to(torch.bfloat16)
pos = (torch.arange(T, device="cuda") % S).to(torch.int32)
inv = 1.0 / (10000.0 ** (torch.arange(0, Dh, 2, dtype=torch.float32) / Dh))
fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
cos_t, sin_t = torch.cos(fr).cuda(), torch.sin(fr).cuda()
inv_freq = inv.cuda().contiguous()


def separate():
    qkv = ops.gemm(x, w)
    ops.rope_inplace(qkv, pos, cos_t, sin_t, H, Dh)
    return qkv


a = separate()
b = ops.gemm_qkv_rope(x, w, pos, inv_freq, 2 * d)
print("max abs diff", (a.float() - b.float()).abs().max().item())
print("gemm only      %.1f us" % (1e3 * timeit(lambda: ops.gemm(x, w))))
print("gemm + rope    %.1f us" % (1e3 * timeit(separate)))
print("fused epilogue %.1f us" % (1e3 * timeit(lambda: ops.gemm_qkv_rope(x, w, pos, inv_freq, 2 * d))))
