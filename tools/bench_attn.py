"""Micro-timings of the attention kernels per kernel generation (CUDA events, L2 flushed). Informational, not bench.py."""
import json
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=8, warm=2):
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


MODES = [(4, 4)]  # the one generation left: two-threads-per-row forward, transposed-score backward with four threads per key row
TAG = "fwd4_bwd4"
res = {}
for name, nseq, S, H in [("bert_64x512", 64, 512, 12), ("vit_256x197", 256, 197, 12)]:
    Dh = 64
    T = nseq * S
    torch.manual_seed(0)
    qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(torch.bfloat16)
    dout = torch.randn(T, H * Dh, device="cuda").to(torch.bfloat16)
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = 1.0 / math.sqrt(Dh)
    ref_out = ref_dqkv = None
    for fm, bm in MODES:
        try:
            out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
            dqkv = ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale)
            torch.cuda.synchronize()
            if ref_out is None:
                ref_out, ref_dqkv = out.float(), dqkv.float()
            eo = (out.float() - ref_out).abs().max().item()
            eg = (dqkv.float() - ref_dqkv).abs().max().item() / ref_dqkv.abs().max().item()
            chk = [out.float().abs().sum().item(), dqkv.float().abs().sum().item()]
            f = timeit(lambda: ops.attn_fwd(qkv, cu, S, H, Dh, scale))
            b = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale))
            res[f"{name}_fwd{fm}_bwd{bm}"] = dict(fwd_ms=f, fwd_tflops=4.0 * T * S * H * Dh / f / 1e9, bwd_ms_incl_delta_finalize=b,
                                                   bwd_tflops=10.0 * T * S * H * Dh / b / 1e9, max_abs_out_vs_first=eo, checksum=chk,
                                                   rel_dqkv_vs_first=eg)
        except Exception as ex:  # keep going: the other generations are still informative
            res[f"{name}_fwd{fm}_bwd{bm}"] = dict(error=str(ex))
        print(name, fm, bm, res[f"{name}_fwd{fm}_bwd{bm}"], flush=True)
    # the kernel the reference calls (layers/attention.py:158-181): FlashAttention-2 varlen qkv-packed, same shape, same box
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
        q3 = qkv.view(T, 3, H, Dh).clone().requires_grad_()
        o = flash_attn_varlen_qkvpacked_func(q3, cu, S, 0.0, softmax_scale=scale, causal=False)
        o.backward(dout.view(T, H, Dh))
        f = timeit(lambda: flash_attn_varlen_qkvpacked_func(q3.detach(), cu, S, 0.0, softmax_scale=scale, causal=False))

        def fa2_fb():
            q3.grad = None
            flash_attn_varlen_qkvpacked_func(q3, cu, S, 0.0, softmax_scale=scale, causal=False).backward(dout.view(T, H, Dh))
        fb = timeit(fa2_fb)
        err = (o.detach().reshape(T, H * Dh).float() - ref_out).abs().max().item() if ref_out is not None else None
        res[f"{name}_flash_attn2"] = dict(fwd_ms=f, fwd_tflops=4.0 * T * S * H * Dh / f / 1e9, fwd_plus_bwd_ms=fb,
                                          bwd_ms_by_difference=fb - f, bwd_tflops=10.0 * T * S * H * Dh / max(fb - f, 1e-6) / 1e9,
                                          max_abs_out_vs_ours=err)
    except Exception as ex:
        res[f"{name}_flash_attn2"] = dict(error=f"{type(ex).__name__}: {ex}"[:200])
    print(name, "flash_attn2", res[f"{name}_flash_attn2"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/bench_attn_{TAG}.json", "w"), indent=1)
