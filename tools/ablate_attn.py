"""Timing ablations of the pipelined attention kernels (CX_ATTN_ABLATE bit mask; results are wrong by design):
1 = no dQ TMA reduce-add, 2 = no exponentials, 4 = no P/dS stores, 8 = no TMEM score reads (the filler uses I2F, i.e. the
SFU: read that bit with care), 16 = 1/8 of the dV/dK/dQ MMAs, 32 = interleaved issue order of the independent MMA chains
(results stay correct).  Forward ablations exist in the 64-key sub-tile kernel (CX_ATTN_FWD=3) only."""
import json
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=6, warm=2):
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


nseq, S, H, Dh = 64, 512, 12, 64
T = nseq * S
torch.manual_seed(0)
qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(torch.bfloat16)
dout = torch.randn(T, H * Dh, device="cuda").to(torch.bfloat16)
cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
scale = 1.0 / math.sqrt(Dh)
os.environ["CX_ATTN_FWD"], os.environ["CX_ATTN_BWD"] = "3", "2"
out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
res = {}
# fixed overheads of the backward wrapper (delta + zero fill + finalize): time them via the pieces
dq_acc = torch.zeros(T, H * Dh, device="cuda")
res["zero_fill_ms"] = timeit(lambda: dq_acc.zero_())
for fm in (3,):
    for ab in (0, 2, 4, 8, 2 | 4, 2 | 8, 2 | 4 | 8):
        os.environ["CX_ATTN_FWD"], os.environ["CX_ATTN_ABLATE"] = str(fm), str(ab)
        res[f"fwd{fm}_ablate{ab}_ms"] = timeit(lambda: ops.attn_fwd(qkv, cu, S, H, Dh, scale))
        print(f"fwd{fm} ablate {ab}: {res[f'fwd{fm}_ablate{ab}_ms']:.4f} ms", flush=True)
os.environ["CX_ATTN_FWD"] = "3"
for ab in (0, 1, 2, 4, 8, 16, 1 | 16, 2 | 4, 2 | 4 | 8, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 16):
    os.environ["CX_ATTN_ABLATE"] = str(ab)
    res[f"bwd2_ablate{ab}_ms"] = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale))
    print(f"bwd2 ablate {ab}: {res[f'bwd2_ablate{ab}_ms']:.4f} ms (incl. delta, zero fill, finalize)", flush=True)
os.environ["CX_ATTN_ABLATE"] = "0"
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ablate_attn.json", "w"), indent=1)
