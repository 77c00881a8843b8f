"""Decode the clock64 event stamps the pipelined attention kernels record through cx_debug_attn_trace (profiling hook)."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import _lib, ops

nseq, S, H, Dh = 64, 512, 12, 64
T = nseq * S
torch.manual_seed(0)
qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(torch.bfloat16)
dout = torch.randn(T, H * Dh, device="cuda").to(torch.bfloat16)
cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
scale = 1.0 / math.sqrt(Dh)
lib = _lib.load()
ncta = ((S + 127) // 128) * H * nseq
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/attn_trace.txt", "w")


def say(*a):
    print(*a, flush=True)
    print(*a, file=log, flush=True)


FWD = {3: "setup done", 4: "mma: Q,K0 landed", 5: "mma: all issued", 22: "softmax: O complete", 23: "softmax: stored", 24: "exit",
       26: "tma: tile0 issued", 27: "tma: tile1 issued", 28: "tma: tile2 issued", 29: "tma: tile3 issued"}
for u in range(8):
    FWD[6 + u] = f"softmax: S({u}) ready"
    FWD[14 + u] = f"softmax: P({u}) arrive"
    FWD[32 + u] = f"mma: P({u}) seen"
BWD = {3: "setup done", 4: "mma: K,V,Q0,dO0 landed", 5: "mma: all issued", 24: "worker: dK/dV complete", 25: "worker: stored", 48: "drain: done",
       50: "exit"}
for i in range(4):
    BWD[8 + 4 * i] = f"worker: S({i}) ready"
    BWD[9 + 4 * i] = f"worker: P({i}) arrive"
    BWD[10 + 4 * i] = f"worker: dP({i}) ready"
    BWD[11 + 4 * i] = f"worker: dS({i}) arrive"
    BWD[28 + 3 * i] = f"drain: dQ({i}) ready"
    BWD[29 + 3 * i] = f"drain: stage free({i})"
    BWD[30 + 3 * i] = f"drain: reduce issued({i})"
    BWD[40 + 2 * i] = f"mma: P({i}) seen"
    BWD[41 + 2 * i] = f"mma: dS({i}) seen"


def report(name, tr, names):
    tr = tr.cpu()
    t0 = tr[:, 0:1]
    rel = (tr - t0).double()
    say(f"== {name}: {tr.shape[0]} CTAs; median / p10 / p90 cycles since CTA entry")
    for slot in sorted(names, key=lambda k: rel[:, k].median().item()):
        v = rel[:, slot]
        say(f"  {names[slot]:28s} {v.median().item():9.0f} {v.quantile(0.1).item():9.0f} {v.quantile(0.9).item():9.0f}")
    gt = tr[:, 2].double()
    say(f"  kernel span by globaltimer: {(gt.max() - gt.min()).item() / 1e3:.1f} us between first and last CTA start")
    sm = tr[:, 1]
    per_sm = torch.bincount(sm, minlength=148).float()
    say(f"  CTAs per SM: min {per_sm.min().item():.0f} max {per_sm.max().item():.0f}")
    # per-SM timeline of one SM: start offsets (us) of its CTAs
    one = (sm == sm[0]).nonzero().flatten()
    starts = sorted(((gt[one] - gt.min()) / 1e3).tolist())
    say("  CTA start times on SM %d (us): %s" % (sm[0].item(), " ".join(f"{x:.1f}" for x in starts)))


for fm in (3, 1 + 2):
    os.environ["CX_ATTN_FWD"], os.environ["CX_ATTN_BWD"] = "3", "2"  # forward stamps exist in the 64-key sub-tile kernel only
    break
for ab in (0,):
    os.environ["CX_ATTN_ABLATE"] = str(ab)
    out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)  # warm
    tr = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
    _lib.check(lib.cx_debug_attn_trace(tr.data_ptr()), "trace on")
    out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
    torch.cuda.synchronize()
    _lib.check(lib.cx_debug_attn_trace(None), "trace off")
    report(f"attn_fwd2 (P in TMEM), ablate={ab}", tr, FWD)
os.environ["CX_ATTN_ABLATE"] = "0"
out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
for ab in (0,):
    os.environ["CX_ATTN_ABLATE"] = str(ab)
    ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale)  # warm
    tr = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
    _lib.check(lib.cx_debug_attn_trace(tr.data_ptr()), "trace on")
    ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale)
    torch.cuda.synchronize()
    _lib.check(lib.cx_debug_attn_trace(None), "trace off")
    report(f"attn_bwd2, ablate={ab}", tr, BWD)
os.environ["CX_ATTN_ABLATE"] = "0"
