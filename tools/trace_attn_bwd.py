"""Phase trace of the attention backward kernel (clock64 stamps per role and query tile).

    python tools/trace_attn_bwd.py --build     # here, no GPU: nvcc -DCX_ATTN_TRACE -> tools/_trace/libcx_trace.so (travels with gpurun)
    python tools/trace_attn_bwd.py             # on the B200: run once with the trace on, print the per-phase medians

The trace library is a separate build of csrc/cx_attn.cu; the product library contains none of the trace code.
Roles: 0 / 1 = first worker warp of query quarters 0 / 2, 2 = MMA warp (its stamp 6 sits between "top" and "p_ready": Q/dO stage
landed), 3 = first dQ-drain warp, 4 = block-level stamps.  VARIANTS takes ablation builds (extra -D flags): round 2 used them to
remove one consumer of the shared-memory port at a time (profiles/r02l_attn_bwd4_ablations.txt).
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_trace")
LIB = os.path.join(OUT, "libcx_trace.so")


VARIANTS = {"": []}   # name -> extra -D flags: A/B builds of schedule variants or ablations go here (CX_TRACE_VARIANT=name selects one)


def build():
    """The plain trace library plus ablation variants (wrong results on purpose: each removes one consumer of a shared resource)."""
    from contrastors_b200 import build as b
    b.build()
    os.makedirs(OUT, exist_ok=True)
    others = [os.path.join(b.OBJ, f[:-3] + ".o") for f in b._sources() if f != "cx_attn.cu"]
    for name, flags in VARIANTS.items():
        obj = os.path.join(OUT, f"cx_attn_trace_{name}.o")
        lib = LIB if not name else LIB.replace(".so", f"_{name}.so")
        subprocess.run([b.NVCC, *b.FLAGS, "-DCX_ATTN_TRACE", *flags, "-c", os.path.join(b.CSRC, "cx_attn.cu"), "-o", obj], check=True)
        subprocess.run([b.NVCC, "-shared", "-o", lib, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        print(lib)


def main():
    import numpy as np
    import torch
    from contrastors_b200 import ops
    variant = os.environ.get("CX_TRACE_VARIANT", "")
    lib = C.CDLL(LIB if not variant else LIB.replace(".so", f"_{variant}.so"))
    lib.cx_attn_trace_set.argtypes = [C.c_void_p]
    lib.cx_attn_bwd.argtypes = [C.c_void_p] * 8 + [C.c_int] * 5 + [C.c_float, C.c_void_p, C.c_void_p]
    lib.cx_last_error.restype = C.c_char_p
    nseq, S, H, Dh = 64, 512, 12, 64
    T = nseq * S
    torch.manual_seed(0)
    qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(torch.bfloat16)
    dout = torch.randn(T, H * Dh, device="cuda").to(torch.bfloat16)
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    scale = 1.0 / math.sqrt(Dh)
    out, lse = ops.attn_fwd(qkv, cu, S, H, Dh, scale)
    ref = ops.attn_bwd(qkv, out, dout, lse, cu, S, H, Dh, scale)
    from contrastors_b200 import _lib
    plib = _lib.load()
    ref_acc = torch.empty(T, H * Dh, device="cuda", dtype=torch.float32)
    ref_dqkv = torch.empty_like(qkv)
    ref_delta = torch.empty(H, T, device="cuda", dtype=torch.float32)
    _lib.check(plib.cx_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), cu.data_ptr(), ref_dqkv.data_ptr(),
                                ref_acc.data_ptr(), ref_delta.data_ptr(), T, nseq, S, H, Dh, scale, 0,
                                torch.cuda.current_stream().cuda_stream), "cx_attn_bwd")
    nblk = (S // 128) * H * nseq
    trace = torch.zeros(nblk * 5 * 8 * 8, dtype=torch.int64, device="cuda")
    dqkv = torch.empty_like(qkv)
    dq_acc = torch.empty(T, H * Dh, device="cuda", dtype=torch.float32)
    delta = torch.empty(H, T, device="cuda", dtype=torch.float32)

    def run():
        rc = lib.cx_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), cu.data_ptr(), dqkv.data_ptr(),
                             dq_acc.data_ptr(), delta.data_ptr(), T, nseq, S, H, Dh, scale, None,
                             torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.cx_last_error()
    run()  # warm, trace off
    torch.cuda.synchronize()
    assert lib.cx_attn_trace_set(trace.data_ptr()) == 0
    run()
    torch.cuda.synchronize()
    lib.cx_attn_trace_set(None)
    for _ in range(30):  # let the clocks ramp before timing
        run()
    evs = []
    for _ in range(11):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b_.record(); torch.cuda.synchronize()
        evs.append(a.elapsed_time(b_) * 1e3)
    print("variant", variant or "plain", "cx_attn_bwd (delta + kernel) us, trace off:", sorted(evs)[5])
    # the drained dk/dv slots match the product library's (dq is finalized by a separate kernel, not compared here)
    HD = H * Dh
    err = (dqkv[:, HD:].float() - ref[:, HD:].float()).abs().max().item() / ref[:, HD:].float().abs().max().item()
    err_q = (dq_acc - ref_acc).abs().max().item() / ref_acc.abs().max().item()
    print("variant", variant or "plain", "max rel err vs product: dk/dv", err, "dq_acc", err_q)
    assert err < 2e-2 and err_q < 2e-2, f"trace build disagrees with the product build: {err} {err_q}"
    t = trace.cpu().numpy().reshape(nblk, 5, 8, 8)
    nq = S // 128
    res = {}

    def med(x):
        return float(np.median(x))
    t0 = t[:, 4, 0, 0]
    res["block_total"] = med(t[:, 4, 0, 4] - t0)
    res["setup_to_sync"] = med(t[:, 4, 0, 1] - t0)
    res["acc_full_at"] = med(t[:, 4, 0, 2] - t0)
    res["worker_epilogue"] = med(t[:, 4, 0, 3] - t[:, 4, 0, 2])
    res["drain_done_at"] = med(t[:, 4, 0, 5] - t0)
    # role 0 / 1 = first worker warp of query quarters 0 / 2 (X then Y in the same warp)
    names = {0: ["top", "s_full", "x_loaded", "x_done", "dp_dq_full", "y_loaded", "y_done", "ds_arrived"],
             2: ["top", "p_ready", "issue1", "ds_ready", "dq_free", "issue2"]}
    names[1] = names[0]
    names[3] = ["top", "dq_full", "loaded", "stage_free_bar", "stored"]
    for role, label in ((0, "worker0"), (1, "worker1"), (2, "mma"), (3, "drain")):
        nm = names[role]
        for i in range(nq):
            row = {"start_at": med(t[:, role, i, 0] - t0)}
            for p in range(1, len(nm)):
                row[nm[p]] = med(t[:, role, i, p] - t[:, role, i, p - 1])
            row["tile_total"] = med(t[:, role, i, len(nm) - 1] - t[:, role, i, 0])
            res[f"{label}_tile{i}"] = row
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "attn_bwd_trace.json"), "w"), indent=1)
    np.save(os.path.join(ROOT, "gpurun_out", "attn_bwd_trace_first296.npy"), t[:296])
    for k, v in res.items():
        print(k, v)


if __name__ == "__main__":
    build() if "--build" in sys.argv else main()
