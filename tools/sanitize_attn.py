"""compute-sanitizer target: one ragged attention forward + backward (run as `compute-sanitizer --tool memcheck python tools/sanitize_attn.py`)."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import ops

H, Dh = 2, 64
lens = [300, 17, 512, 129, 1, 640]
T = sum(lens)
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
torch.manual_seed(0)
qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(torch.bfloat16)
dout = torch.randn(T, H * Dh, device="cuda").to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, cu, max(lens), H, Dh, 1.0 / math.sqrt(Dh))
dqkv = ops.attn_bwd(qkv, out, dout, lse, cu, max(lens), H, Dh, 1.0 / math.sqrt(Dh))
torch.cuda.synchronize()
print("attention fwd+bwd done", float(out.float().abs().sum()), float(dqkv.float().abs().sum()))
