"""Run the InfoNCE forward+backward a few times at the config-2 per-rank shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastors_b200 import ops
n, m, dim = 2048, 16384, 768
g = torch.Generator().manual_seed(1234)
q = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=-1).cuda()
d = torch.nn.functional.normalize(torch.randn(m, dim, generator=g), dim=-1).cuda()
qb, _ = ops.rows_to_bf16(q)
db, _ = ops.rows_to_bf16(d)
ws = ops.infonce_workspace(n, m, dim, "cuda")
dq = torch.empty(n, dim, device="cuda")
dd = torch.empty(m, dim, device="cuda")
st = torch.zeros(4, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    lse, argmax, ll, stats = ops.infonce_fwd(qb, db, dim, 50.0, None, None, None, 0, 8, ws)
    ops.infonce_bwd(qb, db, dim, 50.0, None, None, None, 0, 8, lse, 1.0 / n, None, dq, dd, False, st, ws)
torch.cuda.synchronize()
print("loss", stats[0].item() / n, "dlogit", st[2].item())
