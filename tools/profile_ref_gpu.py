"""One GradCache chunk (64 x 512 tokens per tower) of the UNMODIFIED reference on the GPU (its pure-PyTorch NomicBertModel under
bf16 autocast + its grad_cache_loss, via oracle/ref_tower.py): for the ncu kernel list of bench.py's gpu_baseline leg."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bench._ensure_group("gloo")
ref = bench.ReferenceStep(torch.device("cuda", 0), 64, 64)
for _ in range(reps):
    loss = ref.step()
torch.cuda.synchronize()
print("done", ref.kind, loss)
