"""One GradCache chunk (64 x 512 tokens) of nomic-bert-base: no-grad forward, then forward+backward (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastors_b200 as cb
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = cb.BiEncoder(cb.BiEncoderConfig(encoder=cb.nomic_bert_base())).cuda()
model.train()
ids = torch.randint(0, 30000, (64, 512), device="cuda")
lens = torch.full((64,), 512)
ones = torch.ones(64, 512, dtype=torch.long, device="cuda")
g = torch.randn(64, 768, device="cuda")
for _ in range(reps):
    with torch.no_grad():
        model(ids, attention_mask=ones, seq_lens=lens)
    e = model(ids, attention_mask=ones, seq_lens=lens)["embedding"]
    e.backward(g)
torch.cuda.synchronize()
print("done")
