"""A few SDM C5 steps (10 pairs, bf16, P x K = 64 x 8) for ncu."""
import sys
sys.path.insert(0, '.')
import torch
from prcv2025reid_b200 import synth
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
P, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 8)
dtype = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.bfloat16
npairs = int(sys.argv[4]) if len(sys.argv) > 4 else 10
feats, labels = synth.make_sdm_batch(2002, P, K, n_modalities=5, dtype=dtype, device="cuda")
y = (labels[:, None] == labels[None, :]).float()
pairs = [(a, b) for a in range(5) for b in range(a)][:npairs]
qs = [feats[a].clone().requires_grad_(True) for a, b in pairs]
vs = [feats[b].clone().requires_grad_(True) for a, b in pairs]
for _ in range(4):
    losses = sdm_loss_pairs(qs, vs, [y] * len(pairs), tau=0.2)
    losses.sum().backward()
torch.cuda.synchronize()
print("ok", losses.tolist()[:3])
