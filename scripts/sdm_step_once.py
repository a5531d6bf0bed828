"""C2 / C5 SDM step through reid_sdm_step (SdmStep): a few eager steps for ncu, then the CUDA-graph replay time."""
import sys
sys.path.insert(0, '.')
import torch
from prcv2025reid_b200 import synth
from prcv2025reid_b200.sdm_loss import SdmStep, SdmGraphStep
P, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 2)
dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
feats, labels = synth.make_sdm_batch(2001, P, K, n_modalities=5, dtype=dtype, device="cuda")
y = (labels[:, None] == labels[None, :]).float()
qs = [feats[m] for m in range(1, 5)]
vs = [feats[0]] * 4
step = SdmStep(qs, vs, [y] * 4, tau=0.2)
for _ in range(6):
    losses = step.run()
torch.cuda.synchronize()
g = SdmGraphStep(qs, vs, [y] * 4, tau=0.2)
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    s.record()
    for _ in range(200):
        g.replay()
    e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) * 1e3 / 200)
print("losses", [round(v, 6) for v in losses.tolist()], "graph us/step %.2f (%d launches)" % (best, step.launches))
