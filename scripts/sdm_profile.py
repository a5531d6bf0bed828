"""Where an SDM step spends its time: cProfile of the host side + per-call device times."""
import cProfile, pstats, sys, io
sys.path.insert(0, '.')
import torch
from prcv2025reid_b200 import synth, _cabi
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
P, K, npairs, dtype = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), torch.bfloat16 if sys.argv[4] == "bf16" else torch.float32)
feats, labels = synth.make_sdm_batch(2002, P, K, n_modalities=5, dtype=dtype, device="cuda")
y = (labels[:, None] == labels[None, :]).float()
pairs = [(a, b) for a in range(5) for b in range(a)][:npairs]
qs = [feats[a].clone().requires_grad_(True) for a, b in pairs]
vs = [feats[b].clone().requires_grad_(True) for a, b in pairs]
ys = [y] * len(pairs)
def step():
    losses = sdm_loss_pairs(qs, vs, ys, tau=0.2)
    losses.sum().backward()
    for t in qs + vs:
        t.grad = None
for _ in range(10): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host issue %.1f us/step, with sync %.1f us/step" % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
_cabi.PROFILE = []
for _ in range(10): step()
torch.cuda.synchronize()
agg = {}
for name, a, b in _cabi.PROFILE:
    agg.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
_cabi.PROFILE = None
for k, v in agg.items():
    print(k, "device us: median %.1f min %.1f" % (sorted(v)[len(v) // 2], min(v)))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
