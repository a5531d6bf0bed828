import sys, time
sys.path.insert(0, '.')
import torch
from prcv2025reid_b200 import synth, _cabi
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
P, K, npairs, dtype = 64, 8, int(sys.argv[1]), torch.bfloat16
feats, labels = synth.make_sdm_batch(2002, P, K, n_modalities=5, dtype=dtype, device="cuda")
y = (labels[:, None] == labels[None, :]).float()
pairs = [(a, b) for a in range(5) for b in range(a)][:npairs]
qs = [feats[a].clone().requires_grad_(True) for a, b in pairs]
vs = [feats[b].clone().requires_grad_(True) for a, b in pairs]
ys = [y] * len(pairs)
def stats():
    s = torch.cuda.memory_stats()
    return s["num_device_alloc"], s["num_device_free"], s["reserved_bytes.all.current"] >> 20
tf, tb, tg = [], [], []
for it in range(40):
    t0 = time.perf_counter()
    losses = sdm_loss_pairs(qs, vs, ys, tau=0.2)
    t1 = time.perf_counter()
    losses.sum().backward()
    t2 = time.perf_counter()
    for t in qs + vs:
        t.grad = None
    t3 = time.perf_counter()
    tf.append((t1 - t0) * 1e6); tb.append((t2 - t1) * 1e6); tg.append((t3 - t2) * 1e6)
    if it % 8 == 0: print(it, "fwd %.0f bwd %.0f gradreset %.0f us" % (tf[-1], tb[-1], tg[-1]), stats())
torch.cuda.synchronize()
print("median fwd %.0f bwd %.0f reset %.0f" % (sorted(tf)[20], sorted(tb)[20], sorted(tg)[20]))
# how long does the queue get? sync each step
ts = []
for it in range(20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    losses = sdm_loss_pairs(qs, vs, ys, tau=0.2); losses.sum().backward()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e6, (t2 - t0) * 1e6))
print("synced steps: host %.0f us, total %.0f us" % (sorted(a for a, b in ts)[10], sorted(b for a, b in ts)[10]))
