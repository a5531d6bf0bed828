import sys, torch
sys.path.insert(0, '.')
from prcv2025reid_b200 import engine, synth
Q, k, d = 100000, 4, 512
raw = torch.randn(Q, k, d, device='cuda'); mid = torch.randint(1, 5, (Q, k), device='cuda', dtype=torch.int32)
w = synth.weights_tensor(device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for i in range(12):
    flush.zero_()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); engine.fuse_queries(raw, mid, w); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = sorted(ts[2:])[len(ts[2:]) // 2]
b = Q * (k * d * 4 + d * 6)
print("K2 %.4f ms  %.0f GB/s  frac %.3f" % (ms, b / ms / 1e6, b / ms / 1e6 / 6534.1))
