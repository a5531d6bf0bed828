import sys, time, torch
sys.path.insert(0, '.')
x = torch.empty(820_000_000 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1, 4, 16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step = x.numel() // n
    for i in range(n):
        d[i*step:(i+1)*step].copy_(x[i*step:(i+1)*step], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('H2D pinned %d chunks: %.1f ms  %.1f GB/s' % (n, dt*1e3, x.numel()*4/dt/1e9))
import bench
from prcv2025reid_b200 import engine, synth
seed, n_ids, gpi, k, qpi = bench.WORKLOADS['c4']
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda')
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid); case.gallery_raw = None
w = synth.weights_tensor(device='cuda')
hq, hm, hp, he = [t.cpu().pin_memory() for t in (case.query_raw, case.mod_id, case.q_pid, case.excl)]
print('pinned:', hq.is_pinned(), hq[0:100].is_pinned())
for qb in (32768, 16384):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = engine.retrieve(shard, None, None, hp, he, host_queries=(hq, hm, w), query_block=qb)
        torch.cuda.synchronize(); print('e2e block %d: %.1f ms' % (qb, (time.perf_counter()-t0)*1e3))
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, w)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = engine.retrieve(shard, q32, q16, case.q_pid, case.excl)
    torch.cuda.synchronize(); print('resident: %.1f ms' % ((time.perf_counter()-t0)*1e3))
