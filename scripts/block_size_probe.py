"""C4 on one GPU: step time of engine.retrieve with the default launch schedule (two waves per launch) against larger launches."""
import sys, torch
sys.path.insert(0, '.')
import bench
from prcv2025reid_b200 import engine, synth
seed, n_ids, gpi, k, qpi = bench.WORKLOADS['c4']
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda')
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
case.gallery_raw = None
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
wave = 74 * 256
for qb in (None, 3 * wave, 4 * wave, 6 * wave, None):
    for _ in range(2):
        r = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, query_block=qb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        r = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, query_block=qb)
    b.record(); torch.cuda.synchronize()
    print("query_block %s: %.2f ms per step, mAP %.10f flagged %d" % (qb, a.elapsed_time(b) / 4, r.metrics["mAP"], r.n_flagged))
