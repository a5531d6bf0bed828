"""The launch schedules engine.retrieve picks for 4- and 2-GPU shard sizes (one 100k-query launch on a 250k-row gallery; 4 + 1.3
waves on a 500k-row gallery), run on ONE GPU and checked against the all-fp32 path: identical CMC / top-10, mAP within 1e-4."""
import sys, torch
sys.path.insert(0, '.')
from prcv2025reid_b200 import engine, synth
for n_ids, qpi in ((6250, 16), (12500, 8)):
    case = synth.make_retrieval_case(1005, n_ids, 40, 4, qpi, device='cuda')
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
    case.gallery_raw = None
    q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
    print("gallery %d rows, %d queries, blocks %s" % (shard.G_local, case.Q, engine.resident_query_blocks(case.Q, 148, shard.G_local)))
    a = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode="fused")
    b = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode="exact")
    same = bool(torch.equal(a.top_idx, b.top_idx))
    print("  fused", a.metrics, "flagged", a.n_flagged); print("  exact", b.metrics, " top-10 identical:", same)
    assert a.path == "fused" and abs(a.metrics["mAP"] - b.metrics["mAP"]) <= 1e-4 and all(a.metrics[k] == b.metrics[k] for k in ("R@1", "R@5", "R@10")) and same
    del shard, q32, q16, a, b, case
    torch.cuda.empty_cache()
print("ok")
