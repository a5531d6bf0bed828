"""Debug helper: run the fused path on a workload slice and report flag reasons / candidate counts."""
import sys, torch
sys.path.insert(0, '.')
from prcv2025reid_b200 import engine, synth, _cabi
import bench
w = sys.argv[1] if len(sys.argv) > 1 else 'c3b'
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
seed, n_ids, gpi, k, qpi = bench.WORKLOADS[w]
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
engine._DEBUG_KEEP = {}
res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode='fused', want_ap=True)
d = engine._DEBUG_KEEP
fl = d['flag'].cpu(); cc = d['cand_count'].cpu()
print('flag reasons: overflow %d topk %d cmc %d of %d' % ((fl & 1).ne(0).sum(), (fl & 2).ne(0).sum(), (fl & 4).ne(0).sum(), fl.numel()))
print('cand_count per (q,chunk): mean %.1f max %d  shape %s' % (cc.float().mean(), cc.max(), tuple(cc.shape)))
ex = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode='exact', want_ap=True)
print('fused', res.metrics, 'flagged', res.n_flagged)
print('exact', ex.metrics)
print('mAP diff %.2e  max per-query AP diff %.2e  topk equal %s' % (res.metrics['mAP'] - ex.metrics['mAP'], (res.ap - ex.ap).abs().max().item(), torch.equal(res.top_idx, ex.top_idx)))
