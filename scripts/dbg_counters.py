"""Cycle counters of the fused kernel (needs a build with the counter bit: scripts/build_fused_variant.sh dbg -DREID_DEBUG=8192,
then REID_LIB=prcv2025reid_b200/variants/libreid_dbg.so python scripts/dbg_counters.py)."""
import ctypes, os, sys, torch
sys.path.insert(0, '.')
import bench
from prcv2025reid_b200 import engine, synth, _cabi
bits = 8192
seed, n_ids, gpi, k, qpi = bench.WORKLOADS['c4']
nq = 37888
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid); case.gallery_raw = None
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
raw = ctypes.CDLL(_cabi.LIB_PATH)
out = (ctypes.c_ulonglong * 16)()
engine.retrieve(shard, q32, q16, case.q_pid, case.excl)
torch.cuda.synchronize(); raw.reid_debug_counters(out, 1)
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); engine.retrieve(shard, q32, q16, case.q_pid, case.excl); e.record(); torch.cuda.synchronize()
raw.reid_debug_counters(out, 1)
v = list(out); tiles = max(1, v[5]); wt = max(1, v[9])
print('bits %d step ms %.1f | per tile (cycles): mma_wait_tempty %.0f  mma_wait_full(sum over 8 chunks) %.0f | per warp-tile: scan %.0f wait_tfull %.0f drain %.0f refresh+flush %.0f | drains/warp-tile %.3f refreshes/warp-tile %.3f | max scan+drain of a warp-tile %d' % (
    bits, s.elapsed_time(e), v[2] / tiles, v[3] / tiles, v[0] / wt, v[1] / wt, v[4] / wt, v[7] / wt, v[6] / wt, v[8] / wt, v[10]))
