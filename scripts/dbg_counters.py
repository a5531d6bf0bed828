import ctypes, os, sys, torch
sys.path.insert(0, '.')
os.environ['REID_FUSED_DEBUG'] = str(int(os.environ.get('REID_FUSED_DEBUG', '0')) | 8192)
import bench
from prcv2025reid_b200 import engine, synth, _cabi
seed, n_ids, gpi, k, qpi = bench.WORKLOADS['c4']
nq = 32768
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid); case.gallery_raw = None
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
raw = ctypes.CDLL(_cabi.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
engine.retrieve(shard, q32, q16, case.q_pid, case.excl)
torch.cuda.synchronize(); raw.reid_debug_counters(out, 1)
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); engine.retrieve(shard, q32, q16, case.q_pid, case.excl); e.record(); torch.cuda.synchronize()
raw.reid_debug_counters(out, 1)
v = list(out); tiles = max(1, v[5])
print('step ms %.1f  | per tile-step (cycles): mma_wait_tempty %.0f  mma_wait_full(sum over 8 chunks) %.0f | epi(ew0): busy %.0f wait_tfull %.0f | hit loop: %.0f cycles/tile, %.1f hit columns/tile, %.0f cycles/column' % (
    s.elapsed_time(e), v[2] / tiles, v[3] / tiles, v[0] / (2 * tiles), v[1] / (2 * tiles), v[6] / (2 * tiles), v[7] / (2 * tiles), v[6] / max(1, v[7])))
