"""One rank of an N-way sharded C4 run, emulated on ONE GPU: the positives' scores come from the whole gallery (what the
all_reduce(MAX) delivers), the gallery pass (engine._scan_stage: calibration + fused kernel + cand_select) runs on the first
G/N rows with the gallery-wide budgets of N shards.  Prints the pass time per query block and the candidate volume.
usage: [REID_LIB=variant.so] python scripts/shard_probe.py [world=8] [workload=c4]"""
import sys, torch
sys.path.insert(0, '.')
import bench
from prcv2025reid_b200 import engine, synth, sharding, _cabi
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = sys.argv[2] if len(sys.argv) > 2 else 'c4'
seed, n_ids, gpi, k, qpi = bench.WORKLOADS[w]
nq = min(engine.default_query_block(148), n_ids * qpi)
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
full = engine.prepare_gallery(case.gallery_raw, case.g_pid)
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
S = engine._RankState(nq, full.pmax, q32.device)
pid = case.q_pid.to(torch.int64).contiguous()
ex = case.excl.to(torch.int32).contiguous() if case.excl is not None else None
E = 0 if ex is None else ex.shape[1]
engine._pos_stage(full, S, slice(0, nq), q32, pid, ex, E, group=None, world=1)
r0, r1 = sharding.shard_range(full.G_total, 0, world)
n = r1 - r0
sub = engine.GalleryShard(full.g_f32[:n], full.g_f16[:n], full.g_code[:n], full.sorted_pid, full.order, full.pmax, 0, full.G_total)
engine._DEBUG_KEEP = {}
def run():
    S.pos_above.zero_()
    engine._scan_stage(sub, S, slice(0, nq), q32, q16, ex, E, fused=True, cand_cap=2048, world=world, exact_ap=False)
for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    run()
b.record(); torch.cuda.synchronize()
cc = engine._DEBUG_KEEP["cand_count"].float()
flop = 2.0 * nq * n * 512
ms = a.elapsed_time(b) / 5
print("lib %s  world %d  shard rows %d  queries %d  slots %d" % (_cabi.LIB_PATH.split('/')[-1], world, n, nq, engine._DEBUG_KEEP["n_chunks"]))
print("gallery pass %.3f ms  (%.0f TFLOP/s)   candidates per query: mean %.1f  max %d   selected mean %.1f" % (
    ms, flop / ms / 1e9, cc.sum(1).mean().item(), int(cc.sum(1).max().item()), engine._DEBUG_KEEP["sel_n"].float().mean().item()))
