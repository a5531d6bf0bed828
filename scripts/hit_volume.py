"""Offline estimate of the fused epilogue's hit volume on a workload (debug aid)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from prcv2025reid_b200 import engine, synth
w = sys.argv[1] if len(sys.argv) > 1 else 'c4'
seed, n_ids, gpi, k, qpi = bench.WORKLOADS[w]
nq = 512
case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device='cuda'))
G = shard.G_local
n_chunks = 4
rows_chunk = G // n_chunks
limit = max(rows_chunk / 128, 2048)
S = q32 @ shard.g_f32.T                                  # [nq, G]
pos = (case.g_pid[None, :] == case.q_pid[:nq, None])
f_ex, f_low, nex, npos = [], [], [], []
for qi in range(nq):
    t = torch.sort(S[qi][pos[qi]], descending=True)[0]
    sneg = S[qi][~pos[qi]]
    ranks = (sneg[None, :] > t[:, None]).sum(1).float()   # rows above each threshold (whole gallery)
    ne = int((ranks / n_chunks <= limit).sum())
    f_ex.append(float(ranks[ne - 1] / G) if ne > 0 else 0.0)
    f_low.append(float(ranks[-1] / G))
    nex.append(ne); npos.append(len(t))
f_ex = torch.tensor(f_ex); f_low = torch.tensor(f_low)
print('P %d  n_exact mean %.1f  f_exact mean %.4f  f_low mean %.3f median %.3f' % (npos[0], sum(nex) / nq, f_ex.mean(), f_low.mean(), f_low.median()))
per_tile = 128 * 256
print('hits per CTA tile-step: exact %.0f  sampled(1/64) %.0f  | top-32 level f=%.6f' % (per_tile * f_ex.mean(), per_tile * f_low.mean() / 64, 32 / G))
