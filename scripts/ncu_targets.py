"""One invocation of every kernel of the library at benchmark sizes (the target of the `ncu --set full` captures under
profiles/): C4 gallery shard, one query block through the fused path, the unfused similarity GEMM, SDM C5 and C2 steps.
    ncu --set full --clock-control none --import-source on -k regex:<kernels> -o gpurun_out/rep python scripts/ncu_targets.py"""
import sys
import torch
sys.path.insert(0, '.')
import bench
from prcv2025reid_b200 import engine, synth
from prcv2025reid_b200.sdm_loss import SdmStep, sdm_loss_pairs_labels

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "retrieval"):
    seed, n_ids, gpi, k, qpi = bench.WORKLOADS['c4']
    nq = 37888
    case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, device='cuda', max_queries=nq)
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)          # l2norm_rows (1M rows), pid index
    case.gallery_raw = None
    w = synth.weights_tensor(device='cuda')
    for _ in range(2):
        q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, w)     # mm_fuse_normalize
        res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl)      # pos_scores .. retrieve_fused .. rescore .. metrics
    S = engine.cosine_sim_f16(q16[:8192], shard.g_f16[:100000])            # sim_gemm (unfused K3)
    torch.cuda.synchronize()
    print("retrieval ok", res.metrics)
if which in ("all", "sdm"):
    for P, K, dtype, pairs in ((64, 8, torch.bfloat16, 10), (4, 2, torch.float32, 4)):
        feats, labels = synth.make_sdm_batch(2002 if P == 64 else 2001, P, K, n_modalities=5, dtype=dtype, device="cuda")
        y = (labels[:, None] == labels[None, :]).float()
        pl = [(a, b) for a in range(5) for b in range(a)][:pairs] if pairs > 4 else [(m, 0) for m in range(1, 5)]
        st = SdmStep([feats[a] for a, b in pl], [feats[b] for a, b in pl], [y] * len(pl), tau=0.2)
        for _ in range(2):
            st.run()
        torch.cuda.synchronize()
        print("sdm", P, K, st.losses[:3].tolist())
    # label form (no y): forward + backward through autograd
    feats, labels = synth.make_sdm_batch(2002, 64, 8, n_modalities=5, dtype=torch.bfloat16, device="cuda")
    qs = [f.clone().requires_grad_(True) for f in feats[1:]]
    v = feats[0].clone().requires_grad_(True)
    for _ in range(2):
        losses, status = sdm_loss_pairs_labels(qs, [v] * 4, [labels] * 4, [labels] * 4, tau=0.2)
        losses.sum().backward()
    torch.cuda.synchronize()
    print("sdm label form", losses.tolist())
