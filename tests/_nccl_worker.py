"""Worker of tests/test_zz_nccl_gpu.py: one process per GPU (torchrun), REAL NCCL.  Every rank holds a contiguous
shard of the gallery; engine.retrieve(group=WORLD) runs the fused (sampled) branch on the shard and the four exchange
steps of prcv2025reid_b200/sharding.py; rank 0 compares the result with the oracle's full ranking of the UNSHARDED
gallery (tools/eval_mm_protocol.py:396-455) and with the host-query (sharded upload + all-gather) entry."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, workload, nq = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import bench
    from oracle import retrieval as orc
    from prcv2025reid_b200 import engine, sharding, synth
    seed, n_ids, gpi, k, qpi = bench.WORKLOADS[workload]
    G = n_ids * gpi
    r0, r1 = sharding.shard_range(G, rank, world)
    case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=0.01, n_excl=2, device="cuda", max_queries=nq,
                                     gallery_rows=(r0, r1))
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid, g_offset=r0)
    w = synth.weights_tensor(device="cuda")
    q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, w)
    res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode="fused", group=dist.group.WORLD, want_ap=True)
    host = engine.retrieve(shard, None, None, case.q_pid.cpu().pin_memory(), case.excl.cpu().pin_memory(), mode="fused",
                           group=dist.group.WORLD, want_ap=True, query_block=max(256, nq // 3),
                           host_queries=(case.query_raw.cpu().pin_memory(), case.mod_id.cpu().pin_memory(), w))
    # every rank must hold the same merged result
    t = torch.tensor([res.metrics["mAP"], res.metrics["R@1"], res.metrics["R@5"], res.metrics["R@10"]], device="cuda",
                     dtype=torch.float64)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ti_lo, ti_hi = res.top_idx.clone(), res.top_idx.clone()
    dist.all_reduce(ti_lo, op=dist.ReduceOp.MIN); dist.all_reduce(ti_hi, op=dist.ReduceOp.MAX)
    report = {"world": world, "workload": workload, "queries": nq, "gallery": G, "path": res.path,
              "ranks_agree": bool(torch.equal(lo, hi) and torch.equal(ti_lo, ti_hi)),
              "ranks_metric_spread": (hi - lo).abs().max().item(), "ranks_top_idx_agree": bool(torch.equal(ti_lo, ti_hi)),
              "host_equals_resident": bool(all(abs(res.metrics[m] - host.metrics[m]) < 1e-9 for m in res.metrics)
                                           and torch.equal(res.top_idx, host.top_idx)),
              "flagged": res.n_flagged, "metrics": res.metrics}
    if rank == 0:
        full = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=0.01, n_excl=2, device="cuda", max_queries=nq)
        cw = synth.weights_tensor()
        q = orc.fuse_queries(full.query_raw.cpu(), full.mod_id.cpu(), cw)
        g = orc.l2n(full.gallery_raw.cpu())
        o = orc.rank_and_metrics_counting(q, g, full.q_pid.cpu(), full.g_pid.cpu(), full.excl.cpu(), return_per_query=True)
        v = o["_valid"]
        ap = res.ap.cpu().numpy()
        first = res.pos_above[:, 0].cpu().numpy() + 1
        ti = res.top_idx.cpu().numpy().astype(np.int64)
        n_first = int((np.minimum(first[v], 11) != np.minimum(o["_first"][v], 11)).sum())
        bad_lists = 0
        for qi in np.nonzero((ti[:, :10] != o["_top_idx"][:, :10]).any(axis=1))[0]:
            s = (q[qi:qi + 1] @ g.T).squeeze(0).numpy()
            if any(ti[qi, r] != o["_top_idx"][qi, r] and abs(float(s[ti[qi, r]]) - float(s[o["_top_idx"][qi, r]])) > 2e-6
                   for r in range(10)):
                bad_lists += 1
        report.update(oracle={m: o[m] for m in ("mAP", "R@1", "R@5", "R@10", "num_queries")},
                      d_map=res.metrics["mAP"] - o["mAP"], d_ap_max=float(np.abs(ap[v] - o["_ap"][v]).max()),
                      d_ap_mean=float(np.abs(ap[v] - o["_ap"][v]).mean()), cmc_rank_mismatches=n_first,
                      top10_lists_differing_beyond_ties=bad_lists)
        with open(out_path, "w") as f:
            json.dump(report, f)
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
