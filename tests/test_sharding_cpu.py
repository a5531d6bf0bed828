"""World-size-2 gloo test of the gallery sharding + exchange protocol:
each rank runs the ORACLE on its contiguous gallery shard, the three exchange steps of
prcv2025reid_b200.sharding combine them, and the result must equal the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import retrieval as orc  # noqa: E402
from prcv2025reid_b200 import sharding, synth  # noqa: E402

R = 32


def test_shard_range_partitions_exactly():
    for G in (1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(G, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == G
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.make_retrieval_case(31, 40, 6, 3, 4, excl_frac=0.2, n_excl=2)
    w = synth.weights_tensor()
    q = orc.fuse_queries(case.query_raw, case.mod_id, w)
    g = orc.l2n(case.gallery_raw)
    G, Q = case.G, case.Q
    r0, r1 = sharding.shard_range(G, rank, world)
    S = (q @ g[r0:r1].T)
    for qi in range(Q):                                           # same-image mask on the local shard
        for e in case.excl[qi].tolist():
            if r0 <= e < r1:
                S[qi, e - r0] = -1e9
    pmax = 6
    # step 1: positives' exact scores, owner rank fills, all_reduce(MAX)
    pos = torch.full((Q, pmax), -float("inf"))
    for qi in range(Q):
        idx = torch.nonzero(case.g_pid == case.q_pid[qi]).flatten().tolist()
        for j, gi in enumerate(idx):
            if r0 <= gi < r1 and S[qi, gi - r0] > -1e8:
                pos[qi, j] = S[qi, gi - r0]
    sharding.exchange_pos_scores(pos)
    thr, _ = torch.sort(pos, dim=1, descending=True)
    # step 2: local counts of non-positive rows above each threshold, all_reduce(SUM)
    notpos = (case.g_pid[None, r0:r1] != case.q_pid[:, None]) & (S > -1e8)
    above = torch.zeros(Q, pmax, dtype=torch.int32)
    for j in range(pmax):
        above[:, j] = ((S > thr[:, j:j + 1]) & notpos).sum(1).to(torch.int32)
    sharding.exchange_counts(above)
    # step 3: per-shard top lists (global indices), all_gather
    k = min(R, r1 - r0)
    ts, ti = torch.topk(S, k, dim=1)
    top_s = torch.full((Q, R), -float("inf")); top_i = torch.full((Q, R), -1, dtype=torch.int32)
    top_s[:, :k] = ts; top_i[:, :k] = (ti + r0).to(torch.int32)
    all_s, all_i = sharding.gather_top_lists(top_s, top_i)
    if rank == 0:
        n_pos = torch.isfinite(thr).sum(1)
        aps, first = [], []
        for qi in range(Q):
            npq = int(n_pos[qi])
            if npq == 0:
                continue
            ranks = above[qi, :npq].numpy().astype(np.int64) + np.arange(npq) + 1
            aps.append(float(np.mean((np.arange(npq) + 1) / ranks))); first.append(int(ranks[0]))
        flat_s = all_s.permute(1, 0, 2).reshape(Q, -1); flat_i = all_i.permute(1, 0, 2).reshape(Q, -1)
        order = torch.argsort(flat_s, dim=1, descending=True, stable=True)[:, :10]
        torch.save({"mAP": float(np.mean(aps)), "R1": float(np.mean(np.array(first) <= 1)),
                    "R10": float(np.mean(np.array(first) <= 10)), "n": len(aps),
                    "top10": torch.gather(flat_i, 1, order)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_unsharded_oracle(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    case = synth.make_retrieval_case(31, 40, 6, 3, 4, excl_frac=0.2, n_excl=2)
    w = synth.weights_tensor()
    ref = orc.rank_and_metrics_loop(orc.fuse_queries(case.query_raw, case.mod_id, w), orc.l2n(case.gallery_raw),
                                    case.q_pid, case.g_pid, case.excl, return_per_query=True)
    assert res["n"] == ref["num_queries"]
    assert abs(res["mAP"] - ref["mAP"]) < 1e-9
    assert res["R1"] == ref["R@1"] and res["R10"] == ref["R@10"]
    assert np.array_equal(res["top10"].numpy(), ref["_top_idx"])


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n in (1, 5, 6, 17):
        full = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
        s0, s1, m = sharding.block_slice(n, rank, world)
        part = torch.full((m, 3), -7.0)
        part[:s1 - s0] = full[s0:s1]
        got = sharding.gather_query_block(part, n)
        ok = ok and torch.equal(got, full)
    if rank == 0:
        torch.save({"ok": ok}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_query_block_upload_is_sharded_and_gathered(tmp_path, world):
    """Every rank contributes its block_slice; the gathered block equals the original for ragged sizes."""
    for n in (1, 5, 6, 17, 32768):
        edges = [sharding.block_slice(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and max(e[1] for e in edges) == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        assert all(e[1] - e[0] <= e[2] for e in edges) and len({e[2] for e in edges}) == 1
    out = str(tmp_path / "g.pt")
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.load(out)["ok"]
