"""Host logic of prcv2025reid_b200.engine (gallery installation, query blocks, scratch reuse, the flag -> exact re-run
path, the shard exchange) exercised WITHOUT a GPU: the C library is replaced by tests/_fake_lib.FakeLib, a torch-CPU
stand-in that implements the contract of every entry point as include/reid_b200.h states it, and the results are
compared with the oracle.  The kernels themselves are compared with the same oracle in the -m gpu tests; what is
under test here is everything between the public call and the C ABI."""
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
from tests import _fake_lib  # noqa: E402


def _reference(case):
    q = orc.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    g = orc.l2n(case.gallery_raw)
    return orc.rank_and_metrics_loop(q, g, case.q_pid, case.g_pid, case.excl, return_per_query=True)


def _check(res, ref, Q, fused=False):
    # the fused pass counts deep ranks on fp16-operand scores like the tensor cores do: mAP inside the north star's 1e-4,
    # CMC and the top-10 exact; the all-fp32 pass reproduces the reference's ranking to float64 round-off
    assert res.metrics["num_queries"] == ref["num_queries"]
    assert abs(res.metrics["mAP"] - ref["mAP"]) <= (1e-4 if fused else 1e-9)
    assert [res.metrics[k] for k in ("R@1", "R@5", "R@10")] == [ref[k] for k in ("R@1", "R@5", "R@10")]
    assert np.array_equal(res.top_idx.numpy().astype(np.int64), ref["_top_idx"])
    v = ref["_valid"]
    assert np.abs(res.ap.numpy()[v] - ref["_ap"][v]).max() <= (2e-2 if fused else 1e-12) and (res.ap.numpy()[~v] == -1).all()


@pytest.mark.parametrize("mode", ["fused", "exact"])
@pytest.mark.parametrize("query_block", [32768, 50])
def test_retrieve_host_logic_matches_oracle(monkeypatch, mode, query_block):
    from prcv2025reid_b200 import engine
    fake = _fake_lib.install(monkeypatch, force_flag_every=7)
    case = synth.make_retrieval_case(53, 60, 8, 3, 4, excl_frac=0.2, n_excl=2)          # 480 rows, 240 queries
    q_pid = case.q_pid.clone(); q_pid[::17] = 10_000                                    # queries without a positive
    case.q_pid = q_pid
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
    assert shard.pmax == 8 and shard.G_total == case.G and shard.g_offset == 0
    q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, topk=10, mode=mode, query_block=query_block, want_ap=True)
    _check(res, _reference(case), case.Q, fused=mode == "fused")
    n_blocks = -(-case.Q // query_block)
    if mode == "fused":
        # flagged queries (the forced ones among them) went through ONE all-fp32 pass over a compact copy at the end,
        # then the exchange + metrics were redone
        forced = sum(1 for b0 in range(0, case.Q, query_block) for q in range(min(query_block, case.Q - b0)) if q % 7 == 3)
        assert res.n_flagged >= forced > 0 and res.path == "fused"
        assert fake.calls.count("reid_retrieve_fused") == n_blocks and fake.calls.count("reid_retrieve_exact") == 1
        assert fake.calls.count("reid_metrics_reduce") == 2 and fake.calls.count("reid_rescore_topk") == n_blocks + 1
        assert fake.calls.count("reid_cand_select") == n_blocks + 1 and fake.calls.count("reid_topk_check") == 2
    else:
        assert res.n_flagged == 0 and res.path == "exact"
        assert fake.calls.count("reid_retrieve_exact") == n_blocks and fake.calls.count("reid_metrics_reduce") == 1
    # a second call on the same shard reuses its scratch buffers and gives the same answer
    res2 = engine.retrieve(shard, q32, q16, case.q_pid, None, topk=10, mode=mode, query_block=query_block, want_ap=True)
    case.excl = None
    _check(res2, _reference(case), case.Q, fused=mode == "fused")


def test_prepare_gallery_codes_and_lookup(monkeypatch):
    from prcv2025reid_b200 import engine
    _fake_lib.install(monkeypatch)
    g_pid = torch.tensor([7, 3, 7, 9, 3, 7], dtype=torch.int64)
    shard = engine.prepare_gallery(torch.randn(3, 16), g_pid, g_offset=2)     # this rank holds rows 2..4 of six
    assert shard.G_local == 3 and shard.G_total == 6 and shard.pmax == 3
    assert shard.sorted_pid.tolist() == [3, 3, 7, 7, 7, 9] and shard.order.tolist() == [1, 4, 0, 2, 5, 3]
    assert shard.g_code.tolist() == [2, 5, 0]                                 # first sorted position of pids 7, 9, 3
    assert torch.allclose(shard.g_f32.norm(dim=1), torch.ones(3), atol=1e-6) and shard.g_f16.dtype == torch.float16


def test_fused_slots_fill_the_last_wave():
    from prcv2025reid_b200 import engine
    # whole waves of 74 CTA pairs need no split; the blocks of a partial wave are cut so that it is full too
    assert engine.fused_slots(74 * 256, 1_000_000, 148) == 1 and engine.fused_slots(2 * 74 * 256, 125_000, 148) == 1
    assert engine.fused_slots(24224, 1_000_000, 148) == 3          # 95 blocks = 74 + 21 left -> 3 chunks each
    assert engine.fused_slots(20_000, 100_000, 148) == 8           # C3: 79 blocks = 74 + 5 left (capped at 8)
    assert engine.fused_slots(1, 192, 148) == 1                    # tiny gallery: never split below 4096 rows per chunk
    assert engine.fused_slots(2048, 100_000, 148) == 8
    assert engine.default_query_block(148) == 37888


def test_identities_with_more_than_64_gallery_rows_use_threshold_windows(monkeypatch):
    """Pmax > 64: the fused pass handles 64 thresholds per query and call; deeper positives of the large identities are
    counted by further counting-only passes over those queries (engine._rank_block), results equal to the oracle."""
    from prcv2025reid_b200 import engine
    fake = _fake_lib.install(monkeypatch)
    case = synth.make_ragged_case(71, 12, 1, 150, 2, 2, excl_frac=0.3)
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
    assert shard.pmax == 150
    q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, topk=10, mode="fused", want_ap=True)
    assert res.path == "fused" and fake.calls.count("reid_retrieve_fused") == 1
    assert 1 <= fake.calls.count("reid_retrieve_fused(window)") <= 2
    _check(res, _reference(case), case.Q, fused=True)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _pytest.monkeypatch import MonkeyPatch
    from prcv2025reid_b200 import engine
    mpatch = MonkeyPatch()
    try:
        _fake_lib.install(mpatch)
        case = synth.make_retrieval_case(59, 50, 6, 4, 2, excl_frac=0.2, n_excl=2)
        r0, r1 = sharding.shard_range(case.G, rank, world)
        shard = engine.prepare_gallery(case.gallery_raw[r0:r1], case.g_pid, g_offset=r0)
        q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
        res = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, topk=10, mode="fused", query_block=40,
                              group=dist.group.WORLD, want_ap=True)
        torch.save({"metrics": res.metrics, "top_idx": res.top_idx, "ap": res.ap, "flagged": res.n_flagged}, out + ".%d" % rank)
        dist.barrier()
    finally:
        mpatch.undo()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_retrieve_host_logic_matches_oracle(tmp_path, world):
    """engine.retrieve(group=...) on world gloo ranks, each with its contiguous gallery shard: every rank ends with the
    metrics and the merged top-10 of the unsharded oracle."""
    out = str(tmp_path / "res.pt")
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    case = synth.make_retrieval_case(59, 50, 6, 4, 2, excl_frac=0.2, n_excl=2)
    ref = _reference(case)
    for rank in range(world):
        r = torch.load(out + ".%d" % rank, weights_only=False)
        assert r["metrics"]["num_queries"] == ref["num_queries"]
        assert abs(r["metrics"]["mAP"] - ref["mAP"]) <= 1e-4
        assert [r["metrics"][k] for k in ("R@1", "R@5", "R@10")] == [ref[k] for k in ("R@1", "R@5", "R@10")]
        assert np.array_equal(r["top_idx"].numpy().astype(np.int64), ref["_top_idx"])


def test_topk_ranking_passes_concatenate_to_the_full_argsort(monkeypatch):
    """topk.topk_ranking (ranking core of export_submission_csv, eval_mm_protocol.py:617-625): ranks beyond the exact
    top-32 of a pass come from further passes that exclude the rows already ranked; the concatenation is the argsort."""
    from prcv2025reid_b200 import topk
    fake = _fake_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(5)
    gal = torch.randn(150, 64, generator=g)
    q = torch.nn.functional.normalize(torch.randn(9, 64, generator=g), dim=1)
    want = torch.argsort(q @ torch.nn.functional.normalize(gal, dim=1).T, dim=1, descending=True)
    got = topk.topk_ranking(q, gal, top_k=100, mode="fused")
    assert got.shape == (9, 100) and torch.equal(got.long(), want[:, :100])
    assert fake.calls.count("reid_retrieve_fused") == 4                       # 32 + 32 + 32 + 4
    short = topk.topk_ranking(q, gal[:20], top_k=50, mode="exact")            # fewer rows than top_k: -1 padded
    want20 = torch.argsort(q @ torch.nn.functional.normalize(gal[:20], dim=1).T, dim=1, descending=True)
    assert torch.equal(short[:, :20].long(), want20) and bool((short[:, 20:] == -1).all())


def test_export_submission_csv_host_logic(monkeypatch, tmp_path):
    import csv
    from prcv2025reid_b200 import eval_mm_protocol as emp
    _fake_lib.install(monkeypatch)
    monkeypatch.setattr(emp, "_dev", lambda: torch.device("cpu"))
    case = synth.make_retrieval_case(61, 12, 4, 2, 3, excl_frac=0.0)
    queries, gmeta, ext = synth.case_to_reference_inputs(case)
    path = str(tmp_path / "sub.csv")
    emp.export_submission_csv(queries, orc.l2n(case.gallery_raw), gmeta, ext, dict(synth.DEFAULT_WEIGHTS), path, top_k=20)
    qf = orc.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    want = orc.submission_ranking(qf, orc.l2n(case.gallery_raw), top_k=20)
    rows = list(csv.DictReader(open(path, newline="", encoding="utf-8")))
    assert len(rows) == case.Q and list(rows[0]) == ["query_key", "ranked_gallery_ids"]        # :634-643
    for qi, row in enumerate(rows):
        assert [int(t[1:]) for t in row["ranked_gallery_ids"].split()] == want[qi].tolist()
        pid, mods, _ = row["query_key"].split("|")
        assert int(pid) == int(case.q_pid[qi]) and mods == "+".join(sorted(queries[qi]["modalities"]))


def test_train_eval_dropins_host_logic(monkeypatch):
    """compute_map / compute_cmc / reid_map (train.py:101-138, :451-479) through the engine with the stand-in library,
    against the oracle ports that are pinned to the unmodified train.py functions."""
    from prcv2025reid_b200 import train_eval as te
    _fake_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(3)
    centres = torch.randn(20, 32, generator=g)
    gl = torch.arange(120) // 6
    ql = torch.randint(0, 20, (40,), generator=g); ql[:3] = 999
    gf = centres[gl] + 2.0 * torch.randn(120, 32, generator=g)
    qf = centres[ql.clamp(max=19)] + 2.0 * torch.randn(40, 32, generator=g)
    for k in (1, 5, 100):
        assert te.compute_map(qf, gf, ql, gl, k=k) == pytest.approx(orc.compute_map_oracle(qf, gf, ql, gl, k=k), abs=1e-6)
    for k in (1, 10):
        assert te.compute_cmc(qf, gf, ql, gl, k=k) == pytest.approx(orc.compute_cmc_oracle(qf, gf, ql, gl, k=k), abs=1e-12)
    qn, gn = torch.nn.functional.normalize(qf, dim=1), torch.nn.functional.normalize(gf, dim=1)
    m, t1 = te.reid_map(qn, gn, ql, gl)
    wm, wt1 = orc.reid_map_oracle(qn @ gn.T, ql, gl)
    assert m == pytest.approx(wm, abs=1e-4) and t1 == pytest.approx(wt1, abs=1e-12)
    assert te.compute_map(qf[:0], gf, ql[:0], gl) == 0.0 and te.reid_map(qn[:0], gn, ql[:0], gl) == (0.0, 0.0)


def test_host_query_blocks_partition():
    """The ramped block schedule of host-resident queries: a partition of [0, Q) in order, a small first block, one wave,
    then two-wave blocks, no short tail block."""
    from prcv2025reid_b200 import engine
    for sms in (148, 132, 2):
        wave = max(1, sms // 2) * engine.WAVE_QUERIES
        for Q in (1, 255, 4096, 4097, 20000, 37888, 100000, 250001):
            bl = engine.host_query_blocks(Q, sms)
            assert bl[0][0] == 0 and bl[-1][1] == Q and all(a[1] == b[0] for a, b in zip(bl, bl[1:]))
            sizes = [b - a for a, b in bl]
            assert all(n > 0 for n in sizes) and sizes[0] == min(Q, 16 * engine.WAVE_QUERIES)
            if len(sizes) > 1:
                assert sizes[1] == min(Q - sizes[0], wave)
            assert all(n <= 2 * wave + 2 * wave // 4 for n in sizes[2:])
            if len(sizes) > 3:
                assert sizes[-1] >= 2 * wave // 4
    assert engine.host_query_blocks(100000, 148) == [(0, 4096), (4096, 23040), (23040, 60928), (60928, 100000)]


def test_resident_query_blocks_partition():
    """Blocks of device-resident queries: whole waves, larger launches on smaller shards, no short tail block."""
    from prcv2025reid_b200 import engine
    wave = 74 * engine.WAVE_QUERIES
    for G in (10_000, 125_000, 250_000, 500_000, 1_000_000, 4_000_000):
        for Q in (1, 3000, 37888, 40000, 100000, 1_000_000):
            bl = engine.resident_query_blocks(Q, 148, G)
            assert bl[0][0] == 0 and bl[-1][1] == Q and all(a[1] == b[0] for a, b in zip(bl, bl[1:]))
            assert all((b - a) % wave == 0 for a, b in bl[:-1])
            if len(bl) > 1:
                assert bl[-1][1] - bl[-1][0] >= wave // 2
    assert engine.resident_query_blocks(100000, 148, 1_000_000) == [(0, 37888), (37888, 75776), (75776, 100000)]
    assert engine.resident_query_blocks(100000, 148, 125_000) == [(0, 100000)]
