"""The counting rule of the fused kernel (two-level deep-rank sampling) restated in numpy
(oracle/fused_model.py) and measured against the exact oracle WITHOUT a GPU: the approximation must keep mAP inside the
north star's 1e-4 and leave CMC / the number of valid queries untouched, also when the gallery is split over ranks and
chunks.  (The GPU tests compare the kernel itself with the exact path at 100k and 1M rows.)"""
import numpy as np
import pytest
import torch

from oracle import fused_model as fm
from oracle import retrieval as orc
from prcv2025reid_b200 import sharding, synth


def _case(n_ids, nq):
    case = synth.make_retrieval_case(1005, n_ids, 40, 4, 4, excl_frac=0.02, n_excl=2, max_queries=nq)
    q = orc.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    g = orc.l2n(case.gallery_raw)
    exact = orc.rank_and_metrics_counting(q, g, case.q_pid, case.g_pid, case.excl, return_per_query=True)
    return case, q, g, exact


def _model(case, q, g, world, **kw):
    thr, n_pos = fm.positive_thresholds(q, g, case.q_pid, case.g_pid, case.excl)
    total = np.zeros_like(thr, dtype=np.int64)
    n_exact = []
    for r in range(world):
        r0, r1 = sharding.shard_range(g.shape[0], r, world)
        c = fm.fused_counts(q, g[r0:r1], case.q_pid, case.g_pid[r0:r1], case.excl, n_shards=world, thr=thr, n_pos=n_pos,
                            g_offset=r0, **kw)
        total += c["pos_above"]                                   # sharding.exchange_counts: counts are additive
        n_exact.append(c["n_exact"])
    return fm.metrics_from_counts(total, n_pos), n_pos, np.stack(n_exact)


@pytest.mark.parametrize("n_ids,world", [(1000, 1), (2000, 2)])
def test_deep_rank_sampling_keeps_map_within_the_bar(n_ids, world):
    case, q, g, exact = _case(n_ids, 128)
    m, n_pos, n_exact = _model(case, q, g, world)
    assert (n_exact < n_pos[None, :]).any()                       # the sampled path is exercised
    assert m["num_queries"] == exact["num_queries"]
    assert abs(m["mAP"] - exact["mAP"]) <= 1e-4                   # north star: mAP within 1e-4
    assert [m["R@1"], m["R@5"], m["R@10"]] == [exact["R@1"], exact["R@5"], exact["R@10"]]
    v = m["_ap"] >= 0
    assert np.abs(m["_ap"][v] - exact["_ap"][v]).max() <= 2e-2    # a single query: one head flip of fp16 near-ties at most


def test_without_sampling_only_fp16_rounding_remains():
    case, q, g, exact = _case(1000, 96)
    m, n_pos, n_exact = _model(case, q, g, 1, exact_all=True)      # REID_FUSED_EXACT_COUNTS
    assert np.array_equal(n_exact[0], n_pos)
    assert abs(m["mAP"] - exact["mAP"]) <= 5e-5
    # a shard below 16 x CALIB_ROWS rows is never sampled (reid_retrieve_fused `sample_deep`)
    small = fm.fused_counts(q, g[:20000], case.q_pid, case.g_pid[:20000], case.excl)
    assert np.array_equal(small["n_exact"], small["n_pos"])


def test_counts_are_exact_on_exact_scores():
    """With the score rounding taken out (fp16-representable features) and no sampling, the counting form IS the
    reference's ranking: AP equal to the argsort loop to float64 round-off."""
    case = synth.make_retrieval_case(7, 30, 6, 3, 4, excl_frac=0.2, n_excl=2)
    q = orc.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor()).half().float()
    g = orc.l2n(case.gallery_raw).half().float()
    loop = orc.rank_and_metrics_loop(q, g, case.q_pid, case.g_pid, case.excl, return_per_query=True)
    c = fm.fused_counts(q, g, case.q_pid, case.g_pid, case.excl)
    m = fm.metrics_from_counts(c["pos_above"], c["n_pos"])
    assert m["num_queries"] == loop["num_queries"]
    assert abs(m["mAP"] - loop["mAP"]) <= 1e-12
    assert [m["R@1"], m["R@5"], m["R@10"]] == [loop["R@1"], loop["R@5"], loop["R@10"]]
    v = loop["_valid"]
    assert np.abs(m["_ap"][v] - loop["_ap"][v]).max() <= 1e-12


def test_rescoring_stage_makes_the_head_exact():
    """rescore_topk restated (oracle/fused_model.rescore_stage): the RTOP best rows by fp16 score, re-scored in fp32, give
    the exact top-10 of every unflagged query, exact counts for the positives above the completeness bound, and only
    shrink the error of the modelled AP."""
    case, q, g, exact = _case(1000, 96)
    c = fm.fused_counts(q, g, case.q_pid, case.g_pid, case.excl)
    r = fm.rescore_stage(q, g, case.q_pid, case.g_pid, c, case.excl)
    ok = r["flag"] == 0
    assert ok.mean() >= 0.9                                            # flagged queries are re-run exactly by engine.retrieve
    assert np.array_equal(r["top_idx"][ok][:, :10], exact["_top_idx"][ok])
    before = fm.metrics_from_counts(c["pos_above"], c["n_pos"])
    after = fm.metrics_from_counts(r["pos_above"], c["n_pos"])
    v = exact["_valid"] & ok
    e0 = np.abs(before["_ap"][v] - exact["_ap"][v]).max()
    e1 = np.abs(after["_ap"][v] - exact["_ap"][v]).max()
    assert e1 <= e0 and abs(after["mAP"] - exact["mAP"]) <= 1e-4
    assert [after["R@1"], after["R@5"], after["R@10"]] == [exact["R@1"], exact["R@5"], exact["R@10"]]
    # every positive above the bound has its exact rank
    S = (q @ g.T).numpy()
    for qi in np.nonzero(v)[0][:16]:
        t0 = c["thr"][qi, 0]
        if r["top_score"][qi, fm.KLIST - 1] + 2 * fm.EPS_FP16 < t0:      # clearly above the completeness bound
            use = np.ones(g.shape[0], dtype=bool)
            e = case.excl[qi].numpy(); use[e[e >= 0]] = False
            use &= (case.g_pid.numpy() != int(case.q_pid[qi]))
            assert r["pos_above"][qi, 0] == int((use & (S[qi] > t0)).sum())
