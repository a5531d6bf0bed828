"""Host side of the evaluation protocol (row N4): `build_queries` / `build_gallery` and the MM-1..4 loop of
`run_eval`, against fixtures generated from the UNMODIFIED reference (oracle/make_golden_protocol.py ->
tests/golden/protocol.json) and against the live reference when it is mounted.

The list-of-dicts plumbing of the drop-in module (`_fuse_batch`, `_exclusions`, `rank_and_metrics`,
`run_eval_features`) is exercised here WITHOUT a GPU by swapping the device engine for an oracle-backed stand-in:
what is under test is the host logic (query order, modality ids, weights, same-image exclusion lists, skipped
MM-k, the average), not the kernels -- those are compared with the same fixture in the -m gpu tests."""
import json
import os
import random
import types

import pytest
import torch

from oracle import ref_loader
from oracle import retrieval as orc
from oracle.make_golden_protocol import query_digest
from prcv2025reid_b200 import eval_mm_protocol as emp
from prcv2025reid_b200 import synth

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "protocol.json"), encoding="utf-8"))


@pytest.fixture(scope="module")
def world():
    index, g_feats, g_meta, ext = synth.make_protocol_index(**GOLDEN["index_args"])
    cs = float(g_feats.double().abs().sum()) + float(sum(float(v.double().abs().sum()) for v in ext.table.values()))
    if abs(cs - GOLDEN["checksum"]) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    return index, g_feats, g_meta, ext


@pytest.mark.parametrize("mode", ["lexi_first", "random"])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_build_queries_matches_reference_golden(world, mode, k):
    index = world[0]
    qs = emp.build_queries(index, mode_k=k, rng=random.Random(1000 + k), main_mod_choice=mode)
    want = GOLDEN["build_queries"]["%s/k%d" % (mode, k)]
    sha, rows = query_digest(qs)
    assert len(qs) == want["n"]
    assert rows[:3] == want["first"]
    assert sha == want["sha256"]
    for q in qs:                                   # shape of a query (eval_mm_protocol.py:270-274)
        assert isinstance(q["modalities"], tuple) and list(q["modalities"]) == sorted(q["modalities"])
        assert set(q["samples"]) == set(q["modalities"]) and len(q["modalities"]) == k
        if mode == "lexi_first":
            assert next(iter(q["samples"])) == q["modalities"][0]       # main modality first


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_build_queries_matches_live_reference():
    from collections import defaultdict
    ref = ref_loader.load_reference_eval()
    index, _, _, _ = synth.make_protocol_index(seed=77, n_ids=25, max_per_mod=4, drop_frac=0.4)
    dd = defaultdict(lambda: defaultdict(list))     # the reference's own container type (build_index :71)
    for pid, by in index.items():
        for m, v in by.items():
            dd[pid][m].extend(v)
    for container in (index, dd):
        for mode in ("lexi_first", "random"):
            for k in (1, 2, 3, 4, 5):
                a = emp.build_queries(container, k, random.Random(3), mode)
                b = ref.build_queries(container, k, random.Random(3), mode)
                assert a == b and [list(x["samples"]) for x in a] == [list(x["samples"]) for x in b]
    assert {m for by in dd.values() for m in by} <= {"rgb", "ir", "cpencil", "sketch", "text"}   # no key was created
    assert emp.build_gallery(index) == ref.build_gallery(index)
    assert emp.combos(list("abcd"), 3) == ref.combos(list("abcd"), 3)
    assert emp.pick_one([1, 2, 3], random.Random(9)) == ref.pick_one([1, 2, 3], random.Random(9))


def test_build_queries_edge_cases():
    assert emp.build_queries({}, 2, random.Random(0)) == []
    idx = {5: {"rgb": [{"img_id": "a"}], "ir": [], "text": [{"text": "t", "img_id": "x"}]}}
    assert emp.build_queries(idx, 2, random.Random(0)) == []                        # only one populated modality
    one = emp.build_queries(idx, 1, random.Random(0))
    assert [q["modalities"] for q in one] == [("text",)] and one[0]["pid"] == 5
    assert emp.build_gallery({1: {"ir": [{}]}, 2: {"rgb": [{"img_id": "r"}]}}) == [{"img_id": "r"}]


class _OracleEngine:
    """Stand-in for prcv2025reid_b200.engine with the oracle's CPU arithmetic (TEST ONLY)."""
    GalleryShard = object

    def __init__(self):
        self.installs = 0

    def l2norm_rows(self, x, want_f16=False, eps=1e-12):
        y = orc.l2n(x.float())
        return y, (y.half() if want_f16 else None)

    def fuse_queries(self, rows, mid, w):
        Q, k, D = rows.shape
        f = orc.l2n(rows.float())
        acc = torch.zeros(Q, D)
        for j in range(k):                                    # slots with mod_id < 0 are skipped (reid_b200.h K2)
            use = (mid[:, j] >= 0)
            wj = torch.where(use, w[mid[:, j].clamp(min=0).long()], torch.zeros(Q))
            acc = acc + f[:, j] * wj[:, None] if j else f[:, j] * wj[:, None]
        y = orc.l2n(acc)
        return y, y.half()

    def prepare_gallery(self, gallery, g_pid_all, g_offset=0):
        self.installs += 1
        return types.SimpleNamespace(g_f32=orc.l2n(gallery.float()), g_pid=g_pid_all, G_total=int(g_pid_all.numel()))

    def retrieve(self, shard, q32, q16, q_pid, excl, topk=10, mode="fused", exact_ap=False):
        self.exact_ap_calls = getattr(self, "exact_ap_calls", 0) + int(bool(exact_ap))
        return types.SimpleNamespace(metrics=orc.rank_and_metrics_loop(q32, shard.g_f32, q_pid, shard.g_pid, excl))


@pytest.fixture()
def cpu_engine(monkeypatch):
    eng = _OracleEngine()
    monkeypatch.setattr(emp, "engine", eng)
    monkeypatch.setattr(emp, "_dev", lambda: torch.device("cpu"))
    return eng


@pytest.mark.parametrize("mask", [True, False])
def test_protocol_loop_host_logic_matches_reference_golden(world, cpu_engine, mask):
    index, g_feats, g_meta, ext = world
    want = GOLDEN["run_eval/ignore_same_img=%s" % mask]["results"]
    got = emp.run_eval_features(index, g_feats, g_meta, ext, seed=GOLDEN["run_seed"], ignore_same_img=mask)
    assert cpu_engine.installs == 1                            # the gallery is installed once for the four passes
    assert list(got) == ["MM-1", "MM-2", "MM-3", "MM-4", "AVG(1-4)"]
    for name, w in want.items():
        assert got[name].get("num_queries") == w.get("num_queries"), name
        for key in ("mAP", "R@1", "R@5", "R@10"):
            assert got[name][key] == pytest.approx(w[key], abs=1e-12), (name, key)


def test_protocol_loop_skips_empty_mm_k(world, cpu_engine):
    index, g_feats, g_meta, ext = world
    sparse = {pid: {m: v for m, v in by.items() if m != "text"} for pid, by in index.items()}
    want = GOLDEN["run_eval/no_text"]["results"]
    got = emp.run_eval_features(sparse, g_feats, g_meta, ext, seed=GOLDEN["run_seed"])
    assert got["MM-4"] == {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0, "num_queries": 0} == want["MM-4"]
    for key in ("mAP", "R@1", "R@5", "R@10"):                  # the average leaves MM-4 out (:576)
        assert got["AVG(1-4)"][key] == pytest.approx(want["AVG(1-4)"][key], abs=1e-12)
    empty = emp.run_eval_features({}, g_feats, g_meta, ext)
    assert all(v == 0.0 for v in empty["AVG(1-4)"].values()) and empty["MM-1"]["num_queries"] == 0


@pytest.mark.parametrize("mask", [True, False])
def test_shuffled_gallery_leaves_the_protocol_results_unchanged(world, cpu_engine, mask):
    """`shuffled_gallery` (rows + meta in a seeded random order: the way around the fixed sampling phase of the fused path
    for periodically laid-out caches) changes nothing the protocol reports: the same-image rule follows the image ids."""
    index, g_feats, g_meta, ext = world
    sf, sm = emp.shuffled_gallery(g_feats, g_meta, seed=3)
    assert sf.shape == g_feats.shape and sorted(m["img_id"] for m in sm) == sorted(m["img_id"] for m in g_meta)
    assert [m["img_id"] for m in sm] != [m["img_id"] for m in g_meta]
    j = [m["img_id"] for m in g_meta].index(sm[0]["img_id"])
    assert torch.equal(sf[0], g_feats[j]) and sm[0] is g_meta[j]
    a = emp.run_eval_features(index, g_feats, g_meta, ext, seed=GOLDEN["run_seed"], ignore_same_img=mask)
    b = emp.run_eval_features(index, sf, sm, ext, seed=GOLDEN["run_seed"], ignore_same_img=mask)
    for name in a:
        assert a[name].get("num_queries") == b[name].get("num_queries")
        for key in ("mAP", "R@1", "R@5", "R@10"):
            assert a[name][key] == pytest.approx(b[name][key], abs=1e-12), (name, key)
    with pytest.raises(ValueError):
        emp.shuffled_gallery(g_feats, g_meta[:-1])
    # the other way around the fixed sampling phase: exact_ap travels from the protocol loop to engine.retrieve
    n0 = getattr(cpu_engine, "exact_ap_calls", 0)
    emp.run_eval_features(index, g_feats, g_meta, ext, seed=GOLDEN["run_seed"], ignore_same_img=mask, exact_ap=True)
    assert cpu_engine.exact_ap_calls == n0 + 4 and n0 == 0


def test_rank_and_metrics_rejects_a_shard_of_another_gallery(world, cpu_engine):
    index, g_feats, g_meta, ext = world
    shard = emp.install_gallery(g_feats[:8], g_meta[:8])
    qs = emp.build_queries(index, 2, random.Random(0))
    with pytest.raises(ValueError):
        emp.rank_and_metrics(qs, g_feats, g_meta, ext, dict(synth.DEFAULT_WEIGHTS), shard=shard)


def test_unknown_modality_raises_like_the_reference(world, cpu_engine):
    _, g_feats, g_meta, ext = world
    q = {"pid": 1000, "modalities": ("thermal",), "samples": {"thermal": {"img_path": "x", "img_id": "y"}}}
    with pytest.raises(ValueError):                            # eval_mm_protocol.py:351
        emp.extract_query_feat(q, ext, {})


def test_extract_gallery_feats_cache_hit_and_miss(world, cpu_engine, tmp_path):
    """eval_mm_protocol.py:291-325: a complete cache is returned untouched; a miss encodes, normalises, writes both
    files in the reference's format -- which the unmodified reference then reads back as a cache hit."""
    import numpy as np
    index, g_feats, g_meta, ext = world
    gallery = emp.build_gallery(index)
    calls = []
    orig = ext.encode_rgb
    ext.encode_rgb = lambda key: (calls.append(key), orig(key))[1]
    try:
        feats, meta = emp.extract_gallery_feats(gallery, ext, str(tmp_path / "c"))
        assert len(calls) == len(gallery) and meta == g_meta
        assert torch.equal(feats, orc.l2n(g_feats))                     # (stand-in engine: the oracle's own arithmetic)
        assert sorted(os.listdir(str(tmp_path / "c"))) == ["rgb_feats.npy", "rgb_meta.json"]
        feats2, meta2 = emp.extract_gallery_feats(gallery, ext, str(tmp_path / "c"))    # hit: no encoder call
        assert len(calls) == len(gallery) and torch.equal(feats2, feats) and meta2 == meta
        assert feats2.dtype == torch.float32 and np.load(str(tmp_path / "c" / "rgb_feats.npy")).dtype == np.float32
        if ref_loader.reference_available():
            ref = ref_loader.load_reference_eval()
            rf, rm = ref_loader.quiet(ref.extract_gallery_feats, gallery, ext, str(tmp_path / "c"))   # reads OUR cache
            assert len(calls) == len(gallery) and torch.equal(rf, feats) and rm == meta
            rf2, rm2 = ref_loader.quiet(ref.extract_gallery_feats, gallery, ext, str(tmp_path / "r"))  # writes ITS cache
            f3, m3 = emp.extract_gallery_feats(gallery, ext, str(tmp_path / "r"))                     # we read it
            assert torch.equal(f3, rf2) and m3 == rm2 == meta
            assert torch.equal(rf2, feats)                              # per-image l2n == batched l2n, bit for bit
    finally:
        ext.encode_rgb = orig


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_build_queries_property_vs_live_reference():
    """Random identity indices (missing / empty modalities, any k, both main-modality policies, any seed): the query
    list, the order of the `samples` dicts and the state of the rng afterwards equal the unmodified reference's."""
    from hypothesis import given, settings, strategies as st
    ref = ref_loader.load_reference_eval()
    mods = ["rgb", "ir", "cpencil", "sketch", "text", "thermal"]        # one modality the protocol does not know

    @st.composite
    def indices(draw):
        index = {}
        for pid in draw(st.lists(st.integers(0, 50), unique=True, max_size=6)):
            by = {}
            for m in draw(st.lists(st.sampled_from(mods), unique=True)):
                n = draw(st.integers(0, 3))
                by[m] = [{"img_path": "%s/%d/%d" % (m, pid, j), "img_id": "%d_%d" % (pid, j), "pid": pid} for j in range(n)]
            index[pid] = by
        return index

    @settings(max_examples=150, deadline=None)
    @given(indices(), st.integers(1, 6), st.sampled_from(["lexi_first", "random"]), st.integers(0, 2 ** 20))
    def check(index, k, policy, seed):
        ra, rb = random.Random(seed), random.Random(seed)
        a = emp.build_queries(index, k, ra, policy)
        b = ref.build_queries(index, k, rb, policy)
        assert a == b
        assert [list(q["samples"]) for q in a] == [list(q["samples"]) for q in b]
        assert ra.getstate() == rb.getstate()

    check()
    for policy in ("lexi_first", "random"):             # mode_k = 0 is outside the protocol: same error as the reference
        with pytest.raises(IndexError):
            ref.build_queries({0: {}}, 0, random.Random(0), policy)
        with pytest.raises(IndexError):
            emp.build_queries({0: {}}, 0, random.Random(0), policy)
