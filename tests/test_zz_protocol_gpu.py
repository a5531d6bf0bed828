"""The MM-1..4 protocol loop (run_eval :553-586, row N4) through the CUDA path against the fixture generated from
the unmodified reference (tests/golden/protocol.json).  Kept in its own file, collected after the kernel tests."""
import json
import os
import random

import pytest
import torch

from prcv2025reid_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "protocol.json"), encoding="utf-8"))


@pytest.fixture(scope="module")
def world():
    index, g_feats, g_meta, ext = synth.make_protocol_index(**GOLDEN["index_args"])
    cs = float(g_feats.double().abs().sum()) + float(sum(float(v.double().abs().sum()) for v in ext.table.values()))
    if abs(cs - GOLDEN["checksum"]) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    return index, g_feats, g_meta, ext


def _check(got, want):
    for name, w in want.items():
        assert got[name].get("num_queries") == w.get("num_queries"), name
        assert abs(got[name]["mAP"] - w["mAP"]) <= 1e-4, (name, got[name], w)        # north star: mAP within 1e-4
        if name == "AVG(1-4)":
            for key in ("R@1", "R@5", "R@10"):
                assert abs(got[name][key] - w[key]) <= 1e-12, (name, key)
        else:
            assert [got[name][k] for k in ("R@1", "R@5", "R@10")] == [w[k] for k in ("R@1", "R@5", "R@10")], (name, got[name], w)


@pytest.mark.parametrize("mode", ["exact", "fused"])
@pytest.mark.parametrize("mask", [True, False])
def test_protocol_loop_matches_reference_golden(world, mode, mask):
    from prcv2025reid_b200 import eval_mm_protocol as emp
    index, g_feats, g_meta, ext = world
    got = emp.run_eval_features(index, g_feats, g_meta, ext, seed=GOLDEN["run_seed"], ignore_same_img=mask, mode=mode)
    _check(got, GOLDEN["run_eval/ignore_same_img=%s" % mask]["results"])


def test_installed_gallery_is_reusable_across_query_sets(world):
    from prcv2025reid_b200 import eval_mm_protocol as emp
    index, g_feats, g_meta, ext = world
    w = dict(synth.DEFAULT_WEIGHTS)
    shard = emp.install_gallery(g_feats, g_meta)
    for k in (3, 1, 4, 2):                                   # growing and shrinking query sets over one shard's scratch
        qs = emp.build_queries(index, k, random.Random(k))
        a = emp.rank_and_metrics(qs, g_feats, g_meta, ext, w, shard=shard)
        b = emp.rank_and_metrics(qs, g_feats, g_meta, ext, w)
        assert {x: a[x] for x in a if x != "mAP"} == {x: b[x] for x in b if x != "mAP"}, (k, a, b)
        assert abs(a["mAP"] - b["mAP"]) <= 1e-12, (k, a, b)


# ---------------------------------------------------------------- train-time evaluation loop (row N2)
VGOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_validate.json"), encoding="utf-8"))


@pytest.mark.parametrize("case", ["default", "cpencil_name", "sampled", "include_all", "no_quad", "first_unmatched"])
def test_validate_competition_style_matches_reference_golden(case):
    import types
    from oracle.make_golden_validate import ARGS, CASES, make_world, nest
    from prcv2025reid_b200 import train_eval as te
    kw = dict(CASES[case])
    g_feat, g_id, sets = make_world(**ARGS, rename=kw.pop("rename", None), first_unmatched=kw.pop("first_unmatched", False))
    if case == "default":
        cs = float(g_feat.double().abs().sum()) + float(sum(float(v[0].double().abs().sum()) for v in sets.values()))
        if abs(cs - VGOLD["checksum"]) > 1e-6 * abs(cs):
            pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    cfg = types.SimpleNamespace(eval_include_patterns=kw["include"]) if "include" in kw else None
    if "torch_seed" in kw:
        torch.manual_seed(kw["torch_seed"])
    got = te.validate_competition_style_features(g_feat, g_id, nest(sets), sample_ratio=kw.get("sample_ratio", 1.0), cfg=cfg)
    want = VGOLD["cases"][case]
    assert list(got["detail"]) == list(want["detail"])
    for name, w in want["detail"].items():
        assert abs(got["detail"][name]["mAP"] - w["mAP"]) <= 1e-4, (name, got["detail"][name], w)
        assert abs(got["detail"][name]["Top1"] - w["Top1"]) <= 1e-12, (name, got["detail"][name], w)
    for key in ("map_single", "map_quad", "map_avg2"):
        assert abs(got[key] - want[key]) <= 1e-4, (key, got[key], want[key])
    assert (got["cmc1"], got["cmc5"], got["cmc10"]) == (want["cmc1"], want["cmc5"], want["cmc10"])


def test_extract_gallery_feats_writes_the_reference_cache(world, tmp_path):
    """eval_mm_protocol.py:291-325 on the CUDA path: one K1 pass over all gallery rows, cache files in the reference's format."""
    import numpy as np
    from oracle import retrieval as orc
    from prcv2025reid_b200 import eval_mm_protocol as emp
    index, g_feats, g_meta, ext = world
    gallery = emp.build_gallery(index)
    feats, meta = emp.extract_gallery_feats(gallery, ext, str(tmp_path / "cache"))
    assert meta == g_meta and feats.device.type == "cpu" and feats.dtype == torch.float32
    assert (feats - orc.l2n(g_feats)).abs().max() <= 4e-7                      # K1 bar
    assert np.array_equal(np.load(str(tmp_path / "cache" / "rgb_feats.npy")), feats.numpy())
    feats2, meta2 = emp.extract_gallery_feats(gallery, ext, str(tmp_path / "cache"))     # cache hit
    assert torch.equal(feats2, feats) and meta2 == meta


# ---------------------------------------------------------------- compute_loss SDM section (row N1) vs the unmodified compute_loss
@pytest.mark.parametrize("name", ["full", "ragged", "no_vis", "no_pairs", "missing", "full_bf16", "ragged_bf16", "full_fp16",
                                  "large_bf16"])
def test_sdm_alignment_loss_matches_compute_loss_golden(name):
    """fp32 features: loss and gradients within 1e-5 of the unmodified compute_loss.  bf16 / fp16 features (what the training
    autocast hands over, train.py:852): the loss normalises in that dtype like the reference (sdm_loss.py:31-32), loss within
    1e-3; the gradients are judged like every 16-bit gradient here: against the exact float64 gradient, and never further
    from it than the reference's own 16-bit autograd gradient (the fixture) is."""
    import numpy as np
    from oracle import sdm as osdm
    from oracle.make_golden_alignment import CASES, make_inputs
    from prcv2025reid_b200.sdm_loss import sdm_alignment_loss
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdm_alignment.npz"))
    spec = CASES[name]
    seed, B, d, n_ids, kind, tau = spec[:6]
    feats, masks, labels = make_inputs(seed, B, d, n_ids, kind, *spec[6:])
    half = len(spec) > 6 and spec[6] != "fp32"
    cs = sum(float(f.double().abs().sum()) for f in feats.values() if f is not None)
    if abs(cs - float(z[name + "/checksum"])) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    leaves = {m: (f.cuda().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
    loss = sdm_alignment_loss(leaves, {m: v.cuda() for m, v in masks.items()}, labels.cuda(), tau=tau)
    want = float(z[name + "/loss"])
    tol = 1e-3 if half else 1e-5                                                          # north star: 1e-3 (bf16) / 1e-5 (fp32)
    assert abs(float(loss.detach()) - want) <= tol * max(1.0, abs(want))
    if loss.requires_grad:
        loss.backward()
    if half:
        # exact gradient of the same objective: float64 autograd of the restatement on the 16-bit values
        f64 = {m: (f.double().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
        l64 = osdm.sdm_alignment_oracle(f64, masks, labels, tau=tau)
        l64.backward()
    for m, t in leaves.items():
        key = name + "/grad_" + m
        if key in z.files:
            r = z[key]
            assert t.grad is not None and t.grad.dtype == t.dtype, m
            if half:
                rows = np.abs(r).sum(1) > 0                                              # rows that take part (mask set)
                got = t.grad.float().cpu().numpy()
                assert not got[~rows].any()
                g64 = f64[m].grad.numpy()
                e_got = np.linalg.norm(got - g64) / np.linalg.norm(g64)
                e_ref = np.linalg.norm(r - g64) / np.linalg.norm(g64)
                print("%s grad_%s: |got - exact| %.2e, reference autograd vs exact %.2e, got vs reference %.2e"
                      % (name, m, e_got, e_ref, np.linalg.norm(got - r) / np.linalg.norm(r)))
                assert e_got <= (5e-3 if spec[6] == "bf16" else 1.5e-3) and e_got <= e_ref + 1e-4, m
            else:
                assert np.abs(t.grad.cpu().numpy() - r).max() <= 1e-5 * max(float(np.abs(r).max()), 1e-12), m
        elif t is not None and t.grad is not None:
            assert float(t.grad.abs().sum()) == 0.0, m                                  # no gradient in the reference
