"""The oracle restatements against the golden fixtures produced by the unmodified reference
(and, when /root/reference is present, against the live reference functions)."""
import numpy as np
import pytest
import torch

from oracle import retrieval as orc
from oracle import sdm as osdm
from oracle import ref_loader
from prcv2025reid_b200 import synth
from tests import _golden


@pytest.mark.parametrize("name", _golden.RETRIEVAL_NAMES)
def test_retrieval_oracle_matches_reference_golden(name):
    case, z = _golden.load_retrieval(name)
    w = synth.weights_tensor()
    g = orc.l2n(case.gallery_raw)
    q = orc.fuse_queries(case.query_raw, case.mod_id, w)
    if "q_fused" in z:
        assert np.array_equal(q.numpy(), z["q_fused"])          # bit-exact vs extract_query_feat
        assert np.array_equal(g.numpy(), z["g_norm"])
    else:
        assert np.array_equal(q.numpy()[::7], z["q_fused_s"])
        assert np.array_equal(g.numpy()[::13], z["g_norm_s"])
    loop = orc.rank_and_metrics_loop(q, g, case.q_pid, case.g_pid, case.excl, return_per_query=True)
    gold = z["metrics"]
    assert [loop["mAP"], loop["R@1"], loop["R@5"], loop["R@10"], loop["num_queries"]] == list(gold)
    assert np.array_equal(loop["_top_idx"], z["top10"])
    nomask = orc.rank_and_metrics_loop(q, g, case.q_pid, case.g_pid, None)
    assert [nomask["mAP"], nomask["R@1"], nomask["R@5"], nomask["R@10"], nomask["num_queries"]] == list(z["metrics_nomask"])
    cnt = orc.rank_and_metrics_counting(q, g, case.q_pid, case.g_pid, case.excl)
    assert cnt["num_queries"] == loop["num_queries"]
    assert abs(cnt["mAP"] - loop["mAP"]) < 1e-9
    for k in ("R@1", "R@5", "R@10"):
        assert cnt[k] == loop[k]
    sub = orc.submission_ranking(q, g, top_k=20)
    assert np.array_equal(sub, z["submission"])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_retrieval_oracle_matches_live_reference():
    ref = ref_loader.load_reference_eval()
    case = synth.make_retrieval_case(77, 25, 5, 3, 4, excl_frac=0.2, n_excl=2)
    queries, gmeta, ext = synth.case_to_reference_inputs(case)
    w = synth.weights_tensor()
    g = ref.l2n(case.gallery_raw)
    m = ref_loader.quiet(ref.rank_and_metrics, queries, g, gmeta, ext, dict(synth.DEFAULT_WEIGHTS))
    q = orc.fuse_queries(case.query_raw, case.mod_id, w)
    o = orc.rank_and_metrics_loop(q, orc.l2n(case.gallery_raw), case.q_pid, case.g_pid, case.excl)
    assert m == o


def test_query_without_positive_is_skipped():
    case = synth.make_retrieval_case(5, 10, 3, 2, 2, excl_frac=0.0)
    q_pid = case.q_pid.clone(); q_pid[:4] = 10_000     # ids absent from the gallery
    w = synth.weights_tensor()
    q = orc.fuse_queries(case.query_raw, case.mod_id, w); g = orc.l2n(case.gallery_raw)
    out = orc.rank_and_metrics_loop(q, g, q_pid, case.g_pid, case.excl)
    assert out["num_queries"] == case.Q - 4
    out2 = orc.rank_and_metrics_counting(q, g, q_pid, case.g_pid, case.excl)
    assert out2["num_queries"] == case.Q - 4


def test_zero_row_normalises_to_zero():
    x = torch.zeros(3, 512); x[1] = 1.0
    y = orc.l2n(x)
    assert torch.all(y[0] == 0) and torch.all(y[2] == 0)
    assert abs(float(y[1].norm()) - 1.0) < 1e-6


SDM_NAMES = ["p4k2_tau02", "p4k2_tau01", "p3k2", "ragged", "no_pos", "nan_feat", "quick_check",
             "p64k8_fp32", "p64k8_bf16", "p4k2_bf16"]


@pytest.mark.parametrize("name", SDM_NAMES)
def test_sdm_oracle_matches_reference_golden(name):
    c = _golden.load_sdm()[name]
    q, v, y = _golden.sdm_inputs(c)
    q = q.clone().requires_grad_(True); v = v.clone().requires_grad_(True)
    loss = osdm.sdm_loss_oracle(q, v, y, tau=float(c["tau"]))
    assert float(loss.detach()) == float(c["loss"])
    assert bool(loss.requires_grad) == bool(c["differentiable"])
    if loss.requires_grad:
        loss.backward()
        dq, dv = q.grad.float().numpy(), v.grad.float().numpy()
        if "dq" in c:
            assert np.array_equal(dq, c["dq"]) and np.array_equal(dv, c["dv"])
        else:
            assert np.array_equal(dq[::16], c["dq_s"]) and np.array_equal(dv[::16], c["dv_s"])


@pytest.mark.parametrize("name", ["p4k2_tau02", "p4k2_tau01", "p3k2", "ragged", "quick_check"])
def test_sdm_closed_form_matches_reference_autograd(name):
    c = _golden.load_sdm()[name]
    q, v, y = _golden.sdm_inputs(c)
    loss, dq, dv = osdm.sdm_fwd_bwd_f64(q, v, y, tau=float(c["tau"]))
    assert abs(loss - float(c["loss"])) < 1e-5 * max(1.0, abs(loss))
    assert np.abs(dq - c["dq"]).max() < 1e-5 * np.abs(c["dq"]).max() + 1e-9
    assert np.abs(dv - c["dv"]).max() < 1e-5 * np.abs(c["dv"]).max() + 1e-9


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_sdm_oracle_matches_live_reference():
    sdm = ref_loader.load_reference_sdm().sdm_loss_stable
    g = torch.Generator().manual_seed(123)
    q = torch.randn(24, 512, generator=g); v = torch.randn(20, 512, generator=g)
    lq = torch.randint(0, 6, (24,), generator=g); lv = torch.randint(0, 7, (20,), generator=g)
    y = (lq[:, None] == lv[None, :]).float()
    a = ref_loader.quiet(sdm, q, v, y, tau=0.3)
    b = osdm.sdm_loss_oracle(q, v, y, tau=0.3)
    assert float(a) == float(b)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_sdm_oracle_property_vs_live_reference():
    """Random shapes, label layouts (rows / columns without a positive, no positive at all), temperatures outside the
    clamp range, bf16 inputs and a non-finite feature: loss value, result dtype, differentiability and the autograd
    gradients of the restatement are bit-equal to the unmodified sdm_loss_stable's."""
    from hypothesis import given, settings, strategies as st
    sdm = ref_loader.load_reference_sdm().sdm_loss_stable

    @settings(max_examples=80, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 12), st.integers(1, 12), st.sampled_from([8, 48]),
           st.sampled_from([0.05, 0.15, 0.2, 0.37, 0.5, 0.9]), st.sampled_from(["f32", "bf16", "nan", "nopos"]))
    def check(seed, N, M, d, tau, kind):
        g = torch.Generator().manual_seed(seed)
        q = torch.randn(N, d, generator=g); v = torch.randn(M, d, generator=g)
        lq = torch.randint(0, 4, (N,), generator=g); lv = torch.randint(0, 5, (M,), generator=g)
        y = (lq[:, None] == lv[None, :]).float()
        if kind == "nopos":
            y = torch.zeros_like(y)
        if kind == "nan":
            q[0, 0] = float("nan")
        if kind == "bf16":
            q, v = q.bfloat16(), v.bfloat16()
        qa, va = q.clone().requires_grad_(True), v.clone().requires_grad_(True)
        qb, vb = q.clone().requires_grad_(True), v.clone().requires_grad_(True)
        a = ref_loader.quiet(sdm, qa, va, y, tau=tau)
        b = osdm.sdm_loss_oracle(qb, vb, y, tau=tau)
        assert a.dtype == b.dtype and a.shape == b.shape and a.requires_grad == b.requires_grad
        assert torch.equal(a.detach(), b.detach()) or (torch.isnan(a) and torch.isnan(b))
        if a.requires_grad:
            a.backward(); b.backward()
            assert torch.equal(qa.grad, qb.grad) and torch.equal(va.grad, vb.grad)

    check()


# ---------------------------------------------------------------- train-time evaluator
def _train_eval_case():
    import os
    from oracle.make_golden_train_eval import make_case
    z = np.load(os.path.join(_golden.GOLDEN, "train_eval.npz"))
    qf, gf, ql, gl = make_case()
    cs = float(qf.double().abs().sum()) + float(gf.double().abs().sum())
    if abs(cs - float(z["checksum"])) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    return z, qf, gf, ql, gl


def test_train_eval_oracle_matches_reference_golden():
    z, qf, gf, ql, gl = _train_eval_case()
    for k in (1, 5, 100):
        assert orc.compute_map_oracle(qf, gf, ql, gl, k=k) == float(z["map_k%d" % k])
    for k in (1, 10):
        assert orc.compute_cmc_oracle(qf, gf, ql, gl, k=k) == float(z["cmc_k%d" % k])
    qn = torch.nn.functional.normalize(qf, dim=1); gn = torch.nn.functional.normalize(gf, dim=1)
    assert list(orc.reid_map_oracle(qn @ gn.T, ql, gl)) == list(z["reid_map"])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_train_eval_oracle_matches_live_reference():
    ns = ref_loader.load_reference_train_eval()
    g = torch.Generator().manual_seed(3)
    gl = torch.arange(120) // 4; ql = torch.randint(0, 34, (40,), generator=g)
    c = torch.randn(34, 32, generator=g)
    gf = c[gl] + 2.0 * torch.randn(120, 32, generator=g); qf = c[ql.clamp(max=29)] + 2.0 * torch.randn(40, 32, generator=g)
    for k in (3, 100):
        assert ns["compute_map"](qf, gf, ql, gl, k) == orc.compute_map_oracle(qf, gf, ql, gl, k)
        assert ns["compute_cmc"](qf, gf, ql, gl, k) == orc.compute_cmc_oracle(qf, gf, ql, gl, k)
    qn = torch.nn.functional.normalize(qf, dim=1); gn = torch.nn.functional.normalize(gf, dim=1)
    assert ns["_reid_map"](qn @ gn.T, ql, gl) == orc.reid_map_oracle(qn @ gn.T, ql, gl)


# ---------------------------------------------------------------- gallery store, host logic only
def test_gallery_store_reads_the_reference_cache_format(tmp_path):
    import json, pickle
    from prcv2025reid_b200 import gallery_store
    feats = torch.randn(37, 16)
    meta = [{"img_id": "g%d" % i, "pid": i // 3, "camid": None} for i in range(37)]
    # written exactly as the reference writes it (eval_mm_protocol.py:320-323)
    np.save(str(tmp_path / "rgb_feats.npy"), np.stack([f.numpy() for f in feats], 0))
    json.dump(meta, open(str(tmp_path / "rgb_meta.json"), "w", encoding="utf-8"))
    mm = gallery_store.open_feats(str(tmp_path))
    assert isinstance(mm, np.memmap) and mm.shape == (37, 16) and np.array_equal(np.asarray(mm), feats.numpy())
    assert gallery_store.load_meta(str(tmp_path)) == meta
    # and what this module writes is what the reference reads (:296-302)
    gallery_store.save_cache(str(tmp_path / "out"), feats, meta)
    back = torch.from_numpy(np.load(str(tmp_path / "out" / "rgb_feats.npy"))).float()
    assert torch.equal(back, feats) and json.load(open(str(tmp_path / "out" / "rgb_meta.json"), encoding="utf-8")) == meta
    with open(str(tmp_path / "c.pkl"), "wb") as f:                       # train.py:626-631
        pickle.dump({"g_feat": feats, "g_id": torch.arange(37)}, f)
    gf, gi = gallery_store.load_pickle_cache(str(tmp_path / "c.pkl"))
    assert gf.dtype == np.float32 and gi.dtype == np.int64 and np.array_equal(gf, feats.numpy())


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_retrieval_oracle_property_vs_live_reference():
    """Random tiny jobs -- duplicated gallery rows (exact score ties), zero rows, masked positives, queries without a
    positive, every k: the restatement returns exactly the dict the unmodified rank_and_metrics returns."""
    from hypothesis import given, settings, strategies as st
    ref = ref_loader.load_reference_eval()
    wcfg = dict(synth.DEFAULT_WEIGHTS)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 4), st.integers(2, 24), st.integers(1, 10), st.booleans())
    def check(seed, k, G, Q, mask):
        g = torch.Generator().manual_seed(seed)
        d = 16
        n_ids = max(1, G // 3)
        gallery = torch.randn(G, d, generator=g)
        dup = torch.randint(0, G, (G,), generator=g)
        take = torch.rand(G, generator=g) < 0.3
        gallery = torch.where(take[:, None], gallery[dup], gallery)              # exact duplicates -> tied scores
        if G > 4:
            gallery[int(torch.randint(0, G, (1,), generator=g))] = 0.0            # a zero row (l2n keeps it zero)
        g_pid = torch.randint(0, n_ids, (G,), generator=g)
        q_pid = torch.randint(0, n_ids + 1, (Q,), generator=g)                    # id n_ids: no positive in the gallery
        query = torch.randn(Q, k, d, generator=g)
        mod_id = torch.stack([torch.randperm(4, generator=g)[:k].sort().values for _ in range(Q)]).to(torch.int32)
        E = min(2, k)                                                             # one same-image row per query sample at most
        excl = torch.where(torch.rand(Q, E, generator=g) < 0.4, torch.randint(0, G, (Q, E), generator=g), torch.tensor(-1)).to(torch.int32)
        case = synth.RetrievalCase(gallery, g_pid, query, mod_id, q_pid, excl, k)
        queries, gmeta, ext = synth.case_to_reference_inputs(case)
        gn = ref.l2n(gallery)
        want = ref_loader.quiet(ref.rank_and_metrics, queries, gn, gmeta, ext, wcfg, ignore_same_img=mask)
        qf = orc.fuse_queries(query, mod_id, synth.weights_tensor())
        got = orc.rank_and_metrics_loop(qf, orc.l2n(gallery), q_pid, g_pid, excl if mask else None)
        assert got == want

    check()


# ---------------------------------------------------------------- compute_loss SDM section
ALIGN_CASES = ["full", "ragged", "no_vis", "no_pairs", "missing"]


def _align_case(name):
    import os
    from oracle.make_golden_alignment import CASES, make_inputs
    z = np.load(os.path.join(_golden.GOLDEN, "sdm_alignment.npz"))
    seed, B, d, n_ids, kind, tau = CASES[name]
    feats, masks, labels = make_inputs(seed, B, d, n_ids, kind)
    cs = sum(float(f.double().abs().sum()) for f in feats.values() if f is not None)
    if abs(cs - float(z[name + "/checksum"])) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    grads = {k.split("grad_")[1]: z[k] for k in z.files if k.startswith(name + "/grad_")}
    return feats, masks, labels, tau, float(z[name + "/loss"]), grads


@pytest.mark.parametrize("name", ALIGN_CASES)
def test_alignment_oracle_matches_compute_loss_golden(name):
    """sdm_alignment_oracle vs the fixture produced by the UNMODIFIED compute_loss (models/model.py:512-659)."""
    feats, masks, labels, tau, loss, grads = _align_case(name)
    leaves = {m: (f.clone().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
    got = osdm.sdm_alignment_oracle(leaves, masks, labels, tau=tau)
    assert float(got.detach()) == loss
    if got.requires_grad:
        got.backward()
    have = {m: t.grad.numpy() for m, t in leaves.items() if t is not None and t.grad is not None}
    assert sorted(have) == sorted(grads)
    for m in grads:
        assert np.array_equal(have[m], grads[m]), m


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_alignment_oracle_matches_live_compute_loss():
    from oracle.make_golden_alignment import make_inputs, run_reference
    for seed, kind in ((11, "full"), (12, "ragged"), (13, "missing"), (14, "no_pairs")):
        feats, masks, labels = make_inputs(seed, 10, 96, 3, kind)
        want, grads, out = run_reference(feats, masks, labels, 0.25)
        leaves = {m: (f.clone().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
        got = osdm.sdm_alignment_oracle(leaves, masks, labels, tau=0.25)
        assert float(got.detach()) == float(want)
        assert float(out["total_loss"].detach()) == pytest.approx(0.3 * float(want) + float(out["ce_loss"].detach()))   # :651
        if got.requires_grad:
            got.backward()
            for m, gr in grads.items():
                assert torch.equal(leaves[m].grad, gr), m
