"""Host logic of prcv2025reid_b200.sdm_loss (the autograd Function around reid_sdm_fwd / reid_sdm_bwd, the packed
reid_sdm_pair table, batching of pairs, SdmStep, and the compute_loss section `sdm_alignment_loss`) WITHOUT a GPU: the
C library is replaced by tests/_fake_lib.FakeLib, whose SDM entry points decode the pair table exactly as the header
lays it out and evaluate the oracle.  Compared with the fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from tests import _fake_lib, _golden


@pytest.mark.parametrize("name", ["p4k2_tau02", "p3k2", "ragged", "no_pos", "nan_feat", "quick_check", "p4k2_bf16"])
def test_pairs_function_reproduces_reference_golden(monkeypatch, name):
    from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
    fake = _fake_lib.install(monkeypatch)
    c = _golden.load_sdm()[name]
    q, v, y = _golden.sdm_inputs(c)
    qa, va = q.clone().requires_grad_(True), v.clone().requires_grad_(True)
    loss = sdm_loss_pairs([qa], [va], [y], tau=float(c["tau"]))
    assert loss.shape == (1,) and loss.dtype == torch.float32 and loss.requires_grad       # always attached
    assert float(loss[0].detach()) == pytest.approx(float(c["loss"]), rel=1e-6, abs=1e-7)
    loss.sum().backward()
    if "dq" in c and np.abs(c["dq"]).max() > 0:
        assert np.allclose(qa.grad.float().numpy(), c["dq"], rtol=1e-5, atol=1e-8)
        assert np.allclose(va.grad.float().numpy(), c["dv"], rtol=1e-5, atol=1e-8)
    else:                                                                                   # guard paths: exact zeros
        assert float(qa.grad.abs().sum()) == 0.0 and float(va.grad.abs().sum()) == 0.0
    assert fake.calls.count("reid_sdm_fwd") == 1 and fake.calls.count("reid_sdm_bwd") == 1


def test_pairs_are_batched_and_gradients_routed(monkeypatch):
    """20 pairs of different shapes, non-contiguous inputs, a weighted objective: two forward calls (<= 16 pairs each),
    every gradient lands on its own tensor with its own weight."""
    from oracle import sdm as osdm
    from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
    fake = _fake_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(8)
    qs, vs, ys, ws = [], [], [], []
    for p in range(20):
        N, M = 3 + p % 5, 4 + p % 3
        base = torch.randn(N, 64, generator=g)
        qs.append(base[:, ::2].requires_grad_(True))                    # non-contiguous view, d = 32
        vs.append(torch.randn(M, 32, generator=g).requires_grad_(True))
        ys.append((torch.randint(0, 3, (N, 1), generator=g) == torch.randint(0, 3, (1, M), generator=g)).float())
        ws.append(0.5 + p)
    losses = sdm_loss_pairs(qs, vs, ys, tau=0.2)
    assert losses.shape == (20,) and fake.calls.count("reid_sdm_fwd") == 2
    (losses * torch.tensor(ws)).sum().backward()
    for p in range(20):
        qc, vc = qs[p].detach().clone().requires_grad_(True), vs[p].detach().clone().requires_grad_(True)
        L = osdm.sdm_loss_oracle(qc, vc, ys[p], tau=0.2)
        assert float(losses[p].detach()) == pytest.approx(float(L.detach()), rel=1e-6)
        if L.requires_grad:
            L.backward()
            assert torch.allclose(qs[p].grad, ws[p] * qc.grad, rtol=1e-5, atol=1e-9)
            assert torch.allclose(vs[p].grad, ws[p] * vc.grad, rtol=1e-5, atol=1e-9)
    with pytest.raises(ValueError):
        sdm_loss_pairs([qs[0]], [vs[0]], [ys[1]])                       # y of another shape
    with pytest.raises(TypeError):
        sdm_loss_pairs([qs[0].double()], [vs[0].double()], [ys[0]])     # unsupported dtype
    with pytest.raises(ValueError):
        sdm_loss_pairs([], [], [])


def test_step_object_matches_autograd(monkeypatch):
    from prcv2025reid_b200.sdm_loss import SdmStep, sdm_loss_pairs
    fake = _fake_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(4)
    qs = [torch.randn(8, 32, generator=g) for _ in range(4)]
    vs = [torch.randn(8, 32, generator=g) for _ in range(4)]
    lab = torch.arange(4).repeat_interleave(2)
    ys = [(lab[:, None] == lab[None, :]).float()] * 4
    w = torch.tensor([0.25, 0.25, 0.25, 0.25])                          # the mean of models/model.py:622
    st = SdmStep(qs, vs, ys, tau=0.2, weights=w)
    losses = st.run()
    qa = [q.clone().requires_grad_(True) for q in qs]; va = [v.clone().requires_grad_(True) for v in vs]
    ref = sdm_loss_pairs(qa, va, ys, tau=0.2)
    ref.mean().backward()
    assert torch.allclose(losses, ref.detach(), rtol=1e-6)
    for p in range(4):
        assert torch.allclose(st.dq[p], qa[p].grad, rtol=1e-5, atol=1e-9) and torch.allclose(st.dg[p], va[p].grad, rtol=1e-5, atol=1e-9)
    assert fake.calls.count("reid_sdm_step") == 1 and st.launches == 1
    with pytest.raises(ValueError):
        SdmStep(qs * 5, vs * 5, ys * 5)                                 # more than REID_SDM_MAX_PAIRS


@pytest.mark.parametrize("name", ["full", "ragged", "no_vis", "no_pairs", "missing", "full_bf16", "large_bf16"])
def test_alignment_loss_host_logic_matches_compute_loss_golden(monkeypatch, name):
    """sdm_alignment_loss (mask filtering with one host read, y from labels, pairs without a positive dropped, mean)
    against the fixture produced by the UNMODIFIED compute_loss (models/model.py:512-659)."""
    import os
    from oracle.make_golden_alignment import CASES, make_inputs
    from prcv2025reid_b200.sdm_loss import sdm_alignment_loss
    fake = _fake_lib.install(monkeypatch)
    z = np.load(os.path.join(_golden.GOLDEN, "sdm_alignment.npz"))
    spec = CASES[name]
    seed, B, d, n_ids, kind, tau = spec[:6]
    feats, masks, labels = make_inputs(seed, B, d, n_ids, kind, *spec[6:])
    cs = sum(float(f.double().abs().sum()) for f in feats.values() if f is not None)
    if abs(cs - float(z[name + "/checksum"])) > 1e-6 * abs(cs):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    leaves = {m: (f.clone().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
    # every case runs in the LABEL FORM of the C entry points (any batch size / dtype since the CUDA-core kernels take it
    # too): no y, no row filtering, no host read of the masks
    real_cpu = torch.Tensor.cpu
    monkeypatch.setattr(torch.Tensor, "cpu", lambda self, *a, **k: (_ for _ in ()).throw(AssertionError("host read")))
    loss = sdm_alignment_loss(leaves, masks, labels, tau=tau)
    monkeypatch.setattr(torch.Tensor, "cpu", real_cpu)
    assert fake.calls.count("reid_sdm_fwd") <= 1
    want = float(z[name + "/loss"])
    assert float(loss.detach()) == pytest.approx(want, rel=1e-6, abs=1e-7)
    if loss.requires_grad:
        loss.backward()
    for m, t in leaves.items():
        key = name + "/grad_" + m
        if key in z.files:
            # (the features keep their dtype on the way to the C call: with the oracle behind the stand-in library the
            #  bf16 case reproduces the reference's own bf16 gradient)
            assert t.grad is not None and t.grad.dtype == t.dtype, m
            assert np.allclose(t.grad.float().numpy(), z[key], rtol=1e-5 if t.dtype == torch.float32 else 2e-2, atol=1e-9 if t.dtype == torch.float32 else 1e-4), m
        elif t is not None and t.grad is not None:
            assert float(t.grad.abs().sum()) == 0.0, m


def test_step_object_label_form_and_the_ragged_route(monkeypatch):
    """SdmStep in the label form packs the label / valid pointers into the pair table (one reid_sdm_step call, same result as
    the autograd label form); sdm_alignment_loss issues one label-form call and rejects mismatching modalities."""
    from prcv2025reid_b200.sdm_loss import SdmStep, sdm_alignment_loss, sdm_loss_pairs_labels
    fake = _fake_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(5)
    B = 10
    qs = [torch.randn(B, 32, generator=g) for _ in range(2)]
    vs = [torch.randn(B, 32, generator=g) for _ in range(2)]
    lab = torch.arange(5).repeat_interleave(2)
    rv = torch.ones(B, dtype=torch.bool); rv[3] = False
    cv = torch.ones(B, dtype=torch.bool); cv[[0, 7]] = False
    w = torch.tensor([0.5, 2.0])
    st = SdmStep(qs, vs, None, tau=0.2, weights=w, labels=[(lab, lab, rv, cv), (lab, lab, None, None)])
    losses = st.run()
    qa = [q.clone().requires_grad_(True) for q in qs]; va = [v.clone().requires_grad_(True) for v in vs]
    ref, status = sdm_loss_pairs_labels(qa, va, [lab, lab], [lab, lab], [rv, None], [cv, None], tau=0.2)
    (ref * w).sum().backward()
    assert torch.allclose(losses, ref.detach(), rtol=1e-6) and int(status[0]) == 0
    for p in range(2):
        assert torch.allclose(st.dq[p], qa[p].grad, rtol=1e-5, atol=1e-9) and torch.allclose(st.dg[p], va[p].grad, rtol=1e-5, atol=1e-9)
    assert not st.dq[0][3].any() and not st.dg[0][0].any() and not st.dg[0][7].any()       # masked rows: exact zeros
    assert fake.calls.count("reid_sdm_step") == 1
    with pytest.raises(ValueError):
        SdmStep(qs, vs, None)                                           # neither y nor labels
    with pytest.raises(ValueError):
        sdm_loss_pairs_labels(qs, vs, [lab, lab], [lab, lab], [rv[:4], None], None)        # one valid byte per row
    # sdm_alignment_loss: one label-form call for all modalities; mismatching modalities are rejected (the reference would
    # fail on them too: labels[mod_valid_idx] needs one row per label, sdm_loss.py:75 one dtype)
    feats = {"vis": vs[0], "nir": qs[0], "sk": qs[1]}
    masks = {"vis": cv.float().view(-1, 1), "nir": rv.float().view(-1, 1), "sk": torch.ones(B, 1)}
    n0 = fake.calls.count("reid_sdm_fwd")
    a = sdm_alignment_loss(feats, masks, lab, tau=0.2)
    assert fake.calls.count("reid_sdm_fwd") == n0 + 1 and float(a) > 0
    with pytest.raises(ValueError):
        sdm_alignment_loss(dict(feats, sk=qs[1][:6]), masks, lab, tau=0.2)
    with pytest.raises(TypeError):
        sdm_alignment_loss(dict(feats, sk=qs[1].to(torch.bfloat16)), masks, lab, tau=0.2)
    with pytest.raises(TypeError):
        sdm_alignment_loss({k: v.double() for k, v in feats.items()}, masks, lab, tau=0.2)
