"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE.  Run:  python -m oracle.make_golden      (needs /root/reference)

The reference ships no golden vectors for this path, so parity is pinned
on outputs of the reference's own functions:
  tools/eval_mm_protocol.py : l2n, cosine_sim, extract_query_feat, rank_and_metrics,
                              export_submission_csv  (called through a tensor-serving fake extractor)
  models/sdm_loss.py        : sdm_loss_stable + torch autograd
Inputs of the small cases are stored in the fixture; larger cases store the generator arguments
plus a checksum of the generated inputs so RNG drift is detected instead of mis-reported.
"""
import csv
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from prcv2025reid_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

RETRIEVAL_CASES = {
    # name: (seed, n_ids, gal_per_id, k, queries_per_id, excl_frac, n_excl, store_inputs)
    "mm2_tiny": (11, 12, 4, 2, 6, 0.25, 2, True),
    "mm4_tiny": (12, 12, 4, 4, 2, 0.25, 2, True),
    "mm1_small": (13, 40, 6, 1, 4, 0.10, 1, False),
    "mm2_small": (14, 60, 8, 2, 6, 0.10, 2, False),
    "mm3_small": (15, 60, 8, 3, 4, 0.10, 2, False),
    "mm4_small": (16, 60, 8, 4, 2, 0.10, 2, False),
}


def checksum(*tensors) -> float:
    return float(sum(float(t.double().abs().sum()) for t in tensors))


def reference_retrieval(case, weight_cfg):
    ref = ref_loader.load_reference_eval()
    queries, gmeta, ext = synth.case_to_reference_inputs(case)
    g = ref.l2n(case.gallery_raw)                                     # eval_mm_protocol.py:546
    metrics = ref_loader.quiet(ref.rank_and_metrics, queries, g, gmeta, ext, weight_cfg,
                               ignore_same_img=True)
    metrics_nomask = ref_loader.quiet(ref.rank_and_metrics, queries, g, gmeta, ext, weight_cfg,
                                      ignore_same_img=False)
    qf = torch.stack([ref.extract_query_feat(q, ext, weight_cfg) for q in queries])
    # top-10 under the mask exactly as rank_and_metrics builds it (:401-423)
    gid = {m["img_id"]: i for i, m in enumerate(gmeta)}
    top10 = np.zeros((len(queries), 10), dtype=np.int64)
    top10_val = np.zeros((len(queries), 10), dtype=np.float32)
    for qi, q in enumerate(queries):
        sims = ref.cosine_sim(qf[qi].view(1, -1), g).squeeze(0)
        sm = sims.clone()
        for s in q["samples"].values():
            if s["img_id"] in gid:
                sm[gid[s["img_id"]]] = -1e9
        r = torch.argsort(sm, descending=True)[:10]
        top10[qi] = r.numpy(); top10_val[qi] = sm[r].numpy()
    # submission ranking (no mask), top 20, through the reference's CSV writer (:595-649)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sub.csv")
        ref_loader.quiet(ref.export_submission_csv, queries, g, gmeta, ext, weight_cfg, path, top_k=20)
        sub = []
        with open(path, newline="") as f:
            for row in csv.DictReader(f):
                sub.append([int(t[1:]) for t in row["ranked_gallery_ids"].split()])
    return dict(metrics=metrics, metrics_nomask=metrics_nomask, q_fused=qf.numpy(),
                g_norm=g.numpy(), top10=top10, top10_val=top10_val, submission=np.array(sub))


def make_retrieval():
    weight_cfg = dict(synth.DEFAULT_WEIGHTS)
    for name, (seed, n_ids, gpi, k, qpi, ef, ne, store) in RETRIEVAL_CASES.items():
        case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=ef, n_excl=min(ne, k))
        out = reference_retrieval(case, weight_cfg)
        m, mn = out["metrics"], out["metrics_nomask"]
        payload = dict(
            args=np.array([seed, n_ids, gpi, k, qpi, ne], dtype=np.int64), excl_frac=np.float64(ef),
            checksum=np.float64(checksum(case.gallery_raw, case.query_raw)),
            metrics=np.array([m["mAP"], m["R@1"], m["R@5"], m["R@10"], m["num_queries"]], dtype=np.float64),
            metrics_nomask=np.array([mn["mAP"], mn["R@1"], mn["R@5"], mn["R@10"], mn["num_queries"]], dtype=np.float64),
            top10=out["top10"].astype(np.int32), top10_val=out["top10_val"],
            submission=out["submission"].astype(np.int32),
        )
        if store:
            payload.update(gallery_raw=case.gallery_raw.numpy(), query_raw=case.query_raw.numpy(),
                           mod_id=case.mod_id.numpy(), q_pid=case.q_pid.numpy(), g_pid=case.g_pid.numpy(),
                           excl=case.excl.numpy(), q_fused=out["q_fused"], g_norm=out["g_norm"])
        else:
            # a strided sample of the fused query / normalised gallery rows pins K1/K2
            payload.update(q_fused_s=out["q_fused"][::7], g_norm_s=out["g_norm"][::13])
        np.savez_compressed(os.path.join(GOLDEN, "retrieval_%s.npz" % name), **payload)
        print(name, m)


def _sdm_inputs(N, M, D, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(N, D, generator=g)
    v = torch.randn(M, D, generator=g)
    return q.to(dtype), v.to(dtype)


def _run_sdm(q, v, y, tau):
    sdm = ref_loader.load_reference_sdm().sdm_loss_stable
    q = q.clone().requires_grad_(True); v = v.clone().requires_grad_(True)
    loss = ref_loader.quiet(sdm, q, v, y, tau=tau)
    if loss.requires_grad:
        loss.backward()
        return loss.detach(), q.grad, v.grad, True
    return loss.detach(), torch.zeros_like(q), torch.zeros_like(v), False


def make_sdm():
    cases = {}
    def pk_labels(P, K):
        return torch.arange(P).repeat_interleave(K)
    def add(name, q, v, lq, lv, tau, store_inputs=True):
        y = (lq[:, None] == lv[None, :]).float()
        loss, dq, dv, diff = _run_sdm(q, v, y, tau)
        d = dict(labels_q=lq.numpy(), labels_v=lv.numpy(), tau=np.float64(tau),
                 loss=np.float64(float(loss)), differentiable=np.bool_(diff),
                 dq_norm=np.float64(float(dq.float().norm())), dv_norm=np.float64(float(dv.float().norm())),
                 seed_shape=np.array([0, q.shape[0], v.shape[0], q.shape[1]]),
                 is_bf16=np.bool_(q.dtype == torch.bfloat16))
        if store_inputs:
            d.update(q=q.float().numpy(), v=v.float().numpy(), dq=dq.float().numpy(), dv=dv.float().numpy())
        else:
            d.update(dq_s=dq.float().numpy()[::16], dv_s=dv.float().numpy()[::16])
        cases[name] = d
        print(name, float(loss), float(dq.float().norm()), float(dv.float().norm()))
    # C2 family (SURVEY.md section 8c known answers)
    q, v = _sdm_inputs(8, 8, 512);  add("p4k2_tau02", q, v, pk_labels(4, 2), pk_labels(4, 2), 0.2)
    add("p4k2_tau01", q, v, pk_labels(4, 2), pk_labels(4, 2), 0.1)
    q6, v6 = _sdm_inputs(6, 6, 512); add("p3k2", q6, v6, pk_labels(3, 2), pk_labels(3, 2), 0.2)
    # ragged: rows / columns without positives, N != M
    g = torch.Generator().manual_seed(5)
    qr = torch.randn(10, 512, generator=g); vr = torch.randn(14, 512, generator=g)
    lq = torch.tensor([0, 0, 1, 2, 3, 9, 9, 4, 5, 1]); lv = torch.tensor([0, 1, 1, 2, 7, 7, 8, 4, 4, 4, 6, 6, 0, 3])
    add("ragged", qr, vr, lq, lv, 0.2)
    # no positives at all -> non-differentiable zero (sdm_loss.py:105-106)
    add("no_pos", qr, vr, torch.arange(10), torch.arange(14) + 100, 0.2)
    # non-finite feature -> zero (sdm_loss.py:79-81)
    qn = qr.clone(); qn[3, 7] = float("nan")
    add("nan_feat", qn, vr, lq, lv, 0.2)
    # _quick_check shapes (sdm_loss.py:153-167): D=768, N=16, M=48
    torch.manual_seed(0)
    qq = torch.randn(16, 768); gg = torch.randn(48, 768)
    ql = torch.randint(0, 10, (16,)); gl = torch.randint(0, 10, (48,))
    add("quick_check", qq, gg, ql, gl, 0.2)
    # C5 family
    q5, v5 = _sdm_inputs(512, 512, 512)
    add("p64k8_fp32", q5, v5, pk_labels(64, 8), pk_labels(64, 8), 0.2, store_inputs=False)
    add("p64k8_bf16", q5.bfloat16(), v5.bfloat16(), pk_labels(64, 8), pk_labels(64, 8), 0.2, store_inputs=False)
    q2, v2 = _sdm_inputs(8, 8, 512)
    add("p4k2_bf16", q2.bfloat16(), v2.bfloat16(), pk_labels(4, 2), pk_labels(4, 2), 0.2)
    flat = {}
    for name, d in cases.items():
        for k, val in d.items():
            flat["%s/%s" % (name, k)] = val
    np.savez_compressed(os.path.join(GOLDEN, "sdm_cases.npz"), **flat)


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available; golden fixtures can only be regenerated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(min(8, torch.get_num_threads()))
    make_retrieval()
    make_sdm()
