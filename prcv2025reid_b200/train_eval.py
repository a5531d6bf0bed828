"""Drop-ins for the reference's train-time evaluator (train.py), SURVEY.md section 8f row N2.

    compute_map(query_features, gallery_features, query_labels, gallery_labels, k=100)   train.py:101-126
    compute_cmc(query_features, gallery_features, query_labels, gallery_labels, k=10)    train.py:128-138
    reid_map(q_feat, g_feat, q_ids, g_ids) -> (mAP, top1)     evaluate_one_query step 3 + _reid_map, train.py:451-479, 498-500

Same kernels as the evaluation protocol (normalise K1, fused tcgen05 similarity / ranking kernel, fp32 re-score),
different reductions.  Inputs may be CPU or CUDA tensors; the arithmetic always runs in libreid_b200.so on the
current CUDA device (there is no CPU fallback), results are Python floats like the reference's.
"""
import torch

from . import _cabi, engine, topk
from ._cabi import check, ptr, stream_ptr


def _dev(t):
    if not torch.cuda.is_available():
        raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return t if t.is_cuda else t.to(torch.device("cuda", torch.cuda.current_device()))


def _topk_matches(query_features, gallery_features, query_labels, gallery_labels, k):
    q = _dev(query_features).float()
    g = _dev(gallery_features).float()
    ql = _dev(query_labels).to(torch.int64).contiguous()
    gl = _dev(gallery_labels).to(torch.int64).contiguous()
    qn, _ = engine.l2norm_rows(q)                                  # train.py:108-109 / :130-131
    gn, _ = engine.l2norm_rows(g)
    kk = min(int(k), g.shape[0])
    idx = topk.topk_ranking(qn, gn, kk).contiguous()               # exact `torch.sort(scores, descending=True)[:k]` (:115, :134)
    Q = q.shape[0]
    ap = torch.empty(Q, dtype=torch.float32, device=q.device)
    hit = torch.empty(Q, dtype=torch.int32, device=q.device)
    check(_cabi.lib().reid_topk_label_metrics(ptr(idx), ptr(ql), ptr(gl), Q, idx.shape[1], kk, ptr(ap), ptr(hit),
                                              stream_ptr()), "reid_topk_label_metrics")
    return ap, hit


def compute_map(query_features, gallery_features, query_labels, gallery_labels, k=100):
    """mAP@k (train.py:101-126): queries without a match inside the top-k are skipped; 0.0 when none is left."""
    if query_features.shape[0] == 0:
        return 0.0
    ap, hit = _topk_matches(query_features, gallery_features, query_labels, gallery_labels, k)
    ap = ap.cpu().numpy().astype("float64"); hit = hit.cpu().numpy() > 0
    return float(ap[hit].mean()) if hit.any() else 0.0


def compute_cmc(query_features, gallery_features, query_labels, gallery_labels, k=10):
    """CMC@k (train.py:128-138): fraction of ALL queries with a match inside the top-k."""
    if query_features.shape[0] == 0:
        return 0.0
    _, hit = _topk_matches(query_features, gallery_features, query_labels, gallery_labels, k)
    return float(hit.sum().item()) / query_features.shape[0]


def reid_map(q_feat, g_feat, q_ids, g_ids):
    """(mAP, top1) of `_reid_map(q_feat @ g_feat.T, q_ids, g_ids)` (train.py:451-479, called at :499-500 on
    L2-normalised features): full-ranking AP averaged over the queries that have a match, top-1 over ALL queries."""
    Nq = q_feat.shape[0]
    if Nq == 0:
        return 0.0, 0.0
    q = _dev(q_feat).float().contiguous()
    g = _dev(g_feat).float()
    shard = engine.prepare_gallery(g, _dev(g_ids).to(torch.int64))     # (re-normalising unit rows is the identity up to 1 ulp)
    res = engine.retrieve(shard, q, q.to(torch.float16), _dev(q_ids).to(torch.int64), None, topk=1, mode="fused")
    m = res.metrics
    return float(m["mAP"]), float(m["R@1"] * m["num_queries"] / Nq)
