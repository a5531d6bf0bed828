"""Drop-ins for the reference's train-time evaluator (train.py), SURVEY.md section 8f row N2.

    compute_map(query_features, gallery_features, query_labels, gallery_labels, k=100)   train.py:101-126
    compute_cmc(query_features, gallery_features, query_labels, gallery_labels, k=10)    train.py:128-138
    reid_map(q_feat, g_feat, q_ids, g_ids) -> (mAP, top1)     evaluate_one_query step 3 + _reid_map, train.py:451-479, 498-500
    evaluate_one_query_features(g_feat, g_id, q_feat, q_id)  evaluate_one_query after feature extraction, train.py:481-501
    validate_competition_style_features(...)                 validate_competition_style after feature extraction, train.py:503-631

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


def install_gallery(g_feat, g_ids) -> engine.GalleryShard:
    """The gallery side of `_reid_map` installed once on the device (normalised rows, fp16 operand copy, identity
    index): pass it as `shard=` when several query sets are evaluated against one gallery, like the `cache` dict of
    evaluate_one_query (train.py:488-493) does for the gallery features."""
    return engine.prepare_gallery(_dev(g_feat).float(), _dev(g_ids).to(torch.int64))


def reid_map(q_feat, g_feat, q_ids, g_ids, shard=None):
    """(mAP, top1) of `_reid_map(q_feat @ g_feat.T, q_ids, g_ids)` (train.py:451-479, called at :499-500 on
    L2-normalised features): full-ranking AP averaged over the queries that have a match, top-1 over ALL queries."""
    Nq = q_feat.shape[0]
    if Nq == 0:
        return 0.0, 0.0
    q = _dev(q_feat).float().contiguous()
    if shard is None:
        shard = install_gallery(g_feat, g_ids)                         # (re-normalising unit rows is the identity up to 1 ulp)
    res = engine.retrieve(shard, q, q.to(torch.float16), _dev(q_ids).to(torch.int64), None, topk=1, mode="fused")
    m = res.metrics
    return float(m["mAP"]), float(m["R@1"] * m["num_queries"] / Nq)


# ---------------------------------------------------------------------------------------------
# The callers of `_reid_map`: evaluate_one_query / validate_competition_style (train.py:481-631) after the
# feature-extraction step (`_extract_feats_and_ids`, :428-449, needs the model and the data loaders: out of scope).
# A "query set" is the pair (q_feat [Nq, D] L2-normalised, q_id [Nq]) such a loader yields.
# ---------------------------------------------------------------------------------------------
DEFAULT_INCLUDE = ["single/nir", "single/sk", "single/cp", "single/text", "quad/nir+sk+cp+text"]      # train.py:508


def flatten_query_sets(obj, prefix=""):
    """`_flatten_loaders` (train.py:402-424) for feature sets: nested dicts / lists -> [(name, (q_feat, q_id)), ...]
    with names like 'single/nir' or 'quad/0'; a leaf is a (features, ids) pair of tensors."""
    if isinstance(obj, (tuple, list)) and len(obj) == 2 and all(torch.is_tensor(t) for t in obj):
        yield (prefix.rstrip("/") or "root", (obj[0], obj[1]))
    elif isinstance(obj, dict):
        for k, v in obj.items():
            yield from flatten_query_sets(v, "%s%s/" % (prefix, k))
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            yield from flatten_query_sets(v, "%s%d/" % (prefix, i))
    else:
        raise TypeError("Unsupported query_sets node type: %s at %r" % (type(obj), prefix))


def evaluate_one_query_features(g_feat, g_id, q_feat, q_id, shard=None):
    """evaluate_one_query (train.py:481-501) from extracted features -> {"mAP": float, "Top1": float}."""
    m, t1 = reid_map(q_feat, g_feat, q_id, g_id, shard=shard)                                  # :498-500
    return {"mAP": float(m), "Top1": float(t1)}


def _norm_name(name: str) -> str:
    return name.replace("cpencil", "cp").replace("sketch", "sk")                               # train.py:512-513


def validate_competition_style_features(g_feat, g_id, query_sets, sample_ratio=1.0, cfg=None, epoch=None):
    """validate_competition_style (train.py:503-631) from extracted features.

    query_sets: nested dict / list of (q_feat, q_id) pairs (what `query_loaders` yields after feature extraction).
    Only the sets whose normalised name matches cfg.eval_include_patterns (default DEFAULT_INCLUDE) are evaluated
    (:508-514); with 0 < sample_ratio < 1 each set is cut to `torch.randperm(n)[:int(n * ratio)]` (:554-558, same
    call on the global torch RNG).  Returns {'map_single', 'map_quad', 'map_avg2', 'detail', 'cmc1', 'cmc5',
    'cmc10'}: map_single = sum of the four single-modality mAPs / 4 (a missing one counts as 0, :594-595), map_avg2
    = (map_single + map_quad) / 2, cmc1 = cmc5 = cmc10 = top-1 hit of the FIRST query of the first evaluated set
    (the reference's "simplified CMC", :620-621).  The gallery is installed on the device once; no disk cache and
    no printing."""
    import fnmatch
    include = getattr(cfg, "eval_include_patterns", DEFAULT_INCLUDE)
    pairs = [(n, qs) for n, qs in flatten_query_sets(query_sets)
             if any(fnmatch.fnmatch(_norm_name(n), pat) for pat in include)]                  # :514
    shard = install_gallery(g_feat, g_id) if pairs and g_feat.shape[0] else None
    detail, first = {}, None
    for name, (qf, qi) in pairs:
        if 0.0 < sample_ratio < 1.0:                                                          # :554-558
            idx = torch.randperm(qf.shape[0])[:int(qf.shape[0] * sample_ratio)]
            qf, qi = qf[idx.to(qf.device)], qi[idx.to(qi.device)]
        detail[name] = evaluate_one_query_features(g_feat, g_id, qf, qi, shard=shard)         # :570
        if first is None and qf.shape[0] > 0:
            first = (qf[:1], qi[:1])
    singles = [float(detail.get(k, {}).get("mAP", 0.0)) for k in ("single/nir", "single/sk", "single/cp", "single/text")]
    map_single = sum(singles) / max(1, len([x for x in singles if x == x]))                   # :594-595
    map_quad = float(detail.get("quad/nir+sk+cp+text", {}).get("mAP", 0.0))                   # :598
    cmc1 = 0.0
    if first is not None:                                                                     # :614-621
        cmc1 = reid_map(first[0], g_feat, first[1], g_id, shard=shard)[1]
    return {"map_single": map_single, "map_quad": map_quad, "map_avg2": (map_single + map_quad) / 2.0,
            "detail": detail, "cmc1": cmc1, "cmc5": cmc1, "cmc10": cmc1}
