"""CPU restatement of the reference retrieval + evaluation step.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/tools/eval_mm_protocol.py line by line:
  l2n                    :46-48
  cosine_sim             :50-53
  extract_query_feat     :328-365   (effective path: weighted sum + l2n)
  rank_and_metrics       :389-469
  export ranking core    :617-625

Two forms are provided:
  * `rank_and_metrics_loop`   -- per-query loop, argsort + AP walk exactly as the reference does;
  * `rank_and_metrics_counting` -- vectorised "count items ranked above each positive" form, the
    fast oracle for sizes where the Python loop is infeasible (equality with the loop form and with
    the unmodified reference is asserted in tests/test_oracle_cpu.py).

Pinned against the unmodified reference by oracle/make_golden.py -> tests/golden/retrieval_*.npz.
The arithmetic is torch CPU fp32 (F.normalize / mm / argsort), i.e. the reference's own.
"""
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

MASKED = -1e9  # eval_mm_protocol.py:422


def l2n(x: torch.Tensor) -> torch.Tensor:
    # eval_mm_protocol.py:46-48
    return F.normalize(x, dim=-1)


def cosine_sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # eval_mm_protocol.py:50-53
    return a @ b.T


def fuse_queries(query_raw: torch.Tensor, mod_id: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Batched extract_query_feat (eval_mm_protocol.py:328-365) on pre-extracted features.

    query_raw [Q,k,D] fp32, mod_id [Q,k] (index into weights), weights [n_mod] fp32.
    k == 1: fuse_features_if_any returns the single feature -> l2n(l2n(f))  (:209-210, :359)
    k >= 2: l2n( sum_m w_m * l2n(f_m) )  (:362-365), summed left to right like torch.sum over k<=4.
    """
    Q, k, D = query_raw.shape
    f = l2n(query_raw.float())                       # :353
    if k == 1:
        return l2n(f[:, 0, :])                       # :359
    w = weights[mod_id.long()]                       # :362
    out = torch.empty(Q, D, dtype=torch.float32)
    for s in range(0, Q, 4096):                      # per-row sum over k like (stack*w).sum(0)
        blk = f[s:s + 4096] * w[s:s + 4096, :, None]
        acc = blk[:, 0, :].clone()
        for j in range(1, k):
            acc = acc + blk[:, j, :]
        out[s:s + 4096] = l2n(acc)                   # :365
    return out


def _ap_from_ranks(pos_ranks_sorted: np.ndarray) -> float:
    # eval_mm_protocol.py:444-455 : prec_sum += hit / rank_idx ; AP = prec_sum / num_pos
    prec = 0.0
    for hit, r in enumerate(pos_ranks_sorted.tolist(), start=1):
        prec += hit / r
    return prec / len(pos_ranks_sorted)


def rank_and_metrics_loop(q_feat: torch.Tensor, g_feat: torch.Tensor, q_pid: torch.Tensor,
                          g_pid: torch.Tensor, excl: Optional[torch.Tensor] = None,
                          topk: int = 10, return_per_query: bool = False) -> Dict[str, float]:
    """Per-query restatement of rank_and_metrics (eval_mm_protocol.py:396-469).

    q_feat [Q,D] fused+normalised, g_feat [G,D] normalised, excl [Q,E] gallery indices (-1 pad).
    """
    APs, h1, h5, h10 = [], [], [], []
    top_idx = np.full((q_feat.shape[0], topk), -1, dtype=np.int64)
    valid = np.zeros(q_feat.shape[0], dtype=bool)
    ap_all = np.zeros(q_feat.shape[0])
    for qi in range(q_feat.shape[0]):
        sims = cosine_sim(q_feat[qi:qi + 1], g_feat).squeeze(0)          # :401-402
        mask = torch.ones_like(sims, dtype=torch.bool)                   # :405
        if excl is not None:
            e = excl[qi]
            e = e[e >= 0].long()
            mask[e] = False                                              # :416-418
        sims_masked = sims.clone()
        sims_masked[~mask] = MASKED                                      # :421-422
        ranks = torch.argsort(sims_masked, descending=True)              # :423
        top_idx[qi, :min(topk, ranks.numel())] = ranks[:topk].numpy()
        is_pos = (g_pid == q_pid[qi]) & mask                             # :427
        pos = torch.nonzero(is_pos).flatten()
        if pos.numel() == 0:                                             # :430-432
            continue
        valid[qi] = True
        inv = torch.empty_like(ranks)
        inv[ranks] = torch.arange(1, ranks.numel() + 1)
        pr = np.sort(inv[pos].numpy())
        h1.append(int(pr[0] <= 1)); h5.append(int(pr[0] <= 5)); h10.append(int(pr[0] <= 10))  # :435-441
        ap = _ap_from_ranks(pr)
        APs.append(ap); ap_all[qi] = ap
    out = {
        "mAP": float(np.mean(APs)) if APs else 0.0,                      # :458-461
        "R@1": float(np.mean(h1)) if h1 else 0.0,
        "R@5": float(np.mean(h5)) if h5 else 0.0,
        "R@10": float(np.mean(h10)) if h10 else 0.0,
        "num_queries": len(APs),
    }
    if return_per_query:
        out["_top_idx"] = top_idx
        out["_valid"] = valid
        out["_ap"] = ap_all
    return out


def rank_and_metrics_counting(q_feat: torch.Tensor, g_feat: torch.Tensor, q_pid: torch.Tensor,
                              g_pid: torch.Tensor, excl: Optional[torch.Tensor] = None,
                              topk: int = 10, block: int = 256, return_per_query: bool = False):
    """Vectorised form: rank_j = 1 + #{valid g : s_g > s_pos_j} + #{earlier positives tied}.

    Equal to the loop form whenever no non-positive ties a positive bit-for-bit (argsort's tie
    order is unspecified in the reference, eval_mm_protocol.py:423).  AP in float64 like the
    reference's Python floats.
    """
    Q = q_feat.shape[0]
    ap = np.zeros(Q); valid = np.zeros(Q, dtype=bool); first = np.zeros(Q, dtype=np.int64)
    top_idx = np.full((Q, topk), -1, dtype=np.int64)
    top_val = np.full((Q, topk), MASKED, dtype=np.float32)
    for s in range(0, Q, block):
        e = min(Q, s + block)
        S = cosine_sim(q_feat[s:e], g_feat)                               # [b,G]
        if excl is not None:
            ex = excl[s:e].long()
            rows = torch.arange(e - s)[:, None].expand_as(ex)
            ok = ex >= 0
            S[rows[ok], ex[ok]] = MASKED
        tv, ti = torch.topk(S, min(topk, S.shape[1]), dim=1)
        top_idx[s:e, :ti.shape[1]] = ti.numpy(); top_val[s:e, :tv.shape[1]] = tv.numpy()
        is_pos = (g_pid[None, :] == q_pid[s:e, None]) & (S > MASKED / 2)
        for r in range(e - s):
            ps = S[r][is_pos[r]]
            if ps.numel() == 0:
                continue
            ps, _ = torch.sort(ps, descending=True)
            above = (S[r][None, :] > ps[:, None]).sum(dim=1).numpy()       # strictly greater
            # positives strictly above are already inside `above`; ties among positives are
            # ordered by their sorted position
            ranks = np.empty(len(above), dtype=np.int64)
            psn = ps.numpy()
            for j in range(len(above)):
                tied_before = int(np.sum(psn[:j] == psn[j]))
                ranks[j] = 1 + above[j] + tied_before
            valid[s + r] = True
            first[s + r] = ranks[0]
            ap[s + r] = _ap_from_ranks(ranks)
    v = valid
    out = {
        "mAP": float(np.mean(ap[v])) if v.any() else 0.0,
        "R@1": float(np.mean(first[v] <= 1)) if v.any() else 0.0,
        "R@5": float(np.mean(first[v] <= 5)) if v.any() else 0.0,
        "R@10": float(np.mean(first[v] <= 10)) if v.any() else 0.0,
        "num_queries": int(v.sum()),
    }
    if return_per_query:
        out["_top_idx"] = top_idx; out["_top_val"] = top_val
        out["_valid"] = valid; out["_ap"] = ap; out["_first"] = first
    return out


def submission_ranking(q_feat: torch.Tensor, g_feat: torch.Tensor, top_k: int = 100) -> np.ndarray:
    """Ranking core of export_submission_csv (eval_mm_protocol.py:617-625): NO mask."""
    out = np.empty((q_feat.shape[0], min(top_k, g_feat.shape[0])), dtype=np.int64)
    for qi in range(q_feat.shape[0]):
        sims = cosine_sim(q_feat[qi:qi + 1], g_feat).squeeze(0)
        out[qi] = torch.argsort(sims, descending=True)[:top_k].numpy()
    return out


# ---------------------------------------------------------------------------------------------
# Train-time evaluator: restatement of /root/reference/train.py:101-138 and
# :451-479.  Pinned against the UNMODIFIED reference functions, whose source is extracted from
# train.py with `ast` and executed as-is (oracle.ref_loader.load_reference_train_eval), in
# tests/test_oracle_cpu.py.
# ---------------------------------------------------------------------------------------------
def compute_map_oracle(qf, gf, ql, gl, k=100):
    qf = torch.nn.functional.normalize(qf.float(), p=2, dim=1)          # train.py:105-109
    gf = torch.nn.functional.normalize(gf.float(), p=2, dim=1)
    sim = qf @ gf.t()
    aps = []
    for i in range(sim.shape[0]):                                        # :113-124
        idx = torch.sort(sim[i], descending=True)[1]
        matches = (gl[idx[:k]] == ql[i]).float()
        if matches.sum() > 0:
            prec = torch.cumsum(matches, 0) / torch.arange(1, matches.numel() + 1, dtype=matches.dtype)
            aps.append(prec[matches.bool()].mean().item())
    return float(np.mean(aps)) if aps else 0.0                           # :126


def compute_cmc_oracle(qf, gf, ql, gl, k=10):
    qf = torch.nn.functional.normalize(qf, p=2, dim=1)                   # train.py:130-131
    gf = torch.nn.functional.normalize(gf, p=2, dim=1)
    sim = qf @ gf.t()
    correct = 0
    for i in range(sim.shape[0]):                                        # :134-137
        idx = torch.sort(sim[i], descending=True)[1]
        correct += bool((gl[idx[:k]] == ql[i]).any())
    return correct / sim.shape[0] if sim.shape[0] > 0 else 0.0


def reid_map_oracle(sim, q_ids, g_ids):
    """train.py:451-479 `_reid_map`."""
    Nq = sim.shape[0]
    mAP, top1 = 0.0, 0.0
    ar = torch.arange(sim.shape[1], dtype=torch.float32) + 1.0
    for i in range(Nq):
        order = torch.argsort(sim[i], descending=True)
        matches = (g_ids[order] == q_ids[i]).to(sim.dtype)
        rel = matches.sum().item()
        if rel == 0:
            continue
        prec = torch.cumsum(matches, 0) / ar
        mAP += (torch.sum(prec * matches) / rel).item()
        top1 += matches[0].item()
    valid = max(1, (q_ids.unsqueeze(1) == g_ids.unsqueeze(0)).any(dim=1).sum().item())
    return mAP / valid, top1 / Nq
