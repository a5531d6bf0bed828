"""CPU model of the COUNTING RULE of the fused retrieval kernel (csrc/retrieve_fused.cu).  TEST INFRASTRUCTURE ONLY.

The reference computes AP from a full argsort (tools/eval_mm_protocol.py:423, :444-455).  The fused kernel replaces
the sort by counts -- rank_j = 1 + #{valid non-positive rows scoring above positive j} + j -- and, for a query's DEEP
positives only, estimates that count from a fixed 1/32 row sample (DESIGN.md section 4.1 "Deep-rank sampling").  This
file restates that rule in numpy so that its error against the exact oracle can be measured without a GPU, and so
that a change of its parameters (calibration rows, sample width, exact-count budget) can be assessed before it is
built.  It follows the kernel, not the reference:

  scores             fp32 accumulation of fp16-rounded operands                       retrieve_fused.cu kernel comment (kind::f16 MMA)
  thresholds t_j     the positives' EXACT fp32 scores, sorted descending              reid_pos_scores / reid_pos_sort
  calibration        strided sample of G_local/32 rows (2048 .. 8192), every threshold exact         reid_retrieve_fused host code (`sample_deep`)
  counting classes   calib_split_kernel: threshold j is counted on every row while    retrieve_fused.cu calib_split_kernel
                     the estimated rank INSIDE THE SHARD of thresholds 0..j stays
                     <= limit1 = max(32 * W1 / n_shards, 8 * scale), on the level-1
                     row sample while it stays <= limit2 = max(32 * W2 / n_shards,
                     limit1), else on the level-2 sample; scale = G_local / CALIB_ROWS
  main pass          ordinary rows see thresholds [0, n_exact); rows with             epi_drain32
                     local_row % W1 == 5 see [0, n_l1) and weigh W1 in buckets
                     [n_exact, n_l1); rows with local_row % W2 == 5 see all of them and
                     weigh W2 in buckets [n_l1, n_pos); positives of the query and
                     masked rows are skipped; rows >= G_local are no rows
  pos_above[j]       weighted prefix sum of the bucket histogram                      hist_to_above_kernel

  re-scoring         the RTOP best local rows by fp16 score are re-scored in fp32; cut = KLIST-th best     rescore_topk_kernel, rank.cu
                     fp16 score, bound = cut + eps; a positive with t_j > bound gets its EXACT local
                     count from the re-scored rows, the others keep max(fused count, that lower bound);
                     flags: top-k undecidable (k-th exact score < bound), CMC@10 undecidable
                     (best positive not above the bound and fewer than 10 re-scored rows above it)

Not modelled: the early retirement of deep thresholds inside the calibration pre-pass (a threshold whose RUNNING rank estimate
has passed the level-2 limit on >= 32 sample rows is classed level-2 at once; the model classes every threshold on the counts
of the whole sample -- the two can differ for thresholds within ~18 % of the level-1 / level-2 boundary, where either class
moves AP by < 1e-5), the seed of the candidate threshold taken from the calibration counts, the running candidate thresholds (the model takes the candidate set as complete down to the KLIST-th best
score, which is what the kernel guarantees) and the exact re-run of flagged queries (engine.retrieve).
"""
from typing import Dict, Optional

import numpy as np
import torch

SAMPLE_W = 32        # W1: level-1 row sample
SAMPLE_W2 = 1024     # W2: level-2 row sample
CALIB_ROWS = 8192    # REID_CALIB_ROWS (upper bound of the calibration sample)
CALIB_MIN = 2048     # lower bound; shards below 16 * CALIB_MIN rows are never sampled
SAMPLE_PHASE = 5     # rows with (row % W) == 5
KLIST = 32           # REID_KLIST
RTOP = 32            # REID_RTOP
EPS_FP16 = 2.0 ** -10 + 2.0 ** -13    # engine.EPS_FP16


def positive_thresholds(q_f32: torch.Tensor, g_f32: torch.Tensor, q_pid: torch.Tensor, g_pid: torch.Tensor,
                        excl: Optional[torch.Tensor] = None):
    """reid_pos_scores + reid_pos_sort over the WHOLE gallery (after the all-reduce MAX of sharding.exchange_pos_scores):
    -> (thr [Q, Pmax] fp32 exact scores of the unmasked positives, sorted descending, -inf pad; n_pos [Q])."""
    Q, G = q_f32.shape[0], g_f32.shape[0]
    S32 = (q_f32 @ g_f32.T).numpy()
    gp, qp = g_pid.numpy(), q_pid.numpy()
    is_pos = gp[None, :] == qp[:, None]
    if excl is not None:
        ex = excl.numpy()
        for q in range(Q):
            is_pos[q, ex[q][ex[q] >= 0]] = False
    n_pos = is_pos.sum(1)
    thr = np.full((Q, max(1, int(n_pos.max()))), -np.inf, dtype=np.float32)
    for q in range(Q):
        thr[q, :n_pos[q]] = np.sort(S32[q, is_pos[q]])[::-1]
    return thr, n_pos


def fused_counts(q_f32: torch.Tensor, g_f32: torch.Tensor, q_pid: torch.Tensor, g_pid: torch.Tensor,
                 excl: Optional[torch.Tensor] = None, n_shards: int = 1, exact_all: bool = False,
                 sample_w: int = SAMPLE_W, sample_w2: int = SAMPLE_W2, calib_rows: int = CALIB_ROWS,
                 thr: Optional[np.ndarray] = None, n_pos: Optional[np.ndarray] = None, g_offset: int = 0) -> Dict[str, np.ndarray]:
    """One rank's reid_retrieve_fused -> {"pos_above" [Q, Pmax] int64 (modelled LOCAL counts, additive over ranks),
    "n_pos" [Q], "n_exact" [Q], "n_l1" [Q], "thr" [Q, Pmax], "hits" [Q, 3] (epilogue hit volume by row class)}.

    q_f32 [Q, D] fused + normalised queries; g_f32 [G_local, D] / g_pid [G_local] this rank's normalised gallery rows,
    which are rows [g_offset, g_offset + G_local) of the whole gallery; excl [Q, E] GLOBAL gallery rows masked per query
    (-1 pad); thr / n_pos: the gallery-wide positive thresholds (positive_thresholds; default: computed from this shard,
    i.e. a one-rank job).  exact_all = REID_FUSED_EXACT_COUNTS (no sampling).  sample_w2 = 0: one sampling level only
    (the round-1 rule).  The work decomposition (query blocks, gallery chunks) does not change the counts: a threshold's
    class is the same in every chunk of a shard, chunk starts are multiples of W2 and counts are additive."""
    Q, G = q_f32.shape[0], g_f32.shape[0]
    S16 = (q_f32.half().float() @ g_f32.half().float().T).numpy()          # what the tensor cores score with
    gp, qp = g_pid.numpy(), q_pid.numpy()
    masked = np.zeros((Q, G), dtype=bool)
    if excl is not None:
        ex = excl.numpy().astype(np.int64) - g_offset
        for q in range(Q):
            e = ex[q][(excl.numpy()[q] >= 0) & (ex[q] >= 0) & (ex[q] < G)]
            masked[q, e] = True
    if thr is None:
        thr, n_pos = positive_thresholds(q_f32, g_f32, q_pid, g_pid,
                                         None if excl is None else torch.from_numpy(np.where(excl.numpy() >= 0, excl.numpy() - g_offset, -1)))
    Pmax = thr.shape[1]
    sample_deep = (G >= 16 * CALIB_MIN) and not exact_all
    calib_rows = max(CALIB_MIN, min(calib_rows, G // 32 // 256 * 256))    # 1/32 of the shard in whole tiles
    pos_above = np.zeros((Q, Pmax), dtype=np.int64)
    n_exact = np.zeros(Q, dtype=np.int64)
    n_l1 = np.zeros(Q, dtype=np.int64)
    hits = np.zeros((Q, 3), dtype=np.int64)                                 # epilogue hit volume: [ordinary, level-1, level-2 rows]
    countable = ~masked & (gp[None, :] != qp[:, None])                      # neither masked nor a positive of the query
    if sample_deep:
        stride = G // calib_rows
        cal_rows = np.arange(calib_rows) * stride
        scale = np.float32(G) / np.float32(calib_rows)
        limit1 = max(32.0 * sample_w / max(1, n_shards), 8.0 * float(scale))
        limit2 = max(32.0 * sample_w2 / max(1, n_shards), limit1) if sample_w2 else np.inf
    rows = np.arange(G)
    s1 = (rows % sample_w) == SAMPLE_PHASE
    s2 = ((rows % sample_w2) == SAMPLE_PHASE) if sample_w2 else np.zeros(G, dtype=bool)
    for q in range(Q):
        npq = int(n_pos[q])
        if npq == 0:
            continue
        t = thr[q, :npq]
        ne = n1 = npq
        if sample_deep:
            # calibration pass: bucket histogram of the strided sample, every threshold exact
            cs = S16[q, cal_rows][countable[q, cal_rows]]
            acc = np.array([(cs > tj).sum() for tj in t])                  # prefix sums over buckets <= j
            est = acc.astype(np.float32) * np.float32(scale)
            ok1, ok2 = est <= np.float32(limit1), est <= np.float32(limit2)
            ne = npq if ok1.all() else int(np.argmin(ok1))                # the first threshold over the limit closes the prefix
            n1 = max(ne, npq if ok2.all() else int(np.argmin(ok2)))
        n_exact[q], n_l1[q] = ne, n1
        s, use = S16[q], countable[q]
        # rows the epilogue has to classify (counting side only)
        hits[q, 0] = int((s > t[ne - 1]).sum()) if ne > 0 else 0
        hits[q, 1] = int((s1 & (s > t[n1 - 1])).sum()) if n1 > ne else 0
        hits[q, 2] = int((s2 & (s > t[npq - 1])).sum()) if npq > n1 else 0
        # bucket b of a row = #{j : t_j >= s}: the row lies above thresholds j >= b
        bucket = np.searchsorted(-t, -s, side="right")                     # t descending; t_j >= s  <=>  -t_j <= -s
        hist = np.zeros(npq + 1, dtype=np.int64)
        cls_w = np.where(np.arange(npq + 1) < ne, 1, np.where(np.arange(npq + 1) < n1, sample_w, sample_w2 or sample_w))
        sees = np.where(s2, npq, np.where(s1, n1, ne))                     # thresholds [0, sees) are visible to the row
        counted = use & (bucket < sees)
        np.add.at(hist, bucket[counted], 1)
        pos_above[q, :npq] = np.cumsum(hist[:npq] * cls_w[:npq])
    return {"pos_above": pos_above, "n_pos": n_pos, "n_exact": n_exact, "n_l1": n_l1, "thr": thr, "hits": hits}


def rescore_stage(q_f32: torch.Tensor, g_f32: torch.Tensor, q_pid: torch.Tensor, g_pid: torch.Tensor, counts: Dict[str, np.ndarray],
                  excl: Optional[torch.Tensor] = None, g_offset: int = 0, topk: int = 10, eps: float = EPS_FP16,
                  klist: int = KLIST, rtop: int = RTOP):
    """One rank's reid_rescore_topk applied to the modelled counts of fused_counts (same shard).
    -> {"pos_above" (corrected copy), "flag" [Q] (bit1 top-k undecidable, bit2 CMC@10 undecidable),
        "top_idx" [Q, RTOP] global gallery rows by exact score (-1 pad), "top_score" [Q, RTOP]}."""
    Q, G = q_f32.shape[0], g_f32.shape[0]
    S16 = (q_f32.half().float() @ g_f32.half().float().T).numpy()
    S32 = (q_f32 @ g_f32.T).numpy()
    gp, qp = g_pid.numpy(), q_pid.numpy()
    pos_above = counts["pos_above"].copy()
    thr, n_pos = counts["thr"], counts["n_pos"]
    flag = np.zeros(Q, dtype=np.int32)
    top_idx = np.full((Q, rtop), -1, dtype=np.int64)
    top_score = np.full((Q, rtop), -np.inf, dtype=np.float32)
    for q in range(Q):
        s16 = S16[q].copy()
        if excl is not None:
            e = excl.numpy()[q].astype(np.int64)
            e = e[e >= 0] - g_offset
            s16[e[(e >= 0) & (e < G)]] = -np.inf                      # masked rows are never candidates
        order = np.argsort(-s16, kind="stable")
        total = int(np.isfinite(s16).sum())
        R = min(total, rtop)
        cand = order[:R]
        cut = s16[order[klist - 1]] if total >= klist else -np.inf
        bound = cut + np.float32(eps)
        ex = S32[q, cand]
        o2 = np.lexsort((cand, -ex))                                   # score desc, index asc
        cand, ex = cand[o2], ex[o2]
        top_idx[q, :R] = cand + g_offset
        top_score[q, :R] = ex
        neg = gp[cand] != qp[q]
        for j in range(int(n_pos[q])):
            t = thr[q, j]
            lb = int((neg & (ex > t)).sum())
            if t > bound or cut == -np.inf:
                pos_above[q, j] = lb                                   # exact local count
            else:
                pos_above[q, j] = max(pos_above[q, j], lb)
                if j == 0 and lb < 10:
                    flag[q] |= 4
        if cut > -np.inf and R >= topk and ex[topk - 1] < bound:
            flag[q] |= 2
    return {"pos_above": pos_above, "flag": flag, "top_idx": top_idx, "top_score": top_score}


def metrics_from_counts(pos_above: np.ndarray, n_pos: np.ndarray) -> Dict[str, object]:
    """reid_metrics_reduce (rank.cu metrics_kernel): rank_j = 1 + pos_above[j] + j, AP in float64, queries without
    a positive skipped (eval_mm_protocol.py:430-432, :444-461)."""
    Q = pos_above.shape[0]
    ap = np.full(Q, -1.0)
    first = np.zeros(Q, dtype=np.int64)
    for q in range(Q):
        npq = int(n_pos[q])
        if npq == 0:
            continue
        j = np.arange(npq)
        ap[q] = float(np.sum((j + 1) / (pos_above[q, :npq] + j + 1.0)) / npq)
        first[q] = pos_above[q, 0] + 1
    v = ap >= 0
    return {"mAP": float(ap[v].mean()) if v.any() else 0.0,
            "R@1": float((first[v] <= 1).mean()) if v.any() else 0.0,
            "R@5": float((first[v] <= 5).mean()) if v.any() else 0.0,
            "R@10": float((first[v] <= 10).mean()) if v.any() else 0.0,
            "num_queries": int(v.sum()), "_ap": ap}
