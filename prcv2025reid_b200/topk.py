"""Top-k ranking without a mask: the ranking core of export_submission_csv
(tools/eval_mm_protocol.py:617-625, `argsort(sims, descending=True)[:top_k]`).

The fused retrieval path yields an exactly ordered top-32 per query (REID_RTOP).  Deeper lists are
produced in passes: the rows already ranked are handed back as the query's exclusion list (the same
mechanism as the same-image mask, :421-422), so pass p returns ranks 32p+1 .. 32p+32.  Every pass is
exact (fp32 re-score + completeness check + exact fallback), hence so is the concatenation.
"""
import torch

from . import _cabi, engine


def topk_ranking(q_f32: torch.Tensor, gallery: torch.Tensor, top_k: int = 100, mode: str = "fused") -> torch.Tensor:
    """q_f32 [Q,D] fused+normalised queries (cuda), gallery [G,D] features (cuda) -> [Q, top_k] int32
    gallery indices, best first (-1 pad when G < top_k)."""
    dev = q_f32.device
    G = gallery.shape[0]
    Q = q_f32.shape[0]
    # distinct gallery ids and absent query ids: no positives, only the candidate / re-score machinery runs
    shard = engine.prepare_gallery(gallery, torch.arange(G, device=dev, dtype=torch.int64))
    q16 = q_f32.to(torch.float16)
    q_pid = torch.full((Q,), -1, dtype=torch.int64, device=dev)
    out = torch.full((Q, top_k), -1, dtype=torch.int32, device=dev)
    got = 0
    R = _cabi.RTOP
    while got < min(top_k, G):
        excl = out[:, :got].contiguous() if got else None
        res = engine.retrieve(shard, q_f32, q16, q_pid, excl, topk=min(R, top_k - got, G - got), mode=mode)
        n = res.top_idx.shape[1]
        out[:, got:got + n] = res.top_idx
        got += n
    return out
