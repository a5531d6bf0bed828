"""prcv2025reid_b200 -- B200-native (sm_100a) kernels for the hot path of LingmaFuture/PRCV2025REID:
multimodal query->gallery retrieval + evaluation (tools/eval_mm_protocol.py) and the SDM loss
(models/sdm_loss.py), behind the reference's own Python signatures.

    from prcv2025reid_b200.eval_mm_protocol import l2n, cosine_sim, extract_query_feat, rank_and_metrics
    from prcv2025reid_b200.sdm_loss import sdm_loss_stable, sdm_loss_pairs
    from prcv2025reid_b200 import engine           # tensor-level API (gallery shards, batched queries)
"""
__all__ = ["engine", "eval_mm_protocol", "sdm_loss", "synth"]
