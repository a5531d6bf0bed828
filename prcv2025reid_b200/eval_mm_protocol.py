"""Drop-in for the feature-level functions of the reference's tools/eval_mm_protocol.py.

Same names, argument meaning and error behaviour as the reference (file:line cited per function);
underneath, every tensor operation is a libreid_b200 CUDA kernel.  Inputs may live on the CPU (the
reference keeps `gallery_feats` on the CPU, eval_mm_protocol.py:302,325): they are moved to the
current CUDA device, results are returned where the reference returns them (tensors on the input's
device, metrics as Python floats).  Importing this module without the built library or without a
GPU and calling into it raises -- there is no CPU fallback.
"""
import csv
from typing import Dict, List, Optional

import torch

from . import engine
from .synth import MODALITIES, MOD_ID

ALL_NON_RGB = ["ir", "cpencil", "sketch", "text"]          # eval_mm_protocol.py:35


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def l2n(x: torch.Tensor) -> torch.Tensor:
    """eval_mm_protocol.py:46-48 -- F.normalize(x, dim=-1) (p=2, eps=1e-12)."""
    src = x.device
    shape = x.shape
    out, _ = engine.l2norm_rows(x.reshape(-1, shape[-1]).to(_dev()))
    return out.reshape(shape).to(src)


def cosine_sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """eval_mm_protocol.py:50-53 -- a @ b.T for L2-normalised rows, on the tcgen05 GEMM.

    Operands are rounded to fp16 for the tensor cores (|error| <= 2^-10 for unit rows, typically
    2e-5); the ranking path re-scores in fp32 and does not depend on this function."""
    src = a.device
    dev = _dev()
    a16 = a.reshape(-1, a.shape[-1]).to(dev, torch.float32).to(torch.float16).contiguous()
    b16 = b.reshape(-1, b.shape[-1]).to(dev, torch.float32).to(torch.float16).contiguous()
    return engine.cosine_sim_f16(a16, b16).to(src)


def _encode(extractor, m: str, sample: dict) -> torch.Tensor:
    # eval_mm_protocol.py:341-351
    if m == "ir":
        return extractor.encode_ir(sample["img_path"])
    if m == "cpencil":
        return extractor.encode_cpencil(sample["img_path"])
    if m == "sketch":
        return extractor.encode_sketch(sample["img_path"])
    if m == "text":
        return extractor.encode_text(sample["text"])
    raise ValueError(f"未知模态: {m}")


def _fuse_batch(queries: List[dict], extractor, weight_cfg: Dict[str, float]) -> torch.Tensor:
    """Batched extract_query_feat (eval_mm_protocol.py:328-365) -> [Q, D] fp32 on the CUDA device."""
    dev = _dev()
    raw, mods = [], []
    for q in queries:
        fs, ms = [], []
        for m, sample in q["samples"].items():
            fs.append(_encode(extractor, m, sample).float().view(-1))
            ms.append(m)
        raw.append(fs); mods.append(ms)
    Q = len(queries)
    if Q == 0:
        return torch.empty(0, 0, device=dev)
    D = raw[0][0].numel()
    out = torch.empty(Q, D, dtype=torch.float32, device=dev)
    # per-modality l2n (:353) in one pass, needed as the argument of the extractor's fusion hook
    flat = torch.stack([f for fs in raw for f in fs]).to(dev)
    flat_n, _ = engine.l2norm_rows(flat)
    offs, o = [], 0
    for fs in raw:
        offs.append(o); o += len(fs)
    weighted: Dict[int, List[int]] = {}
    hooked_rows, hooked_idx = [], []
    for qi in range(Q):
        k = len(raw[qi])
        feats_n = [flat_n[offs[qi] + j] for j in range(k)]
        fused = extractor.fuse_features_if_any(feats_n, mods[qi])       # :357
        if fused is not None:
            hooked_rows.append(fused.float().view(-1).to(dev)); hooked_idx.append(qi)   # -> l2n(fused) (:359)
        else:
            weighted.setdefault(k, []).append(qi)
    if hooked_idx:
        hn, _ = engine.l2norm_rows(torch.stack(hooked_rows))
        out[torch.tensor(hooked_idx, device=dev)] = hn
    names = list(MODALITIES) + sorted(set(m for ms in mods for m in ms) - set(MODALITIES))
    ids = {m: i for i, m in enumerate(names)}
    w = torch.tensor([float(weight_cfg.get(m, 1.0)) for m in names], dtype=torch.float32, device=dev)   # :362
    for k, idxs in weighted.items():
        it = torch.tensor(idxs, device=dev)
        rows = torch.stack([flat[offs[qi] + j] for qi in idxs for j in range(k)]).view(len(idxs), k, D)
        mid = torch.tensor([[ids[m] for m in mods[qi]] for qi in idxs], dtype=torch.int32, device=dev)
        if k == 1:      # weighted path with one feature: l2n(w * f) -- keep the weight (no hook result)
            rows = torch.cat([rows, torch.zeros_like(rows)], dim=1)
            mid = torch.cat([mid, torch.full_like(mid, -1)], dim=1)
        q32, _ = engine.fuse_queries(rows, mid, w)
        out[it] = q32
    return out


def extract_query_feat(q: dict, extractor, weight_cfg: Dict[str, float]) -> torch.Tensor:
    """eval_mm_protocol.py:328-365 -- one query -> fused, L2-normalised feature [D]."""
    f = _fuse_batch([q], extractor, weight_cfg)[0]
    probe = next(iter(q["samples"].items()))
    src = _encode(extractor, probe[0], probe[1]).device
    return f.to(src)


def _exclusions(queries, g_imgid, ignore_same_img: bool, dev) -> Optional[torch.Tensor]:
    # eval_mm_protocol.py:408-418: gallery rows whose img_id is one of the query samples' img_ids
    if not ignore_same_img:
        return None
    by_id: Dict[object, List[int]] = {}
    for i, gid in enumerate(g_imgid):
        if gid is not None:
            by_id.setdefault(gid, []).append(i)
    rows, width = [], 0
    for q in queries:
        hit = []
        for s in q["samples"].values():
            iid = s.get("img_id") if "img_id" in s else None
            if iid is not None and iid in by_id:
                hit.extend(by_id[iid])
        hit = sorted(set(hit))
        rows.append(hit); width = max(width, len(hit))
    if width == 0:
        return None
    ex = torch.full((len(queries), width), -1, dtype=torch.int32)
    for i, h in enumerate(rows):
        if h:
            ex[i, :len(h)] = torch.tensor(h, dtype=torch.int32)
    return ex.to(dev)


def rank_and_metrics(queries: List[dict], gallery_feats: torch.Tensor, gallery_meta: List[dict], extractor,
                     weight_cfg: Dict[str, float], ignore_same_img=True, cross_camera=False,
                     mode: str = "fused") -> Dict[str, float]:
    """eval_mm_protocol.py:369-469.  `cross_camera` is accepted and ignored like the reference (:375,392).

    Returns {"mAP","R@1","R@5","R@10","num_queries"}; queries without a positive are skipped (:430-432),
    an empty result gives 0.0 metrics (:458-461)."""
    dev = _dev()
    if len(queries) == 0 or len(gallery_meta) == 0:
        return {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0, "num_queries": 0}
    g_pids = torch.tensor([m["pid"] for m in gallery_meta], dtype=torch.long)                 # :390
    g_imgid = [m.get("img_id", None) for m in gallery_meta]                                   # :391
    shard = engine.prepare_gallery(gallery_feats.to(dev), g_pids.to(dev))
    q32 = _fuse_batch(queries, extractor, weight_cfg)
    q16 = q32.to(torch.float16)
    q_pid = torch.tensor([int(q["pid"]) for q in queries], dtype=torch.long, device=dev)
    excl = _exclusions(queries, g_imgid, ignore_same_img, dev)
    res = engine.retrieve(shard, q32, q16, q_pid, excl, topk=10, mode=mode)
    return res.metrics


def export_submission_csv(queries: List[dict], gallery_feats: torch.Tensor, gallery_meta: List[dict], extractor,
                          weight_cfg: Dict[str, float], output_path: str, top_k: int = 100):
    """eval_mm_protocol.py:595-649 -- ranking WITHOUT the same-image mask (:622-625), CSV columns
    `query_key,ranked_gallery_ids` (:634-648)."""
    from . import topk as topk_mod
    dev = _dev()
    gal = gallery_feats.to(dev, torch.float32).contiguous()
    q32 = _fuse_batch(queries, extractor, weight_cfg)
    ranks = topk_mod.topk_ranking(q32, gal, top_k).cpu().tolist()
    rows = []
    for q, r in zip(queries, ranks):
        mods = "+".join(sorted(q["modalities"]))
        sample_ids = [s["img_id"] for s in q["samples"].values() if "img_id" in s]
        key = f"{q['pid']}|{mods}|{'+'.join(sample_ids)}"
        ids = [gallery_meta[i]["img_id"] for i in r if i >= 0]
        rows.append((key, " ".join(str(g) for g in ids if g is not None)))
    with open(output_path, "w", newline="", encoding="utf-8") as f:
        wtr = csv.writer(f)
        wtr.writerow(["query_key", "ranked_gallery_ids"])
        wtr.writerows(rows)
