"""Drop-in for the feature-level functions of the reference's tools/eval_mm_protocol.py.

Same names, argument meaning and error behaviour as the reference (file:line cited per function);
underneath, every tensor operation is a libreid_b200 CUDA kernel.  Inputs may live on the CPU (the
reference keeps `gallery_feats` on the CPU, eval_mm_protocol.py:302,325): they are moved to the
current CUDA device, results are returned where the reference returns them (tensors on the input's
device, metrics as Python floats).  Importing this module without the built library or without a
GPU and calling into it raises -- there is no CPU fallback.
"""
import csv
import itertools
import random
from typing import Dict, List, Optional, Tuple

import torch

from . import engine
from .synth import MODALITIES, MOD_ID

ALL_NON_RGB = ["ir", "cpencil", "sketch", "text"]          # eval_mm_protocol.py:35


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def pick_one(items: List[dict], rng: random.Random) -> dict:
    """eval_mm_protocol.py:37-39 -- one uniformly drawn sample (one `rng.choice` call, so the stream of a seeded
    `random.Random` is consumed exactly like the reference does)."""
    return rng.choice(items)


def combos(mods: List[str], k: int) -> List[Tuple[str, ...]]:
    """eval_mm_protocol.py:41-44 -- all k-subsets of `mods` in itertools.combinations order."""
    return list(itertools.combinations(mods, k))


def build_queries(index: Dict[int, Dict[str, List[dict]]], mode_k: int, rng: random.Random,
                  main_mod_choice="lexi_first") -> List[dict]:
    """eval_mm_protocol.py:223-276 -- the MM-k query list of an identity index (host-side combinatorics, row N4).

    For every identity (index order) and every k-subset of the non-RGB modalities it HAS samples for (subsets in
    combinations order over ALL_NON_RGB, each sorted by name): one query {"pid", "modalities", "samples"} whose
    `samples` dict holds one drawn sample per modality, main modality first.  The main modality is the
    alphabetically first one ("lexi_first") or one `rng.choice`; draws happen main first, then the rest in
    sorted order -- the same number and order of `rng` calls as the reference, so a seeded rng yields the same
    queries.  Identities with fewer than k populated modalities contribute nothing."""
    out: List[dict] = []
    for pid, by_mod in index.items():
        have = [m for m in ALL_NON_RGB if len(by_mod.get(m, ())) > 0]           # :240 (a missing key is not created)
        for subset in itertools.combinations(have, mode_k):                     # :243
            mods = tuple(sorted(subset))                                        # :244
            main = mods[0] if main_mod_choice == "lexi_first" else rng.choice(mods)   # :247-250
            picked = {main: pick_one(by_mod[main], rng)}                        # :257
            for m in mods:                                                      # :261-267 (every m in `have` is populated)
                if m != main:
                    picked[m] = pick_one(by_mod[m], rng)
            out.append({"pid": pid, "modalities": mods, "samples": picked})     # :270-274
    return out


def build_gallery(index: Dict[int, Dict[str, List[dict]]]) -> List[dict]:
    """eval_mm_protocol.py:280-287 -- every RGB sample of every identity, in index order."""
    return [s for by_mod in index.values() for s in by_mod.get("rgb", [])]


def extract_gallery_feats(gallery: List[dict], extractor, cache_dir: str):
    """eval_mm_protocol.py:291-325 -- the gallery feature cache: `rgb_feats.npy` (fp32 [G, D], rows L2-normalised) +
    `rgb_meta.json` ([{"img_id", "pid", "camid"}]) under cache_dir.  A complete cache is returned as it is (:296-302,
    no device involved); otherwise every item is encoded with `extractor.encode_rgb(item["img_path"])` (the model
    forward, not part of this library), all rows are normalised in ONE K1 pass instead of one `l2n` call per image
    (:310), and both files are written in the reference's format.  -> (feats fp32 CPU tensor [G, D], meta)."""
    import os
    import numpy as np
    from . import gallery_store
    feat_path = os.path.join(cache_dir, gallery_store.FEATS)
    meta_path = os.path.join(cache_dir, gallery_store.META)
    os.makedirs(cache_dir, exist_ok=True)                                                     # :293
    if os.path.exists(feat_path) and os.path.exists(meta_path):                               # :297
        return torch.from_numpy(np.load(feat_path)).float(), gallery_store.load_meta(cache_dir)
    raw = torch.stack([extractor.encode_rgb(item["img_path"]).float().view(-1) for item in gallery])   # :309, :310 (.float().view)
    feats = l2n(raw).cpu()                                                                    # :310, batched
    meta = [{"img_id": item.get("img_id", None), "pid": int(item["pid"]), "camid": item.get("camid", None)}
            for item in gallery]                                                              # :314-318
    gallery_store.save_cache(cache_dir, feats, meta)                                          # :321-323
    return feats.float(), meta


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


def _fuse_batch(queries: List[dict], extractor, weight_cfg: Dict[str, float], want_f16: bool = False):
    """Batched extract_query_feat (eval_mm_protocol.py:328-365) -> [Q, D] fp32 on the CUDA device (and, with want_f16,
    the fp16 tensor-core operand copy the kernels write in the same pass).

    The extractor is duck-typed like the reference's (one `encode_*` call per sample, one `fuse_features_if_any` call per
    query, :341-357): those calls are the only per-query Python left; features travel to the device in ONE copy, the
    per-modality l2n (:353), the weighted sum (:362-364) and the final l2n (:365) are one kernel per distinct k."""
    import numpy as np
    dev = _dev()
    Q = len(queries)
    feats, mods, counts = [], [], []
    for q in queries:
        ms = []
        for m, sample in q["samples"].items():
            feats.append(_encode(extractor, m, sample).float().view(-1))
            ms.append(m)
        mods.append(ms); counts.append(len(ms))
    if Q == 0:
        e = torch.empty(0, 0, device=dev)
        return (e, e.to(torch.float16)) if want_f16 else e
    D = feats[0].numel()
    out = torch.empty(Q, D, dtype=torch.float32, device=dev)
    out16 = torch.empty(Q, D, dtype=torch.float16, device=dev) if want_f16 else None
    flat = torch.stack(feats).to(dev)                                    # one host -> device copy
    flat_n, _ = engine.l2norm_rows(flat)                                 # per-modality l2n (:353): the fusion hook's argument
    rows_n = flat_n.unbind(0)
    counts = np.asarray(counts, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    hooked_rows, hooked_idx, weighted = [], [], []
    for qi in range(Q):
        o, k = int(offs[qi]), int(counts[qi])
        fused = extractor.fuse_features_if_any(list(rows_n[o:o + k]), mods[qi])       # :357
        if fused is not None:
            hooked_rows.append(fused.float().view(-1)); hooked_idx.append(qi)         # -> l2n(fused) (:359)
        else:
            weighted.append(qi)
    if hooked_idx:
        hn, hn16 = engine.l2norm_rows(torch.stack(hooked_rows).to(dev), want_f16=want_f16)
        it = torch.as_tensor(hooked_idx, device=dev)
        out[it] = hn
        if want_f16:
            out16[it] = hn16
    if weighted:
        names = list(MODALITIES) + sorted(set(m for ms in mods for m in ms) - set(MODALITIES))
        ids = {m: i for i, m in enumerate(names)}
        w = torch.tensor([float(weight_cfg.get(m, 1.0)) for m in names], dtype=torch.float32, device=dev)   # :362
        flat_mid = np.fromiter((ids[m] for ms in mods for m in ms), dtype=np.int32, count=int(counts.sum()))
        weighted = np.asarray(weighted, dtype=np.int64)
        for k in np.unique(counts[weighted]).tolist():
            idxs = weighted[counts[weighted] == k]
            rix = offs[idxs][:, None] + np.arange(k)[None, :]                                  # [n, k] rows of `flat`
            rows = flat[torch.from_numpy(rix.reshape(-1)).to(dev)].view(len(idxs), k, D)       # one gather
            mid = torch.from_numpy(flat_mid[rix]).to(dev)
            if k == 1:      # weighted path with one feature: l2n(w * f) -- keep the weight (no hook result)
                rows = torch.cat([rows, torch.zeros_like(rows)], dim=1)
                mid = torch.cat([mid, torch.full_like(mid, -1)], dim=1)
            q32, q16 = engine.fuse_queries(rows, mid, w)
            it = torch.from_numpy(idxs).to(dev)
            out[it] = q32
            if want_f16:
                out16[it] = q16
    return (out, out16) if want_f16 else out


def extract_query_feat(q: dict, extractor, weight_cfg: Dict[str, float]) -> torch.Tensor:
    """eval_mm_protocol.py:328-365 -- one query -> fused, L2-normalised feature [D]."""
    f = _fuse_batch([q], extractor, weight_cfg)[0]
    probe = next(iter(q["samples"].items()))
    src = _encode(extractor, probe[0], probe[1]).device
    return f.to(src)


def _image_index(g_imgid) -> Dict[object, List[int]]:
    by_id: Dict[object, List[int]] = {}
    for i, gid in enumerate(g_imgid):
        if gid is not None:
            by_id.setdefault(gid, []).append(i)
    return by_id


def _exclusions(queries, g_imgid, ignore_same_img: bool, dev, by_id=None) -> Optional[torch.Tensor]:
    # eval_mm_protocol.py:408-418: gallery rows whose img_id is one of the query samples' img_ids.
    # One pass over the gallery ids (cached on the installed gallery), one dictionary probe per query sample.
    if not ignore_same_img:
        return None
    import numpy as np
    if by_id is None:
        by_id = _image_index(g_imgid)
    rows, width = [], 0
    for q in queries:
        hit = []
        for s in q["samples"].values():
            iid = s.get("img_id") if "img_id" in s else None
            if iid is not None:
                h = by_id.get(iid)
                if h:
                    hit.extend(h)
        if len(hit) > 1:
            hit = sorted(set(hit))
        rows.append(hit)
        width = max(width, len(hit))
    if width == 0:
        return None
    ex = np.full((len(queries), width), -1, dtype=np.int32)
    for i, h in enumerate(rows):
        if h:
            ex[i, :len(h)] = h
    return torch.from_numpy(ex).to(dev)


def install_gallery(gallery_feats: torch.Tensor, gallery_meta: List[dict]) -> engine.GalleryShard:
    """The gallery side of rank_and_metrics (eval_mm_protocol.py:390 `g_pids`, :546 `l2n(g_feats)`) installed once
    on the device: normalised fp32 rows, the fp16 tensor-core copy and the identity index.  Pass the result as
    `shard=` to rank_and_metrics when several query sets run against the same gallery (run_eval's MM-1..4 loop)."""
    dev = _dev()
    g_pids = torch.tensor([m["pid"] for m in gallery_meta], dtype=torch.long)                 # :390
    shard = engine.prepare_gallery(gallery_feats.to(dev), g_pids.to(dev))
    shard.img_index = _image_index([m.get("img_id", None) for m in gallery_meta])             # :391, for the same-image rule
    return shard


def shuffled_gallery(gallery_feats: torch.Tensor, gallery_meta: List[dict], seed: int = 0):
    """(gallery_feats, gallery_meta) with the rows in a seeded random order (meta permuted alongside, so that image ids, the
    same-image rule and the submission CSV are unaffected; only the order of exactly tied scores can change).

    Why: beyond the exactly re-scored head, `mode="fused"` counts deep positives on row samples with a FIXED phase (local
    rows = 5 mod 32 / mod 1024 of a shard).  The samples are unbiased unless a row's position modulo 32
    carries information about its score -- e.g. a cache written with exactly 32 images per identity in camera order.  For
    such a gallery pass the shuffled pair to `install_gallery` / `rank_and_metrics` / `export_submission_csv` (or evaluate
    with `exact_ap=True`).  The reference (eval_mm_protocol.py:396-455) is invariant under gallery row order."""
    G = gallery_feats.shape[0]
    if len(gallery_meta) != G:
        raise ValueError("shuffled_gallery: one meta entry per gallery row")
    perm = torch.randperm(G, generator=torch.Generator().manual_seed(int(seed)))
    feats = gallery_feats[perm.to(gallery_feats.device)]
    return feats, [gallery_meta[i] for i in perm.tolist()]


def rank_and_metrics(queries: List[dict], gallery_feats: torch.Tensor, gallery_meta: List[dict], extractor,
                     weight_cfg: Dict[str, float], ignore_same_img=True, cross_camera=False,
                     mode: str = "fused", shard: Optional[engine.GalleryShard] = None,
                     exact_ap: bool = False) -> Dict[str, float]:
    """eval_mm_protocol.py:369-469.  `cross_camera` is accepted and ignored like the reference (:375,392).
    `exact_ap=True` counts every positive's rank on every gallery row (no row samples for deep positives, engine.retrieve).

    Returns {"mAP","R@1","R@5","R@10","num_queries"}; queries without a positive are skipped (:430-432),
    an empty result gives 0.0 metrics (:458-461).  `shard` (optional, from install_gallery on the same
    gallery_feats / gallery_meta) skips the per-call gallery installation."""
    dev = _dev()
    if len(queries) == 0 or len(gallery_meta) == 0:
        return {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0, "num_queries": 0}
    if shard is None:
        shard = install_gallery(gallery_feats, gallery_meta)
    elif shard.G_total != len(gallery_meta):
        raise ValueError("shard was installed for a gallery of %d rows, gallery_meta has %d" % (shard.G_total, len(gallery_meta)))
    q32, q16 = _fuse_batch(queries, extractor, weight_cfg, want_f16=True)
    q_pid = torch.tensor([int(q["pid"]) for q in queries], dtype=torch.long, device=dev)
    by_id = getattr(shard, "img_index", None)
    g_imgid = None if by_id is not None else [m.get("img_id", None) for m in gallery_meta]    # :391
    excl = _exclusions(queries, g_imgid, ignore_same_img, dev, by_id)
    extra = {"exact_ap": True} if exact_ap else {}
    res = engine.retrieve(shard, q32, q16, q_pid, excl, topk=10, mode=mode, **extra)
    return res.metrics


def run_eval_features(index: Dict[int, Dict[str, List[dict]]], gallery_feats: torch.Tensor, gallery_meta: List[dict],
                      extractor, seed: int = 42, weight_cfg: Optional[Dict[str, float]] = None, ignore_same_img=True,
                      cross_camera=False, mode: str = "fused", exact_ap: bool = False) -> Dict[str, Dict[str, float]]:
    """The evaluation loop of run_eval (eval_mm_protocol.py:497-590) from pre-extracted features: the part of
    run_eval after the model / dataset / gallery cache have been loaded (:508-546 are out of scope: they need the
    CLIP weights and the image files).  Seeds `random.Random(seed)` (:498), default weights (:503-504), then for
    k = 1..4 build_queries(index, k, rng, "lexi_first") (:556) and rank_and_metrics (:566-569); MM-k without
    queries gives zero metrics (:559-562); "AVG(1-4)" is the plain mean over the MM-k that had queries (:576-586).
    The gallery is installed on the device once for the four passes."""
    rng = random.Random(seed)                                                                 # :498
    if weight_cfg is None:
        weight_cfg = {"ir": 1.0, "cpencil": 1.0, "sketch": 1.0, "text": 1.2}                  # :504
    shard = install_gallery(gallery_feats, gallery_meta) if len(gallery_meta) else None       # :545-546, once
    results: Dict[str, Dict[str, float]] = {}
    for k in (1, 2, 3, 4):                                                                    # :553
        queries = build_queries(index, mode_k=k, rng=rng, main_mod_choice="lexi_first")       # :556
        if len(queries) == 0:                                                                 # :559-562
            results["MM-%d" % k] = {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0, "num_queries": 0}
            continue
        results["MM-%d" % k] = rank_and_metrics(queries, gallery_feats, gallery_meta, extractor, weight_cfg,
                                                ignore_same_img=ignore_same_img, cross_camera=cross_camera,
                                                mode=mode, shard=shard, exact_ap=exact_ap)
    valid = [results["MM-%d" % k] for k in (1, 2, 3, 4) if results["MM-%d" % k]["num_queries"] > 0]   # :576
    results["AVG(1-4)"] = {key: (sum(r[key] for r in valid) / len(valid) if valid else 0.0)
                           for key in ("mAP", "R@1", "R@5", "R@10")}                          # :577-586
    return results


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
