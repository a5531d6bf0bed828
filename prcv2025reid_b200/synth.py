"""Seeded synthetic "CLIP-shaped" 512-d features for the retrieval path and the SDM loss.

Recipe: feature = c_id + b_modality + sigma_modality * N(0, I), NOT
pre-normalised, so the normalise / fuse kernels do real work.  Noise levels are frozen here so
that MM-k mAP sits in an informative range (neither 0 nor 1).  Modality names follow the
reference evaluator (tools/eval_mm_protocol.py:35): ir, cpencil, sketch, text; gallery = rgb.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

FEAT_DIM = 512
# reference order ALL_NON_RGB = ["ir", "cpencil", "sketch", "text"] (eval_mm_protocol.py:35)
MODALITIES = ("ir", "cpencil", "sketch", "text")
MOD_ID = {m: i for i, m in enumerate(MODALITIES)}
# default weight_cfg of run_eval (eval_mm_protocol.py:504)
DEFAULT_WEIGHTS = {"ir": 1.0, "cpencil": 1.0, "sketch": 1.0, "text": 1.2}

# tuned once at C1 scale (500 ids x 20): MM-1 mAP 0.14, MM-2 0.36, MM-3 0.57, MM-4 0.73; frozen
SIGMA_RGB = 2.5
SIGMA = {"ir": 3.5, "cpencil": 3.5, "sketch": 4.5, "text": 5.5}
BIAS_SCALE = 0.3

# MM-k modality combinations in the order build_queries emits them (sorted tuples of
# itertools.combinations over ALL_NON_RGB, eval_mm_protocol.py:243-244)
def mm_combos(k: int) -> List[tuple]:
    from itertools import combinations
    return [tuple(sorted(c)) for c in combinations(MODALITIES, k)]


@dataclass
class RetrievalCase:
    """Tensor view of one evaluation job."""
    gallery_raw: torch.Tensor      # [G, D] fp32, un-normalised rgb features
    g_pid: torch.Tensor            # [G] int64
    query_raw: torch.Tensor        # [Q, k, D] fp32, un-normalised per-modality features
    mod_id: torch.Tensor           # [Q, k] int32 index into MODALITIES
    q_pid: torch.Tensor            # [Q] int64
    excl: torch.Tensor             # [Q, E] int32 gallery indices to mask (same image), -1 pad
    k: int

    @property
    def Q(self):
        return self.query_raw.shape[0]

    @property
    def G(self):
        return self.gallery_raw.shape[0]


def _chunk_gen(dev, seed: int, stream: int, chunk_index: int) -> torch.Generator:
    g = torch.Generator(device=dev)
    g.manual_seed((seed * 1000003 + stream * 7919 + chunk_index) & 0x7FFFFFFFFFFF)
    return g


def make_centres(seed: int, n_ids: int, dim: int = FEAT_DIM, device="cpu"):
    dev = torch.device(device)
    gen = _chunk_gen(dev, seed, 0, 0)
    centres = torch.randn(n_ids, dim, generator=gen, device=dev)
    bias = {m: BIAS_SCALE * torch.randn(dim, generator=gen, device=dev) for m in ("rgb",) + MODALITIES}
    return centres, bias


GALLERY_CHUNK = 1 << 14


def make_gallery_rows(seed: int, centres, bias, gal_per_id: int, row_start: int, row_end: int) -> torch.Tensor:
    """Rows [row_start, row_end) of the synthetic gallery.  Every GALLERY_CHUNK-aligned block has its
    own generator, so any rank can materialise exactly its shard (chunks are generated whole)."""
    dev = centres.device
    out = torch.empty(row_end - row_start, centres.shape[1], device=dev)
    G = centres.shape[0] * gal_per_id
    c0 = row_start // GALLERY_CHUNK
    c1 = (row_end + GALLERY_CHUNK - 1) // GALLERY_CHUNK
    for c in range(c0, c1):
        s, e = c * GALLERY_CHUNK, min(G, (c + 1) * GALLERY_CHUNK)
        noise = torch.randn(e - s, centres.shape[1], generator=_chunk_gen(dev, seed, 1, c), device=dev)
        pid = torch.arange(s, e, device=dev) // gal_per_id
        rows = centres[pid] + bias["rgb"] + SIGMA_RGB * noise
        a, b = max(s, row_start), min(e, row_end)
        out[a - row_start:b - row_start] = rows[a - s:b - s]
    return out


def make_retrieval_case(seed: int, n_ids: int, gal_per_id: int, k: int, queries_per_id: int,
                        excl_frac: float = 0.01, n_excl: int = 2, device="cpu",
                        dim: int = FEAT_DIM, gallery_rows=None, max_queries=None) -> RetrievalCase:
    """Generate a case.  gallery_rows=(start, end) materialises only that gallery shard (g_pid is
    always the full list); max_queries keeps only the first queries of the workload."""
    dev = torch.device(device)
    centres, bias = make_centres(seed, n_ids, dim, device)
    G = n_ids * gal_per_id
    g_pid = torch.arange(G, device=dev, dtype=torch.int64) // gal_per_id
    r0, r1 = (0, G) if gallery_rows is None else gallery_rows
    gallery = make_gallery_rows(seed, centres, bias, gal_per_id, r0, r1)

    combos = mm_combos(k)
    Q = n_ids * queries_per_id
    if max_queries is not None:
        Q = min(Q, int(max_queries))
    q_pid = torch.arange(Q, device=dev, dtype=torch.int64) // queries_per_id
    # query j of an identity uses combination j % len(combos) (mirrors the C(4,k) combos per id)
    combo_of_q = (torch.arange(Q, device=dev) % queries_per_id) % len(combos)
    combo_tab = torch.tensor([[MOD_ID[m] for m in c] for c in combos], device=dev, dtype=torch.int32)
    mod_id = combo_tab[combo_of_q]                       # [Q, k]
    sig_tab = torch.tensor([SIGMA[m] for m in MODALITIES], device=dev)
    bias_tab = torch.stack([bias[m] for m in MODALITIES])  # [4, D]
    query = torch.empty(Q, k, dim, device=dev)
    for ci, s in enumerate(range(0, Q, GALLERY_CHUNK)):
        e = min(Q, s + GALLERY_CHUNK)
        full = min(n_ids * queries_per_id, s + GALLERY_CHUNK) - s      # chunks are generated whole
        noise = torch.randn(full, k, dim, generator=_chunk_gen(dev, seed, 2, ci), device=dev)[:e - s]
        mid = mod_id[s:e].long()
        query[s:e] = centres[q_pid[s:e]][:, None, :] + bias_tab[mid] + sig_tab[mid][..., None] * noise

    # same-image exclusions: a seeded fraction of queries masks n_excl gallery rows of its own id
    n_excl = min(n_excl, k)
    excl = torch.full((Q, max(1, n_excl)), -1, device=dev, dtype=torch.int32)
    if excl_frac > 0 and n_excl > 0:
        gen = _chunk_gen(dev, seed, 3, 0)
        q_full = n_ids * queries_per_id                      # drawn for the whole workload, then cut
        pick = (torch.rand(q_full, generator=gen, device=dev) < excl_frac)[:Q]
        off = torch.randint(0, gal_per_id, (q_full, n_excl), generator=gen, device=dev)[:Q]
        rows = (q_pid[:, None] * gal_per_id + off).to(torch.int32)
        excl = torch.where(pick[:, None], rows, excl)
    return RetrievalCase(gallery, g_pid, query, mod_id, q_pid, excl, k)


def make_ragged_case(seed: int, n_ids: int, min_rows: int = 1, max_rows: int = 120, k: int = 2, queries_per_id: int = 2,
                     dup_frac: float = 0.0, excl_frac: float = 0.02, n_excl: int = 2, device="cpu",
                     dim: int = FEAT_DIM, heavy_tail: bool = True) -> RetrievalCase:
    """A gallery shaped like a real ReID one (ORBench: 45 113 RGB images of 1 000 identities, a long tail of image
    counts per identity): identity i owns between min_rows and max_rows gallery rows (heavy_tail: most identities are
    small, a few are large), rows are stored in SHUFFLED order (no periodic structure), person ids are non-contiguous,
    and a fraction dup_frac of the rows are bit-identical copies of another row filed under a DIFFERENT identity
    (exact score ties between a positive and a non-positive, between top-list neighbours)."""
    dev = torch.device(device)
    gen = _chunk_gen(dev, seed, 5, 0)
    centres = torch.randn(n_ids, dim, generator=gen, device=dev)
    bias = {m: BIAS_SCALE * torch.randn(dim, generator=gen, device=dev) for m in ("rgb",) + MODALITIES}
    u = torch.rand(n_ids, generator=gen, device=dev)
    if heavy_tail:
        u = u ** 2.5
    rows_of = (min_rows + torch.floor(u * (max_rows - min_rows + 1)).long()).clamp(min_rows, max_rows)
    rows_of[0] = max_rows                                     # the extremes are always present
    rows_of[1 % n_ids] = min_rows
    pid_of_id = 1000 + 7 * torch.arange(n_ids, device=dev, dtype=torch.int64)
    owner = torch.repeat_interleave(torch.arange(n_ids, device=dev), rows_of)
    G = int(owner.numel())
    perm = torch.randperm(G, generator=gen, device=dev)
    owner = owner[perm]
    gallery = centres[owner] + bias["rgb"] + SIGMA_RGB * torch.randn(G, dim, generator=gen, device=dev)
    g_pid = pid_of_id[owner]
    n_dup = int(dup_frac * G)
    if n_dup > 0:
        dst = torch.randperm(G, generator=gen, device=dev)[:n_dup]
        src = torch.randint(0, G, (n_dup,), generator=gen, device=dev)
        keep = ~torch.isin(src, dst)                          # a copy of a row that is itself overwritten would not be a copy
        gallery[dst[keep]] = gallery[src[keep]]               # features copied, identity of the destination row kept
    combos = mm_combos(k)
    Q = n_ids * queries_per_id
    q_id = torch.arange(Q, device=dev) // queries_per_id
    combo_tab = torch.tensor([[MOD_ID[m] for m in c] for c in combos], device=dev, dtype=torch.int32)
    mod_id = combo_tab[(torch.arange(Q, device=dev) % queries_per_id) % len(combos)]
    sig_tab = torch.tensor([SIGMA[m] for m in MODALITIES], device=dev)
    bias_tab = torch.stack([bias[m] for m in MODALITIES])
    mid = mod_id.long()
    query = centres[q_id][:, None, :] + bias_tab[mid] + sig_tab[mid][..., None] * torch.randn(Q, k, dim, generator=gen, device=dev)
    q_pid = pid_of_id[q_id]
    n_excl = max(1, n_excl)
    excl = torch.full((Q, n_excl), -1, device=dev, dtype=torch.int32)
    if excl_frac > 0:
        # same-image rule: a seeded fraction of the queries masks gallery rows of its own identity
        order = torch.argsort(owner, stable=True)              # rows grouped by identity
        start = torch.cumsum(rows_of, 0) - rows_of
        pick = torch.rand(Q, generator=gen, device=dev) < excl_frac
        off = (torch.rand(Q, n_excl, generator=gen, device=dev) * rows_of[q_id][:, None]).long()
        rows = order[(start[q_id][:, None] + off)].to(torch.int32)
        excl = torch.where(pick[:, None], rows, excl)
    return RetrievalCase(gallery, g_pid, query, mod_id, q_pid, excl, k)


def weights_tensor(weight_cfg: Optional[Dict[str, float]] = None, device="cpu") -> torch.Tensor:
    cfg = DEFAULT_WEIGHTS if weight_cfg is None else weight_cfg
    return torch.tensor([float(cfg.get(m, 1.0)) for m in MODALITIES], dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------
# Bridging a RetrievalCase to the reference's list-of-dicts interface (used by tests, the
# oracle and the drop-in demo): a fake extractor that serves pre-extracted features.
# ---------------------------------------------------------------------------------------------
class TensorExtractor:
    """Duck-typed stand-in for eval_mm_protocol.FeatureExtractor (eval_mm_protocol.py:133-219).

    `encode_*` return the stored raw feature for a key; `fuse_features_if_any` reproduces the
    reference's EFFECTIVE behaviour: single feature -> returned as is, several -> None because
    FeatureFusion raises on 1-D inputs and the exception is swallowed.
    """

    def __init__(self, table: Dict[str, torch.Tensor]):
        self.table = table

    def _get(self, key):
        return self.table[key]

    encode_ir = encode_cpencil = encode_sketch = encode_text = encode_rgb = _get

    def fuse_features_if_any(self, modal_feats, modalities):
        if len(modal_feats) <= 1:
            return modal_feats[0] if modal_feats else None
        return None


def case_to_reference_inputs(case: RetrievalCase):
    """-> (queries, gallery_meta, extractor) in the reference's own formats
    (eval_mm_protocol.py:270-274, 314-318)."""
    G, Q = case.G, case.Q
    gallery_meta = [{"img_id": "g%d" % i, "pid": int(p), "camid": None}
                    for i, p in enumerate(case.g_pid.tolist())]
    table = {}
    queries = []
    excl = case.excl.tolist()
    mod = case.mod_id.tolist()
    qp = case.q_pid.tolist()
    qr = case.query_raw.cpu()
    for qi in range(Q):
        samples = {}
        ex = [e for e in excl[qi] if e >= 0]
        for j, mi in enumerate(mod[qi]):
            m = MODALITIES[mi]
            key = "q%d_%s" % (qi, m)
            table[key] = qr[qi, j]
            # the j-th sample carries the img_id of the j-th excluded gallery row (same image)
            img_id = ("g%d" % ex[j]) if j < len(ex) else ("x%d_%s" % (qi, m))
            if m == "text":
                samples[m] = {"text": key, "img_id": img_id}
            else:
                samples[m] = {"img_path": key, "img_id": img_id}
        queries.append({"pid": qp[qi], "modalities": tuple(MODALITIES[mi] for mi in mod[qi]),
                        "samples": samples})
    return queries, gallery_meta, TensorExtractor(table)


# ---------------------------------------------------------------------------------------------
# SDM batches
# ---------------------------------------------------------------------------------------------
def make_sdm_batch(seed: int, P: int, K: int, n_modalities: int = 5, dim: int = FEAT_DIM,
                   dtype=torch.float32, device="cpu"):
    """P identities x K samples, one feature matrix per modality; modality 0 is `vis`."""
    gen = torch.Generator(device=torch.device(device))
    gen.manual_seed(seed)
    labels = torch.arange(P, device=device).repeat_interleave(K)
    centres = torch.randn(P, dim, generator=gen, device=device)
    feats = []
    for _ in range(n_modalities):
        f = centres[labels] + 1.5 * torch.randn(P * K, dim, generator=gen, device=device)
        feats.append(f.to(dtype))
    return feats, labels


# ---------------------------------------------------------------------------------------------
# An identity index in the reference's own format (eval_mm_protocol.py:58-130 `build_index`):
# {pid: {"rgb": [sample, ...], "ir": [...], "cpencil": [...], "sketch": [...], "text": [...]}}
# ---------------------------------------------------------------------------------------------
def make_protocol_index(seed: int, n_ids: int, rgb_per_id: int = 4, max_per_mod: int = 3, drop_frac: float = 0.25,
                        dim: int = FEAT_DIM):
    """-> (index, gallery_feats [G, D] un-normalised, gallery_meta, extractor) for the MM-1..4 protocol loop.

    Every identity has `rgb_per_id` gallery images and 0..max_per_mod samples of each non-RGB modality (a modality
    is absent with probability drop_frac, so identities contribute different numbers of MM-k combinations, some
    none at all).  A non-RGB sample shares its img_id with one of the identity's RGB images with probability 1/3
    (the same-image rule of rank_and_metrics, :408-418).  Features follow the recipe at the top of this file."""
    import random as _random
    rnd = _random.Random(seed)
    gen = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_ids, dim, generator=gen)
    bias = {m: BIAS_SCALE * torch.randn(dim, generator=gen) for m in ("rgb",) + MODALITIES}
    index: Dict[int, Dict[str, List[dict]]] = {}
    table: Dict[str, torch.Tensor] = {}
    gallery_rows, gallery_meta = [], []
    for i in range(n_ids):
        pid = 1000 + 3 * i                       # non-contiguous person ids
        by_mod: Dict[str, List[dict]] = {"rgb": []}
        for j in range(rgb_per_id):
            img_id = "p%d_rgb%d" % (pid, j)
            f = centres[i] + bias["rgb"] + SIGMA_RGB * torch.randn(dim, generator=gen)
            table["rgb/" + img_id] = f
            by_mod["rgb"].append({"img_path": "rgb/" + img_id, "pid": pid, "img_id": img_id, "camid": None})
            gallery_rows.append(f)
            gallery_meta.append({"img_id": img_id, "pid": pid, "camid": None})
        for m in MODALITIES:
            n = 0 if rnd.random() < drop_frac else rnd.randint(1, max_per_mod)
            if n == 0:
                if rnd.random() < 0.5:
                    by_mod[m] = []               # present but empty (build_queries must treat it like a missing key)
                continue
            by_mod[m] = []
            for j in range(n):
                own = "p%d_%s%d" % (pid, m, j)
                img_id = "p%d_rgb%d" % (pid, rnd.randrange(rgb_per_id)) if rnd.random() < 1.0 / 3 else own
                key = "%s/%s" % (m, own)
                table[key] = centres[i] + bias[m] + SIGMA[m] * torch.randn(dim, generator=gen)
                by_mod[m].append({"text": key, "pid": pid, "img_id": img_id} if m == "text" else
                                 {"img_path": key, "pid": pid, "img_id": img_id, "camid": None})
        index[pid] = by_mod
    return index, torch.stack(gallery_rows), gallery_meta, TensorExtractor(table)
