"""Generate tests/golden/train_validate.json from the UNMODIFIED `validate_competition_style` / `evaluate_one_query` /
`_flatten_loaders` / `_reid_map` of train.py (:402-424, :451-631), run on pre-extracted features through fake
loaders (oracle.ref_loader stubs only the model forward, `_extract_feats_and_ids`).  TEST INFRASTRUCTURE.
Run: python -m oracle.make_golden_validate"""
import json
import os
import sys
import tempfile
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
ARGS = dict(seed=97, n_ids=40, per_id=5, d=512)
SET_SIZES = {"single/nir": (60, 4.0), "single/sk": (50, 5.0), "single/cp": (45, 4.5), "single/text": (70, 6.0),
             "double/nir+sk": (30, 3.5), "quad/nir+sk+cp+text": (80, 2.5)}


class FeatureLoader:
    """Looks like a DataLoader to the reference (`.dataset`, iterable, `.batch_size`); items are (feature, id)."""

    def __init__(self, feats, ids):
        self.dataset = [(f, i) for f, i in zip(feats, ids)]
        self.batch_size, self.num_workers, self.pin_memory, self.collate_fn = 16, 0, False, None

    def __iter__(self):
        return iter(self.dataset)


def make_world(seed, n_ids, per_id, d, rename=None, first_unmatched=False):
    """-> (g_feat, g_id, {name: (q_feat, q_id)}) with L2-normalised features (train.py:442)."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_ids, d, generator=g)
    g_id = torch.arange(n_ids * per_id) // per_id
    g_feat = F.normalize(centres[g_id] + 3.0 * torch.randn(g_id.numel(), d, generator=g), dim=1)
    sets = {}
    for name, (n, noise) in SET_SIZES.items():
        q_id = torch.randint(0, n_ids, (n,), generator=g)
        q_id[-2:] = 10_000 + torch.arange(2)                    # two queries without any match per set
        if first_unmatched and name == "single/nir":
            q_id[0] = 20_000                                     # the query the "simplified CMC" (:620-621) looks at
        q_feat = F.normalize(centres[q_id.clamp(max=n_ids - 1)] + noise * torch.randn(n, d, generator=g), dim=1)
        sets[(rename or {}).get(name, name)] = (q_feat, q_id)
    return g_feat, g_id, sets


def nest(sets):
    out = {}
    for name, v in sets.items():
        a, b = name.split("/")
        out.setdefault(a, {})[b] = v
    return out


def run_reference(g_feat, g_id, sets, sample_ratio=1.0, include=None, torch_seed=None):
    ns = ref_loader.load_reference_train_eval()
    loaders = {a: {b: FeatureLoader(*v) for b, v in d.items()} for a, d in nest(sets).items()}
    with tempfile.TemporaryDirectory() as td:
        cfg = types.SimpleNamespace(eval_cache_dir=td, eval_cache_tag="golden")
        if include is not None:
            cfg.eval_include_patterns = include
        model = types.SimpleNamespace(eval=lambda: None)
        if torch_seed is not None:
            torch.manual_seed(torch_seed)
        return ref_loader.quiet(ns["validate_competition_style"], model, FeatureLoader(g_feat, g_id), loaders,
                                torch.device("cpu"), sample_ratio=sample_ratio, cfg=cfg, epoch=3)


CASES = {
    "default": dict(),
    "cpencil_name": dict(rename={"single/cp": "single/cpencil"}),           # matched by the filter, missed by the aggregation
    "sampled": dict(sample_ratio=0.5, torch_seed=123),
    "include_all": dict(include=["*"]),
    "no_quad": dict(include=["single/*"]),
    "first_unmatched": dict(first_unmatched=True),
}


def main():
    out = {"args": ARGS, "set_sizes": {k: list(v) for k, v in SET_SIZES.items()}, "cases": {}}
    g_feat, g_id, sets = make_world(**ARGS)
    out["checksum"] = float(g_feat.double().abs().sum()) + float(sum(float(v[0].double().abs().sum()) for v in sets.values()))
    for name, kw in CASES.items():
        kw = dict(kw)
        g_feat, g_id, sets = make_world(**ARGS, rename=kw.pop("rename", None), first_unmatched=kw.pop("first_unmatched", False))
        res = run_reference(g_feat, g_id, sets, **kw)
        out["cases"][name] = json.loads(json.dumps(res))
        print(name, {k: v for k, v in res.items() if k != "detail"}, list(res["detail"]))
    with open(os.path.join(GOLDEN, "train_validate.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
