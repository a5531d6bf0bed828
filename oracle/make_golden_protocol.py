"""Generate tests/golden/protocol.json from the UNMODIFIED reference: `build_queries` / `build_gallery`
(tools/eval_mm_protocol.py:223-287) and the MM-1..4 evaluation loop of `run_eval` (:553-586, which calls the
reference's own `build_queries` and `rank_and_metrics`; `run_eval` itself cannot run here because it loads the
CLIP checkpoint and the image dataset, :508-538, so its loop body is driven with the same calls on a synthetic
identity index and a tensor-serving extractor).  TEST INFRASTRUCTURE.  Run: python -m oracle.make_golden_protocol"""
import hashlib
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from prcv2025reid_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
INDEX_ARGS = dict(seed=31, n_ids=48, rgb_per_id=4, max_per_mod=3, drop_frac=0.25)
RUN_SEED = 42          # run_eval's default seed (:477)


def query_digest(queries):
    """Order-sensitive digest of a query list: pid, modality tuple and the drawn samples in dict order."""
    rows = [[q["pid"], list(q["modalities"]), [[m, s.get("img_path", s.get("text")), s["img_id"]] for m, s in q["samples"].items()]]
            for q in queries]
    return hashlib.sha256(json.dumps(rows, sort_keys=False).encode()).hexdigest(), rows


def reference_protocol(index, g_feats, g_meta, ext, seed, ignore_same_img=True):
    """run_eval :497-586 with the model / dataset loading removed; every call goes to the reference module."""
    ref = ref_loader.load_reference_eval()
    rng = random.Random(seed)                                                       # :498
    weight_cfg = {"ir": 1.0, "cpencil": 1.0, "sketch": 1.0, "text": 1.2}            # :504
    g = ref.l2n(g_feats)                                                            # :546
    results, digests = {}, {}
    for k in [1, 2, 3, 4]:                                                          # :553
        queries = ref.build_queries(index, mode_k=k, rng=rng, main_mod_choice="lexi_first")
        digests["MM-%d" % k] = query_digest(queries)[0]
        if len(queries) == 0:
            results["MM-%d" % k] = {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0, "num_queries": 0}
            continue
        results["MM-%d" % k] = ref_loader.quiet(ref.rank_and_metrics, queries, g, g_meta, ext, weight_cfg,
                                                ignore_same_img=ignore_same_img, cross_camera=False)
    valid = [results["MM-%d" % k] for k in [1, 2, 3, 4] if results["MM-%d" % k]["num_queries"] > 0]
    results["AVG(1-4)"] = {key: float(np.mean([r[key] for r in valid])) for key in ("mAP", "R@1", "R@5", "R@10")} \
        if valid else {"mAP": 0.0, "R@1": 0.0, "R@5": 0.0, "R@10": 0.0}              # :576-586
    return results, digests


def main():
    ref = ref_loader.load_reference_eval()
    index, g_feats, g_meta, ext = synth.make_protocol_index(**INDEX_ARGS)
    out = {"index_args": INDEX_ARGS, "run_seed": RUN_SEED,
           "checksum": float(g_feats.double().abs().sum()) + float(sum(float(v.double().abs().sum()) for v in ext.table.values())),
           "gallery_img_ids_sha": hashlib.sha256(json.dumps([s["img_id"] for s in ref.build_gallery(index)]).encode()).hexdigest(),
           "build_queries": {}}
    for mode in ("lexi_first", "random"):
        for k in (1, 2, 3, 4):
            qs = ref.build_queries(index, mode_k=k, rng=random.Random(1000 + k), main_mod_choice=mode)
            sha, rows = query_digest(qs)
            out["build_queries"]["%s/k%d" % (mode, k)] = {"n": len(qs), "sha256": sha, "first": rows[:3]}
    for mask in (True, False):
        res, dig = reference_protocol(index, g_feats, g_meta, ext, RUN_SEED, ignore_same_img=mask)
        out["run_eval/ignore_same_img=%s" % mask] = {"results": res, "query_sha256": dig}
        print(mask, json.dumps(res))
    # an index where no identity has 4 modalities: MM-4 has no queries and is left out of the average (:559-562,:576)
    sparse = {pid: {m: v for m, v in by.items() if m != "text"} for pid, by in index.items()}
    res, dig = reference_protocol(sparse, g_feats, g_meta, ext, RUN_SEED)
    out["run_eval/no_text"] = {"results": res, "query_sha256": dig}
    with open(os.path.join(GOLDEN, "protocol.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1)
    print("wrote protocol.json", {k: v["n"] for k, v in out["build_queries"].items()})


if __name__ == "__main__":
    main()
