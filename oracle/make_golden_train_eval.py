"""Generate tests/golden/train_eval.npz from the UNMODIFIED reference train-time evaluator functions
(train.py:101-138 `compute_map` / `compute_cmc`, :451-479 `_reid_map`), extracted from train.py with ast
(oracle.ref_loader.load_reference_train_eval).  TEST INFRASTRUCTURE.  Run: python -m oracle.make_golden_train_eval"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_case(seed=91, n_ids=60, per_id=7, n_q=150, d=512, noise=5.0):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_ids, d, generator=g)
    gl = torch.arange(n_ids * per_id) // per_id
    ql = torch.randint(0, n_ids, (n_q,), generator=g)
    ql[:5] = 10_000                                           # queries without any match
    gf = centres[gl] + noise * torch.randn(gl.numel(), d, generator=g)
    qf = centres[ql.clamp(max=n_ids - 1)] + noise * torch.randn(n_q, d, generator=g)
    return qf, gf, ql, gl


def main():
    ns = ref_loader.load_reference_train_eval()
    qf, gf, ql, gl = make_case()
    out = {"args": np.array([91, 60, 7, 150, 512], dtype=np.int64), "noise": np.float64(5.0),
           "checksum": np.float64(float(qf.double().abs().sum()) + float(gf.double().abs().sum()))}
    for k in (1, 5, 100):
        out["map_k%d" % k] = np.float64(ns["compute_map"](qf, gf, ql, gl, k=k))
    for k in (1, 10):
        out["cmc_k%d" % k] = np.float64(ns["compute_cmc"](qf, gf, ql, gl, k=k))
    qn = torch.nn.functional.normalize(qf, dim=1); gn = torch.nn.functional.normalize(gf, dim=1)
    m, t1 = ns["_reid_map"](qn @ gn.T, ql, gl)
    out["reid_map"] = np.array([m, t1], dtype=np.float64)
    np.savez(os.path.join(GOLDEN, "train_eval.npz"), **out)
    print({k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
