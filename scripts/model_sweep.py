#!/usr/bin/env python
"""Sweep the parameters of the fused kernel's counting rule on the CPU model (oracle/fused_model.py): for every
(calibration rows, sample width) the mAP error against the exact oracle, the share of thresholds counted on the row
sample and the epilogue's hit volume per query -- what a change would cost / save, before it is built.  No GPU.

    python scripts/model_sweep.py [--ids 2500] [--queries 96] [--chunks 4] [--world 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fused_model as fm  # noqa: E402
from oracle import retrieval as orc  # noqa: E402
from prcv2025reid_b200 import sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ids", type=int, default=2500)          # x 40 gallery rows
    ap.add_argument("--queries", type=int, default=96)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--w2", type=int, default=1024, help="level-2 sample width (0 = one sampling level)")
    a = ap.parse_args()
    case = synth.make_retrieval_case(1005, a.ids, 40, 4, 4, excl_frac=0.02, n_excl=2, max_queries=a.queries)
    q = orc.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor())
    g = orc.l2n(case.gallery_raw)
    exact = orc.rank_and_metrics_counting(q, g, case.q_pid, case.g_pid, case.excl, return_per_query=True)
    thr, n_pos = fm.positive_thresholds(q, g, case.q_pid, case.g_pid, case.excl)
    print("gallery %d rows, %d queries, %d rank(s) x %d chunks; exact mAP %.6f" % (g.shape[0], q.shape[0], a.world, a.chunks, exact["mAP"]))
    print("%10s %8s | %11s %11s %10s | %14s %14s" % ("calib_rows", "sample_w", "dmAP", "max dAP", "deep share", "hits/query", "sampled hits"))
    for calib in (1024, 2048, 4096, 8192):
        for sw in (16, 32, 64):
            tot = np.zeros_like(thr, dtype=np.int64)
            deep, hits = [], np.zeros(3)
            for r in range(a.world):
                r0, r1 = sharding.shard_range(g.shape[0], r, a.world)
                c = fm.fused_counts(q, g[r0:r1], case.q_pid, case.g_pid[r0:r1], case.excl, n_shards=a.world, thr=thr, n_pos=n_pos,
                                    g_offset=r0, sample_w=sw, sample_w2=a.w2, calib_rows=calib)
                tot += c["pos_above"]
                deep.append(1.0 - c["n_exact"].sum() / max(1, n_pos.sum()))
                hits += c["hits"].mean(0)
            m = fm.metrics_from_counts(tot, n_pos)
            v = m["_ap"] >= 0
            print("%10d %8d | %+11.2e %11.2e %10.3f | %14.0f %14.0f" % (calib, sw, m["mAP"] - exact["mAP"],
                  np.abs(m["_ap"][v] - exact["_ap"][v]).max(), float(np.mean(deep)), hits[0], hits[1] + hits[2]))


if __name__ == "__main__":
    main()
