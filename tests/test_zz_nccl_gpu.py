"""The multi-GPU exchange with REAL NCCL (the CPU suite covers it on gloo with the stand-in library): torchrun with one
process per GPU, contiguous gallery shards, engine.retrieve(group=WORLD) -- fused (sampled) branch per shard, all-reduce
of the positives' scores / the counts / the flags, all-gather + merge of the top lists, sharded query upload -- compared
with the oracle's ranking of the unsharded gallery.  Skipped on boxes with fewer GPUs than the world size."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,workload,nq", [(2, "c3b", 2048), (8, "c4", 512)])
def test_sharded_retrieval_with_nccl_matches_oracle(tmp_path, world, workload, nq):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "report.json")
    port = 29400 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py"), out, workload, str(nq)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.load(open(out))
    print(json.dumps(rep))
    assert rep["path"] == "fused" and rep["ranks_agree"] and rep["host_equals_resident"]
    assert rep["metrics"]["num_queries"] == rep["oracle"]["num_queries"]
    assert abs(rep["d_map"]) <= 1e-4 and rep["d_ap_max"] <= 2.5e-3 and rep["d_ap_mean"] <= 3e-4
    for k in ("R@1", "R@5", "R@10"):
        assert rep["metrics"][k] == rep["oracle"][k]
    assert rep["cmc_rank_mismatches"] == 0 and rep["top10_lists_differing_beyond_ties"] == 0
