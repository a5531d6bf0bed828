"""SDM fwd+bwd timing (C2 and C5): eager autograd step and CUDA-graph replay."""
import json, sys
sys.path.insert(0, '.')
import torch
import bench
from prcv2025reid_b200 import synth
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
out = {}
for name, (P, K, n, dt) in {"c2_p4k2_fp32_4pairs": (4, 2, 4, torch.float32), "c5_p64k8_bf16_10pairs": (64, 8, 10, torch.bfloat16),
                            "c5_p64k8_bf16_4pairs": (64, 8, 4, torch.bfloat16), "p64k8_fp32_4pairs": (64, 8, 4, torch.float32)}.items():
    us, ab, n, gu = bench.time_sdm(torch, synth, sdm_loss_pairs, P, K, n, dt)
    out[name] = {"us_per_step_eager": round(us, 1), "us_per_step_graph": round(gu, 1), "alg_bytes": ab, "hbm_gbs_graph": round(ab / gu / 1e3, 1)}
print(json.dumps(out, indent=1))
