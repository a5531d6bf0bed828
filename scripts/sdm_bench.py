"""SDM fwd+bwd timing (C2 and C5): autograd step time and device time of the kernels alone."""
import json, sys
sys.path.insert(0, '.')
import torch
import bench
from prcv2025reid_b200 import synth
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs
out = {}
us, ab, n, ku = bench.time_sdm(torch, synth, sdm_loss_pairs, 4, 2, 4, torch.float32)
out["c2_p4k2_fp32_4pairs"] = {"us_per_step": us, "kernels_us": ku, "alg_bytes": ab}
us, ab, n, ku = bench.time_sdm(torch, synth, sdm_loss_pairs, 64, 8, 10, torch.bfloat16)
out["c5_p64k8_bf16_10pairs"] = {"us_per_step": us, "kernels_us": ku, "alg_bytes": ab, "hbm_gbs_kernels": ab / ku / 1e3}
us, ab, n, ku = bench.time_sdm(torch, synth, sdm_loss_pairs, 64, 8, 4, torch.bfloat16)
out["c5_p64k8_bf16_4pairs"] = {"us_per_step": us, "kernels_us": ku, "alg_bytes": ab, "hbm_gbs_kernels": ab / ku / 1e3}
print(json.dumps(out, indent=1))
