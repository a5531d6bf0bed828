"""Small invocations of the fused retrieval path (sampled branch, exclusions, threshold windows) and the SDM kernels for
compute-sanitizer (memcheck / racecheck / synccheck); results are checked against the all-fp32 path."""
import sys
import torch
sys.path.insert(0, '.')
from prcv2025reid_b200 import engine, synth
from prcv2025reid_b200.sdm_loss import sdm_loss_pairs, sdm_loss_pairs_labels

tiny = len(sys.argv) > 1 and sys.argv[1] == "tiny"
case = synth.make_ragged_case(5, 120 if tiny else 900, 1, 150, 2, 1 if tiny else 2, excl_frac=0.1, device="cuda")
shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)
q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, synth.weights_tensor(device="cuda"))
a = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode="fused", want_ap=True)
b = engine.retrieve(shard, q32, q16, case.q_pid, case.excl, mode="exact", want_ap=True)
assert a.path == "fused" and abs(a.metrics["mAP"] - b.metrics["mAP"]) <= 1e-4 and torch.equal(a.top_idx, b.top_idx), (a.metrics, b.metrics)
print("retrieval ok: gallery %d rows, pmax %d, %d queries" % (shard.G_local, shard.pmax, case.Q), a.metrics)
for N, dtype in ((8, torch.float32), (128, torch.bfloat16)):
    feats, labels = synth.make_sdm_batch(7, N // 2, 2, n_modalities=2, dtype=dtype, device="cuda")
    q, v = feats[1].clone().requires_grad_(True), feats[0].clone().requires_grad_(True)
    y = (labels[:, None] == labels[None, :]).float()
    loss = sdm_loss_pairs([q], [v], [y], tau=0.2)
    loss.sum().backward()
    if dtype == torch.bfloat16:
        q2, v2 = feats[1].clone().requires_grad_(True), feats[0].clone().requires_grad_(True)
        valid = torch.ones(N, dtype=torch.bool, device="cuda"); valid[::5] = False
        l2, st = sdm_loss_pairs_labels([q2], [v2], [labels], [labels], [valid], [valid], tau=0.2)
        l2.sum().backward()
    torch.cuda.synchronize()
    print("sdm ok", N, dtype, float(loss))
