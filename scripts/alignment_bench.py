"""Eager cost of the SDM section of compute_loss (models/model.py:556-625) as a training step sees it: forward + backward
of `sdm_alignment_loss` (label form: no y, no row filtering, no host read) against the route it replaced (one host read of
the masks, nonzero / index gathers per modality, dense y, `sdm_loss_pairs`).  Wall-clock per step with a device
synchronisation at the end of every step (the optimiser step that follows needs the gradients), median of `reps`."""
import json
import sys
import time

sys.path.insert(0, '.')
import torch

from prcv2025reid_b200.sdm_loss import sdm_alignment_loss, sdm_loss_pairs


def old_route(feats, masks, labels, tau=0.2):
    names = [m for m in feats if m != "vis"]
    flat = [(masks[m] > 0).reshape(labels.shape[0], -1)[:, 0] for m in ["vis"] + names]
    host = torch.stack(flat).cpu()                                   # the host read
    vis_idx = torch.nonzero(host[0]).flatten().to(labels.device)
    vfeat, vlab = feats["vis"][vis_idx], labels[vis_idx]
    qs, ys = [], []
    for k, m in enumerate(names):
        idx = torch.nonzero(host[k + 1]).flatten()
        if idx.numel() == 0:
            continue
        idx = idx.to(labels.device)
        qs.append(feats[m][idx])
        ys.append((labels[idx].view(-1, 1) == vlab.view(1, -1)).float())
    losses = sdm_loss_pairs(qs, [vfeat] * len(qs), ys, tau)
    has_pos = torch.stack([y.any() for y in ys]) & torch.isfinite(losses)
    kept = torch.where(has_pos, losses, torch.zeros_like(losses))
    return kept.sum() / has_pos.sum().clamp_min(1)


def step_time(fn, feats, masks, labels, reps=200, warm=20):
    ts = []
    for i in range(warm + reps):
        leaves = {m: f.detach().requires_grad_(True) for m, f in feats.items()}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = fn(leaves, masks, labels)
        loss.backward()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e6 * ts[len(ts) // 2], float(loss.detach())


out = {}
for name, (P, K, dt) in {"c2_p4k2_fp32": (4, 2, torch.float32), "p16k2_bf16": (16, 2, torch.bfloat16),
                         "p24k2_fp16": (24, 2, torch.float16), "c5_p64k8_bf16": (64, 8, torch.bfloat16)}.items():
    g = torch.Generator().manual_seed(7)
    B = P * K
    labels = torch.arange(P).repeat_interleave(K).cuda()
    centres = torch.randn(P, 512, generator=g)
    feats = {m: (centres.repeat_interleave(K, 0) + 1.5 * torch.randn(B, 512, generator=g)).to(dt).cuda()
             for m in ("vis", "nir", "sk", "cp", "text")}
    masks = {m: (torch.rand(B, 1, generator=g) > 0.2).float().cuda() for m in feats}
    masks["vis"][:] = 1.0
    a_us, a_loss = step_time(sdm_alignment_loss, feats, masks, labels)
    b_us, b_loss = step_time(old_route, feats, masks, labels)
    out[name] = {"rows": B, "dtype": str(dt).replace("torch.", ""), "label_form_us_per_step": round(a_us, 1),
                 "host_read_route_us_per_step": round(b_us, 1), "loss_label_form": a_loss, "loss_host_read_route": b_loss}
print(json.dumps(out, indent=1))
