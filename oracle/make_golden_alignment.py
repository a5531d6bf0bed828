"""Generate tests/golden/sdm_alignment.npz from the UNMODIFIED `compute_loss` of models/model.py (:512-659; its SDM
section :552-635 is the caller of sdm_loss_stable, row N1), run through oracle.ref_loader.load_reference_compute_loss
with a stand-in `self`.  Stored per case: the inputs, the reference's `sdm_loss` and its autograd gradients with
respect to every modality's features.  TEST INFRASTRUCTURE.  Run: python -m oracle.make_golden_alignment"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODS = ("vis", "nir", "sk", "cp", "text")


def make_self(tau=0.2, epoch=5, warmup=0, weight=0.3, training=True):
    """What compute_loss reads from the model object."""
    return types.SimpleNamespace(ce_loss=torch.nn.CrossEntropyLoss(), current_epoch=epoch, ce_weight=1.0,
                                 config=types.SimpleNamespace(sdm_weight_warmup_epochs=warmup),
                                 contrastive_weight=weight, sdm_temperature=tau, training=training)


DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def make_inputs(seed, B, d, n_ids, kind, dtype="fp32"):
    """dtype: what the training autocast (train.py:852) hands to compute_loss as `raw_modality_features`."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, n_ids, (B,), generator=g)
    feats = {m: torch.randn(B, d, generator=g).to(DTYPES[dtype]) for m in MODS}
    masks = {m: (torch.rand(B, 1, generator=g) > 0.3).float() for m in MODS}
    if kind == "ragged":
        masks["cp"] = torch.zeros(B, 1)                          # a modality without valid rows (:597)
        masks["sk"] = torch.zeros(B, 1); masks["sk"][0] = 1.0    # a single valid row
        masks["text"] = masks["text"].squeeze(-1)                # 1-D mask (:571 handles both)
    elif kind == "no_vis":
        masks["vis"] = torch.zeros(B, 1)                         # :572-574
    elif kind == "no_pairs":
        labels = torch.arange(B)                                 # nothing shares an identity with a DIFFERENT row ...
        masks = {m: torch.zeros(B, 1) for m in MODS}
        masks["vis"][: B // 2] = 1.0                             # ... and vis / non-vis rows are disjoint: no positive (:608)
        for m in MODS[1:]:
            masks[m][B // 2:] = 1.0
    elif kind == "missing":
        feats["nir"] = None                                      # :593-594
        del masks["sk"]
    return feats, masks, labels


CASES = {  # name: (seed, B, d, n_ids, kind, tau[, dtype])
    "full": (1, 12, 512, 4, "full", 0.2),
    "ragged": (2, 12, 512, 4, "ragged", 0.2),
    "no_vis": (3, 8, 128, 3, "no_vis", 0.2),
    "no_pairs": (4, 8, 128, 3, "no_pairs", 0.2),
    "missing": (5, 10, 128, 3, "missing", 0.1),
    # features in the autocast dtype: the loss normalises in that dtype (sdm_loss.py:31-32)
    "full_bf16": (6, 16, 512, 5, "full", 0.2, "bf16"),
    "ragged_bf16": (7, 12, 512, 4, "ragged", 0.2, "bf16"),
    "full_fp16": (8, 16, 512, 5, "full", 0.2, "fp16"),
    "large_bf16": (9, 96, 512, 12, "full", 0.2, "bf16"),       # valid rows >= 64: the tcgen05 path
}


def run_reference(feats, masks, labels, tau):
    compute_loss = ref_loader.load_reference_compute_loss()
    leaves = {m: (f.clone().requires_grad_(True) if f is not None else None) for m, f in feats.items()}
    outputs = {"logits": torch.zeros(labels.shape[0], int(labels.max()) + 1), "raw_modality_features": leaves,
               "feature_masks": masks}
    out = compute_loss(make_self(tau), outputs, labels)
    loss = out["sdm_loss"]
    grads = {}
    if loss.requires_grad:
        loss.backward()
        grads = {m: t.grad for m, t in leaves.items() if t is not None and t.grad is not None}
    return loss.detach(), grads, out


def main():
    payload = {}
    for name, spec in CASES.items():
        seed, B, d, n_ids, kind, tau = spec[:6]
        feats, masks, labels = make_inputs(seed, B, d, n_ids, kind, *spec[6:])
        loss, grads, out = run_reference(feats, masks, labels, tau)
        payload[name + "/loss"] = np.float32(float(loss))
        payload[name + "/args"] = np.array([seed, B, d, n_ids], dtype=np.int64)
        payload[name + "/tau"] = np.float64(tau)
        payload[name + "/checksum"] = np.float64(sum(float(f.double().abs().sum()) for f in feats.values() if f is not None))
        payload[name + "/total"] = np.float32(float(out["total_loss"].detach()))
        for m, gr in grads.items():
            payload[name + "/grad_" + m] = gr.float().numpy()
        print(name, float(loss), sorted(grads))
    np.savez_compressed(os.path.join(GOLDEN, "sdm_alignment.npz"), **payload)


if __name__ == "__main__":
    main()
