"""CPU restatement of the reference SDM loss.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/sdm_loss.py:13-149 (`sdm_loss_stable`), without the wall-clock
throttled prints and without the `_last_*_time` function attributes (non-numeric behaviour).
Numeric behaviour kept, by reference line:
  :28      tau clamp to [0.15, 0.5]
  :31-32   F.normalize(dim=1, eps) IN THE INPUT DTYPE
  :75      .float()
  :79-81   non-finite feature  -> non-differentiable 0 in qry.dtype
  :86      S = q g^T / tau      :89-91 non-finite S -> 0      :94 clamp(-20, 20)
  :101-106 no row with a positive -> zeros([])
  :34-70   one-sided CE over rows with >=1 positive (target uniform over positives), mean over
           valid rows, non-finite rows dropped
  :121-123 0.5 * (CE(S, y) + CE(S^T, y^T))
  :145-147 NaN / Inf / negative -> non-differentiable 0

`sdm_loss_oracle` is differentiable through torch autograd exactly like the reference.
`sdm_fwd_bwd_f64` is an independent float64 closed form (loss + analytic gradients) used to
cross-check both the reference autograd and the CUDA kernels.
Pinned by oracle/make_golden.py -> tests/golden/sdm_*.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F


def effective_tau(tau: float) -> float:
    return max(0.15, min(0.5, tau))          # sdm_loss.py:28


def _one_side_ce(S, y):
    # sdm_loss.py:34-70
    row_pos = y.sum(dim=1)
    valid = row_pos > 0
    if not bool(valid.any()):
        return torch.tensor(0.0, device=S.device, dtype=S.dtype)
    S_valid = S[valid].clamp(-20.0, 20.0)
    y_valid = y[valid]
    pos = (y_valid > 0).float()
    pos_sum = pos.sum(dim=1, keepdim=True).clamp_min(1.0)
    q = pos / pos_sum
    log_p = F.log_softmax(S_valid, dim=1)
    ce = -(q * log_p).sum(dim=1)
    fin = torch.isfinite(ce)
    if not bool(fin.all()):
        ce = ce[fin]
        if ce.numel() == 0:
            return torch.tensor(0.0, device=S.device, dtype=S.dtype)
    return ce.mean()


def sdm_loss_oracle(qry, gal, y, tau=0.2, eps=1e-8):
    t = effective_tau(tau)
    qry = F.normalize(qry, dim=1, eps=eps)   # :31-32 (input dtype)
    gal = F.normalize(gal, dim=1, eps=eps)
    qf, gf = qry.float(), gal.float()        # :75
    for feat in (qf, gf):                    # :78-81
        if not bool(torch.isfinite(feat).all()):
            return torch.tensor(0.0, device=qry.device, dtype=qry.dtype)
    S = qf @ gf.t() / t                      # :86
    if not bool(torch.isfinite(S).all()):    # :89-91
        return torch.tensor(0.0, device=qry.device, dtype=qry.dtype)
    S = torch.clamp(S, min=-20.0, max=20.0)  # :94
    pos_cnt = y.sum(dim=1)                   # :101-106
    if int((pos_cnt == 0).sum()) == pos_cnt.numel():
        return torch.zeros([], device=qry.device, dtype=qry.dtype)
    res = 0.5 * (_one_side_ce(S, y) + _one_side_ce(S.t(), y.t()))   # :121-123
    if bool(torch.isnan(res)) or bool(torch.isinf(res)) or bool(res < 0):   # :145-147
        return torch.tensor(0.0, device=qry.device, dtype=qry.dtype)
    return res


def sdm_fwd_bwd_f64(qry, gal, y, tau=0.2, eps=1e-8):
    """float64 closed form: returns (loss, dqry, dgal) as numpy arrays (grad of loss, upstream 1).

    dL/dS = 0.5 * [ 1_R (softmax_row(S) - q_row) / |R|  +  1_C (softmax_col(S) - q_col) / |C| ]
    (SURVEY.md section 8a row S6); clamp is treated as inactive (|S| <= 1/0.15 for unit rows).
    Inputs are first rounded through the reference's dtype path (normalise in input dtype).
    """
    t = effective_tau(tau)
    qn = F.normalize(qry, dim=1, eps=eps).double().numpy()
    gn = F.normalize(gal, dim=1, eps=eps).double().numpy()
    Y = (y.double().numpy() > 0).astype(np.float64)
    S = qn @ gn.T / t
    def side(S, Y):
        cnt = Y.sum(1)
        v = cnt > 0
        if not v.any():
            return 0.0, np.zeros_like(S)
        m = S.max(1, keepdims=True)
        lse = m + np.log(np.exp(S - m).sum(1, keepdims=True))
        logp = S - lse
        q = Y / np.maximum(cnt, 1.0)[:, None]
        ce = -(q * logp).sum(1)
        n = v.sum()
        dS = (np.exp(logp) - q) * v[:, None] / n
        return ce[v].mean(), dS
    l1, d1 = side(S, Y)
    l2, d2 = side(S.T, Y.T)
    loss = 0.5 * (l1 + l2)
    dS = 0.5 * (d1 + d2.T) / t
    dqn = dS @ gn
    dgn = dS.T @ qn
    # normalisation Jacobian, with the reference's input-dtype rounding ignored (float64 of raw)
    def back_norm(x_raw, xn, dxn):
        x = x_raw.double().numpy()
        nrm = np.maximum(np.linalg.norm(x, axis=1, keepdims=True), eps)
        xh = x / nrm
        return (dxn - xh * (dxn * xh).sum(1, keepdims=True)) / nrm
    return loss, back_norm(qry, qn, dqn), back_norm(gal, gn, dgn)


def sdm_alignment_oracle(raw_modality_features, feature_masks, labels, tau=0.2):
    """CPU restatement of the SDM section of compute_loss, /root/reference/models/model.py:556-625
    (mask filtering :570-602, y :605, no-positive skip :608-613, finite check :617-618, mean :621-625)."""
    vis = raw_modality_features.get("vis"); vmask = feature_masks.get("vis")
    if vis is None or vmask is None:
        return torch.tensor(0.0)
    vidx = (vmask > 0).squeeze(-1) if vmask.dim() > 1 else (vmask > 0)
    if vidx.sum() == 0:
        return torch.tensor(0.0)
    vfeat, vlab = vis[vidx], labels[vidx]
    out = []
    for name, feat in raw_modality_features.items():
        if name == "vis":
            continue
        mask = feature_masks.get(name)
        if feat is None or mask is None:
            continue
        idx = (mask > 0).squeeze(-1) if mask.dim() > 1 else (mask > 0)
        if idx.sum() == 0:
            continue
        y = (labels[idx].view(-1, 1) == vlab.view(1, -1)).float()
        if y.numel() == 0 or y.sum() == 0:
            continue
        L = sdm_loss_oracle(feat[idx], vfeat, y, tau=tau)
        if torch.isfinite(L):
            out.append(L)
    return torch.stack(out).mean() if out else torch.tensor(0.0)
