"""Drop-in for the reference's models/sdm_loss.py.

`sdm_loss_stable(qry, gal, y, tau=0.2, eps=1e-8)` keeps the reference signature (sdm_loss.py:13).
Forward and backward are one CUDA launch each (csrc/sdm.cu); `sdm_loss_pairs` runs all modality
pairs of a training step (the four `sdm_loss_stable` calls of models/model.py:586-622) in ONE
forward launch and ONE backward launch.  No host synchronisation happens on the numeric path: the
reference's guards (sdm_loss.py:79-81, 89-91, 105-106, 145-147) are evaluated on the device.

Documented differences from the reference (non-numeric):
  * no prints and no `_last_*_time` function attributes (sdm_loss.py:108-139);
  * the result is always fp32 and always attached to the graph; on a guard path the reference
    returns a zero WITHOUT grad_fn (so inputs get no gradient), here the value is the same zero and
    the gradients are exact zeros.
"""
from typing import List, Sequence

import torch

from . import _cabi
from ._cabi import SdmPair, check, stream_ptr


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _cabi.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return _cabi.DTYPE_BF16
    raise TypeError("sdm_loss: features must be float32 or bfloat16 (got %s)" % t.dtype)


def _make_pairs(qrys, gals, ys, losses, status, saved, grad=None, dq=None, dg=None):
    n = len(qrys)
    arr = (SdmPair * n)()
    for i in range(n):
        arr[i].qry = qrys[i].data_ptr(); arr[i].gal = gals[i].data_ptr(); arr[i].y = ys[i].data_ptr()
        arr[i].N = qrys[i].shape[0]; arr[i].M = gals[i].shape[0]
        arr[i].loss = losses[i:i + 1].data_ptr(); arr[i].status = status[i:i + 1].data_ptr()
        arr[i].saved = saved[i].data_ptr()
        if grad is not None:
            arr[i].grad_out = grad[i:i + 1].data_ptr()
            arr[i].dqry = dq[i].data_ptr(); arr[i].dgal = dg[i].data_ptr()
    return arr


class _SdmPairsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tau, eps, n, *tensors):
        qrys, gals, ys = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        L = _cabi.lib()
        dev = qrys[0].device
        d = qrys[0].shape[1]
        code = _dtype_code(qrys[0])
        qs = [q.detach().contiguous() for q in qrys]
        gs = [g.detach().contiguous() for g in gals]
        yy = [y.detach().to(torch.float32).contiguous() for y in ys]
        for q, g, y in zip(qs, gs, yy):
            if q.dtype != qs[0].dtype or g.dtype != qs[0].dtype or q.shape[1] != d or g.shape[1] != d:
                raise TypeError("sdm_loss: all features of a batch must share dtype and width")
            if tuple(y.shape) != (q.shape[0], g.shape[0]):
                raise ValueError("sdm_loss: y must be [N, M]")
        losses = torch.empty(n, dtype=torch.float32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        saved = [torch.empty(L.reid_sdm_saved_floats(q.shape[0], g.shape[0], d), dtype=torch.float32, device=dev)
                 for q, g in zip(qs, gs)]
        arr = _make_pairs(qs, gs, yy, losses, status, saved)
        check(L.reid_sdm_fwd(arr, n, code, d, float(tau), float(eps), stream_ptr()), "reid_sdm_fwd")
        ctx.n, ctx.tau, ctx.eps, ctx.code, ctx.d = n, float(tau), float(eps), code, d
        ctx.keep = (qs, gs, yy, losses, status, saved)
        ctx.status = status
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        qs, gs, yy, losses, status, saved = ctx.keep
        L = _cabi.lib()
        grad = grad_losses.detach().to(torch.float32).contiguous()
        dq = [torch.empty_like(q) for q in qs]
        dg = [torch.empty_like(g) for g in gs]
        arr = _make_pairs(qs, gs, yy, losses, status, saved, grad, dq, dg)
        check(L.reid_sdm_bwd(arr, ctx.n, ctx.code, ctx.d, ctx.tau, ctx.eps, stream_ptr()), "reid_sdm_bwd")
        return (None, None, None) + tuple(dq) + tuple(dg) + (None,) * ctx.n


def sdm_loss_pairs(qrys: Sequence[torch.Tensor], gals: Sequence[torch.Tensor], ys: Sequence[torch.Tensor],
                   tau: float = 0.2, eps: float = 1e-8) -> torch.Tensor:
    """All modality pairs in one launch: returns losses [n_pairs] fp32 (differentiable)."""
    n = len(qrys)
    if not (n == len(gals) == len(ys)) or n == 0:
        raise ValueError("sdm_loss_pairs: need equally long, non-empty lists")
    out: List[torch.Tensor] = []
    for s in range(0, n, _cabi.SDM_MAX_PAIRS):
        e = min(n, s + _cabi.SDM_MAX_PAIRS)
        out.append(_SdmPairsFn.apply(tau, eps, e - s, *qrys[s:e], *gals[s:e], *ys[s:e]))
    return out[0] if len(out) == 1 else torch.cat(out)


def sdm_loss_stable(qry, gal, y, tau=0.2, eps=1e-8):
    """models/sdm_loss.py:13 -- symmetric SDM cross-entropy, scalar fp32 tensor."""
    if not qry.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        return sdm_loss_stable(qry.to(dev), gal.to(dev), y.to(dev), tau, eps).to(qry.device)
    return sdm_loss_pairs([qry], [gal], [y.to(qry.device)], tau, eps)[0]
