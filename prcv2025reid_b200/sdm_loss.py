"""Drop-in for the reference's models/sdm_loss.py.

`sdm_loss_stable(qry, gal, y, tau=0.2, eps=1e-8)` keeps the reference signature (sdm_loss.py:13).
`sdm_loss_pairs` runs all modality pairs of a training step (the four `sdm_loss_stable` calls of
models/model.py:586-622) together: one forward and one backward launch on the fp32 CUDA-core path
(csrc/sdm.cu: fp32 inputs, small or odd shapes), normalise + forward and one backward launch on the
bf16 tcgen05 path (csrc/sdm_tc.cu: 64 <= N, M <= 512).  No host synchronisation happens on the numeric path: the
reference's guards (sdm_loss.py:79-81, 89-91, 105-106, 145-147) are evaluated on the device.

Documented differences from the reference (non-numeric):
  * no prints and no `_last_*_time` function attributes (sdm_loss.py:108-139);
  * the result is always fp32 and always attached to the graph; on a guard path the reference
    returns a zero WITHOUT grad_fn (so inputs get no gradient), here the value is the same zero and
    the gradients are exact zeros.
"""
import ctypes
from typing import List, Sequence

import torch

from . import _cabi
from ._cabi import check


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _cabi.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return _cabi.DTYPE_BF16
    if t.dtype == torch.float16:
        return _cabi.DTYPE_F16
    raise TypeError("sdm_loss: features must be float32, bfloat16 or float16 (got %s)" % t.dtype)


_SAVED_FLOATS = {}


def _saved_floats(L, N, M, d):
    key = (N, M, d)
    v = _SAVED_FLOATS.get(key)
    if v is None:
        v = _SAVED_FLOATS[key] = (L.reid_sdm_saved_floats(N, M, d) + 63) // 64 * 64
    return v


def _raw_stream(dev):
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device()))


PAIR_WORDS = 14      # sizeof(reid_sdm_pair) / 8


def _pair_words(qs, gs, ys, losses, status, saved, labels=None):
    """The `reid_sdm_pair` array (include/reid_b200.h) as PAIR_WORDS 64-bit words per pair (N and M share one);
    the three backward-only slots stay zero.  ys[i] is None in the label form: labels[i] = (row_label, col_label,
    row_valid or None, col_valid or None)."""
    words = []
    lp, sp = losses.data_ptr(), status.data_ptr()
    for i, (q, g, y) in enumerate(zip(qs, gs, ys)):
        lab = [0, 0, 0, 0]
        if y is None:
            lab = [t.data_ptr() if t is not None else 0 for t in labels[i]]
        words += [q.data_ptr(), g.data_ptr(), y.data_ptr() if y is not None else 0, q.shape[0] | (g.shape[0] << 32),
                  lp + 4 * i, sp + 4 * i, saved[i].data_ptr(), 0, 0, 0] + lab
    return words


def _pair_table(qs, gs, ys, losses, status, saved, grad=None, dq=None, dg=None, labels=None):
    words = _pair_words(qs, gs, ys, losses, status, saved, labels)
    if grad is not None:
        gp = grad.data_ptr()
        for i in range(len(qs)):
            words[PAIR_WORDS * i + 7] = gp + 4 * i
            words[PAIR_WORDS * i + 8] = dq[i].data_ptr()
            words[PAIR_WORDS * i + 9] = dg[i].data_ptr()
    return (ctypes.c_uint64 * len(words))(*words)


class _SdmPairsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tau, eps, n, *tensors):
        qrys, gals, ys = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        labels = tensors[3 * n] if len(tensors) > 3 * n else None      # label form: list of (row_label, col_label, row_valid, col_valid)
        align = len(tensors) > 3 * n + 1 and tensors[3 * n + 1] == "align"   # reduce like models/model.py:608-625 (see below)
        L = _cabi.lib()
        dev = qrys[0].device
        d = qrys[0].shape[1]
        dt = qrys[0].dtype
        code = _dtype_code(qrys[0])
        qs = [q if q.is_contiguous() else q.contiguous() for q in qrys]
        gs = [g if g.is_contiguous() else g.contiguous() for g in gals]
        yy = [None if y is None else (y if (y.dtype == torch.float32 and y.is_contiguous()) else y.to(torch.float32).contiguous())
              for y in ys]
        for q, g, y in zip(qs, gs, yy):
            if q.dtype != dt or g.dtype != dt or q.shape[1] != d or g.shape[1] != d:
                raise TypeError("sdm_loss: all features of a batch must share dtype and width")
            if y is not None and (y.shape[0] != q.shape[0] or y.shape[1] != g.shape[0]):
                raise ValueError("sdm_loss: y must be [N, M]")
        losses = torch.empty(n, dtype=torch.float32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        sizes = [_saved_floats(L, q.shape[0], g.shape[0], d) for q, g in zip(qs, gs)]
        saved = torch.empty(sum(sizes), dtype=torch.float32, device=dev).split(sizes)
        words = _pair_words(qs, gs, yy, losses, status, saved, labels)
        arr = (ctypes.c_uint64 * len(words))(*words)
        if L.reid_sdm_uses_tensor_cores(arr, n, code, d):
            _cabi.LAUNCH_COUNT["n"] += 1                              # normalise/pack + forward = 2 launches
        check(L.reid_sdm_fwd(arr, n, code, d, float(tau), float(eps), _raw_stream(dev)), "reid_sdm_fwd")
        ctx.n, ctx.tau, ctx.eps, ctx.code, ctx.d = n, float(tau), float(eps), code, d
        # NOTE: the output must not be reachable from ctx (output -> grad_fn -> ctx -> output would keep every
        # step's buffers alive until the cyclic GC runs); the backward only needs the device-side status bits
        for i in range(n):
            words[PAIR_WORDS * i + 4] = words[PAIR_WORDS * i + 5]     # (the loss slot is not written again)
        ctx.keep = (qs, gs, yy, status, saved, words, labels)
        ctx.n_in = len(tensors)
        ctx.align = None
        ctx.mark_non_differentiable(status)
        if align:
            # the tail of the compute_loss section inside the Function (no autograd nodes for it): pairs without a positive
            # (status bit 3) or with a non-finite loss are skipped, the rest averaged (:608-625); d total / d loss[p] =
            # has_pos[p] / count is applied in the backward
            has_pos = ((status & 8) == 0) & torch.isfinite(losses)
            cnt = has_pos.sum().clamp_min(1)
            ctx.align = (has_pos, cnt)
            return torch.where(has_pos, losses, torch.zeros_like(losses)).sum() / cnt, status
        return losses, status

    @staticmethod
    def backward(ctx, grad_losses, _grad_status=None):
        qs, gs, yy, status, saved, words, labels = ctx.keep
        L = _cabi.lib()
        n = ctx.n
        grad = grad_losses
        if ctx.align is not None:
            has_pos, cnt = ctx.align
            grad = torch.where(has_pos, grad.to(torch.float32) / cnt, torch.zeros((), dtype=torch.float32, device=grad.device))
        if grad.dtype != torch.float32 or not grad.is_contiguous():
            grad = grad.to(torch.float32).contiguous()
        q0, g0 = qs[0], gs[0]
        if all(q.shape == q0.shape for q in qs) and all(g.shape == g0.shape for g in gs):
            dq = torch.empty((n,) + tuple(q0.shape), dtype=q0.dtype, device=q0.device).unbind(0)
            dg = torch.empty((n,) + tuple(g0.shape), dtype=g0.dtype, device=g0.device).unbind(0)
        else:
            nq = [q.numel() for q in qs]
            ng = [g.numel() for g in gs]
            parts = torch.empty(sum(nq) + sum(ng), dtype=q0.dtype, device=q0.device).split(nq + ng)
            dq = [p.view(q.shape) for p, q in zip(parts[:n], qs)]
            dg = [p.view(g.shape) for p, g in zip(parts[n:], gs)]
        gp = grad.data_ptr()
        w = list(words)
        for i in range(n):
            w[PAIR_WORDS * i + 7] = gp + 4 * i
            w[PAIR_WORDS * i + 8] = dq[i].data_ptr()
            w[PAIR_WORDS * i + 9] = dg[i].data_ptr()
        arr = (ctypes.c_uint64 * len(w))(*w)
        check(L.reid_sdm_bwd(arr, n, ctx.code, ctx.d, ctx.tau, ctx.eps, _raw_stream(q0.device)), "reid_sdm_bwd")
        return (None, None, None) + tuple(dq) + tuple(dg) + (None,) * (ctx.n_in - 2 * n)


class SdmStep:
    """Forward + backward of all pairs of a training step through ONE C call (`reid_sdm_step`), without autograd.

    The objective is `sum_p weights[p] * loss[p]` (weights default to 1: the `losses.sum()` of a step; the mean of
    models/model.py:622 is weights = 1/n).  Buffers are allocated once: `losses` [n] fp32, `dq[p]` / `dg[p]` in the
    input dtype hold the step's results after `run()`.  On the small-batch path the step is a single kernel launch;
    on the tcgen05 path three (pack, forward, backward)."""

    def __init__(self, qrys, gals, ys, tau=0.2, eps=1e-8, weights=None, labels=None):
        """`labels` (label form, `ys` = None or a list of None): per pair (row_label, col_label, row_valid or None,
        col_valid or None) as `sdm_loss_pairs_labels` takes them; the tensors are kept (and read at every `run`)."""
        n = len(qrys)
        if ys is None:
            ys = [None] * n
        if not (n == len(gals) == len(ys)) or n == 0 or n > _cabi.SDM_MAX_PAIRS:
            raise ValueError("SdmStep: need equally long lists of 1..%d pairs" % _cabi.SDM_MAX_PAIRS)
        L = _cabi.lib()
        dev, dt, d = qrys[0].device, qrys[0].dtype, qrys[0].shape[1]
        self.code = _dtype_code(qrys[0])
        self.q = [q.detach().contiguous() for q in qrys]
        self.g = [g.detach().contiguous() for g in gals]
        self.y = [None if y is None else y.detach().to(torch.float32).contiguous() for y in ys]
        self.labels = None
        if any(y is None for y in self.y):
            if labels is None or len(labels) != n or not all(y is None for y in self.y):
                raise ValueError("SdmStep: give y for every pair, or labels for every pair")
            self.labels = _label_tuples(self.q, self.g, [l[0] for l in labels], [l[1] for l in labels],
                                        [l[2] for l in labels], [l[3] for l in labels])
        for q, g, y in zip(self.q, self.g, self.y):
            if q.dtype != dt or g.dtype != dt or q.shape[1] != d or g.shape[1] != d:
                raise TypeError("sdm_loss: all features of a batch must share dtype and width")
            if y is not None and (y.shape[0] != q.shape[0] or y.shape[1] != g.shape[0]):
                raise ValueError("sdm_loss: y must be [N, M]")
        self.n, self.d, self.tau, self.eps = n, d, float(tau), float(eps)
        self.losses = torch.empty(n, dtype=torch.float32, device=dev)
        self.status = torch.empty(n, dtype=torch.int32, device=dev)
        self.weights = (torch.ones(n, dtype=torch.float32, device=dev) if weights is None
                        else weights.detach().to(device=dev, dtype=torch.float32).contiguous().clone())
        sizes = [_saved_floats(L, q.shape[0], g.shape[0], d) for q, g in zip(self.q, self.g)]
        self.saved = torch.empty(sum(sizes), dtype=torch.float32, device=dev).split(sizes)
        self.dq = [torch.empty_like(q) for q in self.q]
        self.dg = [torch.empty_like(g) for g in self.g]
        self.table = _pair_table(self.q, self.g, self.y, self.losses, self.status, self.saved, self.weights, self.dq, self.dg,
                                 self.labels)
        self.launches = L.reid_sdm_step_launches(self.table, n, self.code, d)

    def run(self):
        """Enqueue the step on the current stream; returns `losses` (results are stream-ordered, no host sync)."""
        _cabi.LAUNCH_COUNT["n"] += self.launches
        check(_cabi.lib().reid_sdm_step(self.table, self.n, self.code, self.d, self.tau, self.eps, _raw_stream(self.losses.device)),
              "reid_sdm_step")
        return self.losses


class SdmGraphStep:
    """One SDM training step (all pairs: forward, objective = sum of the pair losses, backward) as a CUDA graph.

    The step is launch-bound at the reference's batch sizes (P x K = 4 x 2: a few KB of data), so the whole
    forward + backward (`SdmStep`: one kernel for small pairs, pack + forward + backward on the tcgen05 path; no
    autograd glue kernels) is recorded once and replayed with a single graph launch.  Inputs are copied into the
    graph's static buffers (`load`), gradients are read from `dq` / `dg`."""

    def __init__(self, qrys, gals, ys, tau=0.2, eps=1e-8, warmup=3, weights=None):
        self.step = SdmStep([q.detach().clone() for q in qrys], [g.detach().clone() for g in gals],
                            [y.detach().clone() for y in ys], tau, eps, weights)
        self.q, self.g, self.y = self.step.q, self.step.g, self.step.y
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step.run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = self.step.run()
        self.dq, self.dg = self.step.dq, self.step.dg
        self.launches = self.step.launches

    def load(self, qrys, gals, ys=None):
        with torch.no_grad():
            for dst, src in zip(self.q + self.g, list(qrys) + list(gals)):
                dst.copy_(src)
            if ys is not None:
                for dst, src in zip(self.y, ys):
                    dst.copy_(src)

    def replay(self):
        _cabi.LAUNCH_COUNT["n"] += self.launches
        self.graph.replay()
        return self.losses


def sdm_loss_pairs(qrys: Sequence[torch.Tensor], gals: Sequence[torch.Tensor], ys: Sequence[torch.Tensor],
                   tau: float = 0.2, eps: float = 1e-8) -> torch.Tensor:
    """All modality pairs in one launch: returns losses [n_pairs] fp32 (differentiable)."""
    n = len(qrys)
    if not (n == len(gals) == len(ys)) or n == 0:
        raise ValueError("sdm_loss_pairs: need equally long, non-empty lists")
    out: List[torch.Tensor] = []
    for s in range(0, n, _cabi.SDM_MAX_PAIRS):
        e = min(n, s + _cabi.SDM_MAX_PAIRS)
        out.append(_SdmPairsFn.apply(tau, eps, e - s, *qrys[s:e], *gals[s:e], *ys[s:e])[0])
    return out[0] if len(out) == 1 else torch.cat(out)


def label_form_supported(n_rows: int, d: int, dtype) -> bool:
    """Shapes the label form of the C entry points serves (include/reid_b200.h, reid_sdm_pair): every code path of the
    library takes it -- the tcgen05 path (bf16, 64..512 rows), the one-CTA small-batch kernels and the general CUDA-core
    kernels (csrc/sdm.cu; round 2: the tcgen05 path only)."""
    return dtype in (torch.float32, torch.bfloat16, torch.float16) and n_rows >= 1 and d >= 1


def _label_tuples(qrys, gals, row_labels, col_labels, row_valid=None, col_valid=None):
    labs = []
    for i in range(len(qrys)):
        rl = row_labels[i].to(torch.int64).contiguous()
        cl = col_labels[i].to(torch.int64).contiguous()
        rv = None if row_valid is None or row_valid[i] is None else row_valid[i].to(torch.uint8).contiguous()
        cv = None if col_valid is None or col_valid[i] is None else col_valid[i].to(torch.uint8).contiguous()
        if rl.numel() != qrys[i].shape[0] or cl.numel() != gals[i].shape[0]:
            raise ValueError("sdm_loss_pairs_labels: one label per feature row")
        if (rv is not None and rv.numel() != rl.numel()) or (cv is not None and cv.numel() != cl.numel()):
            raise ValueError("sdm_loss_pairs_labels: one valid byte per feature row")
        labs.append((rl, cl, rv, cv))
    return labs


def sdm_loss_pairs_labels(qrys: Sequence[torch.Tensor], gals: Sequence[torch.Tensor], row_labels, col_labels,
                          row_valid=None, col_valid=None, tau: float = 0.2, eps: float = 1e-8):
    """The pairs of a step in the LABEL FORM: y is never materialised, y[i][j] = valid_i & valid_j & (label_i == label_j),
    rows whose valid byte is 0 are left out of the loss altogether (the feature masks of models/model.py:570-602).
    -> (losses [n] fp32 differentiable, status [n] int32: bit 3 = the pair has no positive, include/reid_b200.h)."""
    n = len(qrys)
    if n == 0 or n > _cabi.SDM_MAX_PAIRS or not (n == len(gals) == len(row_labels) == len(col_labels)):
        raise ValueError("sdm_loss_pairs_labels: need equally long lists of 1..%d pairs" % _cabi.SDM_MAX_PAIRS)
    labs = _label_tuples(qrys, gals, row_labels, col_labels, row_valid, col_valid)
    return _SdmPairsFn.apply(tau, eps, n, *qrys, *gals, *([None] * n), labs)


def sdm_loss_stable(qry, gal, y, tau=0.2, eps=1e-8):
    """models/sdm_loss.py:13 -- symmetric SDM cross-entropy, scalar fp32 tensor."""
    if not qry.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        return sdm_loss_stable(qry.to(dev), gal.to(dev), y.to(dev), tau, eps).to(qry.device)
    return sdm_loss_pairs([qry], [gal], [y.to(qry.device)], tau, eps)[0]


def sdm_alignment_loss(raw_modality_features, feature_masks, labels, tau=0.2, eps=1e-8):
    """The SDM section of `CLIPBasedMultiModalReIDModel.compute_loss` (models/model.py:556-622): every non-vis
    modality is aligned to vis over the rows whose feature mask is set, pairs without a positive are skipped and
    the remaining per-modality losses are averaged (zero when none remains).

    The reference issues one `sdm_loss_stable` call and >= 3 host synchronisations per modality (`.sum() == 0`
    at :572/:597/:608 plus those inside the loss).  Here all modalities go through ONE launch sequence in the LABEL FORM
    of the kernels: labels [B], masks, features [B, d] in; no y, no row filtering, NO host synchronisation, whatever the
    batch size -- the tcgen05 path for bf16 batches of 64..512 rows, the CUDA-core kernels otherwise.  Features keep the
    dtype the caller passes (bf16 / fp16 under the training autocast of train.py:852, fp32 otherwise): like the reference,
    the loss normalises in that dtype (sdm_loss.py:31-32) and only then computes in fp32 (:75)."""
    vis = raw_modality_features.get("vis")
    vmask = feature_masks.get("vis")
    dev = labels.device
    zero = torch.tensor(0.0, device=dev, dtype=torch.float32)
    if vis is None or vmask is None:
        return zero                                                              # :566-568
    names = [m for m, f in raw_modality_features.items()
             if m != "vis" and f is not None and feature_masks.get(m) is not None]                # :586-592
    B = labels.shape[0]
    if not names or B == 0:
        return zero                                                              # nothing to align
    if vis.dim() != 2 or vis.shape[0] != B:
        raise ValueError("sdm_alignment_loss: vis features must be [B, d] with one row per label")
    for m in names:
        f = raw_modality_features[m]
        if f.dtype != vis.dtype:
            raise TypeError("sdm_alignment_loss: %s features are %s but vis features are %s" % (m, f.dtype, vis.dtype))
        if f.shape != vis.shape:
            raise ValueError("sdm_alignment_loss: %s features are %s but vis features are %s" % (m, tuple(f.shape), tuple(vis.shape)))
    _dtype_code(vis)                                                             # (fp32 / bf16 / fp16, else TypeError)
    # all feature masks as ONE [1 + n, B] byte tensor (three launches instead of three per modality); row i is pair i's
    # row_valid, row 0 (vis) every pair's col_valid
    cols = [feature_masks[m].reshape(B, -1)[:, 0] for m in ["vis"] + names]
    if all(c.dtype == cols[0].dtype for c in cols):
        valid = (torch.stack(cols) > 0).to(torch.uint8)
    else:
        valid = torch.stack([c > 0 for c in cols]).to(torch.uint8)
    lab = labels.to(torch.int64).contiguous()
    # masked rows are left out inside the kernels (a modality or a vis side without a valid row ends as "no positive",
    # :572-574 / :597-598); "no positive" / finiteness (:608-618) come back as device-side status bits
    n = len(names)
    if n <= _cabi.SDM_MAX_PAIRS:
        labs = [(lab, lab, valid[1 + i], valid[0]) for i in range(n)]
        return _SdmPairsFn.apply(tau, eps, n, *[raw_modality_features[m] for m in names], *([vis] * n), *([None] * n),
                                 labs, "align")[0]
    out = []
    for s0 in range(0, n, _cabi.SDM_MAX_PAIRS):
        k = len(names[s0:s0 + _cabi.SDM_MAX_PAIRS])
        labs = [(lab, lab, valid[1 + s0 + i], valid[0]) for i in range(k)]
        out.append(_SdmPairsFn.apply(tau, eps, k, *[raw_modality_features[m] for m in names[s0:s0 + k]], *([vis] * k),
                                     *([None] * k), labs))
    losses = torch.cat([o[0] for o in out])
    status = torch.cat([o[1] for o in out])
    has_pos = ((status & 8) == 0) & torch.isfinite(losses)                       # :608-618
    kept = torch.where(has_pos, losses, torch.zeros_like(losses))                # (a skipped loss must not leak a NaN)
    return kept.sum() / has_pos.sum().clamp_min(1)                               # :621-625
