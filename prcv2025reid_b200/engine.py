"""Tensor-level host side of the retrieval + evaluation path (one gallery shard per process).

Mirrors the math of tools/eval_mm_protocol.py (reference) on device tensors:
    prepare_gallery   <- build_gallery / extract_gallery_feats cache + `g_feats = l2n(g_feats)` (:546)
    fuse_queries      <- extract_query_feat (:328-365), batched
    retrieve          <- the body of rank_and_metrics (:396-469) for a batch of queries
PyTorch is used only for device memory, streams and (optionally) torch.distributed; every
arithmetic step is a call into libreid_b200.so.  There is no CPU / eager fallback.
"""
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _cabi, sharding
from ._cabi import check, ptr, stream_ptr

# bound on |fp16 tensor-core score - fp32 score| for unit-norm rows: both operands are rounded to
# fp16 (relative 2^-11 each), so |err| <= (2*2^-11 + 2^-22) * sum|q_i g_i| <= 2^-10 (Cauchy-Schwarz),
# plus fp32 accumulation slack.  Top-k lists and CMC never depend on it: queries whose top-k / CMC cannot be
# decided within this bound are re-run through the all-fp32 kernel.  AP does: ranks beyond the re-scored head are
# counted on fp16 scores (a row within ~1e-5 of a positive's score may fall on either side of it) and, for deep
# positives of large shards, on row samples (include/reid_b200.h, reid_retrieve_fused) -- per-query AP moves by
# <= ~2e-3 in the worst case and mAP by ~1e-5 (asserted against the oracle in tests/test_fused_oracle_gpu.py);
# `exact_ap=True` switches the sampling off.
EPS_FP16 = 2.0 ** -10 + 2.0 ** -13
WAVE_QUERIES = 256           # queries per work item of the fused kernel (one CTA pair)
_DEBUG_KEEP = None   # set to a dict to keep the flags / candidate counts of the last block (debug scripts)


def _f32c(t):
    return t.detach().to(dtype=torch.float32).contiguous()


@dataclass
class GalleryShard:
    g_f32: torch.Tensor        # [G_local, D] fp32, L2-normalised rows
    g_f16: torch.Tensor        # [G_local, D] fp16 copy (tensor-core operand)
    g_code: torch.Tensor       # [G_local] int32 pid code of every local row
    sorted_pid: torch.Tensor   # [G_total] int64, gallery pids sorted
    order: torch.Tensor        # [G_total] int32, gallery row of every sorted position
    pmax: int                  # largest number of gallery rows sharing one pid
    g_offset: int
    G_total: int
    _bufs: dict = None         # work buffers reused across query blocks / calls (no allocator churn)
    img_index: dict = None     # img_id -> gallery rows (set by eval_mm_protocol.install_gallery: the same-image rule)

    def buf(self, name, shape, dtype, zero=False):
        """A cached scratch tensor; contents are undefined unless zero=True.  Stream-ordered reuse only."""
        if self._bufs is None:
            self._bufs = {}
        t = self._bufs.get(name)
        shape = tuple(int(x) for x in shape)
        if t is None or t.dtype != dtype or t.numel() < int(torch.Size(shape).numel()) or t.device != self.g_f32.device:
            t = torch.empty(max(1, int(torch.Size(shape).numel())), dtype=dtype, device=self.g_f32.device)
            self._bufs[name] = t
        v = t[:int(torch.Size(shape).numel())].view(shape)
        if zero:
            v.zero_()
        return v

    @property
    def G_local(self):
        return self.g_f32.shape[0]

    @property
    def d(self):
        return self.g_f32.shape[1]


def l2norm_rows(x: torch.Tensor, want_f16: bool = False, eps: float = 1e-12):
    """K1.  x [rows, D] fp32 cuda -> (fp32 normalised, fp16 copy or None)."""
    x = _f32c(x)
    rows, d = x.shape
    out = torch.empty_like(x)
    out16 = torch.empty(rows, d, dtype=torch.float16, device=x.device) if want_f16 else None
    check(_cabi.lib().reid_l2norm_rows(ptr(x), ptr(out), ptr(out16), rows, d, eps, stream_ptr()), "reid_l2norm_rows")
    return out, out16


def fuse_queries(query_raw: torch.Tensor, mod_id: torch.Tensor, weights: torch.Tensor):
    """K2.  query_raw [Q,k,D] fp32, mod_id [Q,k] int32 (index into weights, <0 = empty slot),
    weights [n_mod] fp32 -> (q_f32 [Q,D], q_f16 [Q,D])."""
    query_raw = _f32c(query_raw)
    Q, k, d = query_raw.shape
    mod_id = mod_id.to(device=query_raw.device, dtype=torch.int32).contiguous()
    weights = weights.to(device=query_raw.device, dtype=torch.float32).contiguous()
    q32 = torch.empty(Q, d, dtype=torch.float32, device=query_raw.device)
    q16 = torch.empty(Q, d, dtype=torch.float16, device=query_raw.device)
    check(_cabi.lib().reid_mm_fuse_normalize(ptr(query_raw), ptr(mod_id), ptr(weights), weights.numel(), ptr(q32),
                                             ptr(q16), Q, k, d, stream_ptr()), "reid_mm_fuse_normalize")
    return q32, q16


def cosine_sim_f16(q_f16: torch.Tensor, g_f16: torch.Tensor) -> torch.Tensor:
    """K3 (unfused).  [Q,D] x [G,D] fp16 -> S [Q,G] fp32 via the tcgen05 GEMM."""
    Q, d = q_f16.shape
    G = g_f16.shape[0]
    S = torch.empty(Q, G, dtype=torch.float32, device=q_f16.device)
    check(_cabi.lib().reid_sim_gemm(ptr(q_f16), ptr(g_f16), ptr(S), Q, G, d, G, stream_ptr()), "reid_sim_gemm")
    return S


def prepare_gallery(gallery: torch.Tensor, g_pid_all: torch.Tensor, g_offset: int = 0) -> GalleryShard:
    """Normalise the local gallery rows and build the identity index over the WHOLE gallery's pids.

    gallery [G_local, D] (cuda, any float dtype), g_pid_all [G_total] int64 (replicated on every rank).
    """
    g32, g16 = l2norm_rows(gallery, want_f16=True)
    return install_normalised(g32, g16, g_pid_all, g_offset)


def install_normalised(g_f32: torch.Tensor, g_f16: torch.Tensor, g_pid_all: torch.Tensor, g_offset: int = 0) -> GalleryShard:
    """A shard whose rows are ALREADY normalised (K1 output kept on disk: gallery_store's pre-normalised store): only the
    identity index is built.  g_f32 [G_local, D] fp32, g_f16 its fp16 copy, g_pid_all [G_total] int64."""
    L = _cabi.lib()
    dev = g_f32.device
    assert g_f32.dtype == torch.float32 and g_f16.dtype == torch.float16 and g_f32.shape == g_f16.shape
    g_pid_all = g_pid_all.to(device=dev, dtype=torch.int64).contiguous()
    G_total = g_pid_all.numel()
    sorted_pid = torch.empty_like(g_pid_all)
    order = torch.empty(G_total, dtype=torch.int32, device=dev)
    max_run = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = L.reid_workspace_bytes(0, 0, G_total, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(L.reid_pid_index_build(ptr(g_pid_all), G_total, ptr(sorted_pid), ptr(order), ptr(max_run), ptr(ws),
                                 ws_bytes, stream_ptr()), "reid_pid_index_build")
    G_local = g_f32.shape[0]
    g_code = torch.empty(G_local, dtype=torch.int32, device=dev)
    local_pid = g_pid_all[g_offset:g_offset + G_local].contiguous()
    check(L.reid_pid_lookup(ptr(sorted_pid), G_total, ptr(local_pid), G_local, ptr(g_code), None, stream_ptr()),
          "reid_pid_lookup")
    pmax = int(max_run.item())          # one-time host read when a gallery is installed
    return GalleryShard(g_f32.contiguous(), g_f16.contiguous(), g_code, sorted_pid, order, max(1, pmax), int(g_offset), int(G_total))


def fused_slots(n_queries: int, G_local: int, sms: int) -> int:
    """Candidate slots per query (= gallery chunks of the query blocks in the last, partial wave of the persistent
    CTA pairs; reid_retrieve_fused runs all whole waves against the whole shard)."""
    units = max(1, sms // 2)
    left = (-(-n_queries // WAVE_QUERIES)) % units
    if left == 0:
        return 1
    return max(1, min(8, units // left, G_local // 4096))


def default_query_block(sms: int) -> int:
    """Two whole waves of the fused kernel's work items."""
    return 2 * max(1, sms // 2) * WAVE_QUERIES


def resident_query_blocks(Q: int, sms: int, G_local: int, cand_cap: int = 2048):
    """Query blocks [(start, end)] when the queries are already on the device.  Every launch of the gallery pass ends with
    a tail (the persistent CTA pairs of its last wave do not finish together) and carries a calibration launch of its own;
    on a 1M-row shard a wave takes ~14 ms and two waves per launch are plenty, on a 125k-row shard (8 GPUs) a wave takes
    ~2 ms and the tails of three launches cost ~3 % of the step: the launch grows as the shard shrinks (about 4M query x
    kilo-rows per launch), bounded by the candidate buffers (16 KB per query and slot, at most 12 GB per block).  Blocks are
    whole waves; a tail shorter than half a wave joins the last block."""
    wave = max(1, sms // 2) * WAVE_QUERIES
    waves = max(2, min(16, int(round(2.0 * 1_000_000 / max(1, G_local)))))
    while waves > 2 and waves * wave * int(cand_cap) * 8 * 2 > (12 << 30):
        waves -= 1
    block = waves * wave
    out, b0 = [], 0
    while b0 < Q:
        b1 = min(Q, b0 + block)
        if 0 < Q - b1 < wave // 2:
            b1 = Q
        out.append((b0, b1)); b0 = b1
    return out


def host_query_blocks(Q: int, sms: int):
    """Query blocks [(start, end)] for HOST-resident queries.  The upload of a block overlaps the kernels of the block
    before it, but nothing hides the upload of the FIRST block: it is kept small (16 work items' worth of queries, run as
    a chunked partial wave), the second block is one wave -- its upload (8 KB per query over PCIe) is shorter than the
    first block's gallery pass --, the rest are the usual two waves; a short tail joins the last block."""
    wave = max(1, sms // 2) * WAVE_QUERIES
    big = 2 * wave
    sizes, left = [], Q
    for n in (16 * WAVE_QUERIES, wave):
        if left > 0:
            sizes.append(min(left, n)); left -= sizes[-1]
    while left > 0:
        n = min(left, big)
        if 0 < left - n < big // 4:
            n = left                                        # (a tail shorter than half a wave is not worth a block of its own)
        sizes.append(n); left -= n
    out, b0 = [], 0
    for n in sizes:
        out.append((b0, b0 + n)); b0 += n
    return out


@dataclass
class RetrievalResult:
    metrics: Dict[str, float]
    top_idx: torch.Tensor            # [Q, topk] int32 global gallery index (-1 pad)
    top_score: torch.Tensor          # [Q, topk] fp32 exact scores
    ap: Optional[torch.Tensor]       # [Q] float64, -1 for skipped queries
    n_flagged: int                   # queries re-run through the exact kernel
    pos_above: torch.Tensor          # [Q, Pmax] int32 (gallery-wide)
    n_pos: torch.Tensor              # [Q] int32
    path: str = ""                   # "fused" (tcgen05) or "exact" (fp32 SIMT): which kernel ranked the queries


RESCORE_KX_ONE_SHARD = 32     # completeness cut-off of the re-scorer: the kx-th best approximate score of a shard ...
RESCORE_KX_SHARDED = 16       # ... with several shards the MAX over the shards' cut-offs is used and 16 suffice for top-10


class _RankState:
    """Per-call work arrays of `retrieve`, all sized for the WHOLE query batch so that the exchange steps over the shards
    run once per call instead of once per query block: q_code [Q], pos_thr [Q, Pmax] (exact positive scores, sorted),
    the selected candidates sel_* [Q, RTOP] and one flat int32 buffer [pos_above | lb0 | flag] (a single all-reduce)."""

    def __init__(self, Q, Pmax, dev):
        self.Q, self.Pmax = Q, Pmax
        self.counts = torch.zeros(Q * (Pmax + 2), dtype=torch.int32, device=dev)
        self.pos_above = self.counts[:Q * Pmax].view(Q, Pmax)            # rows ranked above each positive (this shard)
        self.lb0 = self.counts[Q * Pmax:Q * (Pmax + 1)]                  # re-scored rows above the best positive (this shard)
        self.flag = self.counts[Q * (Pmax + 1):]                         # candidate-buffer overflow
        self.n_pos = torch.empty(Q, dtype=torch.int32, device=dev)
        self.q_code = torch.empty(Q, dtype=torch.int32, device=dev)
        self.pos_thr = torch.empty(Q, Pmax, dtype=torch.float32, device=dev)
        self.top_score = torch.empty(Q, _cabi.RTOP, dtype=torch.float32, device=dev)
        self.top_idx = torch.empty(Q, _cabi.RTOP, dtype=torch.int32, device=dev)
        self.bound = torch.empty(Q, dtype=torch.float32, device=dev)     # completeness cut-off of the re-scored head (gallery-wide)
        self.sel_score = torch.empty(Q, _cabi.RTOP, dtype=torch.float32, device=dev)
        self.sel_idx = torch.empty(Q, _cabi.RTOP, dtype=torch.int32, device=dev)
        self.sel_n = torch.empty(Q, dtype=torch.int32, device=dev)


def _pos_stage(shard, S, sl, q32, pid, ex, E, *, group, world):
    """Exact fp32 scores of the positives of queries `sl` (sorted descending -> S.pos_thr, S.n_pos); with several shards the
    owner rank of a positive holds its score, the others -inf: all_reduce(MAX)."""
    L = _cabi.lib()
    st = stream_ptr()
    nb = pid.shape[0]
    q_code, pos_thr, n_pos = S.q_code[sl], S.pos_thr[sl], S.n_pos[sl]
    q_count = shard.buf("q_count", (nb,), torch.int32)
    check(L.reid_pid_lookup(ptr(shard.sorted_pid), shard.G_total, ptr(pid), nb, ptr(q_code), ptr(q_count), st),
          "reid_pid_lookup")
    check(L.reid_pos_scores(ptr(q32), ptr(shard.g_f32), ptr(shard.order), ptr(q_code), ptr(q_count), ptr(ex), E,
                            nb, shard.G_local, shard.g_offset, shard.d, S.Pmax, ptr(pos_thr), st), "reid_pos_scores")
    if world > 1:
        sharding.exchange_pos_scores(pos_thr, group)
    check(L.reid_pos_sort(ptr(pos_thr), ptr(n_pos), nb, S.Pmax, st), "reid_pos_sort")


def _scan_stage(shard, S, sl, q32_b, q16_b, ex_b, E, *, fused, cand_cap, world, exact_ap, n_slots=None):
    """The gallery pass of one query block (<= two waves of work items) on the current stream: local counts ->
    S.pos_above, candidates -> the best RTOP by approximate score (S.sel_*) and the shard's completeness cut-off (S.bound),
    candidate-buffer overflow -> S.flag.  No host synchronisation unless an identity has more than 64 gallery rows."""
    L = _cabi.lib()
    st = stream_ptr()
    d, Pmax = shard.d, S.Pmax
    q_code, pos_thr, n_pos_b, pos_above_b = S.q_code[sl], S.pos_thr[sl], S.n_pos[sl], S.pos_above[sl]
    nb = q_code.shape[0]
    sms = L.reid_device_sm_count()
    if fused:
        n_chunks = int(n_slots) if n_slots else fused_slots(nb, shard.G_local, sms)
    else:
        n_chunks = max(1, min(16, (2 * sms) // max(1, -(-nb // 8)), shard.G_local // 1024 or 1))
    cap = max(int(cand_cap), 64) // 4 * 4
    cand_score = shard.buf("cand_score", (nb, n_chunks, cap), torch.float32)
    cand_idx = shard.buf("cand_idx", (nb, n_chunks, cap), torch.int32)
    cand_count = shard.buf("cand_count", (nb, n_chunks), torch.int32, zero=True)
    cand_thr = shard.buf("cand_thr", (nb,), torch.float32) if fused else None
    if fused:
        ws_bytes = L.reid_workspace_bytes(1, nb, shard.G_local, d)
        ws = shard.buf("fused_ws", (ws_bytes,), torch.uint8)
        flags = (_cabi.FUSED_EXACT_COUNTS if exact_ap else 0) | (_cabi.FUSED_KLIST16 if world > 1 else 0)
        PW = _cabi.FUSED_PMAX
        check(L.reid_retrieve_fused(ptr(q16_b), ptr(shard.g_f16), ptr(q_code), ptr(shard.g_code), ptr(ex_b), E,
                                    ptr(pos_thr), ptr(n_pos_b), nb, shard.G_local, shard.g_offset, d, min(Pmax, PW), Pmax,
                                    n_chunks, world, cap, flags, ptr(pos_above_b), ptr(cand_score), ptr(cand_idx),
                                    ptr(cand_count), ptr(cand_thr), ptr(ws), ws_bytes, st), "reid_retrieve_fused")
        for w0 in range(PW, Pmax, PW):
            # identities with more than 64 gallery rows: thresholds [w0, w0 + 64) of the queries that have them, in a
            # counting-only pass over a compact copy of those queries (one host read: how many there are)
            sel = torch.nonzero(n_pos_b > w0).flatten()
            ns = int(sel.numel())
            if ns == 0:
                break
            pw = min(PW, Pmax - w0)
            thr_w = pos_thr[sel, w0:w0 + pw].contiguous()
            np_w = (n_pos_b[sel] - w0).clamp_(max=pw).to(torch.int32)
            above_w = torch.zeros(ns, pw, dtype=torch.int32, device=sel.device)
            q16_w = q16_b[sel].contiguous()
            code_w = q_code[sel].contiguous()
            ex_w = ex_b[sel].contiguous() if ex_b is not None else None
            ws_w = torch.empty(L.reid_workspace_bytes(1, ns, shard.G_local, d), dtype=torch.uint8, device=sel.device)
            check(L.reid_retrieve_fused(ptr(q16_w), ptr(shard.g_f16), ptr(code_w), ptr(shard.g_code), ptr(ex_w), E,
                                        ptr(thr_w), ptr(np_w), ns, shard.G_local, shard.g_offset, d, pw, pw,
                                        fused_slots(ns, shard.G_local, sms), world, cap,
                                        flags | _cabi.FUSED_NO_CANDIDATES, ptr(above_w), None, None, None, None,
                                        ptr(ws_w), ws_w.numel(), st), "reid_retrieve_fused(window)")
            pos_above_b[sel, w0:w0 + pw] += above_w
    else:
        check(L.reid_retrieve_exact(ptr(q32_b), ptr(shard.g_f32), ptr(q_code), ptr(shard.g_code), ptr(ex_b), E,
                                    ptr(pos_thr), ptr(n_pos_b), None, nb, nb, shard.G_local, shard.g_offset, d, Pmax,
                                    n_chunks, cap, ptr(pos_above_b), ptr(cand_score), ptr(cand_idx), ptr(cand_count), st),
              "reid_retrieve_exact")
    kx = RESCORE_KX_ONE_SHARD if world == 1 else RESCORE_KX_SHARDED
    check(L.reid_cand_select(ptr(cand_score), ptr(cand_idx), ptr(cand_count), ptr(cand_thr), nb, n_chunks, cap, kx,
                             ptr(S.sel_score[sl]), ptr(S.sel_idx[sl]), ptr(S.sel_n[sl]), ptr(S.bound[sl]), ptr(S.flag[sl]), st),
          "reid_cand_select")
    if _DEBUG_KEEP is not None:
        _DEBUG_KEEP.update(cand_count=cand_count.clone(), n_chunks=n_chunks, sel_n=S.sel_n[sl].clone())


def _rescore_stage(shard, S, sl, q32, *, eps):
    """Exact fp32 re-score of the selected rows of queries `sl` at or above S.bound (after the exchange of the cut-offs:
    only the rows that can still reach the gallery-wide head) -> the shard's exact top list, exact local counts of the
    positives above the bound, S.lb0."""
    L = _cabi.lib()
    nb = S.q_code[sl].shape[0]
    check(L.reid_rescore_topk(ptr(q32), ptr(shard.g_f32), ptr(S.q_code[sl]), ptr(shard.g_code), ptr(S.pos_thr[sl]),
                              ptr(S.n_pos[sl]), ptr(S.sel_score[sl]), ptr(S.sel_idx[sl]), ptr(S.sel_n[sl]), ptr(S.bound[sl]),
                              nb, shard.G_local, shard.g_offset, shard.d, S.Pmax, float(eps), ptr(S.pos_above[sl]),
                              ptr(S.top_score[sl]), ptr(S.top_idx[sl]), ptr(S.lb0[sl]), stream_ptr()), "reid_rescore_topk")


def retrieve(shard: GalleryShard, q_f32: Optional[torch.Tensor], q_f16: Optional[torch.Tensor], q_pid: torch.Tensor,
             excl: Optional[torch.Tensor] = None, topk: int = 10, mode: str = "fused", eps: float = EPS_FP16,
             cand_cap: int = 2048, query_block: Optional[int] = None, group=None, want_ap: bool = False,
             host_queries=None, exact_ap: bool = False, n_slots: Optional[int] = None) -> RetrievalResult:
    """Ranking statistics of a batch of queries against the gallery shard(s).

    mode "fused": tcgen05 GEMM with the counting / candidate epilogue, exact fp32 re-score of the
    candidates, exact fp32 re-run of the (rare) queries whose top-k / CMC is not decidable within eps.
    mode "exact": everything through the fp32 SIMT kernel.
    exact_ap: count every positive's rank on every gallery row (no row sampling for deep positives; slower).
    group: a torch.distributed process group whose ranks hold disjoint contiguous gallery shards.  The shards exchange
    FOUR times per call (not per query block): the positives' scores before the gallery passes, the completeness
    cut-offs after them, then the counts and the top lists (sharding.py).
    host_queries: (query_raw [Q,k,D], mod_id [Q,k], weights) in pinned HOST memory instead of q_f32/q_f16;
    q_pid / excl may then be host tensors too.  Query blocks are copied on a side stream while the previous
    block computes (H2D overlapped with the kernels); with a process group every rank uploads and fuses only its
    1/world slice of a block and the fused block is all-gathered over NVLink (sharding.gather_query_block) -- the
    positives' scores are then exchanged per block.
    The host is synchronised ONCE, by the read of the metrics (plus once per query block when an identity has more than
    64 gallery rows, and once more in the rare case that queries were flagged for the exact re-run).
    """
    assert mode in ("fused", "exact")
    assert 1 <= topk <= _cabi.RTOP
    L = _cabi.lib()
    dev = shard.g_f32.device
    d = shard.d
    Pmax = shard.pmax
    Q = q_pid.shape[0]
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist_mod
        world = dist_mod.get_world_size(group)
        rank = dist_mod.get_rank(group)
    st = stream_ptr()
    sms = L.reid_device_sm_count()
    E = 0 if excl is None or excl.numel() == 0 else excl.shape[1]
    if E == 0:
        excl = None
    auto_blocks = query_block is None
    if query_block is None:
        query_block = default_query_block(sms)

    S = _RankState(Q, Pmax, dev)
    use_fused = (mode == "fused" and Pmax <= 2048 and d % 64 == 0 and d <= 512 and shard.G_local <= (1 << 22)
                 and (host_queries is not None or q_f16 is not None))

    if auto_blocks:
        blocks = host_query_blocks(Q, sms) if host_queries is not None else resident_query_blocks(Q, sms, shard.G_local, cand_cap)
    else:
        blocks = [(b0, min(Q, b0 + query_block)) for b0 in range(0, Q, query_block)]
    staged = {}
    kept = []                                 # (q32, pid, excl) of every block: the re-scorer and the exact re-run read them
    if host_queries is not None:
        h_raw, h_mod, weights = host_queries
        weights = weights.to(dev)
        # staging sets (two, alternating) are allocated on the compute stream, once, at the size of the largest block
        mmax = max(sharding.block_slice(b1 - b0, 0, world)[2] for b0, b1 in blocks)
        nmax = max(b1 - b0 for b0, b1 in blocks)
        sets = []
        for tag in ("stage0_", "stage1_"):
            sets.append((shard.buf(tag + "raw", (mmax,) + tuple(h_raw.shape[1:]), h_raw.dtype),
                         shard.buf(tag + "mod", (mmax,) + tuple(h_mod.shape[1:]), h_mod.dtype),
                         shard.buf(tag + "pid", (nmax,), q_pid.dtype),
                         shard.buf(tag + "excl", (nmax,) + tuple(excl.shape[1:]), excl.dtype) if excl is not None else None))
        copy_stream = torch.cuda.Stream(device=dev)
        copy_stream.wait_stream(torch.cuda.current_stream())      # (the buffers may have just been allocated)
        done_ev = {}                                          # compute-stream event after block bi has consumed its staging set

        def stage(bi):
            # upload this rank's slice of block bi (the features: 8 KB per query at k = 4) and the whole block's ids
            b0, b1 = blocks[bi]
            n = b1 - b0
            s0, s1, m = sharding.block_slice(n, rank, world)
            raw, mod, pid, ex = sets[bi & 1]
            raw, mod, pid = raw[:m], mod[:m], pid[:n]
            ex = ex[:n] if ex is not None else None
            with torch.cuda.stream(copy_stream):
                if bi - 2 in done_ev:
                    copy_stream.wait_event(done_ev.pop(bi - 2))   # the set is free once block bi-2 has finished
                if s1 > s0:
                    raw[:s1 - s0].copy_(h_raw[b0 + s0:b0 + s1], non_blocking=True)
                    mod[:s1 - s0].copy_(h_mod[b0 + s0:b0 + s1], non_blocking=True)
                if s1 - s0 < m:
                    mod[s1 - s0:].fill_(-1)                      # unused slot rows: empty queries (dropped after the gather)
                pid.copy_(q_pid[b0:b1], non_blocking=True)
                if ex is not None:
                    ex.copy_(excl[b0:b1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            staged[bi] = ([raw, mod, pid, ex], ev)
        stage(0)
    else:
        q_pid = q_pid.to(device=dev, dtype=torch.int64).contiguous()
        if excl is not None:
            excl = excl.to(device=dev, dtype=torch.int32).contiguous()
        # resident queries: the positives' scores of the WHOLE batch up front (one exchange over the shards)
        if Q > 0:
            _pos_stage(shard, S, slice(0, Q), q_f32, q_pid, excl, E, group=group, world=world)

    for bi, (b0, b1) in enumerate(blocks):
        nb = b1 - b0
        sl = slice(b0, b1)
        if host_queries is not None:
            if bi + 1 < len(blocks):
                stage(bi + 1)                                    # next block's H2D overlaps this block's kernels
            (raw_b, mod_b, pid_b, ex_b), ev = staged.pop(bi)
            torch.cuda.current_stream().wait_event(ev)
            q32_b, q16_b = fuse_queries(raw_b, mod_b, weights)
            if world > 1:
                # every rank fused 1/world of the block: the fp32 rows travel over NVLink, the fp16 operand copy is
                # re-derived from them (the same round-to-nearest conversion K2 applies)
                q32_b = sharding.gather_query_block(q32_b, nb, group)
                q16_b = q32_b.to(torch.float16)
            pid_b = pid_b.to(torch.int64)
            ex_b = ex_b.to(torch.int32) if ex_b is not None else None
            _pos_stage(shard, S, sl, q32_b, pid_b, ex_b, E, group=group, world=world)
        else:
            q32_b, q16_b = q_f32[sl], (q_f16[sl] if q_f16 is not None else None)
            pid_b, ex_b = q_pid[sl], (excl[sl] if excl is not None else None)
        _scan_stage(shard, S, sl, q32_b, q16_b, ex_b, E, fused=use_fused, cand_cap=cand_cap, world=world,
                    exact_ap=exact_ap, n_slots=n_slots)
        if host_queries is not None:
            # (the staging set is overwritten two blocks later)
            kept.append((q32_b, pid_b.clone(), ex_b.clone() if ex_b is not None else None))
            done_ev[bi] = torch.cuda.Event(); done_ev[bi].record()
        else:
            kept.append((q32_b, pid_b, ex_b))

    # the completeness cut-offs of all blocks are exchanged together, then every shard re-scores what can still reach the head
    if world > 1:
        sharding.exchange_bound(S.bound, group)                     # the best shard's kx-th best approximate score
    eps_rs = float(eps) if use_fused else 0.0
    for (b0, b1), k in zip(blocks, kept):
        _rescore_stage(shard, S, slice(b0, b1), k[0], eps=eps_rs)
    t0 = S.pos_thr[:, 0].contiguous()                               # best positive's exact score

    eps_check = eps_rs

    def finish():
        """Exchange over the shards, decidability check, metrics; -> (pos_above gallery-wide, top lists, metrics incl. flag count)."""
        cnt = S.counts
        if world > 1:
            cnt = sharding.exchange_counts(S.counts.clone(), group)         # counts are additive over shards (one all-reduce)
            # the global top-k is contained in the union of the shards' (exactly ordered) top-k lists
            all_s, all_i = sharding.gather_top_lists(S.top_score[:, :topk], S.top_idx[:, :topk], group)
            out_s = torch.empty(Q, topk, dtype=torch.float32, device=dev)
            out_i = torch.empty(Q, topk, dtype=torch.int32, device=dev)
            check(L.reid_merge_topk(ptr(all_s), ptr(all_i), world, Q, topk, topk, ptr(out_s), ptr(out_i), st), "reid_merge_topk")
        else:
            out_s, out_i = S.top_score[:, :topk].contiguous(), S.top_idx[:, :topk].contiguous()   # one shard: already ordered
        pa = cnt[:Q * Pmax].view(Q, Pmax)
        lb0_g, fl_g = cnt[Q * Pmax:Q * (Pmax + 1)], cnt[Q * (Pmax + 1):]
        if world == 1:
            fl_g = fl_g.clone()                                             # (topk_check adds its bits; S.flag keeps the overflow bit)
        # every rank holds the same gallery-wide data here, so every rank derives the same flags
        check(L.reid_topk_check(ptr(out_s), topk, topk, ptr(S.bound), eps_check, ptr(t0), ptr(S.n_pos), 1, ptr(lb0_g), Q,
                                ptr(fl_g), st), "reid_topk_check")
        out = torch.empty(6, dtype=torch.float64, device=dev)
        ap = torch.empty(Q, dtype=torch.float64, device=dev) if want_ap else None
        check(L.reid_metrics_reduce(ptr(pa), ptr(S.n_pos), Q, Pmax, ptr(out), ptr(ap), st), "reid_metrics_reduce")
        out[5] = torch.count_nonzero(fl_g)
        return pa, out_s, out_i, ap, fl_g, out.cpu().tolist()               # the step's result: D2H read

    pa, out_s, out_i, ap, fl_g, m = finish()
    n_flagged = int(round(m[5]))
    if n_flagged and use_fused:
        # rare: top-k / CMC of these queries is not decidable from fp16 scores within eps (or a candidate buffer
        # overflowed): all-fp32 re-run on a compact copy, results written back, exchange + metrics redone
        sel = torch.nonzero(fl_g).flatten()
        q32_s = torch.cat([k[0] for k in kept])[sel].contiguous()
        pid_s = torch.cat([k[1] for k in kept])[sel].contiguous()
        ex_s = torch.cat([k[2] for k in kept])[sel].contiguous() if E else None
        ns = int(sel.numel())
        X = _RankState(ns, Pmax, dev)
        _pos_stage(shard, X, slice(0, ns), q32_s, pid_s, ex_s, E, group=group, world=world)
        _scan_stage(shard, X, slice(0, ns), q32_s, None, ex_s, E, fused=False, cand_cap=cand_cap, world=world, exact_ap=True)
        if world > 1:
            sharding.exchange_bound(X.bound, group)
        _rescore_stage(shard, X, slice(0, ns), q32_s, eps=0.0)
        S.pos_above[sel] = X.pos_above; S.n_pos[sel] = X.n_pos; S.top_score[sel] = X.top_score; S.top_idx[sel] = X.top_idx
        S.lb0[sel] = X.lb0; t0[sel] = X.pos_thr[:, 0]
        S.bound[sel] = float("-inf")                 # ranked in fp32: exact by construction, nothing left to decide
        S.flag.zero_()
        pa, out_s, out_i, ap, fl_g, m = finish()
    metrics = {"mAP": m[0], "R@1": m[1], "R@5": m[2], "R@10": m[3], "num_queries": int(round(m[4]))}
    return RetrievalResult(metrics, out_i, out_s, ap, n_flagged, pa, S.n_pos, "fused" if use_fused else "exact")
