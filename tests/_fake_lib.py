"""A CPU stand-in for libreid_b200.so, for testing the HOST logic of prcv2025reid_b200.engine without a GPU.  TEST ONLY.

Every method implements the CONTRACT of the entry point of the same name as written in include/reid_b200.h, with
torch CPU arithmetic and tensors instead of device pointers (`install()` swaps `engine.ptr` for the identity and
`engine.stream_ptr` for a no-op, so `engine` hands the tensors / views themselves to these methods and in-place
results land where the real kernels would write them).  The "fused" pass scores with the fp16 operand copies like the
tensor cores do, so the engine's flag -> exact re-run path is exercised for real.  Nothing here is product code, and the
product never loads it: `prcv2025reid_b200._cabi.lib()` raises when the real library is missing.
"""
import numpy as np
import torch
import torch.nn.functional as F

KLIST = 32
RTOP = 32
NEG_INF = float("-inf")


class FakeLib:
    def __init__(self, sm_count=148, force_flag_every=0):
        self.sm_count = sm_count
        self.calls = []
        # > 0: the re-scorer of the FUSED pass declares every n-th query of a block undecidable and leaves garbage in its
        # outputs, as a kernel that gave up on the query might -- the host must repair it through the exact pass
        self.force_flag_every = force_flag_every
        self._last_fused = False

    # ------------------------------------------------------------------ utilities
    def _log(self, name):
        self.calls.append(name)
        return 0

    def reid_device_sm_count(self):
        return self.sm_count

    def reid_workspace_bytes(self, which, Q, G, d):
        return 64

    def reid_strerror(self, rc):
        return b"fake"

    # ------------------------------------------------------------------ K1 / K2
    def reid_l2norm_rows(self, x, out32, out16, rows, d, eps, st):
        y = F.normalize(x.float(), dim=-1, eps=eps)
        if out32 is not None:
            out32.copy_(y)
        if out16 is not None:
            out16.copy_(y.half())
        return self._log("reid_l2norm_rows")

    def reid_mm_fuse_normalize(self, feats, mod_id, w, n_mod, out32, out16, Q, k, d, st):
        f = F.normalize(feats.float(), dim=-1)
        if k == 1:
            acc = f[:, 0]
        else:
            acc = None
            for j in range(k):
                use = mod_id[:, j] >= 0
                wj = torch.where(use, w[mod_id[:, j].clamp(min=0).long()], torch.zeros(Q))
                term = f[:, j] * wj[:, None]
                acc = term if acc is None else acc + term
        y = F.normalize(acc, dim=-1)
        out32.copy_(y)
        if out16 is not None:
            out16.copy_(y.half())
        return self._log("reid_mm_fuse_normalize")

    # ------------------------------------------------------------------ identity index
    def reid_pid_index_build(self, g_pid, G, sorted_pid, order, max_run, ws, ws_bytes, st):
        s, o = torch.sort(g_pid, stable=True)
        sorted_pid.copy_(s); order.copy_(o.to(torch.int32))
        _, counts = torch.unique_consecutive(s, return_counts=True)
        max_run.fill_(int(counts.max()))
        return self._log("reid_pid_index_build")

    def reid_pid_lookup(self, sorted_pid, G, pids, n, code, count, st):
        lo = torch.searchsorted(sorted_pid, pids, right=False)
        hi = torch.searchsorted(sorted_pid, pids, right=True)
        found = hi > lo
        code.copy_(torch.where(found, lo, torch.full_like(lo, -1)).to(torch.int32))
        if count is not None:
            count.copy_((hi - lo).to(torch.int32))
        return self._log("reid_pid_lookup")

    # ------------------------------------------------------------------ positives
    @staticmethod
    def _masked(excl, E, nb, G_local, g_offset):
        m = torch.zeros(nb, G_local, dtype=torch.bool)
        if excl is not None and E > 0:
            loc = excl.long() - g_offset
            ok = (excl >= 0) & (loc >= 0) & (loc < G_local)
            rows = torch.arange(nb)[:, None].expand_as(loc)
            m[rows[ok], loc[ok]] = True
        return m

    def reid_pos_scores(self, q32, g32, order, q_code, q_count, excl, E, Q, G_local, g_offset, d, Pmax, pos_score, st):
        pos_score.fill_(NEG_INF)
        masked = self._masked(excl, E, Q, G_local, g_offset)
        for q in range(Q):
            c, n = int(q_code[q]), int(q_count[q])
            for j in range(min(n, Pmax) if c >= 0 else 0):
                row = int(order[c + j]) - g_offset
                if 0 <= row < G_local and not masked[q, row]:
                    pos_score[q, j] = torch.dot(q32[q], g32[row])
        return self._log("reid_pos_scores")

    def reid_pos_sort(self, pos_score, n_pos, Q, Pmax, st):
        s, _ = torch.sort(pos_score, dim=1, descending=True)
        pos_score.copy_(s)
        n_pos.copy_(torch.isfinite(s).sum(1).to(torch.int32))
        return self._log("reid_pos_sort")

    # ------------------------------------------------------------------ the ranking step
    def _rank_pass(self, S, sel, q_code, g_code, masked, pos_thr, n_pos, Pmax, n_chunks, cap, rows_per_chunk, pos_above,
                   cand_score, cand_idx, cand_count, cand_thr):
        G = S.shape[1]
        for q in sel:
            s = S[q].clone()
            ok = ~masked[q]
            nonpos = ok & (g_code != q_code[q])
            for j in range(min(int(n_pos[q]), Pmax)):
                pos_above[q, j] += int((nonpos & (s > pos_thr[q, j])).sum())
            s_ok = torch.where(ok, s, torch.full_like(s, NEG_INF))
            n_ok = int(ok.sum())
            thr = torch.topk(s_ok, KLIST).values[-1] if n_ok >= KLIST else torch.tensor(NEG_INF)
            if cand_thr is not None:
                cand_thr[q] = thr
            keep = torch.nonzero(ok & (s_ok >= thr)).flatten()         # complete down to the KLIST-th best row
            for c in range(n_chunks):
                mine = keep[(keep // rows_per_chunk) == c]
                cand_count[q, c] = mine.numel()
                m = min(mine.numel(), cap)
                cand_score[q, c, :m] = s[mine[:m]]
                cand_idx[q, c, :m] = mine[:m].to(torch.int32)

    def reid_retrieve_fused(self, q16, g16, q_code, g_code, excl, E, pos_thr, n_pos, Q, G_local, g_offset, d, Pmax, pos_stride,
                            n_chunks, n_shards, cap, flags, pos_above, cand_score, cand_idx, cand_count, cand_thr, ws, ws_bytes, st):
        # (pos_thr / pos_above arrive as [Q, pos_stride] tensors; this call covers their first Pmax columns)
        assert Pmax <= 64 and (pos_stride == 0 or pos_stride >= Pmax) and pos_thr.shape[1] == (pos_stride or Pmax)
        self._last_fused = True
        S = q16.float() @ g16.float().T                                # fp16 operands, fp32 accumulation
        rpc = -(-G_local // n_chunks)
        rpc = -(-rpc // 1024) * 1024
        if flags & 2:                                                  # REID_FUSED_NO_CANDIDATES: counting only
            assert cand_score is None and cand_idx is None and cand_count is None
            nb = Q
            cand_score, cand_idx = torch.empty(nb, n_chunks, cap), torch.empty(nb, n_chunks, cap, dtype=torch.int32)
            cand_count, cand_thr = torch.zeros(nb, n_chunks, dtype=torch.int32), None
        self._rank_pass(S, range(Q), q_code, g_code, self._masked(excl, E, Q, G_local, g_offset), pos_thr, n_pos, Pmax, n_chunks, cap,
                        rpc, pos_above, cand_score, cand_idx, cand_count, cand_thr)
        return self._log("reid_retrieve_fused" if not (flags & 2) else "reid_retrieve_fused(window)")

    def reid_retrieve_exact(self, q32, g32, q_code, g_code, excl, E, pos_thr, n_pos, q_sel, n_sel, Q, G_local, g_offset, d, Pmax,
                            n_chunks, cap, pos_above, cand_score, cand_idx, cand_count, st):
        self._last_fused = False
        sel = range(Q) if q_sel is None else [int(v) for v in q_sel[:n_sel]]
        S = q32 @ g32.T
        rpc = -(-G_local // n_chunks)
        self._rank_pass(S, sel, q_code, g_code, self._masked(excl, E, Q, G_local, g_offset), pos_thr, n_pos, Pmax, n_chunks, cap,
                        rpc, pos_above, cand_score, cand_idx, cand_count, None)
        return self._log("reid_retrieve_exact" if q_sel is None else "reid_retrieve_exact(sel)")

    def reid_cand_select(self, cand_score, cand_idx, cand_count, cand_thr, Q, n_chunks, cap, kx, sel_score, sel_idx, sel_n,
                         sel_cut, sel_flag, st):
        for q in range(Q):
            keep = float(cand_thr[q]) if cand_thr is not None else NEG_INF
            sc, ix, overflow = [], [], False
            for c in range(n_chunks):
                n = int(cand_count[q, c])
                if n > cap:
                    n, overflow = cap, True
                sc.append(cand_score[q, c, :n]); ix.append(cand_idx[q, c, :n])
            sc, ix = torch.cat(sc), torch.cat(ix)
            m = sc >= keep
            sc, ix = sc[m], ix[m]
            o = torch.from_numpy(np.lexsort((ix.numpy(), -sc.numpy())))      # score desc, index asc (ranks_before of the kernels)
            sc, ix = sc[o], ix[o]
            total = sc.numel()
            R = min(total, RTOP)
            sel_score[q].fill_(NEG_INF); sel_idx[q].fill_(-1)
            sel_score[q, :R] = sc[:R]; sel_idx[q, :R] = ix[:R].to(torch.int32)
            sel_n[q] = R
            sel_cut[q] = float(sc[kx - 1]) if total >= kx else NEG_INF
            sel_flag[q] = 1 if overflow else 0
            if self.force_flag_every and self._last_fused and q % self.force_flag_every == 3:
                sel_flag[q] = 1                                        # "this query's candidate buffer overflowed"
        return self._log("reid_cand_select")

    def reid_rescore_topk(self, q32, g32, q_code, g_code, pos_thr, n_pos, sel_score, sel_idx, sel_n, bound, Q, G_local, g_offset,
                          d, Pmax, eps, pos_above, top_score, top_idx, lb0, st):
        for q in range(Q):
            b = float(bound[q])
            n = int(sel_n[q])
            R = int((sel_score[q, :n] >= b).sum())                     # the prefix of the descending list at or above the bound
            gi = sel_idx[q, :R].long()
            ex = (g32[gi] @ q32[q]) if R else torch.empty(0)
            o2 = sorted(range(R), key=lambda r: (-float(ex[r]), int(gi[r])))
            ex, gi = ex[o2], gi[o2]
            top_score[q].fill_(NEG_INF); top_idx[q].fill_(-1)
            top_score[q, :R] = ex
            top_idx[q, :R] = (gi + g_offset).to(torch.int32)
            lim = b + eps
            neg = g_code[gi] != q_code[q] if R else torch.zeros(0, dtype=torch.bool)
            lb0[q] = 0
            for j in range(min(int(n_pos[q]), Pmax)):
                t = float(pos_thr[q, j])
                lb = int((neg & (ex > t)).sum())
                if t > lim or b == NEG_INF:
                    pos_above[q, j] = lb
                else:
                    pos_above[q, j] = max(int(pos_above[q, j]), lb)
                if j == 0:
                    lb0[q] = lb
            if self.force_flag_every and eps > 0 and q % self.force_flag_every == 3:
                # a kernel that gave up on the query leaves garbage: the host must repair it through the exact pass
                pos_above[q].fill_(12345); top_idx[q].fill_(-7); top_score[q].fill_(9.0)
        return self._log("reid_rescore_topk")

    def reid_topk_check(self, top, list_len, topk, bound, eps, pos_thr, n_pos, Pmax, lb0, Q, flag, st):
        for q in range(Q):
            b = float(bound[q])
            f = 1 if int(flag[q]) else 0
            if b > NEG_INF:
                lim = b + eps
                if not float(top[q, topk - 1]) >= lim:
                    f |= 2
                t0 = float(pos_thr[q]) if pos_thr.dim() == 1 else float(pos_thr[q, 0])
                if int(n_pos[q]) > 0 and not t0 > lim and int(lb0[q]) < 10:
                    f |= 4
            flag[q] = f
        return self._log("reid_topk_check")

    def reid_merge_topk(self, scores, idx, n_lists, Q, list_len, topk, out_s, out_i, st):
        s = scores.view(n_lists, Q, list_len).permute(1, 0, 2).reshape(Q, -1)
        i = idx.view(n_lists, Q, list_len).permute(1, 0, 2).reshape(Q, -1)
        for q in range(Q):
            ent = sorted(((-float(a), int(b)) for a, b in zip(s[q], i[q]) if int(b) >= 0))[:topk]
            out_s[q].fill_(NEG_INF); out_i[q].fill_(-1)
            for r, (a, b) in enumerate(ent):
                out_s[q, r] = -a; out_i[q, r] = b
        return self._log("reid_merge_topk")

    def reid_metrics_reduce(self, pos_above, n_pos, Q, Pmax, out, ap, st):
        acc = torch.zeros(5, dtype=torch.float64)
        for q in range(Q):
            n = min(int(n_pos[q]), Pmax)
            if ap is not None:
                ap[q] = -1.0
            if n == 0:
                continue
            a = sum((j + 1) / (int(pos_above[q, j]) + j + 1) for j in range(n)) / n
            first = int(pos_above[q, 0]) + 1
            acc += torch.tensor([a, first <= 1, first <= 5, first <= 10, 1.0], dtype=torch.float64)
            if ap is not None:
                ap[q] = a
        out[:5].copy_(torch.cat([acc[:4] / acc[4] if acc[4] > 0 else torch.zeros(4, dtype=torch.float64), acc[4:]]))
        return self._log("reid_metrics_reduce")


    def reid_topk_label_metrics(self, top_idx, q_label, g_label, Q, list_len, k, ap, hit, st):
        for q in range(Q):                                             # train.py:116-124 / :133-136 (fp32 like the reference)
            row = top_idx[q, :k].long()
            m = (row >= 0) & (g_label[row.clamp(min=0)] == q_label[q])
            n = int(m.sum())
            hit[q] = 1 if n else 0
            if n:
                prec = torch.cumsum(m.float(), 0) / torch.arange(1, k + 1, dtype=torch.float32)
                ap[q] = (prec * m.float()).sum() / n
            else:
                ap[q] = -1.0
        return self._log("reid_topk_label_metrics")


    # ------------------------------------------------------------------ SDM (pairs arrive as the packed reid_sdm_pair array)
    @staticmethod
    def _tensor(addr, shape, dtype):
        """A torch view of host memory at `addr` (the stand-in of a device pointer)."""
        import ctypes
        import numpy as np
        n = 1
        for v in shape:
            n *= int(v)
        if dtype == torch.bfloat16:
            arr = np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(addr))
            return torch.from_numpy(arr).view(torch.bfloat16).view(*shape)
        if dtype == torch.float16:
            arr = np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(addr))
            return torch.from_numpy(arr).view(torch.float16).view(*shape)
        ct = {torch.float32: ctypes.c_float, torch.int32: ctypes.c_int32, torch.int64: ctypes.c_int64, torch.uint8: ctypes.c_uint8}[dtype]
        return torch.from_numpy(np.ctypeslib.as_array((ct * n).from_address(addr))).view(*shape)

    def _pairs(self, arr, n, code, d):
        dt = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}[code]
        W = 14                                                          # sizeof(reid_sdm_pair) / 8
        out = []
        for p in range(n):
            w = [int(arr[W * p + i]) for i in range(W)]
            N, M = w[3] & 0xFFFFFFFF, w[3] >> 32
            P = dict(q=self._tensor(w[0], (N, d), dt), g=self._tensor(w[1], (M, d), dt),
                     y=self._tensor(w[2], (N, M), torch.float32) if w[2] else None,
                     loss=self._tensor(w[4], (1,), torch.float32), status=self._tensor(w[5], (1,), torch.int32),
                     grad=self._tensor(w[7], (1,), torch.float32) if w[7] else None,
                     dq=self._tensor(w[8], (N, d), dt) if w[8] else None, dg=self._tensor(w[9], (M, d), dt) if w[9] else None,
                     rows=None, cols=None)
            if P["y"] is None:                                          # label form: rows that take part, y from the labels
                rl, cl = self._tensor(w[10], (N,), torch.int64), self._tensor(w[11], (M,), torch.int64)
                rv = self._tensor(w[12], (N,), torch.uint8).bool() if w[12] else torch.ones(N, dtype=torch.bool)
                cv = self._tensor(w[13], (M,), torch.uint8).bool() if w[13] else torch.ones(M, dtype=torch.bool)
                P["rows"], P["cols"] = torch.nonzero(rv).flatten(), torch.nonzero(cv).flatten()
                P["y"] = (rl[P["rows"]][:, None] == cl[P["cols"]][None, :]).float()
            out.append(P)
        return out

    def reid_sdm_saved_floats(self, N, M, d):
        return 64

    def reid_sdm_uses_tensor_cores(self, arr, n, code, d):
        """bf16, 64 <= N, M <= 512 multiples of 8, d % 64 == 0, d <= 512 (include/reid_b200.h); dense y or the label form"""
        def ok(p):
            N, M = int(arr[14 * p + 3]) & 0xFFFFFFFF, int(arr[14 * p + 3]) >> 32
            return all(64 <= v <= 512 and v % 8 == 0 for v in (N, M))
        return 1 if code == 1 and d % 64 == 0 and 64 <= d <= 512 and all(ok(p) for p in range(n)) else 0

    def reid_sdm_step_launches(self, arr, n, code, d):
        return 1

    def _sdm(self, arr, n, code, d, tau, eps, fwd, bwd):
        from oracle import sdm as osdm
        for P in self._pairs(arr, n, code, d):
            with torch.enable_grad():                                  # (autograd is off inside Function.backward)
                q = P["q"].clone().requires_grad_(True); g = P["g"].clone().requires_grad_(True)
                if P["rows"] is not None:                              # label form: the reference filters the rows first
                    if P["rows"].numel() == 0 or P["cols"].numel() == 0 or not bool(P["y"].any()):
                        L = torch.zeros([])
                    else:
                        L = osdm.sdm_loss_oracle(q[P["rows"]], g[P["cols"]], P["y"], tau=tau, eps=eps)
                else:
                    L = osdm.sdm_loss_oracle(q, g, P["y"], tau=tau, eps=eps)
            if fwd:
                P["loss"][0] = float(L.detach())
                P["status"][0] = 0 if L.requires_grad else (9 if not bool(P["y"].any()) else 1)   # bit0: the reference's zero; bit3: no positive
            if bwd:
                if L.requires_grad:
                    with torch.enable_grad():
                        L.backward()
                    P["dq"].copy_(q.grad * P["grad"][0]); P["dg"].copy_(g.grad * P["grad"][0])
                else:
                    P["dq"].zero_(); P["dg"].zero_()                   # guard path: exact-zero gradients
        return 0

    def reid_sdm_fwd(self, arr, n, code, d, tau, eps, st):
        self._log("reid_sdm_fwd")
        return self._sdm(arr, n, code, d, tau, eps, True, False)

    def reid_sdm_bwd(self, arr, n, code, d, tau, eps, st):
        self._log("reid_sdm_bwd")
        return self._sdm(arr, n, code, d, tau, eps, False, True)

    def reid_sdm_step(self, arr, n, code, d, tau, eps, st):
        self._log("reid_sdm_step")
        return self._sdm(arr, n, code, d, tau, eps, True, True)


def install(monkeypatch, fake=None, **kw):
    """Route prcv2025reid_b200.engine to a FakeLib (tensors instead of device pointers).  -> the FakeLib."""
    from prcv2025reid_b200 import _cabi, engine
    fake = fake or FakeLib(**kw)
    monkeypatch.setattr(_cabi, "lib", lambda: fake)
    monkeypatch.setattr(engine, "ptr", lambda t: t)
    monkeypatch.setattr(engine, "stream_ptr", lambda: None)
    chk = lambda rc, what="": None if rc == 0 else (_ for _ in ()).throw(RuntimeError(what))   # noqa: E731
    monkeypatch.setattr(engine, "check", chk)
    from prcv2025reid_b200 import train_eval
    monkeypatch.setattr(train_eval, "ptr", lambda t: t)
    monkeypatch.setattr(train_eval, "stream_ptr", lambda: None)
    monkeypatch.setattr(train_eval, "check", chk)
    monkeypatch.setattr(train_eval, "_dev", lambda t: t)
    from prcv2025reid_b200 import sdm_loss
    monkeypatch.setattr(sdm_loss, "check", chk)
    monkeypatch.setattr(sdm_loss, "_raw_stream", lambda dev: None)
    monkeypatch.setattr(sdm_loss, "_SAVED_FLOATS", {})
    return fake
