#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3a|c3b|c1]

Metric: queries/sec for MM-4 retrieval + evaluation, 512-d features, 100k queries x 1M gallery
(BASELINE config C4).  One "step" = one pass of the hot path over the whole
query batch: MM-4 fusion + normalisation (K2), exact positive scores, the fused tcgen05 similarity /
ranking kernel (K3+K4), candidate re-scoring, shard merge and the metric reduction.  The gallery is
installed once (normalise + identity index; reported as gallery_prepare_ms, like the reference's
cached rgb_feats.npy) and, at N > 1, sharded contiguously over the ranks (strong scaling: the job is
fixed, the per-GPU gallery shrinks).  `value` times the step with inputs resident in HBM; `e2e`
times the same step through the public tensor API with the query-side inputs in pinned HOST memory
(H2D inside the timed region) and the metric dict read back (D2H).

--impl reference times the reference algorithm's CPU path (the oracle port of
tools/eval_mm_protocol.py:369-469, torch CPU, all host threads) on a bounded query sample of the
same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (seed, n_ids, gal_per_id, k, queries_per_id)   SURVEY.md section 8d
    "c4": (1005, 25000, 40, 4, 4),      # MM-4, Q=100k, G=1M
    "c3a": (1003, 5000, 20, 3, 4),      # MM-3, Q=20k,  G=100k
    "c3b": (1004, 5000, 20, 4, 4),      # MM-4, Q=20k,  G=100k
    "c1": (1001, 500, 20, 2, 6),        # MM-2, Q=3k,   G=10k
}
FEAT_DIM = 512


def workload_name(w):
    seed, n_ids, gpi, k, qpi = WORKLOADS[w]
    return "MM-%d retrieval+eval, %d queries x %d gallery, 512-d synthetic CLIP-shaped features (seed %d)" % (
        k, n_ids * qpi, n_ids * gpi, seed)


EXCL_FRAC, N_EXCL, TOPK = 0.01, 2, 10


def bench_config(w):
    """The workload description BOTH arms print (identical dicts: the driver compares them)."""
    seed, n_ids, gpi, k, qpi = WORKLOADS[w]
    G, Q = n_ids * gpi, n_ids * qpi
    return {"workload": workload_name(w), "k": k, "topk": TOPK, "excl_frac": EXCL_FRAC, "n_excl": N_EXCL,
            "l2": "inputs larger than L2, no flush needed (query features %d MB fp32, gallery %d MB fp16 / %d MB fp32 per "
                  "step; L2 = 126 MB)" % ((Q * k * FEAT_DIM * 4) >> 20, (G * FEAT_DIM * 2) >> 20, (G * FEAT_DIM * 4) >> 20)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """CPU arm: the reference algorithm (oracle port) on the host cores, bounded query sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import retrieval as orc
    from prcv2025reid_b200 import synth
    seed, n_ids, gpi, k, qpi = WORKLOADS[args.workload]
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(max(1, avail))
    cores = torch.get_num_threads()
    centres, bias = synth.make_centres(seed, n_ids)
    G = n_ids * gpi
    gallery = synth.make_gallery_rows(seed, centres, bias, gpi, 0, G)
    g_pid = torch.arange(G, dtype=torch.int64) // gpi
    g = orc.l2n(gallery)                                         # eval_mm_protocol.py:546
    del gallery
    sample = args.ref_queries
    n_need = sample * (args.steps + args.warmup)
    # the first queries of the workload, generated with the full workload's identity centres
    small = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=EXCL_FRAC, n_excl=N_EXCL, gallery_rows=(0, 0),
                                      max_queries=n_need)
    w = synth.weights_tensor()
    times = []
    for it in range(args.steps + args.warmup):
        sl = slice(it * sample, (it + 1) * sample)
        t0 = time.perf_counter()
        q = orc.fuse_queries(small.query_raw[sl], small.mod_id[sl], w)
        orc.rank_and_metrics_loop(q, g, small.q_pid[sl], g_pid, small.excl[sl])
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    qps = sample / t
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "%d queries per step (same-image exclusion list included) against the full %d-row gallery, "
                                   "per-query loop of eval_mm_protocol.py:396-455 (torch CPU GEMV + argsort + AP walk)" % (sample, G)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def time_sdm(torch, synth, sdm_loss_pairs, P, K, n_pairs, dtype, iters=50):
    """SDM forward + backward of one training step (all pairs).  Returns
    (eager us/step through autograd, algorithmic bytes, pairs, device us/step of the same step replayed as a CUDA graph)."""
    from prcv2025reid_b200.sdm_loss import SdmGraphStep
    feats, labels = synth.make_sdm_batch(2001 if P == 4 else 2002, P, K, n_modalities=5, dtype=dtype, device="cuda")
    y = (labels[:, None] == labels[None, :]).float()
    pairs = [(a, b) for a in range(5) for b in range(a)][:n_pairs] if n_pairs > 4 else [(m, 0) for m in range(1, 5)]
    qs = [feats[a].clone().requires_grad_(True) for a, b in pairs]
    vs = [feats[b].clone().requires_grad_(True) for a, b in pairs]
    ys = [y] * len(pairs)

    def step():
        losses = sdm_loss_pairs(qs, vs, ys, tau=0.2)
        losses.sum().backward()
        for t in qs + vs:
            t.grad = None
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        step()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / iters
    # the same step as one CUDA-graph launch: device time of the kernels without the host-side launch gaps
    g = SdmGraphStep(qs, vs, ys, tau=0.2)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    graph_us = s.elapsed_time(e) * 1e3 / iters
    N = P * K
    esz = 2 if dtype == torch.bfloat16 else 4
    alg_bytes = len(pairs) * (3 * (2 * N) * FEAT_DIM * esz + 2 * N * N * 4)
    return us, alg_bytes, len(pairs), graph_us


def torch_gpu_baseline(torch, engine, shard, case, weights, nq, iters=3):
    """The reference ALGORITHM (fp32 similarity rows, same-image mask, full argsort, AP over the full ranking;
    eval_mm_protocol.py:401-455, vectorised over a block of queries) with plain torch CUDA ops on the same GPU: a
    second, stronger baseline next to the CPU arm.  Not the product path and not timed into `value`."""
    q32, _ = engine.fuse_queries(case.query_raw[:nq], case.mod_id[:nq], weights)
    g32, g_pid, q_pid, excl = shard.g_f32, case.g_pid, case.q_pid[:nq], case.excl[:nq]
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    rows = torch.arange(nq, device=q32.device)[:, None].expand_as(excl)

    def run():
        S = q32 @ g32.T                                                     # :401
        ok = excl >= 0
        S[rows[ok], excl[ok].long()] = -1e9                                 # :421-422
        order = torch.argsort(S, dim=1, descending=True)                    # :423
        match = (g_pid[order] == q_pid[:, None]) & (S.gather(1, order) > -1e8)   # :427
        npos = match.sum(1)
        prec = match.cumsum(1) / torch.arange(1, S.shape[1] + 1, device=S.device)
        ap = (prec * match).sum(1) / npos.clamp_min(1)                      # :444-455
        valid = npos > 0
        return float(ap[valid].double().mean()), float(match[:, :1].any(1)[valid].float().mean())
    run()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        m = run()
    e.record()
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = old
    ms = s.elapsed_time(e) / iters
    return {"value": nq / (ms * 1e-3), "unit": "queries/s", "kind": "reference algorithm, torch CUDA ops (fp32 matmul, full argsort), same GPU",
            "sample": "first %d queries of the workload against the full gallery shard" % nq, "mAP_on_sample": m[0], "R@1_on_sample": m[1]}


def parity_check(torch, engine, synth, shard, case, weights, args, group, rank, world, dev):
    """Oracle check inside the run: the first `parity_queries` queries through engine.retrieve (every rank takes part),
    rank 0 compares with oracle.rank_and_metrics_counting on the unsharded gallery.  CMC must be identical, the top-10
    lists identical modulo 2e-6 ties of the reference's own fp32 scores, mAP within 1e-4."""
    import numpy as np
    nq = min(args.parity_queries, case.Q)
    seed, n_ids, gpi, k, qpi = WORKLOADS[args.workload]
    q32, q16 = engine.fuse_queries(case.query_raw[:nq], case.mod_id[:nq], weights)
    r = engine.retrieve(shard, q32, q16, case.q_pid[:nq], case.excl[:nq], topk=TOPK, mode=args.mode, group=group, want_ap=True)
    if rank != 0:
        return None
    from oracle import retrieval as orc
    centres, bias = synth.make_centres(seed, n_ids, device=dev)
    G = n_ids * gpi
    g = orc.l2n(synth.make_gallery_rows(seed, centres, bias, gpi, 0, G).cpu())
    q = orc.fuse_queries(case.query_raw[:nq].cpu(), case.mod_id[:nq].cpu(), weights.cpu())
    t0 = time.perf_counter()
    o = orc.rank_and_metrics_counting(q, g, case.q_pid[:nq].cpu(), case.g_pid.cpu(), case.excl[:nq].cpu(), return_per_query=True)
    oracle_s = time.perf_counter() - t0
    v = o["_valid"]
    ap = r.ap.cpu().numpy()
    first = r.pos_above[:, 0].cpu().numpy() + 1
    ti = r.top_idx.cpu().numpy().astype(np.int64)
    beyond = 0
    for qi in np.nonzero((ti[:, :TOPK] != o["_top_idx"][:, :TOPK]).any(axis=1))[0]:
        sc = (q[qi:qi + 1] @ g.T).squeeze(0).numpy()
        if any(ti[qi, j] != o["_top_idx"][qi, j] and abs(float(sc[ti[qi, j]]) - float(sc[o["_top_idx"][qi, j]])) > 2e-6 for j in range(TOPK)):
            beyond += 1
    d_ap = np.abs(ap[v] - o["_ap"][v])
    out = {"queries": nq, "oracle": "oracle.retrieval.rank_and_metrics_counting (CPU fp32, full ranking of the unsharded gallery)",
           "mAP": r.metrics["mAP"], "mAP_oracle": o["mAP"], "d_mAP": r.metrics["mAP"] - o["mAP"],
           "cmc": [r.metrics["R@1"], r.metrics["R@5"], r.metrics["R@10"]], "cmc_oracle": [o["R@1"], o["R@5"], o["R@10"]],
           "cmc_rank_mismatches": int((np.minimum(first[v], 11) != np.minimum(o["_first"][v], 11)).sum()), "top10_lists_differing_beyond_2e-6_ties": beyond,
           "per_query_dAP_max": float(d_ap.max()), "per_query_dAP_mean": float(d_ap.mean()), "oracle_seconds": round(oracle_s, 2)}
    out["ok"] = bool(abs(out["d_mAP"]) <= 1e-4 and out["cmc"] == out["cmc_oracle"] and beyond == 0)
    return out


def secondary_workloads(torch, engine, synth, _cabi, weights, dev, steps=5):
    """The other BASELINE retrieval configs on one GPU (C3a, C3b: 20k x 100k; C1: 3k x 10k), same step as the headline,
    plus the dict-level drop-in `rank_and_metrics(queries, ...)` on C1 (the call a user of the reference makes)."""
    out = {}
    for w in ("c3a", "c3b", "c1"):
        seed, n_ids, gpi, k, qpi = WORKLOADS[w]
        G, Q = n_ids * gpi, n_ids * qpi
        case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=EXCL_FRAC, n_excl=N_EXCL, device=dev)
        shard = engine.prepare_gallery(case.gallery_raw, case.g_pid)

        def step():
            q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, weights)
            return engine.retrieve(shard, q32, q16, case.q_pid, case.excl, topk=TOPK)
        for _ in range(3):
            res = step()
        torch.cuda.synchronize()
        _cabi.PROFILE = []
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            res = step()
        e.record()
        torch.cuda.synchronize()
        prof, _cabi.PROFILE = _cabi.PROFILE, None
        ms = s.elapsed_time(e) / steps
        fused_ms = sum(a.elapsed_time(b) for n, a, b in prof if n == "reid_retrieve_fused") / steps
        tf = 2.0 * Q * G * FEAT_DIM / (fused_ms * 1e-3) / 1e12 if fused_ms > 0 else None
        out[w] = {"workload": workload_name(w), "queries_per_sec": Q / (ms * 1e-3), "ms_per_step": ms,
                  "fused_kernel_ms": fused_ms, "fused_tflops": tf, "metrics": res.metrics, "path": res.path,
                  "flagged_queries": res.n_flagged}
        if w == "c1":
            out["c1_dropin"] = dropin_c1(torch, synth, case)
        del case, shard
        torch.cuda.empty_cache()
    return out


def dropin_c1(torch, synth, case):
    """rank_and_metrics(queries, gallery_feats, gallery_meta, extractor, weight_cfg) -- the reference's own signature
    (tools/eval_mm_protocol.py:369) with its list-of-dicts inputs on the HOST -- timed end to end beside the oracle port
    of the reference's per-query CPU loop on the same inputs."""
    from oracle import retrieval as orc
    from prcv2025reid_b200 import eval_mm_protocol as emp
    cpu = synth.RetrievalCase(case.gallery_raw.cpu(), case.g_pid.cpu(), case.query_raw.cpu(), case.mod_id.cpu(),
                              case.q_pid.cpu(), case.excl.cpu(), case.k)
    queries, gmeta, ext = synth.case_to_reference_inputs(cpu)
    g = emp.l2n(cpu.gallery_raw)
    wcfg = dict(synth.DEFAULT_WEIGHTS)
    emp.rank_and_metrics(queries, g, gmeta, ext, wcfg, ignore_same_img=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        m = emp.rank_and_metrics(queries, g, gmeta, ext, wcfg, ignore_same_img=True)
    torch.cuda.synchronize()
    ours = (time.perf_counter() - t0) / n
    nq = 300
    t0 = time.perf_counter()
    q = orc.fuse_queries(cpu.query_raw[:nq], cpu.mod_id[:nq], synth.weights_tensor())
    orc.rank_and_metrics_loop(q, orc.l2n(cpu.gallery_raw), cpu.q_pid[:nq], cpu.g_pid, cpu.excl[:nq])
    ref = (time.perf_counter() - t0) / nq
    return {"call": "rank_and_metrics(queries, gallery_feats, gallery_meta, extractor, weight_cfg, ignore_same_img=True)",
            "queries_per_sec": len(queries) / ours, "seconds_per_call": ours, "metrics": m,
            "cpu_port_queries_per_sec": 1.0 / ref, "cpu_port_sample": "first %d queries, per-query loop" % nq}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-queries", type=int, default=24, help="queries per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-queries", type=int, default=128)
    ap.add_argument("--parity-queries", type=int, default=256, help="queries of the in-run oracle check (0 = skip)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C1 / C3 secondary workloads")
    ap.add_argument("--no-sdm", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fused", choices=["fused", "exact"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from prcv2025reid_b200 import _cabi, engine, synth
    from prcv2025reid_b200 import sharding
    from prcv2025reid_b200.sdm_loss import sdm_loss_pairs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    seed, n_ids, gpi, k, qpi = WORKLOADS[args.workload]
    G, Q = n_ids * gpi, n_ids * qpi
    r0, r1 = sharding.shard_range(G, rank, world)
    case = synth.make_retrieval_case(seed, n_ids, gpi, k, qpi, excl_frac=EXCL_FRAC, n_excl=N_EXCL, device=dev,
                                     gallery_rows=(r0, r1))
    weights = synth.weights_tensor(device=dev)
    torch.cuda.synchronize()
    # K1 alone (gallery normalisation, one-time): HBM roofline of the pass
    k1 = []
    for _ in range(4):
        s_ = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
        s_.record(); _n32, _n16 = engine.l2norm_rows(case.gallery_raw, want_f16=True); e_.record()
        torch.cuda.synchronize(); k1.append(s_.elapsed_time(e_))
    del _n32, _n16
    k1_ms = min(k1[1:])
    k1_bytes = (r1 - r0) * FEAT_DIM * (4 + 4 + 2)
    t0 = time.perf_counter()
    shard = engine.prepare_gallery(case.gallery_raw, case.g_pid, g_offset=r0)
    torch.cuda.synchronize()
    gallery_prepare_ms = (time.perf_counter() - t0) * 1e3
    case.gallery_raw = None
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier(group)
        torch.cuda.synchronize()

    def step_resident():
        q32, q16 = engine.fuse_queries(case.query_raw, case.mod_id, weights)
        return engine.retrieve(shard, q32, q16, case.q_pid, case.excl, topk=10, mode=args.mode, group=group)

    # host-side (pinned) copies of the query-side inputs for the end-to-end number
    h_query = case.query_raw.cpu().pin_memory()
    h_mod = case.mod_id.cpu().pin_memory()
    h_pid = case.q_pid.cpu().pin_memory()
    h_excl = case.excl.cpu().pin_memory()
    # whole-job bytes: every query's features are uploaded by exactly one rank (engine.retrieve shards the upload
    # and all-gathers the fused block over NVLink); the small id / exclusion arrays go to every rank
    nbytes = lambda t: t.numel() * t.element_size()
    h2d_bytes = nbytes(h_query) + nbytes(h_mod) + world * (nbytes(h_pid) + nbytes(h_excl))

    def step_e2e():
        # public tensor API with HOST inputs: block-wise H2D on a side stream, metrics read back (D2H)
        return engine.retrieve(shard, None, None, h_pid, h_excl, topk=10, mode=args.mode, group=group,
                               host_queries=(h_query, h_mod, weights))

    def timed(fn, steps, profile=False):
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()                      # nvidia-smi needs ~0.3 s to start: the warm-up steps run the same load
        for _ in range(args.warmup):
            res = fn()
        barrier()
        _cabi.LAUNCH_COUNT["n"] = 0
        if profile:
            _cabi.PROFILE = []
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            res = fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        prof = _cabi.PROFILE
        _cabi.PROFILE = None
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item()) / steps, res, _cabi.LAUNCH_COUNT["n"] // steps, prof, clocks

    ms_step, res, launches, prof, clocks = timed(step_resident, args.steps, profile=True)
    ms_e2e, res2, _, _, _ = timed(step_e2e, args.steps)

    # per-kernel device time over the timed region (CUDA events on the launching stream)
    per_kernel = {}
    for name, s, e in prof:
        d = per_kernel.setdefault(name, [0.0, 0])
        d[0] += s.elapsed_time(e); d[1] += 1
    dom = "reid_retrieve_fused" if args.mode == "fused" else "reid_retrieve_exact"
    roofline = None
    if dom in per_kernel:
        tot_ms, n_calls = per_kernel[dom]
        flops_total = 2.0 * Q * (r1 - r0) * FEAT_DIM * args.steps     # this rank's shard
        achieved = flops_total / (tot_ms * 1e-3) / 1e12
        peak, which = 1380.3, "measured (MEASURED_PEAKS.json bf16_tflops_sustained; kernel timed inside a long step)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak = float(mp.get("bf16_tflops_sustained", peak))
        except Exception:
            peak, which = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
        traffic, traffic_src = None, None
        try:      # DRAM bytes per launch from the newest committed ncu --set full capture of this exact configuration
            import glob
            for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
                tr = json.load(open(path)).get(dom)
                if tr and tr["workload"] == args.workload and tr["n_gpus"] == world:
                    traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"] + " [" + os.path.basename(path) + "]"
                    break
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes (dram read+write per launch)",
                    "traffic_source": traffic_src, "peak_source": which,
                    "algorithmic_flop_per_launch": flops_total / max(1, n_calls),
                    "avg_launch_ms": tot_ms / max(1, n_calls), "launches": n_calls,
                    "share_of_step": tot_ms / (ms_step * args.steps)}
    kernel_ms = {n: round(v[0] / args.steps, 4) for n, v in sorted(per_kernel.items())}
    hbm_peak = 6534.1
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", hbm_peak))
    except Exception:
        pass
    hbm_kernels = {"reid_l2norm_rows": {"ms": round(k1_ms, 4), "algorithmic_bytes": k1_bytes,
                                        "gbs": k1_bytes / (k1_ms * 1e-3) / 1e9, "frac_of_hbm_peak": k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak,
                                        "note": "gallery shard: 4 B read + 4 B fp32 + 2 B fp16 written per element"}}
    if "reid_mm_fuse_normalize" in per_kernel:
        k2_ms = per_kernel["reid_mm_fuse_normalize"][0] / args.steps
        k2_bytes = Q * (k * FEAT_DIM * 4 + FEAT_DIM * (4 + 2))
        hbm_kernels["reid_mm_fuse_normalize"] = {"ms": round(k2_ms, 4), "algorithmic_bytes": k2_bytes,
                                                 "gbs": k2_bytes / (k2_ms * 1e-3) / 1e9,
                                                 "frac_of_hbm_peak": k2_bytes / (k2_ms * 1e-3) / 1e9 / hbm_peak,
                                                 "note": "k x 2 KB read + 3 KB (fp32 + fp16) written per query"}

    # ---- in-run parity: the first queries of the workload through the same engine call (same shards, same exchange)
    #      against the oracle's full ranking walk on the CPU (tools/eval_mm_protocol.py:396-455); outside the timed region
    parity = None
    if args.parity_queries > 0:
        parity = parity_check(torch, engine, synth, shard, case, weights, args, group, rank, world, dev)

    line = {
        "metric": "queries_per_sec", "value": Q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world,
        "metrics": res.metrics, "parity": parity,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate / f32 re-score",
        "data": "synthetic",
        "config": bench_config(args.workload),
        "run_info": {"mode": args.mode, "path": res.path, "gallery_rows_per_gpu": r1 - r0,
                     "gallery_prepare_ms": round(gallery_prepare_ms, 2), "flagged_queries": res.n_flagged,
                     "query_block": engine.default_query_block(_cabi.lib().reid_device_sm_count())},
        "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 48, "ms_per_step": ms_e2e, "metrics_equal_resident": all(abs(res2.metrics[m_] - res.metrics[m_]) <= 1e-12 for m_ in res.metrics)},
        "gpu_launches": launches, "kernel_ms_per_step": kernel_ms, "roofline": roofline,
        "hbm_kernels": hbm_kernels, "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_secondary and args.workload == "c4":
        del h_query, res, res2
        torch.cuda.empty_cache()
        line["secondary"] = secondary_workloads(torch, engine, synth, _cabi, weights, dev)

    if rank == 0 and world == 1:
        if not args.no_sdm:
            sdm = {}
            note = ("us_per_step_eager = autograd step on the stream (host-bound: Python + launches); "
                    "us_per_step_graph = the same forward+backward (reid_sdm_step, objective = sum of the pair losses) "
                    "replayed as one CUDA graph (device time)")
            us, ab, npairs, gu = time_sdm(torch, synth, sdm_loss_pairs, 4, 2, 4, torch.float32)
            sdm["c2_p4k2_fp32_4pairs"] = {"us_per_step_eager": us, "us_per_step_graph": gu, "algorithmic_bytes": ab,
                                          "hbm_gbs_graph": ab / (gu * 1e-6) / 1e9, "frac_of_hbm_peak_graph": ab / (gu * 1e-6) / 1e9 / hbm_peak,
                                          "kernels": "graph: sdm_small_step (forward + backward in ONE launch, fp32 SIMT, 1 CTA per pair); eager: sdm_small_fwd + sdm_small_bwd",
                                          "note": note}
            us, ab, npairs, gu = time_sdm(torch, synth, sdm_loss_pairs, 64, 8, 10, torch.bfloat16)
            sdm["c5_p64k8_bf16_10pairs"] = {"us_per_step_eager": us, "us_per_step_graph": gu, "algorithmic_bytes": ab,
                                            "hbm_gbs_graph": ab / (gu * 1e-6) / 1e9, "frac_of_hbm_peak_graph": ab / (gu * 1e-6) / 1e9 / hbm_peak,
                                            "tensor_tflops_graph": 10 * 3 * 2.0 * 512 * 512 * 512 / (gu * 1e-6) / 1e12,
                                            "kernels": "tc_prep + tc_fwd + tc_bwd (tcgen05, bf16)"}
            line["sdm"] = sdm
        if not args.no_cpu_baseline:
            try:
                line["torch_gpu_baseline"] = torch_gpu_baseline(torch, engine, shard, case, weights, 128 if G > 200000 else 1024)
            except Exception as ex:                                          # (never let the optional arm break the bench line)
                line["torch_gpu_baseline"] = {"unavailable": str(ex)[:200]}
            torch.cuda.empty_cache()
            from oracle import retrieval as orc
            nq = args.cpu_baseline_queries
            g_cpu = shard.g_f32.cpu()
            q32, _ = engine.fuse_queries(case.query_raw[:nq], case.mod_id[:nq], weights)
            qraw, qmod = case.query_raw[:nq].cpu(), case.mod_id[:nq].cpu()
            t0 = time.perf_counter()
            qc = orc.fuse_queries(qraw, qmod, weights.cpu())
            o = orc.rank_and_metrics_loop(qc, g_cpu, case.q_pid[:nq].cpu(), case.g_pid.cpu(), case.excl[:nq].cpu())
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nq / dt, "unit": "queries/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": "first %d queries of the workload against the full gallery, per-query "
                                              "loop of eval_mm_protocol.py:396-455 (oracle port, torch CPU)" % nq}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
