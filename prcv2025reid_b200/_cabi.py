"""ctypes binding of libreid_b200.so (include/reid_b200.h).  No CPU fallback: importing the
kernels without the built library raises."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER, Structure

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REID_LIB") or os.path.join(HERE, "libreid_b200.so")

KLIST = 32
RTOP = 32
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
SDM_MAX_PAIRS = 16
FUSED_EXACT_COUNTS, FUSED_NO_CANDIDATES, FUSED_KLIST16 = 1, 2, 4
FUSED_PMAX = 64           # thresholds per query and reid_retrieve_fused call


class ReidError(RuntimeError):
    pass


class SdmPair(Structure):
    _fields_ = [("qry", c_void_p), ("gal", c_void_p), ("y", c_void_p), ("N", c_int32), ("M", c_int32),
                ("loss", c_void_p), ("status", c_void_p), ("saved", c_void_p), ("grad_out", c_void_p),
                ("dqry", c_void_p), ("dgal", c_void_p), ("row_label", c_void_p), ("col_label", c_void_p),
                ("row_valid", c_void_p), ("col_valid", c_void_p)]


_SIGS = {
    "reid_strerror": (c_char_p, [c_int]),
    "reid_abi_version": (c_int, []),
    "reid_device_sm_count": (c_int, []),
    "reid_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int]),
    "reid_l2norm_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "reid_mm_fuse_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "reid_sim_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_void_p]),
    "reid_pid_index_build": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "reid_pid_lookup": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "reid_pos_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                c_int64, c_int, c_int, c_void_p, c_void_p]),
    "reid_pos_sort": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "reid_retrieve_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "reid_retrieve_exact": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "reid_cand_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "reid_rescore_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "reid_topk_check": (c_int, [c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p, c_int64,
                                c_void_p, c_void_p]),
    "reid_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "reid_metrics_reduce": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "reid_topk_label_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "reid_sdm_saved_floats": (c_size_t, [c_int, c_int, c_int]),
    "reid_sdm_uses_tensor_cores": (c_int, [c_void_p, c_int, c_int, c_int]),
    "reid_sdm_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "reid_sdm_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "reid_sdm_step": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "reid_sdm_step_launches": (c_int, [c_void_p, c_int, c_int, c_int]),
}

EXPORTED_SYMBOLS = sorted(_SIGS)
_lib = None

# kernels of THIS library launched per C call (library plumbing such as the CUB sort is not counted)
_LAUNCHES_PER_CALL = {
    "reid_l2norm_rows": 1, "reid_mm_fuse_normalize": 1, "reid_sim_gemm": 1, "reid_pid_index_build": 2,
    "reid_pid_lookup": 1, "reid_pos_scores": 1, "reid_pos_sort": 1, "reid_retrieve_fused": 4,
    "reid_retrieve_exact": 1, "reid_cand_select": 1, "reid_rescore_topk": 1, "reid_topk_check": 1, "reid_merge_topk": 1, "reid_metrics_reduce": 2,
    "reid_topk_label_metrics": 1, "reid_sdm_fwd": 1, "reid_sdm_bwd": 1,   # (tcgen05 path: fwd = 2 launches, counted in sdm_loss.py)
    "reid_sdm_step": 0,                                                   # (counted in sdm_loss.py: reid_sdm_step_launches)
}
LAUNCH_COUNT = {"n": 0}
# optional per-kernel device timing: set PROFILE = [] and every kernel call appends (name, start, end) events
PROFILE = None


class _Timed:
    def __init__(self, name, fn):
        self.name, self.fn, self.n = name, fn, _LAUNCHES_PER_CALL[name]

    def __call__(self, *a):
        LAUNCH_COUNT["n"] += self.n
        if PROFILE is None:
            return self.fn(*a)
        import torch
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        rc = self.fn(*a)
        e.record()
        PROFILE.append((self.name, s, e))
        return rc


class _Lib:
    def __init__(self, cdll):
        for name in _SIGS:
            fn = getattr(cdll, name)
            setattr(self, name, _Timed(name, fn) if name in _LAUNCHES_PER_CALL else fn)


def lib():
    """The loaded library; raises ReidError when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ReidError("libreid_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _Lib(L)
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().reid_strerror(rc).decode()
        extra = ""
        if rc == -2:
            try:
                import torch
                torch.cuda.synchronize()
            except Exception as e:  # surface the CUDA error text
                extra = " (%s)" % e
        raise ReidError("%s failed: %s [%d]%s" % (what or "libreid_b200 call", msg, rc, extra))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
