"""The C ABI (include/reid_b200.h): the built library loads without a GPU and exports every declared entry point;
the Python binding table covers the header; the product path refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "reid_b200.h")


def _declared():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reid_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from prcv2025reid_b200 import _cabi, build
    build.build_library()                          # nvcc cross-compiles for sm_100a without a GPU; no-op when up to date
    return ctypes.CDLL(_cabi.LIB_PATH)


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20 and "reid_retrieve_fused" in names and "reid_sdm_fwd" in names
    for n in names:
        assert getattr(lib, n) is not None, n


def test_binding_table_covers_the_header():
    from prcv2025reid_b200 import _cabi
    assert set(_declared()) <= set(_cabi.EXPORTED_SYMBOLS) | {"reid_sdm_pair"}
    assert ctypes.sizeof(_cabi.SdmPair) == 80      # struct reid_sdm_pair: 10 x 8 bytes (N and M share a word)


def test_host_only_entry_points(lib):
    lib.reid_strerror.restype = ctypes.c_char_p
    lib.reid_abi_version.restype = ctypes.c_int
    assert lib.reid_abi_version() >= 1
    assert lib.reid_strerror(0) and lib.reid_strerror(-1)
    lib.reid_sdm_saved_floats.restype = ctypes.c_size_t
    n = lib.reid_sdm_saved_floats(512, 512, 512)
    assert n >= 512 * 512                           # at least S itself
    lib.reid_workspace_bytes.restype = ctypes.c_size_t
    lib.reid_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    assert lib.reid_workspace_bytes(1, 1000, 100000, 512) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_product_path_fails_loudly_without_a_gpu():
    from prcv2025reid_b200 import eval_mm_protocol as emp
    from prcv2025reid_b200 import train_eval
    from prcv2025reid_b200.sdm_loss import sdm_loss_stable
    with pytest.raises(RuntimeError):
        emp.l2n(torch.randn(3, 8))
    with pytest.raises(RuntimeError):
        sdm_loss_stable(torch.randn(4, 8), torch.randn(4, 8), torch.eye(4))
    with pytest.raises(RuntimeError):
        train_eval.compute_cmc(torch.randn(4, 8), torch.randn(6, 8), torch.arange(4), torch.arange(6))
