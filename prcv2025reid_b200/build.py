"""Build libreid_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library is a plain C-ABI shared object, see include/reid_b200.h)."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libreid_b200.so")
SOURCES = ["normalize.cu", "pos_index.cu", "rank.cu", "sdm.cu", "sdm_tc.cu", "sim_gemm.cu", "retrieve_fused.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "reid_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    extra = os.environ.get("REID_NVCC_EXTRA", "").split()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o])
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs)
    return LIB


def build_native_host(force=False):
    """examples/abi_host: a plain C++ host of the C ABI (device memory from cudaMalloc, no torch), linked against the
    in-tree library with a relative rpath so that it runs from the repository copy on the GPU box."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "examples", "abi_host.cpp")
    out = os.path.join(root, "examples", "abi_host")
    lib = build_library()
    if force or _stale(out, [src, lib, os.path.join(root, "include", "reid_b200.h")]):
        cmd = [_nvcc(), "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(root, "include"), src, "-o", out,
               "-L", HERE, "-lreid_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../prcv2025reid_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
