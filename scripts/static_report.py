#!/usr/bin/env python
"""Static (no GPU) evidence of what the library compiles to: per kernel the registers / spills / shared memory
ptxas reports for sm_100a, and per translation unit the count of the SASS mnemonics that prove which hardware
paths the kernels use (B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma, UTMALDG = TMA tensor loads,
UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, SYNCS = mbarrier traffic).

    python scripts/static_report.py > profiles/<round>_static_ptxas_sass.txt
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prcv2025reid_b200 import build  # noqa: E402

MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "LDTM", "STTM",
             "UTCATOMSWS", "SYNCS", "REDUX", "HMMA", "LDGSTS", "MUFU.EX2", "MUFU.RCP", "ATOMS", "REDS", "RED.E", "ATOMG"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return [re.sub(r"\(anonymous namespace\)::", "", o) for o in out]
    except Exception:
        return names


def main():
    tmp = tempfile.mkdtemp(prefix="reid_static_")
    print("# static build report: nvcc %s" % " ".join(build.NVCC_FLAGS))
    print("# (cross-compiled without a GPU; nothing here is a timing)\n")
    for src in build.SOURCES:
        s = os.path.join(build.CSRC, src)
        o = os.path.join(tmp, src.replace(".cu", ".o"))
        r = subprocess.run([build._nvcc()] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            print("## %s: nvcc failed\n%s" % (src, r.stderr))
            continue
        print("## %s" % src)
        rows, cur = [], None
        for line in r.stderr.split("\n"):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = {"name": m.group(1)}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["spill"] = int(m.group(1)), int(m.group(2)) + int(m.group(3))
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
            if m:
                cur["regs"], cur["bars"], cur["smem"] = int(m.group(1)), m.group(2) or "0", m.group(3) or "0"
        names = demangle([x["name"] for x in rows])
        print("%-88s %5s %5s %6s %6s %9s" % ("kernel", "regs", "bars", "stack", "spill", "static_smem"))
        for x, n in zip(rows, names):
            n = re.sub(r"\(.*$", "", n)
            print("%-88s %5s %5s %6s %6s %9s" % (n[:88], x.get("regs", "?"), x.get("bars", "?"), x.get("stack", "?"),
                                                  x.get("spill", "?"), x.get("smem", "?")))
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        cnt = collections.Counter()
        n_inst = 0
        for line in sass.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if not m:
                continue
            n_inst += 1
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + ".") or op.startswith(k + "_"):
                    cnt[k] += 1
        print("SASS: %d instructions; %s\n" % (n_inst, ", ".join("%s %d" % (k, cnt[k]) for k in MNEMONICS if cnt[k]) or "-"))


if __name__ == "__main__":
    main()
