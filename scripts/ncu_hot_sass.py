"""Top stalled SASS instructions of one kernel in an .ncu-rep (source page): scripts/ncu_hot_sass.py rep kernel_regex [n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
bodies, body = [], []
for r in rows[2:]:                      # several captured launches of the kernel: keep the one with the most samples
    if len(r) != len(hdr) or r[isamp] == "# Samples":
        if body:
            bodies.append(body); body = []
        continue
    body.append(r)
if body:
    bodies.append(body)
body = max(bodies, key=lambda b: sum(int(r[isamp]) for r in b))
tot = sum(int(r[isamp]) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp]))[:n]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  #%-4d %-70s %s" % (int(r[isamp]), 100.0 * int(r[isamp]) / max(1, tot), k, r[isrc].strip()[:70], st))
