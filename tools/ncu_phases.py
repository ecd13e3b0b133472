#!/usr/bin/env python3
"""Per-'phase' (grouped by executions per pair) stall-reason breakdown from an ncu source-page CSV.
usage: ncu_phases.py report.ncu-rep kernel-regex units(=pairs per launch)"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict
rep, pat, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
i_src, i_inst = h.index("Source"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: h.index(c) for c in stalls}
body, seen = [], set()
for l in rows[2:]:
    if len(l) <= i_inst or not l[i_inst].isdigit():
        continue
    if l[0] in seen:
        break
    seen.add(l[0]); body.append(l)
grp = defaultdict(lambda: [0, 0, Counter()])
for l in body:
    k = round(int(l[i_inst]) / units, 1)
    key = "5.x (phase B loop)" if 4.5 <= k <= 6 else "1.x/0.9x (once per pair)" if 0.85 <= k <= 2.1 else "0.5-0.8 (per-strand / first-pass)" if 0.45 <= k < 0.85 else "rare"
    g = grp[key]
    g[0] += 1; g[1] += int(l[i_inst])
    for c in stalls:
        g[2][c] += int(l[idx[c]] or 0)
tot = sum(sum(g[2].values()) for g in grp.values())
for key, g in grp.items():
    s = sum(g[2].values())
    print("%-34s lines %4d inst/unit %7.1f samples %5.1f%%  cyc/inst %.2f" % (key, g[0], g[1] / units, 100.0 * s / tot, (s / tot) / (g[1] / sum(x[1] for x in grp.values())) if g[1] else 0))
    print("     " + "  ".join("%s %.0f%%" % (c.replace("stall_", ""), 100.0 * v / s) for c, v in g[2].most_common(8) if s))
