#!/usr/bin/env python3
"""Instructions and stall samples of a kernel per barrier-delimited phase (one line per BAR.SYNC / BAR.RED segment
of the SASS), from an .ncu-rep captured with --import-source on.
usage: ncu_barrier_phases.py report.ncu-rep kernel-regex [blocks]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
blocks = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r)][0]
h = rows[hi]
ie, sp, src = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
data = [r for r in rows[hi + 1:] if len(r) > max(ie, sp)]
num = lambda x: float(x or 0)
tot, ts = sum(num(r[ie]) for r in data), sum(num(r[sp]) for r in data)
seg, cur = [], [0, 0, 0, 0]
for i, r in enumerate(data):
    cur[0] += num(r[ie]); cur[1] += num(r[sp]); cur[3] = i
    if "BAR.SYNC" in r[src] or "BAR.RED" in r[src]:
        seg.append(cur); cur = [0, 0, i + 1, i + 1]
seg.append(cur)
print("warp instructions %.0f%s, samples %.0f" % (tot, " (%.0f per block)" % (tot / blocks) if blocks else "", ts))
for s in seg:
    print("sass %5d-%5d  inst %5.1f%%%s  samples %5.1f%%  (of which waiting at the phase's first instruction %4.1f%%)"
          % (s[2], s[3], 100 * s[0] / tot, "  %7.0f/block" % (s[0] / blocks) if blocks else "", 100 * s[1] / ts, 100 * num(data[s[2]][sp]) / ts))
