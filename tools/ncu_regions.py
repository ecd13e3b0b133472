#!/usr/bin/env python3
"""Per-pair instruction counts of k_reads from an .ncu-rep (source page): total, the phase-B loop, what comes
before and after it; optionally the SASS listing with executions per pair.
usage: ncu_regions.py report.ncu-rep [pairs_per_launch] [listing.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread']:
    print("%-70s %s" % (k, d.get(k)))
for k in rows[0]:
    if 'issue_stalled' in k and 'per_issue_active' in k and float(d[k]) > 0.25:
        print("  stall %-40s %.2f" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(d[k])))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
for i, l in enumerate(lines):
    if 'Source' in l and 'Instructions Executed' in l:
        h, start = l, i + 1
        break
i_src, i_inst, i_thr, i_samp = h.index('Source'), h.index('Instructions Executed'), h.index('Avg. Threads Executed'), h.index('Warp Stall Sampling (All Samples)')
body = [l for l in lines[start:] if len(l) > i_inst]
cnt = [int(l[i_inst] or 0) / pairs for l in body]
b = [k for k, c in enumerate(cnt) if c > 4.5]
print("instructions per pair: %.0f = %.0f before the phase-B loop + %.0f in it + %.0f after it (functions called from it included)"
      % (sum(cnt), sum(cnt[:b[0]]), sum(cnt[b[0]:b[-1] + 1]), sum(cnt[b[-1] + 1:])))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for k, l in enumerate(body):
            f.write("%5d %6.2f %5s %6s  %s\n" % (k, cnt[k], l[i_thr][:5], l[i_samp], l[i_src].strip()[:110]))
