#!/usr/bin/env python3
"""Summarise an .ncu-rep: key metrics per kernel + the hottest source lines / SASS.
usage: ncu_summary.py report.ncu-rep [kernel-substring] [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' , 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    if pat and pat not in r[hdr.index('Kernel Name')]:
        continue
    for w in want:
        if w in hdr:
            print("%-90s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + pat] if pat else []),
                     capture_output=True, text=True).stdout
# the source page prints one table per kernel
blocks = src.split('"Kernel Name"')
for b in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    name = lines[0][1] if len(lines[0]) > 1 else "?"
    if pat and pat not in name:
        continue
    h = lines[1]
    try:
        i_src, i_samp, i_inst = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    except ValueError:
        continue
    body = [l for l in lines[2:] if len(l) > i_inst]
    tot_s = sum(int(l[i_samp] or 0) for l in body) or 1
    tot_i = sum(int(l[i_inst] or 0) for l in body) or 1
    print("kernel", name, "samples", tot_s, "warp-inst", tot_i, "sass lines", len(body))
    ranked = sorted(range(len(body)), key=lambda k: -int(body[k][i_samp] or 0))[:topn]
    for k in sorted(ranked):
        l = body[k]
        print("%5d %6.2f%% samp %6.2f%% inst  %s" % (k, 100.0 * int(l[i_samp] or 0) / tot_s, 100.0 * int(l[i_inst] or 0) / tot_i, l[i_src].strip()[:110]))
