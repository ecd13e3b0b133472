#!/usr/bin/env python3
"""Haplotype materialisation (k_materialize) on a 200 Mb chromosome with 1 % substitutions and
0.1 % indels per site: run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:k_materialize` to get its time and traffic; prints the host-side wall time of set_haplotypes too."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jackalope_b200 as J  # noqa: E402

g = J.random_genome(1, 200_000_000, seed=1)
haps = J.random_haplotypes(g, 2, sub_rate=0.01, indel_rate=0.001, seed=2)
ctx = J.Context(0)
ctx.set_genome(g)
t0 = time.perf_counter()
ctx.set_haplotypes(haps)
dt = time.perf_counter() - t0
m = haps.muts[0][0]
print("2 haplotypes x 200 Mb, %d mutation records each: set_haplotypes %.3f s (uploads of the records included)" % (m.old_pos.size, dt))
print("(byte-for-byte parity of the materialised chromosomes is tests/test_gpu_parity.py::test_materialize_matches_oracle)")
