#!/usr/bin/env python3
"""The PacBio leg of bench.py on its 3.1 Gb genome, alone, with the library's host-side trace (JLP_TRACE=1): where the
wall time of a device-resident pacbio() call goes (first preparation, waiting for later ones, allocations, device)."""
import os
import sys
import time

os.environ["JLP_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import jackalope_b200 as J  # noqa: E402

lens, L, kw, full_pairs = bench.workload("human_pe150_hs25", 3_100_000_000)
total = int(lens.sum())
pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
flat = pinned.numpy()
bench.make_genome_into(flat, lens, 20261018)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
ctx = J.Context(0)
nthr = min(len(os.sched_getaffinity(0)), 32)
J.pacbio(g, "", 1 << 13, seed=1, ctx=ctx, sink="device", n_threads=nthr)
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
for rep in range(6):
    t0 = time.perf_counter()
    st = J.pacbio(g, "", n_reads, seed=2 + rep, ctx=ctx, sink="device", n_threads=nthr)
    print("run %d: %.4f s, kernels %.1f ms, %d reads" % (rep, time.perf_counter() - t0, st["reads_ms"], st["pairs"]), flush=True)
