#!/usr/bin/env python3
"""PacBio reads: default pacbio() arguments on a 100 Mb genome (10 x 10 Mb), reads left on the device and written to files
on tmpfs.  Prints one JSON line; kept under profiles/.  (The unmodified reference is timed by bench.py, pacbio.cpu_baseline.)"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jackalope_b200 as J  # noqa: E402

threads = min(os.cpu_count() or 1, 32)
g = J.random_genome(10, 10_000_000, seed=3)
ctx = J.Context(0)
n = 1 << 17
J.pacbio(g, "", 4096, seed=1, ctx=ctx, sink="device", n_threads=threads)
t0 = time.perf_counter()
st = J.pacbio(g, "", n, seed=2, ctx=ctx, sink="device", n_threads=threads)
t_dev = time.perf_counter() - t0
out = {"workload": "10 x 10 Mb genome, pacbio() defaults", "reads": n, "bases": st["bytes_out"][0] / 2,
       "device": {"kernel_ms": st["reads_ms"], "kernel_reads_per_s": n / (st["reads_ms"] / 1e3), "wall_s": t_dev,
                  "reads_per_s": n / t_dev, "fastq_GBps_kernel": st["bytes_out"][0] / (st["reads_ms"] / 1e3) / 1e9,
                  "host_threads": threads}}
d = tempfile.mkdtemp(dir="/dev/shm")
try:
    J.pacbio(g, os.path.join(d, "w"), 1 << 15, seed=1, ctx=ctx, n_threads=threads, overwrite=True)      # buffers
    J.pacbio(g, os.path.join(d, "w"), 1 << 15, seed=1, ctx=ctx, n_threads=threads, compress=True, overwrite=True)
    t0 = time.perf_counter()
    J.pacbio(g, os.path.join(d, "p"), n, seed=2, ctx=ctx, n_threads=threads)
    t_f = time.perf_counter() - t0
    out["files"] = {"wall_s": t_f, "reads_per_s": n / t_f, "bytes": os.path.getsize(os.path.join(d, "p_R1.fq"))}
    t0 = time.perf_counter()
    J.pacbio(g, os.path.join(d, "z"), n, seed=2, ctx=ctx, n_threads=threads, compress=True)
    t_z = time.perf_counter() - t0
    out["files_bgzf_device"] = {"wall_s": t_z, "reads_per_s": n / t_z, "bytes": os.path.getsize(os.path.join(d, "z_R1.fq.gz"))}
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
