#!/usr/bin/env python3
"""The files leg of bench.py alone (3.1 Gb genome, PE150 HS25, 2^24 pairs into <prefix>_R{1,2}.fq on tmpfs), three plain
runs and two device-BGZF runs: for comparing ways of filling the page cache and writer thread counts.

    python tools/files_probe.py [writer_threads] [launches]"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import jackalope_b200 as J  # noqa: E402
import torch  # noqa: E402

nthr = int(sys.argv[1]) if len(sys.argv) > 1 else min(len(os.sched_getaffinity(0)), 32)
n_l = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lens, L, kw, full_pairs = bench.workload("human_pe150_hs25", 3_100_000_000)
total = int(lens.sum())
pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
flat = pinned.numpy()
bench.make_genome_into(flat, lens, 20261018)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
ctx = J.Context(0)
B = 1 << 20
d = tempfile.mkdtemp(dir="/dev/shm")
out = {"writer_threads": nthr, "pairs": n_l * B, "plain_s": [], "bgzf_s": []}
try:
    J.illumina(g, os.path.join(d, "w"), 2 * B, L, True, seed=1, ctx=ctx, batch_pairs=B, n_threads=nthr, overwrite=True, **kw)
    for rep in range(3):
        for f in os.listdir(d):
            os.unlink(os.path.join(d, f))
        t0 = time.perf_counter()
        J.illumina(g, os.path.join(d, "r"), 2 * n_l * B, L, True, seed=2, ctx=ctx, batch_pairs=B, n_threads=nthr, overwrite=True, **kw)
        out["plain_s"].append(round(time.perf_counter() - t0, 4))
    sz = os.path.getsize(os.path.join(d, "r_R1.fq")) + os.path.getsize(os.path.join(d, "r_R2.fq"))
    out["plain_GBps"] = round(sz / min(out["plain_s"]) / 1e9, 2)
    out["plain_pairs_per_s"] = n_l * B / min(out["plain_s"])
    for rep in range(2):
        for f in os.listdir(d):
            os.unlink(os.path.join(d, f))
        t0 = time.perf_counter()
        J.illumina(g, os.path.join(d, "z"), 2 * n_l * B, L, True, seed=2, ctx=ctx, batch_pairs=B, n_threads=nthr, compress=True, overwrite=True, **kw)
        out["bgzf_s"].append(round(time.perf_counter() - t0, 4))
    out["bgzf_pairs_per_s"] = n_l * B / min(out["bgzf_s"])
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
