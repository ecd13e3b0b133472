#!/usr/bin/env python3
"""What a FIRST illumina() call on a fresh context costs against a warm one, by batch size: the pinned host buffers of the
three-stage pipeline (3 slots x 2 ends x batch bytes) are allocated by the first call, and pinning is slow on these VMs.
PE150 HS25 on a 500 Mb genome, stream sink (every batch lands in pinned host memory), 3e7 pairs and 5e5 pairs.

    python tools/cold_call_probe.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import jackalope_b200 as J  # noqa: E402
import torch  # noqa: E402

lens, L, kw, _ = bench.workload("human_pe150_hs25", 500_000_000)
total = int(lens.sum())
pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
flat = pinned.numpy()
bench.make_genome_into(flat, lens, 20261018)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
J.Context(0).close()        # CUDA context creation is not what is measured
out = []
sink = lambda *a: None
for n_pairs in (30_000_000, 500_000):
    for B in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
        ctx = J.Context(0)
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            st = J.illumina(g, "", 2 * n_pairs, L, True, seed=3 + rep, ctx=ctx, sink=sink, batch_pairs=B, **kw)
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.close()
        t_close = time.perf_counter() - t0
        out.append({"pairs": n_pairs, "batch_pairs": B, "first_call_s": round(ts[0], 4), "warm_call_s": round(min(ts[1:]), 4),
                    "warm_pairs_per_s": n_pairs / min(ts[1:]), "close_s": round(t_close, 4), "device_ms": st["device_ms"]})
        print(json.dumps(out[-1]), flush=True)
