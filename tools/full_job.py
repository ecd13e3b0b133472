#!/usr/bin/env python3
"""The whole BASELINE.json configs[4] job on one B200: 3.1 Gb genome, 30x coverage, PE150 HS25 =
3.1e8 read pairs (~204 GB of FASTQ).  Legs: FASTQ left in HBM, FASTQ streamed into the
library's pinned host buffers (nothing is written to disk: no file system here takes 200 GB), and the same
stream compressed on the device (compress = 6 and 1: BGZF bytes reach the host).
Prints one JSON line; run under gpurun, output kept in profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import jackalope_b200 as J  # noqa: E402

lens, L, kw, full_pairs = bench.workload("human_pe150_hs25", int(3.1e9))
total = int(lens.sum())
pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
flat = pinned.numpy()
bench.make_genome_into(flat, lens, 20261018)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
g.flat = lambda: (flat, off.astype(np.uint64))
ctx = J.Context(0)
J.illumina(g, "", 2 * (1 << 21), L, True, seed=1, ctx=ctx, sink="device", **kw)         # warm-up, buffers
t0 = time.perf_counter()
st = J.illumina(g, "", 2 * full_pairs, L, True, seed=2, ctx=ctx, sink="device", **kw)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
seen = [0, 0]


def sink(job, end, buf):
    seen[end] += len(buf)


J.illumina(g, "", 2 * (1 << 21), L, True, seed=1, ctx=ctx, sink=sink, **kw)
seen[0] = seen[1] = 0
ctx._genome = None
t0 = time.perf_counter()
st2 = J.illumina(g, "", 2 * full_pairs, L, True, seed=2, ctx=ctx, sink=sink, **kw)
torch.cuda.synchronize()
t_e2e = time.perf_counter() - t0
assert seen == st2["bytes_out"] == st["bytes_out"]
bg = {}
for level in (6, 1):
    zk = dict(compress=level, comp_engine="device")
    J.illumina(g, "", 2 * (1 << 21), L, True, seed=1, ctx=ctx, sink=sink, **kw, **zk)
    seen[0] = seen[1] = 0
    ctx._genome = None
    t0 = time.perf_counter()
    st3 = J.illumina(g, "", 2 * full_pairs, L, True, seed=2, ctx=ctx, sink=sink, **kw, **zk)
    torch.cuda.synchronize()
    t_z = time.perf_counter() - t0
    assert st3["bytes_out"] == st["bytes_out"] and seen[0] == st3["z_bytes"][0] + 28
    bg["level%d" % level] = {"end_to_end_s": t_z, "end_to_end_pairs_per_s": full_pairs / t_z, "bgzf_bytes": sum(seen),
                             "ratio": sum(st3["z_bytes"]) / sum(st3["bytes_out"]), "k_bgzf_ms_per_batch": st3["bgzf_ms"] / st3["batches"]}
seen[0], seen[1] = st2["bytes_out"]
print(json.dumps({"bgzf": bg, "workload": "3.1 Gb genome, 30x, PE150 HS25", "pairs": full_pairs, "fastq_bytes": sum(seen),
                  "device_resident_s": t_dev, "device_resident_pairs_per_s": full_pairs / t_dev,
                  "device_event_ms": st["run_ms"], "batches": st["batches"],
                  "end_to_end_s": t_e2e, "end_to_end_pairs_per_s": full_pairs / t_e2e,
                  "d2h_GBps": sum(seen) / t_e2e / 1e9}))
