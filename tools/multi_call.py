#!/usr/bin/env python3
"""ONE illumina() call on all visible GPUs (jlp_ctx_create_multi): the human-scale workload of bench.py (3.1 Gb genome,
PE150 HS25) into one ordered pair of files, plain and BGZF, and device-only; plus a byte check of the head and tail of
the plain files against single-device slices of the same run.

    python tools/multi_call.py [pairs] [out_dir]        (default 6e7 pairs into /dev/shm)
"""
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

n_pairs = int(float(sys.argv[1])) if len(sys.argv) > 1 else 60_000_000
base = sys.argv[2] if len(sys.argv) > 2 else "/dev/shm"
lens, L, kw, full_pairs = bench.workload("human_pe150_hs25", 3_100_000_000)
total = int(lens.sum())
pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
flat = pinned.numpy()
bench.make_genome_into(flat, lens, 20261018)
off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
g.flat = lambda: (flat, off.astype(np.uint64))
n_gpus = torch.cuda.device_count()
m = J.Context(devices="all")
one = J.Context(0)
out = {"n_gpus": n_gpus, "pairs": n_pairs, "threads": bench.host_threads()}
d = tempfile.mkdtemp(dir=base)
try:
    nthr = min(bench.host_threads(), 64)
    J.illumina(g, os.path.join(d, "w"), 2 * (1 << 20) * n_gpus, L, True, seed=1, ctx=m, n_threads=nthr, overwrite=True, **kw)   # warm-up: buffers
    st = J.illumina(g, "", 2 * n_pairs, L, True, seed=5, ctx=m, sink="device", **kw)
    t0 = time.perf_counter()
    st = J.illumina(g, "", 2 * n_pairs, L, True, seed=5, ctx=m, sink="device", **kw)
    out["device_only_s"] = time.perf_counter() - t0
    out["device_only_pairs_per_s"] = n_pairs / out["device_only_s"]
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    m._genome = None
    t0 = time.perf_counter()
    J.illumina(g, os.path.join(d, "p"), 2 * n_pairs, L, True, seed=5, ctx=m, n_threads=nthr, overwrite=True, **kw)
    out["files_plain_s"] = time.perf_counter() - t0
    sz = [os.path.getsize(os.path.join(d, "p_R%d.fq" % r)) for r in (1, 2)]
    out["files_plain_bytes"] = sz
    out["files_plain_pairs_per_s"] = n_pairs / out["files_plain_s"]
    out["files_plain_GBps"] = sum(sz) / out["files_plain_s"] / 1e9
    assert sz == st["bytes_out"], (sz, st["bytes_out"])
    # head and tail of the one file set against single-device slices of the same run
    S = n_pairs // 4096
    for k in (0, S - 1):
        r1, r2, _ = J.illumina(g, "", 2 * n_pairs, L, True, seed=5, ctx=one, sink="memory", shard=(k, S), **kw)
        for r, want in ((1, r1), (2, r2)):
            with open(os.path.join(d, "p_R%d.fq" % r), "rb") as fh:
                if k:
                    fh.seek(-len(want), os.SEEK_END)
                got = fh.read(len(want))
            assert got == want, "slice %d of R%d differs from the single-device run" % (k, r)
    out["files_identical_to_single_device_slices"] = True
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    J.illumina(g, os.path.join(d, "w"), 2 * (1 << 20) * n_gpus, L, True, seed=1, ctx=m, n_threads=nthr, compress=True, overwrite=True, **kw)   # warm-up: the coder's buffers
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    m._genome = None
    t0 = time.perf_counter()
    J.illumina(g, os.path.join(d, "z"), 2 * n_pairs, L, True, seed=5, ctx=m, n_threads=nthr, compress=True, overwrite=True, **kw)
    out["files_bgzf_s"] = time.perf_counter() - t0
    zs = [os.path.getsize(os.path.join(d, "z_R%d.fq.gz" % r)) for r in (1, 2)]
    out["files_bgzf_bytes"] = zs
    out["files_bgzf_pairs_per_s"] = n_pairs / out["files_bgzf_s"]
    out["leftover_part_files"] = [f for f in os.listdir(d) if ".part" in f]
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out, indent=1))
