#!/usr/bin/env python3
"""Where the time of a big haplotype run goes (BASELINE configs[3]: 96 haplotypes x 500 Mb): host copy of the
mutation records (jlp_add_haplotype), their upload + materialisation on first touch, generation."""
import json
import sys
import time
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jackalope_b200 as J  # noqa: E402
from jackalope_b200.genome import random_mutations  # noqa: E402
from concurrent.futures import ThreadPoolExecutor  # noqa: E402

n_haps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
g = J.random_genome(20, 25_000_000, seed=105)


def one(h):
    rng = np.random.default_rng([106, h])
    return [random_mutations(s, rng, 0.001, 0.0001, want_edits=False)[0] for s in g.seqs]


t0 = time.perf_counter()
with ThreadPoolExecutor(max_workers=16) as ex:
    muts = list(ex.map(one, range(n_haps)))
haps = J.Haplotypes(g, ["hap%d" % i for i in range(n_haps)], muts)
out = {"n_haps": n_haps, "build_records_s": time.perf_counter() - t0}
ctx = J.Context(0)
probs = (1.0 / np.arange(1, n_haps + 1)).tolist()
n_pairs = 500_000_000 * 10 // 300
kw = dict(seq_sys="HS25", haplotype_probs=probs, sep_files=True)
J.illumina(g, "", 200_000, 150, True, seed=1, ctx=ctx, sink="device", seq_sys="HS25")
for rep in range(2):
    ctx._genome = ctx._haps = None
    t0 = time.perf_counter()
    ctx.set_haplotypes(haps)
    t1 = time.perf_counter()
    ctx._check(ctx.lib.jlp_materialize_haplotypes(ctx.h), "materialize")
    t2 = time.perf_counter()
    st = J.illumina(haps, "", 2 * n_pairs, 150, True, seed=3, ctx=ctx, sink="device", **kw)
    t3 = time.perf_counter()
    n = [0]

    def sink(job, end, buf):
        n[0] += len(buf)

    st2 = J.illumina(haps, "", 2 * n_pairs, 150, True, seed=3, ctx=ctx, sink=sink, **kw)
    t4 = time.perf_counter()
    out["rep%d" % rep] = {"set_haplotypes_s": t1 - t0, "materialize_all_s": t2 - t1, "generate_device_only_s": t3 - t2,
                          "generate_stream_s": t4 - t3, "device_ms": st["device_ms"], "batches": st["batches"], "bytes": n[0]}
print(json.dumps(out, indent=1))
