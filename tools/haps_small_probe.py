"""BASELINE configs[1] (8 haplotypes on 10 x 1 Mb, PE150) as bench.py's extra workload runs it -- one whole illumina() call with
the genome and haplotype records uploaded and the haplotypes materialised inside -- several times, split into its parts:
    python tools/haps_small_probe.py [repetitions]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jackalope_b200 as J

ctx = J.Context(0)
g = J.random_genome(10, 1_000_000, seed=101)
haps = J.random_haplotypes(g, 8, sub_rate=0.01, indel_rate=0.001, seed=103)
n_pairs = 8 * (10_000_000 * 10 // 300)
out = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    n = [0, 0]

    def sink(job, end, buf):
        n[end] += len(buf)

    ctx._genome = ctx._haps = None
    t0 = time.perf_counter()
    ctx.set_haplotypes(haps)
    t1 = time.perf_counter()
    ctx._check(ctx.lib.jlp_materialize_haplotypes(ctx.h), "materialize")
    t2 = time.perf_counter()
    st = J.illumina(haps, "", 2 * n_pairs, 150, True, seed=7 + rep, ctx=ctx, sink=sink, seq_sys="HS25", frag_mean=400)
    t3 = time.perf_counter()
    ctx._genome = ctx._haps = None
    t4 = time.perf_counter()
    st = J.illumina(haps, "", 2 * n_pairs, 150, True, seed=7 + rep, ctx=ctx, sink=sink, seq_sys="HS25", frag_mean=400)
    t5 = time.perf_counter()
    out.append({"set_haplotypes_ms": 1e3 * (t1 - t0), "materialize_ms": 1e3 * (t2 - t1), "generate_ms": 1e3 * (t3 - t2), "whole_call_ms": 1e3 * (t5 - t4)})
print(json.dumps(out, indent=1))
