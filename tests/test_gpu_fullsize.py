"""Parity at the sizes bench.py reports: BASELINE.json configs[4] at its full 3.1 Gb / 3.1e8 pairs and configs[3] at
its full 96 haplotypes x 500 Mb -- 64-bit template addresses, 9-digit start coordinates, the 24-entry / 1920-entry
group search, the deferred per-chromosome genome upload and 48 Gb of materialised haplotypes are all in play.
The library generates any contiguous pair-index range of a job on request (shard = (index, count)); the oracle
generates exactly the same range from the same seed, materialising (get_chrom_full restatement) only the
(haplotype, chromosome) groups that range can touch.  Slices: the first, one that crosses a chromosome boundary,
the last; for configs[3] three of the 96 haplotypes' files.  Reference behaviour: IlluminaHaplotypes::one_read,
/root/reference/src/hts_illumina.cpp:495-536; chrom_indels_frag :191-226."""
import ctypes as C

import numpy as np
import pytest

import jackalope_b200 as J
from jackalope_b200 import _lib
from common import explain_diff, first_diff, lazy_hap_sequences, oracle_jobs, oracle_run

pytestmark = pytest.mark.gpu

HUMAN_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]


def big_genome(total, weights, seed):
    """Uniform TCAG chromosomes with the given length proportions, in ONE contiguous buffer (no second copy)."""
    w = np.asarray(weights, dtype=np.float64)
    lens = np.floor(w / w.sum() * total).astype(np.int64)
    lens[0] += total - lens.sum()
    flat = np.empty(total, dtype=np.uint8)
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"TCAG", dtype=np.uint8)
    for o in range(0, total, 1 << 26):
        n = min(1 << 26, total - o)
        np.take(lut, rng.integers(0, 2 ** 63, size=(n + 7) // 8, dtype=np.int64).view(np.uint8)[:n] & 3, out=flat[o:o + n])
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    g = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
    g.flat = lambda: (flat, off.astype(np.uint64))
    return g


def many_haplotypes(g, n_haps, sub_rate, indel_rate, seed):
    """random_haplotypes with one generator per haplotype, spread over host threads (numpy's sorts release the GIL)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from jackalope_b200.genome import random_mutations

    def one(h):
        rng = np.random.default_rng([seed, h])
        return [random_mutations(seq, rng, sub_rate, indel_rate, want_edits=False)[0] for seq in g.seqs]

    with ThreadPoolExecutor(max_workers=min(16, len(os.sched_getaffinity(0)))) as ex:
        muts = list(ex.map(one, range(n_haps)))
    return J.Haplotypes(g, ["hap%d" % i for i in range(n_haps)], muts)


def shard_bounds(jl, jh, k, n):
    lo, hi = C.c_uint64(), C.c_uint64()
    assert _lib.lib().jlp_shard_range(jl, jh, k, n, C.byref(lo), C.byref(hi)) == 0
    return lo.value, hi.value


def test_config5_human_scale_slices_of_the_full_job():
    total = 3_100_000_000
    g = big_genome(total, HUMAN_MB, seed=20261018)
    L, seed = 150, 77
    n_pairs = total * 30 // (2 * L)                       # 3.1e8 pairs, 30x
    n_reads = 2 * n_pairs
    kw = dict(seq_sys="HS25")
    S = n_pairs // 2000
    # a slice that straddles the boundary between chromosome 9 and 10, one whose starts have 9 digits on chrom0
    from oracle.compare import _prepare, group_counts, DEFAULTS
    a = dict(DEFAULTS)
    a.update(kw)
    p = _prepare(g, "x", n_reads, L, True, a["frag_mean"], a["frag_sd"], a["matepair"], a["seq_sys"], a["profile1"], a["profile2"],
                 a["ins_prob1"], a["del_prob1"], a["ins_prob2"], a["del_prob2"], a["frag_len_min"], a["frag_len_max"],
                 a["haplotype_probs"], a["barcodes"], a["prob_dup"], a["sep_files"], a["compress"], a["comp_method"], a["n_threads"],
                 a["read_pool_size"], a["show_progress"], True, seed, None, None, check_files=False)[0]
    off = np.concatenate(([0], np.cumsum(group_counts(p, g, False)))).astype(np.int64)
    k_cross = None
    for k in range(int(off[10]) * S // n_pairs - 2, int(off[10]) * S // n_pairs + 3):
        lo, hi = shard_bounds(0, n_pairs, k, S)
        if lo < off[10] < hi:
            k_cross = k
    assert k_cross is not None
    ctx = J.Context(0)
    try:
        for k in (0, k_cross, S // 3, S - 1):
            # a fresh deferred upload per slice: only the chromosomes the slice reads cross PCIe
            ctx._genome = None
            r1, r2, st = J.illumina(g, "", n_reads, L, True, seed=seed, ctx=ctx, sink="memory", shard=(k, S), **kw)
            lo, hi = shard_bounds(0, n_pairs, k, S)
            o = oracle_run(g, n_reads, L, True, seed, lo=lo, hi=hi, **kw)
            d1, d2 = first_diff(r1, o["r1"]), first_diff(r2, o["r2"])
            assert d1 is None and d2 is None, "slice %d (pairs %d..%d):\nR1 %s\nR2 %s" % (k, lo, hi, explain_diff(r1, o["r1"]), explain_diff(r2, o["r2"]))
            assert st["pairs"] == hi - lo and r1.count(b"\n") == 4 * (hi - lo)
            assert st["h2d_bytes"] < 0.6e9, "the deferred upload copied more than the touched chromosomes"
            if k == 0:       # chrom0 is 248 Mb: most of its start coordinates have 9 digits
                assert sum(len(x.split(b"-")[2]) == 9 for x in r1.split(b"\n")[0:-1:4]) > 500
    finally:
        ctx.close()


def test_config4_96_haplotypes_500Mb_sep_files_full_size():
    total = 500_000_000
    g = big_genome(total, [1.0] * 20, seed=105)            # 20 chromosomes x 25 Mb
    haps = many_haplotypes(g, 96, sub_rate=0.001, indel_rate=0.0001, seed=106)
    probs = (1.0 / np.arange(1, 97)).tolist()
    L, seed = 150, 10
    n_pairs = total * 10 // (2 * L)                        # 10x of the 500 Mb genome over the multiplexed library
    n_reads = 2 * n_pairs
    kw = dict(haplotype_probs=probs, sep_files=True, seq_sys="HS25")
    ctx = J.Context(0)
    try:
        ctx.set_haplotypes(haps)                           # 96 x 500 Mb = 48 Gb materialised in HBM
        # spot check of the materialisation against the oracle's get_chrom_full
        lazy = lazy_hap_sequences(haps)
        for h, c in ((0, 0), (57, 13), (95, 19)):
            assert ctx.haplotype_chrom(h, c) == lazy(h, c)
        # the whole job through the stream sink: one job per haplotype, R1/R2 aligned, counts add up
        lines = np.zeros((96, 2), dtype=np.int64)

        def sink(job, end, buf):
            lines[job, end] += bytes(buf).count(b"\n")

        st = J.illumina(haps, "", n_reads, L, True, seed=seed, ctx=ctx, sink=sink, **kw)
        assert st["pairs"] == n_pairs and lines.sum() == 8 * n_pairs and np.all(lines[:, 0] == lines[:, 1])
        jobs = oracle_jobs(haps, n_reads, L, True, seed, **kw)
        assert [int(x) // 4 for x in lines[:, 0]] == [jh - jl for jl, jh in jobs]
        # slices of three haplotypes' files against the oracle
        S = max(1, (jobs[95][1] - jobs[95][0]) // 1500)
        for k in (0, S // 2, S - 1):
            got = {}

            def take(job, end, buf):
                got[(job, end)] = got.get((job, end), b"") + bytes(buf)

            J.illumina(haps, "", n_reads, L, True, seed=seed, ctx=ctx, sink=take, shard=(k, S), **kw)
            for h in (0, 41, 95):
                jl, jh = jobs[h]
                lo, hi = shard_bounds(jl, jh, k, S)
                o = oracle_run(haps, n_reads, L, True, seed, lo=lo, hi=hi, hap_seqs=lazy, only_job=(jl, jh), **kw)
                d1, d2 = first_diff(got[(h, 0)], o["r1"]), first_diff(got[(h, 1)], o["r2"])
                assert d1 is None and d2 is None, "hap %d slice %d (pairs %d..%d of job %d..%d):\nR1 %s\nR2 %s" % (
                    h, k, lo, hi, jl, jh, explain_diff(got[(h, 0)], o["r1"]), explain_diff(got[(h, 1)], o["r2"]))
                assert o["r1"].startswith(b"@hap%d-" % h) and hi > lo
    finally:
        ctx.close()
