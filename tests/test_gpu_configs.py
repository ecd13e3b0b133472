"""BASELINE.json configs[0..3] as parity cases on the GPU.

Each config runs at its named size (config 4 with the genome scaled, stated below) through
the C ABI.  Byte-for-byte comparison with the oracle is done on slices of the job -- the
library generates any contiguous pair-index range on request (shard = (index, count)), and the
oracle generates exactly the same range -- at the start, in the middle, across group
boundaries and at the end; the whole job is checked through size-independent properties:
record counts, R1/R2 alignment, output independent of batch size (checksum), per-position
quality histograms against the profile, mismatch rate against the quality model."""
import ctypes as C
import hashlib

import numpy as np
import pytest
from scipy import stats

import jackalope_b200 as J
from jackalope_b200 import _lib
from common import first_diff, oracle_run

pytestmark = pytest.mark.gpu


def oracle_shard(obj, n_reads, L, paired, seed, shard, hap_seqs=None, **kw):
    """Oracle output of shard (k, n) of every job of the run."""
    lib = _lib.lib()
    from oracle.compare import oracle_jobs
    out1, out2 = b"", b""
    for (jl, jh) in oracle_jobs(obj, n_reads, L, paired, seed, **kw):
        lo, hi = C.c_uint64(), C.c_uint64()
        assert lib.jlp_shard_range(jl, jh, shard[0], shard[1], C.byref(lo), C.byref(hi)) == 0
        if hi.value > lo.value:
            o = oracle_run(obj, n_reads, L, paired, seed, lo=lo.value, hi=hi.value, hap_seqs=hap_seqs, only_job=(jl, jh), **kw)
            out1 += o["r1"]
            out2 += o["r2"]
    return out1, out2


def check_slices(ctx, obj, n_reads, L, paired, seed, n_shards, which, hap_seqs=None, **kw):
    for k in which:
        r1, r2, st = J.illumina(obj, "", n_reads, L, paired, seed=seed, ctx=ctx, sink="memory", shard=(k, n_shards), **kw)
        o1, o2 = oracle_shard(obj, n_reads, L, paired, seed, (k, n_shards), hap_seqs=hap_seqs, **kw)
        d1, d2 = first_diff(r1, o1), first_diff(r2, o2)
        assert d1 is None and d2 is None, "shard %d: R1 byte %r, R2 byte %r" % (k, d1, d2)
        assert len(r1) > 0


def digest_run(ctx, obj, n_reads, L, paired, seed, **kw):
    """sha256 of R1 and R2 streamed batch by batch + record counts."""
    h = [hashlib.sha256(), hashlib.sha256()]
    lines = [0, 0]

    def sink(job, end, buf):
        b = bytes(buf)
        h[end].update(b)
        lines[end] += b.count(b"\n")

    st = J.illumina(obj, "", n_reads, L, paired, seed=seed, ctx=ctx, sink=sink, **kw)
    return h[0].hexdigest(), h[1].hexdigest(), lines, st


def test_config1_ref_10x1Mb_1e6_reads_pe100_hs25(ctx):
    g = J.random_genome(10, 1_000_000, seed=101)
    kw = dict(seq_sys="HS25")
    a1, a2, lines, st = digest_run(ctx, g, 1_000_000, 100, True, 7, **kw)
    assert lines == [4 * 500_000, 4 * 500_000] and st["pairs"] == 500_000
    b1, b2, _, st2 = digest_run(ctx, g, 1_000_000, 100, True, 7, batch_pairs=77_777, **kw)
    assert (a1, a2) == (b1, b2) and st2["batches"] == 7
    check_slices(ctx, g, 1_000_000, 100, True, 7, 250, [0, 24, 25, 125, 249], **kw)     # 2000 pairs each; 25 starts chrom1


def test_config2_haplotypes_8x10Mb_pe150(ctx):
    g = J.random_genome(10, 1_000_000, seed=102)
    haps = J.random_haplotypes(g, 8, sub_rate=0.01, indel_rate=0.001, seed=103)
    n_pairs = 8 * (10_000_000 * 10 // 300)
    kw = dict(frag_mean=400, seq_sys="HS25")
    a1, a2, lines, st = digest_run(ctx, haps, 2 * n_pairs, 150, True, 8, **kw)
    assert lines == [4 * n_pairs, 4 * n_pairs]
    ctx2_out = [ctx.haplotype_chrom(3, 7)]
    from common import hap_sequences
    hs = hap_sequences(haps)
    assert ctx2_out[0] == hs[3][7]
    n_sh = n_pairs // 1500
    check_slices(ctx, haps, 2 * n_pairs, 150, True, 8, n_sh, [0, n_sh // 8, n_sh // 8 + 1, n_sh // 2, n_sh - 1], hap_seqs=hs, **kw)


def test_config3_matepair_100Mb_dups_indels(ctx):
    g = J.random_genome(20, 5_000_000, seed=104)
    n_pairs = 100_000_000 * 10 // 300
    kw = dict(matepair=True, frag_mean=3000, frag_sd=500, prob_dup=0.02, ins_prob1=9e-4, del_prob1=1.1e-3,
              ins_prob2=1.5e-3, del_prob2=2.3e-3, seq_sys="HS25")
    a1, a2, lines, st = digest_run(ctx, g, 2 * n_pairs, 150, True, 9, **kw)
    assert lines == [4 * n_pairs, 4 * n_pairs]
    b1, b2, _, _ = digest_run(ctx, g, 2 * n_pairs, 150, True, 9, batch_pairs=400_000, **kw)
    assert (a1, a2) == (b1, b2)
    n_sh = n_pairs // 1500
    check_slices(ctx, g, 2 * n_pairs, 150, True, 9, n_sh, [0, 1, n_sh // 3, n_sh - 1], **kw)


def test_config4_96_haplotypes_uneven_sep_files(ctx, tmp_path):
    """configs[3] with the genome scaled from 500 Mb to 24 Mb (96 x 24 Mb = 2.3 Gb materialised;
    the full 48 Gb needs no other code path).  192 output files."""
    g = J.random_genome(12, 2_000_000, seed=105)
    haps = J.random_haplotypes(g, 96, sub_rate=0.005, indel_rate=0.0005, seed=106)
    probs = (1.0 / np.arange(1, 97)).tolist()
    n_reads = 600_000
    kw = dict(haplotype_probs=probs, sep_files=True, seq_sys="HS25")
    pre = str(tmp_path / "mx")
    J.illumina(haps, pre, n_reads, 150, True, seed=10, ctx=ctx, n_threads=4, overwrite=True, **kw)
    tot = [0, 0]
    for h in haps.hap_names:
        a = open("%s_%s_R1.fq" % (pre, h), "rb").read()
        b = open("%s_%s_R2.fq" % (pre, h), "rb").read()
        assert a.count(b"\n") == b.count(b"\n") and a.count(b"\n") % 4 == 0
        assert a == b"" or a.startswith(b"@" + h.encode() + b"-")
        tot[0] += a.count(b"\n") // 4
        tot[1] += b.count(b"\n") // 4
    assert tot == [n_reads // 2, n_reads // 2]
    # reads per haplotype follow haplotype_probs
    cnt = np.array([open("%s_%s_R1.fq" % (pre, h), "rb").read().count(b"\n") // 4 for h in haps.hap_names])
    p = np.array(probs) / np.sum(probs)
    assert stats.chisquare(cnt, p * cnt.sum()).pvalue > 1e-4
    from common import hap_sequences
    hs = hap_sequences(haps)
    check_slices(ctx, haps, n_reads, 150, True, 10, 40, [0, 39], hap_seqs=hs, **kw)


def test_config4_compressed_sep_files(ctx, tmp_path):
    """The same multiplexed run with compress = TRUE: 192 BGZF files written by the GPU; every one inflates to
    the plain run's file."""
    import gzip
    g = J.random_genome(12, 2_000_000, seed=105)
    haps = J.random_haplotypes(g, 96, sub_rate=0.005, indel_rate=0.0005, seed=106)
    probs = (1.0 / np.arange(1, 97)).tolist()
    n_reads = 600_000
    kw = dict(haplotype_probs=probs, sep_files=True, seq_sys="HS25")
    pre, zpre = str(tmp_path / "mx"), str(tmp_path / "mz")
    J.illumina(haps, pre, n_reads, 150, True, seed=10, ctx=ctx, n_threads=4, **kw)
    J.illumina(haps, zpre, n_reads, 150, True, seed=10, ctx=ctx, n_threads=4, compress=True, **kw)
    plain = comp = 0
    for h in haps.hap_names:
        for r in (1, 2):
            a = open("%s_%s_R%d.fq" % (pre, h, r), "rb").read()
            z = open("%s_%s_R%d.fq.gz" % (zpre, h, r), "rb").read()
            assert z[-28:-24] == b"\x1f\x8b\x08\x04" and gzip.decompress(z) == a     # ends with the BGZF EOF block
            plain += len(a)
            comp += len(z)
    assert comp < 0.37 * plain


def test_config5_slice_compressed_stream(ctx):
    """A 1e6-pair slice of the human-scale workload shape (24 chromosomes, PE150 HS25) streamed as BGZF at both
    device levels: the inflated stream has the plain stream's digest.  Every chunk the sink receives is a whole
    number of BGZF members, walked by their BSIZE fields."""
    import struct
    import zlib
    g = J.random_genome(24, 500_000, seed=107)
    n_reads = 2_000_000
    a1, a2, lines, st = digest_run(ctx, g, n_reads, 150, True, 11, seq_sys="HS25", batch_pairs=300_000)
    for level in (1, 6):
        h = [hashlib.sha256(), hashlib.sha256()]
        nz = [0, 0]

        def sink(job, end, buf):
            mv = memoryview(bytes(buf))
            nz[end] += len(mv)
            p = 0
            while p < len(mv):
                bsize = struct.unpack_from("<H", mv, p + 16)[0] + 1
                body = zlib.decompress(mv[p + 18:p + bsize - 8], -15)
                crc, isize = struct.unpack_from("<II", mv, p + bsize - 8)
                assert isize == len(body) and crc == zlib.crc32(body)
                h[end].update(body)
                p += bsize
            assert p == len(mv)

        stz = J.illumina(g, "", n_reads, 150, True, seed=11, ctx=ctx, sink=sink, seq_sys="HS25", compress=level,
                         comp_engine="device", batch_pairs=300_000)
        assert (h[0].hexdigest(), h[1].hexdigest()) == (a1, a2)
        assert stz["d2h_bytes"] == sum(stz["z_bytes"]) and nz[0] == stz["z_bytes"][0] + 28
        assert sum(stz["z_bytes"]) < (0.36 if level == 6 else 0.40) * sum(st["bytes_out"])


def test_model_statistics_at_scale(ctx):
    """4e5 PE150 HS25 pairs: per-position quality histograms against the profile, mismatch rate per
    position against sum_q P(q) 10^(-q/10), strand balance, fragment lengths against the Gamma law."""
    L, n_pairs = 150, 400_000
    g = J.random_genome(4, 2_500_000, seed=107)
    kw = dict(seq_sys="HS25", ins_prob1=0, del_prob1=0, ins_prob2=0, del_prob2=0, prob_dup=0)
    r1, r2, _ = J.illumina(g, "", 2 * n_pairs, L, True, seed=11, ctx=ctx, sink="memory", **kw)
    comp = np.zeros(256, dtype=np.uint8)
    comp[np.frombuffer(b"TCAG", np.uint8)] = np.frombuffer(b"AGTC", np.uint8)
    seqs = {n.encode(): s for n, s in zip(g.names, g.seqs)}
    for end, fq in enumerate((r1, r2)):
        prof = J.read_profile(None, "HS25", L, end + 1)
        lines = fq.split(b"\n")
        ids, reads, quals = lines[0:-1:4], lines[1:-1:4], lines[3:-1:4]
        assert len(reads) == n_pairs and all(len(r) == L for r in reads[:1000])
        R = np.frombuffer(b"".join(reads), np.uint8).reshape(n_pairs, L)
        Q = np.frombuffer(b"".join(quals), np.uint8).reshape(n_pairs, L) - 33
        # templates from the ID lines
        T = np.empty_like(R)
        rev = np.zeros(n_pairs, dtype=bool)
        for i, idl in enumerate(ids):
            f = idl.split(b"-")
            start = int(f[2])
            t = seqs[f[1]][start:start + L]
            if f[3][:1] == b"R":
                t = comp[t][::-1]
                rev[i] = True
            T[i] = t
        mism = R != T
        code = np.full(256, 4, dtype=np.int64)
        code[np.frombuffer(b"TCAG", np.uint8)] = np.arange(4)
        tc = code[T]
        pmin = 1.0
        for pos in range(0, L, 7):
            exp_q = np.zeros(64)
            exp_m = 0.0
            for nt in range(4):
                w = np.mean(tc[:, pos] == nt)
                qs, pr = prof["quals"][nt][pos], prof["qual_probs"][nt][pos]
                np.add.at(exp_q, qs, w * pr)
                exp_m += w * np.sum(pr * np.where(qs == 0, 1.0, 10.0 ** (-qs / 10.0)))
            obs = np.bincount(Q[:, pos], minlength=64).astype(np.float64)
            keep = exp_q * n_pairs >= 5
            assert obs[~keep].sum() <= 0.002 * n_pairs
            pq = stats.chisquare(obs[keep], exp_q[keep] / exp_q[keep].sum() * obs[keep].sum()).pvalue
            pm = stats.binomtest(int(mism[:, pos].sum()), n_pairs, exp_m).pvalue
            pmin = min(pmin, pq, pm)
        assert pmin > 1e-3 / (2 * (L // 7 + 1) * 2), pmin
        assert stats.binomtest(int(rev.sum()), n_pairs, 0.5).pvalue > 1e-4
    # fragment lengths: forward start + L .. reverse start: frag_len = rev_start + L - fwd_start
    s1 = np.array([int(x.split(b"-")[2]) for x in r1.split(b"\n")[0:-1:4]])
    s2 = np.array([int(x.split(b"-")[2]) for x in r2.split(b"\n")[0:-1:4]])
    fl = np.abs(s2 - s1) + L
    u = np.random.default_rng(1).random(fl.size)
    assert stats.kstest(fl + u, lambda x: np.where(x < L, 0.0, stats.gamma.cdf(x, a=16.0, scale=25.0))).pvalue > 1e-4
