"""GPU parity tests proper: every call goes through the C ABI (jackalope_b200._lib)
and is compared byte for byte with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import jackalope_b200 as J
from common import fastq_records, first_diff, hap_sequences, oracle_run

pytestmark = pytest.mark.gpu


def small_genome(seed=1, n=3, length=5000, with_n=True):
    g = J.random_genome(n, length, seed=seed)
    if with_n:
        s = g.seqs[0].copy()
        s[100:160] = ord("N")
        s[1000] = ord("x")
        g.seqs[0] = s
    return g


def check(ctx, obj, n_reads, L, paired, seed, **kw):
    r1, r2, st = J.illumina(obj, "", n_reads, L, paired, seed=seed, ctx=ctx, sink="memory", **kw)
    o = oracle_run(obj, n_reads, L, paired, seed, **kw)
    d1, d2 = first_diff(r1, o["r1"]), first_diff(r2, o["r2"])
    assert d1 is None, "R1 differs at byte %d: gpu=%r oracle=%r" % (d1, r1[max(0, d1 - 80):d1 + 40], o["r1"][max(0, d1 - 80):d1 + 40])
    assert d2 is None, "R2 differs at byte %d: gpu=%r oracle=%r" % (d2, r2[max(0, d2 - 80):d2 + 40], o["r2"][max(0, d2 - 80):d2 + 40])
    return r1, r2, st


def test_materialize_matches_oracle(ctx):
    g = small_genome(seed=3, n=4, length=20000, with_n=False)
    haps = J.random_haplotypes(g, 3, sub_rate=0.02, indel_rate=0.01, seed=5)
    ctx.set_haplotypes(haps)
    want = hap_sequences(haps)
    for h in range(3):
        for c in range(4):
            assert ctx.haplotype_chrom(h, c) == want[h][c]


@pytest.mark.parametrize("paired,matepair", [(False, False), (True, False), (True, True)])
def test_ref_default_args(ctx, paired, matepair):
    g = small_genome()
    check(ctx, g, 2000, 100, paired, seed=11, matepair=matepair)


def test_ref_pe150_hs25(ctx):
    g = small_genome(seed=2, n=5, length=20000)
    check(ctx, g, 6000, 150, True, seed=12, seq_sys="HS25")


def test_high_indels_dups_barcode(ctx):
    g = small_genome(seed=4)
    check(ctx, g, 3000, 100, True, seed=13, ins_prob1=0.02, del_prob1=0.03, ins_prob2=0.05, del_prob2=0.01,
          prob_dup=0.4, read_pool_size=14, barcodes=["ACGTTG"])


def test_short_fragments_and_chromosomes(ctx):
    # fragments shorter than the read, and a chromosome shorter than any fragment
    g = J.RefGenome(["a", "b", "c"], [J.random_genome(1, 60, seed=5).seqs[0], J.random_genome(1, 4000, seed=6).seqs[0],
                                      J.random_genome(1, 130, seed=7).seqs[0]])
    check(ctx, g, 3000, 100, True, seed=14, frag_mean=120, frag_sd=40, frag_len_min=20, ins_prob1=0.01, del_prob1=0.01)


def test_haplotypes_pooled_and_sep(ctx):
    g = small_genome(seed=8, n=3, length=8000, with_n=False)
    haps = J.random_haplotypes(g, 4, sub_rate=0.02, indel_rate=0.005, seed=9)
    check(ctx, haps, 4000, 100, True, seed=15, haplotype_probs=[1, 2, 0, 4], barcodes=["AC", "GT", "TT", "CA"])
    check(ctx, haps, 4000, 100, True, seed=16, haplotype_probs=[1, 2, 0.5, 4], sep_files=True)
    check(ctx, haps, 1500, 100, False, seed=17)


def test_batches_and_shards_do_not_change_output(ctx):
    g = small_genome(seed=10)
    r1, r2, _ = check(ctx, g, 5000, 100, True, seed=18, prob_dup=0.3)
    b1, b2, st = J.illumina(g, "", 5000, 100, True, seed=18, ctx=ctx, sink="memory", prob_dup=0.3, batch_pairs=333)
    assert (b1, b2) == (r1, r2) and st["batches"] == 8
    parts = [J.illumina(g, "", 5000, 100, True, seed=18, ctx=ctx, sink="memory", prob_dup=0.3, batch_pairs=400,
                        shard=(i, 3)) for i in range(3)]
    assert b"".join(p[0] for p in parts) == r1 and b"".join(p[1] for p in parts) == r2


def test_files_and_sep_files(ctx, tmp_path):
    g = small_genome(seed=11, with_n=False)
    haps = J.random_haplotypes(g, 3, seed=12)
    pre = str(tmp_path / "reads")
    assert J.illumina(g, pre, 1000, 100, True, seed=19, ctx=ctx) is None
    r1, r2, _ = J.illumina(g, "", 1000, 100, True, seed=19, ctx=ctx, sink="memory")
    assert open(pre + "_R1.fq", "rb").read() == r1 and open(pre + "_R2.fq", "rb").read() == r2
    with pytest.raises(J.JackalopeError):
        J.illumina(g, pre, 1000, 100, True, seed=19, ctx=ctx)           # exists, overwrite = FALSE
    J.illumina(haps, pre, 900, 100, True, seed=20, ctx=ctx, sep_files=True, overwrite=True)
    m1, m2, _ = J.illumina(haps, "", 900, 100, True, seed=20, ctx=ctx, sep_files=True, sink="memory")
    cat1 = b"".join(open("%s_%s_R1.fq" % (pre, h), "rb").read() for h in haps.hap_names)
    cat2 = b"".join(open("%s_%s_R2.fq" % (pre, h), "rb").read() for h in haps.hap_names)
    assert cat1 == m1 and cat2 == m2
    # shape checks of tests/testthat/test-sequencer.R:31-77,172-272
    for fq in (r1, r2, m1, m2):
        recs = fastq_records(fq)
        assert all(r[0].startswith(b"@") and r[2] == b"+" and len(r[1]) == len(r[3]) for r in recs)
    assert len(fastq_records(r1)) == len(fastq_records(r2)) == 500


# ---- golden vectors (bytes produced by the unmodified reference, tests/golden/make_golden.py)

from golden.cases import CASES, load  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(ctx, name):
    make, n_reads, L, paired, seed, kw = CASES[name]
    want1, want2, _ = load(name)
    r1, r2, st = J.illumina(make(), "", n_reads, L, paired, seed=seed, ctx=ctx, sink="memory", **kw)
    assert first_diff(r1, want1) is None and first_diff(r2, want2) is None
    assert st["kernel_launches"] >= 4


def test_pair_geometry_known_answers(ctx, tmp_path):
    """tests/testthat/test-sequencer.R:81-161: error-free custom profile (quality 255),
    chromosome C25 N150 T25, fragment forced to 200, no indels."""
    prof = str(tmp_path / "prof.txt")
    with open(prof, "w") as fh:
        for nt in "ACGT":
            for pos in range(100):
                fh.write("%s\t%d\t255\n%s\t%d\t1000000\n" % (nt, pos, nt, pos))
    g = J.RefGenome(["chrom0"], [b"C" * 25 + b"N" * 150 + b"T" * 25])
    common = dict(profile1=prof, profile2=prof, frag_mean=400, frag_sd=100, frag_len_min=200, frag_len_max=200,
                  ins_prob1=0, del_prob1=0, ins_prob2=0, del_prob2=0, ctx=ctx, sink="memory")
    for matepair, want in ((False, {b"C" * 25 + b"N" * 75, b"A" * 25 + b"N" * 75}),
                           (True, {b"N" * 75 + b"T" * 25, b"N" * 75 + b"G" * 25})):
        r1, r2, _ = J.illumina(g, "", 1000, 100, True, seed=5, matepair=matepair, **common)
        for fq in (r1, r2):
            recs = fastq_records(fq)
            assert len(recs) == 500 and {r[1] for r in recs} == want
            assert all(len(r[3]) == 100 for r in recs)


def test_writer_threads_and_many_batches(ctx, tmp_path):
    g = small_genome(seed=31)
    pre = str(tmp_path / "w")
    r1, r2, _ = J.illumina(g, "", 6000, 100, True, seed=32, ctx=ctx, sink="memory")
    for nthr in (1, 4):
        J.illumina(g, pre, 6000, 100, True, seed=32, ctx=ctx, batch_pairs=500, n_threads=nthr, overwrite=True)
        assert open(pre + "_R1.fq", "rb").read() == r1 and open(pre + "_R2.fq", "rb").read() == r2


def test_stream_sink_hands_out_the_same_bytes(ctx):
    g = small_genome(seed=21)
    r1, r2, _ = J.illumina(g, "", 3000, 100, True, seed=22, ctx=ctx, sink="memory")
    got = {0: bytearray(), 1: bytearray()}
    st = J.illumina(g, "", 3000, 100, True, seed=22, ctx=ctx, batch_pairs=256,
                    sink=lambda job, end, buf: got[end].extend(bytes(buf)))
    assert bytes(got[0]) == r1 and bytes(got[1]) == r2 and st["batches"] == 6 and st["d2h_bytes"] == len(r1) + len(r2)
    dev = J.illumina(g, "", 3000, 100, True, seed=22, ctx=ctx, sink="device")
    assert dev["bytes_out"] == [len(r1), len(r2)] and dev["d2h_bytes"] == 0 and dev["run_ms"] > 0


def test_errors_surface_like_the_reference(ctx, tmp_path):
    g = small_genome(seed=23)
    with pytest.raises(J.JackalopeError):          # haplotype_probs with a ref_genome
        J.illumina(g, "", 100, 100, True, seed=1, ctx=ctx, sink="memory", haplotype_probs=[1.0])
    with pytest.raises(J.JackalopeError):          # params == NULL -> JLP_ERR_ARG with a message
        ctx._check(ctx.lib.jlp_illumina_ref(ctx.h, None, None), "jlp_illumina_ref")
    blocker = tmp_path / "blocker"
    blocker.write_text("a file where a directory is needed")
    with pytest.raises((RuntimeError, OSError)):   # unopenable output file (src/io.h:288-290)
        J.illumina(g, str(blocker / "reads"), 100, 100, True, seed=1, ctx=ctx, overwrite=True)
    with pytest.raises(J.JackalopeError, match="gzip cannot be performed using multiple threads"):
        J.illumina(g, str(tmp_path / "z"), 100, 100, True, seed=1, ctx=ctx, compress=True, comp_method="gzip", n_threads=2)


def bgzf_blocks(z: bytes):
    """Walk the BGZF blocks of a file: (compressed size, uncompressed size) per block."""
    import struct
    out, p = [], 0
    while p < len(z):
        assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 10:p + 16] == b"\x06\x00BC\x02\x00"
        bsize = struct.unpack_from("<H", z, p + 16)[0] + 1
        isize = struct.unpack_from("<I", z, p + bsize - 4)[0]
        out.append((bsize, isize))
        p += bsize
    assert p == len(z)
    return out


def test_compressed_output(ctx, tmp_path):
    """compress > 0 (write_reads_cpp_, src/hts.h:441-500): <prefix>_R{1,2}.fq.gz, bgzip or gzip; the
    decompressed bytes are the uncompressed run's."""
    import gzip
    g = small_genome(seed=41)
    r1, r2, _ = J.illumina(g, "", 8000, 100, True, seed=42, ctx=ctx, sink="memory")
    for method, nthr, level in (("bgzip", 1, True), ("bgzip", 4, 3), ("gzip", 1, 9)):
        pre = str(tmp_path / ("z_%s_%d" % (method, nthr)))
        J.illumina(g, pre, 8000, 100, True, seed=42, ctx=ctx, compress=level, comp_method=method, n_threads=nthr,
                   batch_pairs=1500)
        z1, z2 = open(pre + "_R1.fq.gz", "rb").read(), open(pre + "_R2.fq.gz", "rb").read()
        assert gzip.decompress(z1) == r1 and gzip.decompress(z2) == r2
        if method == "bgzip":
            for z in (z1, z2):
                blocks = bgzf_blocks(z)
                assert blocks[-1] == (28, 0) and all(b <= 65536 and 0 < i <= 0xff00 for b, i in blocks[:-1])
        with pytest.raises(J.JackalopeError, match="already exists"):
            J.illumina(g, pre, 8000, 100, True, seed=42, ctx=ctx, compress=level, comp_method=method, n_threads=nthr)
    haps = J.random_haplotypes(g, 2, seed=43)
    pre = str(tmp_path / "zh")
    J.illumina(haps, pre, 3000, 100, True, seed=44, ctx=ctx, compress=True, sep_files=True, n_threads=2)
    m1, _, _ = J.illumina(haps, "", 3000, 100, True, seed=44, ctx=ctx, sep_files=True, sink="memory")
    assert b"".join(gzip.decompress(open("%s_%s_R1.fq.gz" % (pre, h), "rb").read()) for h in haps.hap_names) == m1


def test_long_names_take_the_long_id_path(ctx):
    """ID lines longer than the 64 bytes the plan record inlines are written by the fallback path."""
    base = J.random_genome(2, 3000, seed=51)
    g = J.RefGenome(["chromosome_with_a_very_long_name_%d_" % i + "x" * 40 for i in range(2)], base.seqs)
    check(ctx, g, 1200, 100, True, seed=52)
    haps = J.random_haplotypes(g, 2, seed=53, names=["haplotype_" + "y" * 30, "h"])
    check(ctx, haps, 1200, 100, True, seed=54, sep_files=True, barcodes=["ACGT", "TTGCA"])


@pytest.mark.parametrize("L,seq_sys,paired", [(250, "MSv1", True), (250, "MSv3", False), (36, "GA1", True), (75, "NS50", True),
                                              (50, "MinS", False), (125, "HS25", True), (150, "HSXt", True)])
def test_other_profiles_and_read_lengths(ctx, L, seq_sys, paired):
    g = small_genome(seed=55, n=2, length=6000)
    check(ctx, g, 1500, L, paired, seed=56 + L, seq_sys=seq_sys, frag_mean=max(400, 2 * L), frag_sd=60,
          ins_prob1=0.004, del_prob1=0.006, ins_prob2=0.005, del_prob2=0.003)


def test_heavy_deletions_long_spans(ctx):
    """deletion-heavy ends consume templates much longer than the read (span > 256 positions)"""
    g = small_genome(seed=57, n=2, length=9000, with_n=False)
    check(ctx, g, 1000, 100, True, seed=58, frag_mean=900, frag_sd=100, del_prob1=0.7, ins_prob1=0.05, del_prob2=0.5,
          ins_prob2=0.3)


def test_single_chromosome_shorter_than_read(ctx):
    g = J.RefGenome(["tiny"], [J.random_genome(1, 40, seed=59).seqs[0]])
    check(ctx, g, 400, 100, True, seed=60)
    check(ctx, g, 200, 100, False, seed=61, barcodes=["ACG"])


def test_degenerate_counts(ctx, tmp_path):
    """an odd n_reads is floored to whole pairs (src/hts.h:334-336); zero pairs give empty files; a haplotype
    with probability 0 gets empty files of its own with sep_files"""
    g = small_genome(seed=71, with_n=False)
    pre = str(tmp_path / "d")
    J.illumina(g, pre, 1, 100, True, seed=72, ctx=ctx)
    assert open(pre + "_R1.fq", "rb").read() == b"" and open(pre + "_R2.fq", "rb").read() == b""
    r1, r2, st = J.illumina(g, "", 7, 100, True, seed=73, ctx=ctx, sink="memory")
    assert len(fastq_records(r1)) == len(fastq_records(r2)) == 3 and st["pairs"] == 3
    check(ctx, g, 7, 100, True, seed=73)
    check(ctx, g, 1, 100, False, seed=74)
    haps = J.random_haplotypes(g, 3, seed=75)
    J.illumina(haps, pre, 600, 100, True, seed=76, ctx=ctx, sep_files=True, haplotype_probs=[1, 0, 1], overwrite=True)
    assert open("%s_hap1_R1.fq" % pre, "rb").read() == b""
    assert open("%s_hap0_R1.fq" % pre, "rb").read().count(b"\n") + open("%s_hap2_R1.fq" % pre, "rb").read().count(b"\n") == 4 * 300
    check(ctx, haps, 600, 100, True, seed=76, sep_files=True, haplotype_probs=[1, 0, 1])


def test_create_genome_on_device(ctx):
    """create_genome: chromosomes generated in HBM are the oracle's byte for byte;
    the genome is resident, so illumina() on it uploads nothing and still matches the oracle."""
    from scipy import stats
    from oracle import harness as H
    for pi in ([0.25, 0.25, 0.25, 0.25], [0.1, 0.2, 0.3, 0.4], [0, 1, 0, 3]):
        g = J.create_genome(3, 7001, pi_tcag=pi, seed=81, ctx=ctx)
        assert g.names == ["chrom0", "chrom1", "chrom2"]
        for c in range(3):
            assert g.chrom(c) == H.create_chrom(81, c, 7001, pi)
    g = J.create_genome(5, 40000, len_sd=8000, pi_tcag=[0.1, 0.2, 0.3, 0.4], seed=82, ctx=ctx)
    assert len(set(int(s) for s in g.sizes())) == 5 and min(g.sizes()) >= 1
    cnt = np.bincount(np.concatenate(g.seqs), minlength=256)[np.frombuffer(b"TCAG", np.uint8)]
    assert stats.chisquare(cnt, np.array([0.1, 0.2, 0.3, 0.4]) * cnt.sum()).pvalue > 1e-4
    before = J.illumina(g, "", 2, 100, True, seed=1, ctx=ctx, sink="device")["h2d_bytes"]
    r1, r2, st = check(ctx, g, 3000, 100, True, seed=83)
    assert st["h2d_bytes"] - before < 1_000_000          # tables only: the genome was never uploaded
    with pytest.raises(J.JackalopeError):
        J.create_genome(0, 100, ctx=ctx)
    with pytest.raises(J.JackalopeError):
        J.create_genome(2, 100, pi_tcag=[0, 0, 0, 0], ctx=ctx)


def _flat_profile_file(path, L, quals=(2, 20, 30, 38), counts=(10, 100, 1000, 5000)):
    """An ART-format profile with the same four qualities at every position (cumulative counts)."""
    cum = np.cumsum(counts)
    with open(path, "w") as fh:
        for nt in "ACGT":
            for pos in range(L):
                fh.write("%s\t%d\t%s\n%s\t%d\t%s\n" % (nt, pos, "\t".join(map(str, quals)), nt, pos, "\t".join(map(str, cum))))


def test_long_reads_custom_profile(ctx, tmp_path):
    """Read lengths beyond one round of the per-lane work (256 positions per end in phase A, 256 bytes of template
    window per cp.async round): a custom 300-position profile, paired, with indels and duplicates, byte for byte."""
    prof = str(tmp_path / "p300.txt")
    _flat_profile_file(prof, 300)
    g = J.random_genome(3, 50_000, seed=61)
    kw = dict(profile1=prof, profile2=prof, frag_mean=700, frag_sd=120, ins_prob1=0.002, del_prob1=0.002, ins_prob2=0.003,
              del_prob2=0.001, prob_dup=0.1)
    r1, r2, st = J.illumina(g, "", 6000, 300, True, seed=62, ctx=ctx, sink="memory", **kw)
    o = oracle_run(g, 6000, 300, True, 62, **kw)
    assert first_diff(r1, o["r1"]) is None and first_diff(r2, o["r2"]) is None
    assert all(len(x) == 300 for x in r1.split(b"\n")[1:4 * 50:4])


def test_read_length_too_long_for_the_read_kernel_is_reported(ctx, tmp_path):
    """A read whose record does not fit the shared memory of an SM (several thousand positions): a clear
    JLP_ERR_UNSUPPORTED instead of a CUDA launch failure (the reference has no such limit)."""
    prof = str(tmp_path / "p60k.txt")
    _flat_profile_file(prof, 60_000, quals=(30,), counts=(1,))
    g = J.random_genome(1, 200_000, seed=63)
    with pytest.raises(RuntimeError, match="shared memory"):
        J.illumina(g, "", 10, 60_000, False, seed=64, ctx=ctx, sink="memory", profile1=prof, frag_mean=80_000, frag_sd=100)
    # 20 000 positions still fit (one warp per CTA, tables in global memory) and must be right
    prof = str(tmp_path / "p20k.txt")
    _flat_profile_file(prof, 20_000, quals=(20, 35), counts=(1, 9))
    kw = dict(profile1=prof, frag_mean=30_000, frag_sd=500)
    r1, _, _ = J.illumina(g, "", 12, 20_000, False, seed=65, ctx=ctx, sink="memory", **kw)
    o = oracle_run(g, 12, 20_000, False, 65, **kw)
    assert first_diff(r1, o["r1"]) is None and r1.count(b"\n") == 48
