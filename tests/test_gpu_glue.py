"""The Rcpp glue (integration/*.cpp) exercised through the STOCK .Call wrappers of the reference
(/root/reference/src/RcppExports.cpp:109-243): oracle/_ref/libjlp_glue.so holds wrappers + glue, linked with
libjlp_b200.so (oracle/Makefile).  Each test calls _jackalope_illumina_{ref,hap}_cpp with one "SEXP" per argument,
as R's .Call does, on a RefGenome / HapSet built with the package's own classes, and byte-compares the files with
the oracle run on the seed the glue drew from the (stub) R RNG."""
import os

import numpy as np
import pytest

import jackalope_b200 as J
from common import first_diff, oracle_run
from oracle import harness as H

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.have_glue(), reason="oracle/_ref/libjlp_glue.so not built")]


def profiles(L, paired=True):
    p1 = J.flatten_profile(J.read_profile(None, "HS25", L, 1))
    p2 = J.flatten_profile(J.read_profile(None, "HS25", L, 2)) if paired else None
    return p1, p2


def read(path):
    with open(path, "rb") as f:
        return f.read()


def test_illumina_ref_cpp_through_stock_wrapper(tmp_path):
    g = J.random_genome(3, 40_000, seed=31)
    gg = H.GlueGenome(g.names, [g.chrom(c) for c in range(3)])
    p1, p2 = profiles(100)
    pre = str(tmp_path / "glue")
    seed = H.glue_illumina(gg, paired=True, matepair=False, out_prefix=pre, n_reads=6000, prof1=p1, prof2=p2, r_seed=77,
                           barcodes=["ACGT"], n_threads=2)
    o = oracle_run(g, 6000, 100, True, seed, seq_sys="HS25", barcodes=["ACGT"])
    assert first_diff(read(pre + "_R1.fq"), o["r1"]) is None and first_diff(read(pre + "_R2.fq"), o["r2"]) is None
    assert read(pre + "_R1.fq").count(b"\n") == 4 * 3000
    # another R seed, another run; the same R seed, the same bytes (set.seed governs the output)
    seed2 = H.glue_illumina(gg, paired=True, matepair=False, out_prefix=str(tmp_path / "g2"), n_reads=6000, prof1=p1, prof2=p2,
                            r_seed=78, barcodes=["ACGT"])
    assert seed2 != seed and read(str(tmp_path / "g2_R1.fq")) != o["r1"]
    H.glue_illumina(gg, paired=True, matepair=False, out_prefix=str(tmp_path / "g3"), n_reads=6000, prof1=p1, prof2=p2, r_seed=77,
                    barcodes=["ACGT"])
    assert read(str(tmp_path / "g3_R2.fq")) == o["r2"]


def test_illumina_ref_cpp_single_end_and_matepair(tmp_path):
    g = J.random_genome(2, 30_000, seed=32)
    gg = H.GlueGenome(g.names, [g.chrom(c) for c in range(2)])
    p1, p2 = profiles(100)
    pre = str(tmp_path / "se")
    seed = H.glue_illumina(gg, paired=False, matepair=False, out_prefix=pre, n_reads=3001, prof1=p1, prof2=None, r_seed=5)
    o = oracle_run(g, 3001, 100, False, seed, seq_sys="HS25")
    assert first_diff(read(pre + "_R1.fq"), o["r1"]) is None and not os.path.exists(pre + "_R2.fq")
    pre = str(tmp_path / "mp")
    seed = H.glue_illumina(gg, paired=True, matepair=True, out_prefix=pre, n_reads=4000, prof1=p1, prof2=p2, r_seed=6,
                           shape=36.0, scale=3000.0 / 36.0, prob_dup=0.3, read_pool_size=50)
    o = oracle_run(g, 4000, 100, True, seed, seq_sys="HS25", matepair=True, frag_mean=3000, frag_sd=500, prob_dup=0.3,
                   read_pool_size=50)
    assert first_diff(read(pre + "_R1.fq"), o["r1"]) is None and first_diff(read(pre + "_R2.fq"), o["r2"]) is None


def test_illumina_hap_cpp_through_stock_wrapper(tmp_path):
    g = J.random_genome(3, 30_000, seed=33)
    haps, edits = J.random_haplotypes(g, 3, sub_rate=0.01, indel_rate=0.003, seed=34, return_edits=True)
    gg = H.GlueGenome(g.names, [g.chrom(c) for c in range(3)])
    hs = H.GlueHapSet(gg, haps.hap_names, edits)
    for h in range(3):
        for c in range(3):
            assert hs.lib.jglue_hap_chrom_size(hs.h, h, c) == haps.muts[h][c].chrom_size
    p1, p2 = profiles(150)
    probs = [1.0, 3.0, 0.5]
    # pooled
    pre = str(tmp_path / "pool")
    seed = H.glue_illumina(hs, paired=True, matepair=False, out_prefix=pre, n_reads=5000, prof1=p1, prof2=p2, r_seed=9,
                           hap_probs=probs, barcodes=["AC", "GT", "TTA"])
    o = oracle_run(haps, 5000, 150, True, seed, seq_sys="HS25", haplotype_probs=probs, barcodes=["AC", "GT", "TTA"])
    assert first_diff(read(pre + "_R1.fq"), o["r1"]) is None and first_diff(read(pre + "_R2.fq"), o["r2"]) is None
    # one file pair per haplotype, compressed by the device (compress = 6 -> BGZF members; inflates to the oracle's bytes)
    import gzip
    pre = str(tmp_path / "sep")
    seed = H.glue_illumina(hs, paired=True, matepair=False, out_prefix=pre, n_reads=5000, prof1=p1, prof2=p2, r_seed=10,
                           hap_probs=probs, sep_files=True, compress=6, n_threads=2)
    o = oracle_run(haps, 5000, 150, True, seed, seq_sys="HS25", haplotype_probs=probs, sep_files=True)
    got1 = b"".join(gzip.decompress(read("%s_%s_R1.fq.gz" % (pre, h))) for h in haps.hap_names)
    got2 = b"".join(gzip.decompress(read("%s_%s_R2.fq.gz" % (pre, h))) for h in haps.hap_names)
    assert first_diff(got1, o["r1"]) is None and first_diff(got2, o["r2"]) is None


def test_errors_become_r_errors(tmp_path):
    """Rcpp::stop in the glue -> END_RCPP of the stock wrapper -> an R error with the reference's text."""
    g = J.random_genome(1, 5_000, seed=35)
    gg = H.GlueGenome(g.names, [g.chrom(0)])
    p1, p2 = profiles(100)
    with pytest.raises(RuntimeError, match="barcode"):
        H.glue_illumina(gg, paired=True, matepair=False, out_prefix=str(tmp_path / "e"), n_reads=100, prof1=p1, prof2=p2, r_seed=1,
                        barcodes=["ACGT" * 25])
    with pytest.raises(RuntimeError, match="Unable to open file"):                          # src/io.h:288-290
        H.glue_illumina(gg, paired=False, matepair=False, out_prefix=str(tmp_path / "no" / "such" / "dir" / "x"), n_reads=100,
                        prof1=p1, prof2=None, r_seed=1)


def test_pacbio_ref_cpp_through_stock_wrapper(tmp_path):
    from oracle.harness_pacbio import DEFAULTS as D
    g = J.random_genome(2, 60_000, seed=36)
    gg = H.GlueGenome(g.names, [g.chrom(c) for c in range(2)])
    lib = H.glue_lib()
    arr = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    cn, cs, sq, nm = arr(D["chi2_params_n"]), arr(D["chi2_params_s"]), arr(D["sqrt_params"]), arr(D["norm_params"])
    ln = D["lognorm_read_length"]
    import ctypes as C
    err = C.create_string_buffer(512)
    pre = str(tmp_path / "pb")
    lib.jglue_set_r_seed(3)
    seed = lib.jglue_seed_after(3)
    rc = lib.jglue_pacbio_ref(gg.h, pre.encode(), 300, 1, 100, 0.0, ln[2], ln[0], ln[1], 50.0, 40, H._ptr(cn, H.f64p),
                              H._ptr(cs, H.f64p), H._ptr(sq, H.f64p), H._ptr(nm, H.f64p), 0.2, 0.11, 0.04, 0.01, err, 512)
    assert rc == 0, err.value
    got = read(pre + "_R1.fq")
    want, _ = J.pacbio(g, "", 300, seed=seed, sink="memory")
    assert got == want and got.count(b"\n") == 4 * 300
