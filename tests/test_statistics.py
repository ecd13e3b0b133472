"""Statistical equivalence with the UNMODIFIED reference run end to end on its own pcg64
streams (oracle/_ref/libjlp_ref.so).  The reference's tests pin none of this
(SURVEY.md section 4), so the thresholds are ours: every test below is a two-sample test
at alpha = 0.001, Bonferroni-corrected over the cells it looks at.

What is compared: per-position quality histograms per end, per-position mismatch counts,
the substitution matrix, fragment lengths (inverse-CDF table vs libstdc++'s
gamma_distribution), fragment starts, strand balance, duplicate rate, indel rates, and
reads per chromosome.  The CUDA path is byte-identical to the oracle
(tests/test_gpu_parity.py), so equivalence of the oracle carries over to it; the GPU
suite repeats the analytic checks at a larger size (tests/test_gpu_statistics.py).
"""
import os
import tempfile

import numpy as np
import pytest
from scipy import stats

import jackalope_b200 as J
from oracle import harness as H
from oracle.compare import fastq_records, oracle_run

ALPHA = 1e-3
needs_ref = pytest.mark.skipif(not H.have_ref(False), reason="oracle/_ref not built")
COMP = bytes.maketrans(b"TCAG", b"AGTC")


def summarize(r1, r2, genome, L):
    """Per-record facts parsed from a pair of FASTQ streams."""
    seqs = {n.encode(): genome.chrom(i) for i, n in enumerate(genome.names)}
    out = dict(qual=[np.zeros((L, 64), dtype=np.int64) for _ in range(2)], mism=[np.zeros(L, dtype=np.int64) for _ in range(2)],
               bases=[0, 0], sub=np.zeros((4, 4), dtype=np.int64), frag_len=[], frag_start=[], first_rev=0, n=0,
               chrom={}, indel_reads=[0, 0], keys=[])
    a, b = fastq_records(r1), fastq_records(r2)
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        ida, idb = ra[0].split(b"-"), rb[0].split(b"-")
        chrom = ida[1]
        sa, sb = int(ida[2]), int(idb[2])
        rev_a = ida[3][:1] == b"R"
        assert (idb[3][:1] == b"R") != rev_a
        out["n"] += 1
        out["first_rev"] += rev_a
        out["chrom"][chrom] = out["chrom"].get(chrom, 0) + 1
        fwd_start, rev_start = (sb, sa) if rev_a else (sa, sb)
        out["keys"].append((chrom, fwd_start, rev_start))
        clean = True
        for e, (rec, start, rev) in enumerate(((ra, sa, rev_a), (rb, sb, not rev_a))):
            read, qual = rec[1], rec[3]
            q = np.frombuffer(qual, np.uint8) - 33
            out["qual"][e][np.arange(len(q)), q] += 1
            tmpl = seqs[chrom][start:start + len(read)]
            if rev:
                tmpl = tmpl.translate(COMP)[::-1]
            x, y = np.frombuffer(read, np.uint8), np.frombuffer(tmpl, np.uint8)
            if len(x) != len(y):
                out["indel_reads"][e] += 1
                clean = False
                continue
            d = x != y
            if d.sum() > 4:                      # an indel shifted the read against its template (true
                                                 # substitutions average ~0.5 per read); a looser cut lets
                                                 # shifted reads in, whose clustered mismatches break the
                                                 # independence the chi-square tests assume
                out["indel_reads"][e] += 1
                clean = False
                continue
            out["mism"][e][np.nonzero(d)[0]] += 1
            out["bases"][e] += len(x)
            for i in np.nonzero(d)[0]:
                out["sub"][b"TCAG".find(bytes([y[i]])), b"TCAG".find(bytes([x[i]]))] += 1
        if clean:
            out["frag_len"].append(rev_start + L - fwd_start)
            out["frag_start"].append(fwd_start)
    return out


def chi2_two_sample(x, y):
    """p-value of H0: the two count vectors come from the same distribution."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    keep = (x + y) >= 10
    if keep.sum() < 2:
        return 1.0
    x, y = x[keep], y[keep]
    return stats.chi2_contingency(np.vstack([x, y]))[1]


@pytest.fixture(scope="module")
def runs():
    if not H.have_ref(False):
        pytest.skip("oracle/_ref not built")
    L, n_pairs = 100, 30000
    g = J.random_genome(4, [40000, 60000, 100000, 20000], seed=31)
    kw = dict(prob_dup=0.05, ins_prob1=0.001, del_prob1=0.002, ins_prob2=0.002, del_prob2=0.001)
    o = oracle_run(g, 2 * n_pairs, L, True, seed=77, **kw)
    prof1, prof2 = o["profiles"]
    ref = H.RefGenomeH(g.names, [g.chrom(c) for c in range(g.n_chroms())])
    with tempfile.TemporaryDirectory() as d:
        pre = os.path.join(d, "r")
        H.ref_illumina_ref(ref, paired=True, matepair=False, out_prefix=pre, n_reads=2 * n_pairs, prob_dup=0.05,
                           n_threads=1, read_pool_size=1000, shape=16.0, scale=25.0, frag_len_min=L,
                           frag_len_max=2 ** 32 - 1, prof1=prof1, prof2=prof2, ins_prob=[0.001, 0.002],
                           del_prob=[0.002, 0.001], r_seed=5)
        f1, f2 = open(pre + "_R1.fq", "rb").read(), open(pre + "_R2.fq", "rb").read()
    return summarize(o["r1"], o["r2"], g, L), summarize(f1, f2, g, L), L, n_pairs, g


def test_quality_histograms_per_position(runs):
    a, b, L, _, _ = runs
    for e in range(2):
        ps = [chi2_two_sample(a["qual"][e][pos], b["qual"][e][pos]) for pos in range(L)]
        assert min(ps) > ALPHA / (2 * L), (e, int(np.argmin(ps)), min(ps))


def test_mismatch_counts_and_substitution_matrix(runs):
    a, b, L, _, _ = runs
    for e in range(2):
        # pooled over blocks of 10 positions so that cells have counts
        xa, xb = a["mism"][e].reshape(-1, 10).sum(1), b["mism"][e].reshape(-1, 10).sum(1)
        na, nb = a["bases"][e] / 10, b["bases"][e] / 10
        for i in range(xa.size):
            tab = [[xa[i], na / L * 10 - xa[i]], [xb[i], nb / L * 10 - xb[i]]]
            assert stats.chi2_contingency(tab)[1] > ALPHA / (2 * xa.size), (e, i, tab)
    assert a["sub"].sum() > 1000 and np.all(np.diag(a["sub"]) == 0) and np.all(np.diag(b["sub"]) == 0)
    off = ~np.eye(4, dtype=bool)
    assert chi2_two_sample(a["sub"][off], b["sub"][off]) > ALPHA


def test_fragment_lengths_and_starts(runs):
    a, b, L, _, g = runs
    fa, fb = np.array(a["frag_len"]), np.array(b["frag_len"])
    assert stats.ks_2samp(fa, fb).pvalue > ALPHA
    assert abs(fa.mean() - fb.mean()) < 3.0 and fa.min() >= L
    # against the analytic law too: floor(Gamma(16, 25)) clamped below at L
    pk = stats.kstest(fa + np.random.default_rng(1).random(fa.size), lambda x: np.where(
        x < L, 0.0, stats.gamma.cdf(x, a=16.0, scale=25.0))).pvalue
    assert pk > ALPHA
    sa, sb = np.array(a["frag_start"]), np.array(b["frag_start"])
    assert stats.ks_2samp(sa, sb).pvalue > ALPHA


def test_strand_duplicates_indels_and_chromosome_shares(runs):
    a, b, L, n_pairs, g = runs
    for s in (a, b):
        assert s["n"] == n_pairs
        assert stats.binomtest(int(s["first_rev"]), s["n"], 0.5).pvalue > ALPHA / 2
    # duplicates: consecutive pairs re-reading one fragment share both coordinates
    def dups(s):
        k = s["keys"]
        return sum(1 for i in range(1, len(k)) if k[i] == k[i - 1])
    da, db = dups(a), dups(b)
    assert stats.chi2_contingency([[da, n_pairs - da], [db, n_pairs - db]])[1] > ALPHA
    assert 0.02 * n_pairs < da < 0.05 * n_pairs     # prob_dup = 0.05, less the re-reads whose indels moved a coordinate
    for e in range(2):
        ia, ib = a["indel_reads"][e], b["indel_reads"][e]
        assert ia > 100 and stats.chi2_contingency([[ia, n_pairs - ia], [ib, n_pairs - ib]])[1] > ALPHA
    names = [n.encode() for n in g.names]
    ca, cb = [a["chrom"].get(n, 0) for n in names], [b["chrom"].get(n, 0) for n in names]
    assert chi2_two_sample(ca, cb) > ALPHA
    assert stats.chisquare(ca, np.asarray(g.sizes(), dtype=np.float64) / g.sizes().sum() * n_pairs).pvalue > ALPHA
