"""The host samplers of the PacBio generator (read length, number of passes and their split: jlp_pacbio_sample, no
device needed) against the reference's own samplers run on real pcg64 streams (PacBioReadLenSampler::sample,
PacBioPassSampler::sample through oracle/_ref): two-sample Kolmogorov-Smirnov and chi-square tests.  These quantities
are the statistical tier of the PacBio parity protocol -- the reference draws them from std::lognormal_distribution
and std::chi_squared_distribution, whose draw counts depend on cached state -- everything after them is replayed
bit for bit (tests/test_pacbio_oracle.py, tests/test_gpu_pacbio.py)."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import jackalope_b200 as J
from oracle import harness as H
from oracle.harness_pacbio import DEFAULTS

needs_ref = pytest.mark.skipif(not H.have_ref(False), reason="oracle/_ref not built (no /root/reference here)")
u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
ALPHA = 1e-3


def ref_samples(n, seed, lognorm=DEFAULTS["lognorm_read_length"], min_len=50, max_passes=40):
    lib = H.ref_lib(False)
    lib.jrefpb_sample_lengths_passes.argtypes = [C.c_uint64, C.c_uint64] + [C.c_double] * 4 + [C.c_uint64, f64p, f64p, u64p, u64p, f64p, f64p]
    rl, sp = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pl, pr = np.zeros(n, np.float64), np.zeros(n, np.float64)
    cn = np.array(DEFAULTS["chi2_params_n"], np.float64)
    cs = np.array(DEFAULTS["chi2_params_s"], np.float64)
    assert lib.jrefpb_sample_lengths_passes(n, seed, lognorm[2], lognorm[0], lognorm[1], float(min_len), max_passes,
                                            cn.ctypes.data_as(f64p), cs.ctypes.data_as(f64p), rl.ctypes.data_as(u64p),
                                            sp.ctypes.data_as(u64p), pl.ctypes.data_as(f64p), pr.ctypes.data_as(f64p)) == 0
    return rl, sp, pl, pr


@needs_ref
@pytest.mark.parametrize("lognorm,min_len", [(DEFAULTS["lognorm_read_length"], 50), ((0.4, -3000.0, 5000.0), 800)])
def test_read_length_and_passes_follow_the_reference(lognorm, min_len):
    n = 60000
    a = J.sample_read_plan(n, 10 ** 9, seed=5, lognorm_read_length=lognorm, min_read_length=min_len)
    b = ref_samples(n, 11, lognorm, min_len)
    assert stats.ks_2samp(a[0].astype(float), b[0].astype(float)).pvalue > ALPHA                      # read lengths
    assert (a[0] >= min_len).all() and (b[0] >= min_len).all()
    # passes (left + right determines the capped number of passes up to its fractional part) and the split fraction
    assert stats.ks_2samp(a[2] + a[3], b[2] + b[3]).pvalue > ALPHA
    fa, fb = a[1] / a[0], b[1] / b[0]
    assert stats.ks_2samp(fa, fb).pvalue > ALPHA
    # which side got the extra pass
    ta = np.array([(a[2] > a[3]).sum(), (a[2] < a[3]).sum(), (a[2] == a[3]).sum()])
    tb = np.array([(b[2] > b[3]).sum(), (b[2] < b[3]).sum(), (b[2] == b[3]).sum()])
    keep = (ta + tb) > 10
    assert stats.chi2_contingency(np.vstack([ta[keep], tb[keep]]))[1] > ALPHA
    # conditional on the read length: short and long reads separately (the chi-square parameters depend on it)
    for lo, hi in ((0, np.median(b[0])), (np.median(b[0]), 1e18)):
        ma, mb = (a[0] >= lo) & (a[0] < hi), (b[0] >= lo) & (b[0] < hi)
        assert stats.ks_2samp((a[2] + a[3])[ma], (b[2] + b[3])[mb]).pvalue > ALPHA


def test_samplers_are_addressed_by_read_index_and_clamp_to_the_chromosome():
    a = J.sample_read_plan(2000, 10 ** 9, seed=9)
    b = J.sample_read_plan(1000, 10 ** 9, seed=9)
    assert all(np.array_equal(x[:1000], y) for x, y in zip(a, b))
    c = J.sample_read_plan(500, 3000, seed=9)
    assert (c[0] <= 3000).all() and (c[0] == 3000).any() and (c[1] <= c[0]).all()
    d = J.sample_read_plan(3000, 10 ** 9, seed=3, custom_read_lengths=[[100, 1], [2000, 3]])
    assert set(np.unique(d[0])) == {100, 2000} and 0.70 < (d[0] == 2000).mean() < 0.80
    with pytest.raises(J.JackalopeError):
        J.sample_read_plan(10, 1000, ins_prob=0.6, del_prob=0.5)


def test_check_pacbio_args_mirrors_the_reference():
    """R/hts_pacbio.R:8-110: the argument checks of pacbio()."""
    g = J.random_genome(2, 500, seed=1)
    haps = J.random_haplotypes(g, 2, seed=2)
    D = dict(n_reads=10, haplotype_probs=None, sep_files=False, compress=False, comp_method="bgzip", n_threads=1,
             read_pool_size=100, chi2_params_s=DEFAULTS["chi2_params_s"], chi2_params_n=DEFAULTS["chi2_params_n"], max_passes=40,
             sqrt_params=DEFAULTS["sqrt_params"], norm_params=DEFAULTS["norm_params"], prob_thresh=0.2, ins_prob=0.11,
             del_prob=0.04, sub_prob=0.01, min_read_length=50, lognorm_read_length=DEFAULTS["lognorm_read_length"],
             custom_read_lengths=None, prob_dup=0.0, show_progress=False)

    def call(obj=g, **kw):
        a = dict(D)
        a.update(kw)
        J.check_pacbio_args(obj, **a)

    call()
    call(haps, haplotype_probs=[1, 2], sep_files=True, custom_read_lengths=[[100, 1], [200, 0]])
    call(custom_read_lengths=[100, 200, 300])
    for bad in (dict(obj="x"), dict(n_reads=0), dict(n_threads=0), dict(read_pool_size=0), dict(max_passes=0),
                dict(min_read_length=0), dict(prob_thresh=1.5), dict(ins_prob=-0.1), dict(prob_dup=2),
                dict(ins_prob=0.5, del_prob=0.4, sub_prob=0.2), dict(chi2_params_s=(1, 2, 3)), dict(chi2_params_n=(1, 2)),
                dict(lognorm_read_length=(1, 2)), dict(sqrt_params=(1,)), dict(norm_params=(0, 0.2, 1)),
                dict(custom_read_lengths=[[1, 2, 3]]), dict(custom_read_lengths=[[100, -1], [200, 1]]),
                dict(custom_read_lengths=[[100, 0], [200, 0]]), dict(haplotype_probs=[1, 2]), dict(obj=haps, haplotype_probs=[1]),
                dict(obj=haps, haplotype_probs=[0, 0]), dict(sep_files=1), dict(compress=10), dict(comp_method="xz"),
                dict(show_progress=1)):
        with pytest.raises(J.JackalopeError):
            call(**bad)
