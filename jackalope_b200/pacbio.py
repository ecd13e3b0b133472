"""pacbio(): the PacBio read generator behind the reference's argument surface
(/root/reference/R/hts_pacbio.R: pacbio(), check_pacbio_args()); SURVEY.md section 8f rank 3."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .genome import Haplotypes, RefGenome
from .illumina import _single_integer, _single_number, default_context
from .profiles import JackalopeError


def _err(par, what):
    raise JackalopeError("\nFor the PacBio sequencer, argument `%s` must be %s." % (par, what))


def check_pacbio_args(obj, n_reads, haplotype_probs, sep_files, compress, comp_method, n_threads, read_pool_size,
                      chi2_params_s, chi2_params_n, max_passes, sqrt_params, norm_params, prob_thresh, ins_prob,
                      del_prob, sub_prob, min_read_length, lognorm_read_length, custom_read_lengths, prob_dup,
                      show_progress):
    """R/hts_pacbio.R:8-110."""
    if not isinstance(obj, (RefGenome, Haplotypes)):
        raise JackalopeError("\nWhen providing info for the PacBio sequencer, the object providing the sequence information "
                             "should be of class \"ref_genome\" or \"haplotypes\".")
    for name, v in (("n_reads", n_reads), ("n_threads", n_threads), ("read_pool_size", read_pool_size),
                    ("max_passes", max_passes), ("min_read_length", min_read_length)):
        if not _single_integer(v, 1):
            _err(name, "a single integer >= 1")
    for name, v in (("prob_thresh", prob_thresh), ("ins_prob", ins_prob), ("del_prob", del_prob), ("sub_prob", sub_prob),
                    ("prob_dup", prob_dup)):
        if not _single_number(v, 0, 1):
            _err(name, "a single number in range [0,1].")
    if ins_prob + del_prob + sub_prob > 1:
        raise JackalopeError("\nWhen providing info for the PacBio sequencer, the insertion, deletion, and substitution "
                             "probabilities cannot sum to > 1.")
    for name, v, k in (("chi2_params_s", chi2_params_s, 5), ("chi2_params_n", chi2_params_n, 3),
                       ("lognorm_read_length", lognorm_read_length, 3), ("sqrt_params", sqrt_params, 2),
                       ("norm_params", norm_params, 2)):
        try:
            ok = len(v) == k and all(isinstance(float(x), float) for x in v)
        except Exception:
            ok = False
        if not ok:
            _err(name, "a numeric vector of length %d." % k)
    if custom_read_lengths is not None:
        a = np.asarray(custom_read_lengths, dtype=np.float64)
        if a.ndim == 2:
            if a.shape[1] != 2:
                raise JackalopeError("\nWhen providing info for the PacBio sequencer, if the `custom_read_lengths` argument is "
                                     "a matrix, it should have exactly 2 columns.")
            if (a[:, 1] < 0).any() or (a[:, 1] == 0).all():
                raise JackalopeError("\nWhen providing info for the PacBio sequencer, if the `custom_read_lengths` argument is "
                                     "a matrix, it should have exactly 2 columns, and the second column should contain no "
                                     "values < 0 and at least one value > 0.")
        elif a.ndim != 1:
            _err("custom_read_lengths", "NULL, a matrix, or a numeric vector.")
    if haplotype_probs is not None:
        hp = np.asarray(haplotype_probs, dtype=np.float64)
        if not isinstance(obj, Haplotypes) or hp.ndim != 1 or len(hp) != obj.n_haps() or (hp < 0).any() or not (hp > 0).any():
            _err("haplotype_probs", "NULL or a numeric vector of the same length as the number of haplotypes, with no "
                                    "values < 0 and at least one value > 0")
    if not isinstance(sep_files, bool):
        _err("sep_files", "a single logical.")
    if not isinstance(compress, bool) and not _single_integer(compress, 1, 9):
        _err("compress", "a single logical or integer from 1 to 9")
    if comp_method not in ("gzip", "bgzip"):
        _err("comp_method", "\"gzip\" or \"bgzip\"")
    if not isinstance(show_progress, bool):
        _err("show_progress", "a single logical.")


def _params(obj, out_prefix, n_reads, chi2_params_s, chi2_params_n, max_passes, sqrt_params, norm_params, prob_thresh,
            ins_prob, del_prob, sub_prob, min_read_length, lognorm_read_length, custom_read_lengths, prob_dup,
            haplotype_probs, sep_files, compress, comp_method, n_threads, read_pool_size, seed, batch_reads, comp_engine,
            shard=None):
    keep = []
    p = _lib.PacbioParams()
    p.out_prefix = (out_prefix or "").encode()
    p.sep_files, p.compress, p.comp_method = int(sep_files), int(compress), comp_method.encode()
    p.n_reads, p.n_threads, p.read_pool_size = int(n_reads), int(n_threads), int(read_pool_size)
    if isinstance(obj, Haplotypes):
        hp = np.ascontiguousarray(haplotype_probs if haplotype_probs is not None else [1.0] * obj.n_haps(), dtype=np.float64)
        keep.append(hp)
        p.haplotype_probs = hp.ctypes.data_as(_lib.f64p)
    p.prob_dup = float(prob_dup)
    p.sigma, p.loc, p.scale = (float(x) for x in lognorm_read_length)
    p.min_read_len = float(min_read_length)
    if custom_read_lengths is not None:
        a = np.asarray(custom_read_lengths, dtype=np.float64)
        lens = np.ascontiguousarray(a[:, 0] if a.ndim == 2 else a, dtype=np.uint64)
        probs = np.ascontiguousarray(a[:, 1] if a.ndim == 2 else np.ones(len(a)), dtype=np.float64)
        keep += [lens, probs]
        p.read_lens, p.read_probs, p.n_custom = lens.ctypes.data_as(_lib.u64p), probs.ctypes.data_as(_lib.f64p), len(lens)
    p.max_passes = int(max_passes)
    p.chi2_params_n = (C.c_double * 3)(*map(float, chi2_params_n))
    p.chi2_params_s = (C.c_double * 5)(*map(float, chi2_params_s))
    p.sqrt_params = (C.c_double * 2)(*map(float, sqrt_params))
    p.norm_params = (C.c_double * 2)(*map(float, norm_params))
    p.prob_thresh, p.prob_ins, p.prob_del, p.prob_subst = float(prob_thresh), float(ins_prob), float(del_prob), float(sub_prob)
    p.seed = int(seed) & (2 ** 64 - 1)
    p.batch_reads = int(batch_reads or 0)
    p.comp_engine = {"auto": 0, "host": 1, "device": 2}[comp_engine]
    p.shard_index, p.shard_count = (int(shard[0]), int(shard[1])) if shard else (0, 1)
    return p, keep


def pacbio(obj, out_prefix, n_reads,
           chi2_params_s=(0.01214, -5.12, 675, 48303.0732881, 1.4691051212330266),
           chi2_params_n=(0.00189237136, 2.53944970, 5500), max_passes=40, sqrt_params=(0.5, 0.2247),
           norm_params=(0, 0.2), prob_thresh=0.2, ins_prob=0.11, del_prob=0.04, sub_prob=0.01, min_read_length=50,
           lognorm_read_length=(0.200110276521, -10075.4363813, 17922.611306), custom_read_lengths=None, prob_dup=0.0,
           haplotype_probs=None, sep_files=False, compress=False, comp_method="bgzip", n_threads=1, read_pool_size=100,
           show_progress=False, overwrite=False, *, seed=None, device=0, ctx=None, batch_reads=None, sink="files",
           comp_engine="auto", want_plan=False, shard=None):
    """Create and write PacBio reads to FASTQ file(s); positional and keyword arguments up to ``overwrite`` are the
    reference's.  ``shard=(index, count)``: this call generates its share of every job (whole pools of
    ``read_pool_size`` reads; the shards' outputs concatenate to the unsharded run).  ``sink``: "files" (returns None), "memory" (returns (fastq_bytes, stats)), "device" (generate and
    drop on the GPU; returns stats).  ``want_plan=True`` (with sink="memory") also returns what was drawn per read
    before the per-base work: dict(group, read_len, split_pos, passes_left, passes_right)."""
    check_pacbio_args(obj, n_reads, haplotype_probs, sep_files, compress, comp_method, n_threads, read_pool_size,
                      chi2_params_s, chi2_params_n, max_passes, sqrt_params, norm_params, prob_thresh, ins_prob, del_prob,
                      sub_prob, min_read_length, lognorm_read_length, custom_read_lengths, prob_dup, show_progress)
    if comp_engine not in ("auto", "host", "device"):
        raise JackalopeError("comp_engine must be \"auto\", \"host\" or \"device\"")
    is_haps = isinstance(obj, Haplotypes)
    if not is_haps:
        sep_files = False
    out_prefix = os.path.expanduser(out_prefix) if out_prefix else ""
    if isinstance(compress, bool):
        compress = 6 if compress else 0
    if sink == "files":
        fns = (["%s_R1.fq" % out_prefix] if not sep_files else ["%s_%s_R1.fq" % (out_prefix, h) for h in obj.hap_names])
        full = [f + ".gz" for f in fns] if compress else fns
        for d in {os.path.dirname(f) for f in full}:
            if d and not os.path.isdir(d):
                os.makedirs(d, exist_ok=True)
        if not overwrite:
            for f in full:
                if os.path.exists(f):
                    raise JackalopeError("\nFile %s already exists." % f)
    if n_threads > 1 and compress > 0 and comp_method == "gzip":
        raise JackalopeError("\nCompression using gzip cannot be performed using multiple threads. "
                             "Please use bgzip compression instead.")
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    p, keep = _params(obj, out_prefix, n_reads, chi2_params_s, chi2_params_n, max_passes, sqrt_params, norm_params,
                      prob_thresh, ins_prob, del_prob, sub_prob, min_read_length, lognorm_read_length, custom_read_lengths,
                      prob_dup, haplotype_probs, sep_files, compress, comp_method, n_threads, read_pool_size, seed,
                      batch_reads, comp_engine, shard)
    ctx = ctx or default_context(device)
    if is_haps:
        ctx.set_haplotypes(obj)
    else:
        ctx.set_genome(obj, wait=True)
    lib = ctx.lib
    stats = _lib.RunStats()
    if sink == "files":
        ctx._check(lib.jlp_pacbio(ctx.h, int(is_haps), C.byref(p), C.byref(stats)), "pacbio")
        return None
    if sink == "device":
        n = C.c_uint64()
        ctx._check(lib.jlp_pacbio_to_memory(ctx.h, int(is_haps), C.byref(p), None, 0, C.byref(n), C.byref(stats)), "pacbio")
        return stats.as_dict()
    if sink != "memory":
        raise JackalopeError("sink must be \"files\", \"memory\" or \"device\"")
    n = int(n_reads)
    plan = dict(group=np.zeros(n, np.uint64), read_len=np.zeros(n, np.uint64), split_pos=np.zeros(n, np.uint64),
                passes_left=np.zeros(n, np.float64), passes_right=np.zeros(n, np.float64))
    ctx._check(lib.jlp_pacbio_read_plan(ctx.h, int(is_haps), C.byref(p), plan["group"].ctypes.data_as(_lib.u64p),
                                        plan["read_len"].ctypes.data_as(_lib.u64p), plan["split_pos"].ctypes.data_as(_lib.u64p),
                                        plan["passes_left"].ctypes.data_as(_lib.f64p),
                                        plan["passes_right"].ctypes.data_as(_lib.f64p)), "pacbio")
    cap = int(plan["read_len"].sum()) * 2 + n * 160 + 64
    out = C.create_string_buffer(cap)
    ln = C.c_uint64()
    ctx._check(lib.jlp_pacbio_to_memory(ctx.h, int(is_haps), C.byref(p), out, cap, C.byref(ln), C.byref(stats)), "pacbio")
    res = (out.raw[:ln.value], stats.as_dict())
    return res + (plan,) if want_plan else res


def sample_read_plan(n, chrom_len, seed=1, **kw):
    """The host samplers alone (no device): read_len, split_pos, passes_left, passes_right of n reads."""
    d = dict(chi2_params_s=(0.01214, -5.12, 675, 48303.0732881, 1.4691051212330266), chi2_params_n=(0.00189237136, 2.53944970, 5500),
             max_passes=40, sqrt_params=(0.5, 0.2247), norm_params=(0, 0.2), prob_thresh=0.2, ins_prob=0.11, del_prob=0.04,
             sub_prob=0.01, min_read_length=50, lognorm_read_length=(0.200110276521, -10075.4363813, 17922.611306),
             custom_read_lengths=None)
    d.update(kw)
    g = RefGenome(["c"], [np.frombuffer(b"A", dtype=np.uint8)])
    p, keep = _params(g, "", n, d["chi2_params_s"], d["chi2_params_n"], d["max_passes"], d["sqrt_params"], d["norm_params"],
                      d["prob_thresh"], d["ins_prob"], d["del_prob"], d["sub_prob"], d["min_read_length"], d["lognorm_read_length"],
                      d["custom_read_lengths"], 0.0, None, False, 0, "bgzip", 1, 100, seed, None, "auto")
    rl, sp = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    pl, pr = np.zeros(n, np.float64), np.zeros(n, np.float64)
    rc = _lib.lib().jlp_pacbio_sample(C.byref(p), n, int(chrom_len), rl.ctypes.data_as(_lib.u64p), sp.ctypes.data_as(_lib.u64p),
                                      pl.ctypes.data_as(_lib.f64p), pr.ctypes.data_as(_lib.f64p))
    if rc != 0:
        raise JackalopeError("jlp_pacbio_sample failed (%d)" % rc)
    return rl, sp, pl, pr
