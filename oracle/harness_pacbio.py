"""TEST INFRASTRUCTURE ONLY.  ctypes access to the PacBio parts of the oracle (oracle/jlp_oracle.c,
orc_pacbio_generate) and of the unmodified reference (oracle/ref_driver_pacbio.cpp)."""
import ctypes as C

import numpy as np

from oracle import harness as H

u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)

# the defaults of pacbio(), /root/reference/R/hts_pacbio.R
DEFAULTS = dict(chi2_params_s=(0.01214, -5.12, 675, 48303.0732881, 1.4691051212330266),
                chi2_params_n=(0.00189237136, 2.53944970, 5500), max_passes=40, sqrt_params=(0.5, 0.2247),
                norm_params=(0.0, 0.2), prob_thresh=0.2, ins_prob=0.11, del_prob=0.04, sub_prob=0.01,
                min_read_length=50, lognorm_read_length=(0.200110276521, -10075.4363813, 17922.611306))


class OrcPbJob(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("job_lo", C.c_uint64), ("job_hi", C.c_uint64),
        ("sqrt_params", C.c_double * 2), ("norm_params", C.c_double * 2),
        ("prob_thresh", C.c_double), ("prob_ins", C.c_double), ("prob_del", C.c_double), ("prob_subst", C.c_double),
        ("n_groups", C.c_uint64), ("group_off", u64p), ("group_seq", C.POINTER(C.c_char_p)), ("group_len", u64p),
        ("group_genome_name", C.POINTER(C.c_char_p)), ("group_chrom_name", C.POINTER(C.c_char_p)),
        ("read_len", u64p), ("split_pos", u64p), ("passes_left", f64p), ("passes_right", f64p),
        ("beyond_template_is_n", C.c_int32), ("prob_dup", C.c_double), ("pool_reads", C.c_uint64),
    ]


def _orc():
    lib = H.oracle()
    lib.orc_pb_min_exp.restype = C.c_double
    lib.orc_pb_min_exp.argtypes = [f64p, f64p, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.orc_pacbio_generate.argtypes = [C.POINTER(OrcPbJob), C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64, u64p,
                                        u64p, u64p, C.c_uint64, u64p, u64p]
    return lib


def _arr(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def min_exp(sqrt_params, norm_params, prob_thresh, ins, dele, sub, ref=False):
    sp, npar = _arr(sqrt_params, np.float64), _arr(norm_params, np.float64)
    if ref:
        lib = H.ref_lib(True)
        lib.jrefpb_min_exp.restype = C.c_double
        lib.jrefpb_min_exp.argtypes = [f64p, f64p, C.c_double, C.c_double, C.c_double, C.c_double]
        return lib.jrefpb_min_exp(sp.ctypes.data_as(f64p), npar.ctypes.data_as(f64p), prob_thresh, ins, dele, sub)
    return _orc().orc_pb_min_exp(sp.ctypes.data_as(f64p), npar.ctypes.data_as(f64p), prob_thresh, ins, dele, sub)


def split_passes(passes, read_length):
    """(split_pos, passes_left, passes_right) from the capped number of passes, as the tail of
    PacBioPassSampler::sample computes them (src/hts_pacbio.h:184-199)."""
    import math
    frac, wholes = math.modf(passes)
    prop_left = frac if int(wholes) & 1 == 0 else 1 - frac
    split = int(math.floor(read_length * prop_left + 0.5))          # std::round of a non-negative value
    if int(wholes) & 1 == 0:
        return split, float(math.ceil(passes)), float(math.floor(passes))
    return split, float(math.floor(passes)), float(math.ceil(passes))


def generate(names, seqs, genome_name, counts, read_len, split_pos, passes_left, passes_right, seed, want_ledger=True,
             beyond_template_is_n=False, prob_dup=0.0, pool_reads=100, **model):
    """Oracle output for sum(counts) reads, counts[c] of them from chromosome c in order."""
    m = dict(DEFAULTS)
    m.update(model)
    lib = _orc()
    n = int(np.sum(counts))
    off = _arr(np.concatenate(([0], np.cumsum(counts))), np.uint64)
    J = OrcPbJob()
    J.seed, J.job_lo, J.job_hi = seed, 0, n
    J.sqrt_params = (C.c_double * 2)(*m["sqrt_params"])
    J.norm_params = (C.c_double * 2)(*m["norm_params"])
    J.prob_thresh, J.prob_ins, J.prob_del, J.prob_subst = m["prob_thresh"], m["ins_prob"], m["del_prob"], m["sub_prob"]
    J.n_groups = len(seqs)
    J.group_off = off.ctypes.data_as(u64p)
    keep = [H._strs([bytes(s) for s in seqs]), H._strs(genome_name if isinstance(genome_name, (list, tuple)) else [genome_name] * len(seqs)), H._strs(names)]
    J.group_seq, J.group_genome_name, J.group_chrom_name = keep
    lens = _arr([len(s) for s in seqs], np.uint64)
    J.group_len = lens.ctypes.data_as(u64p)
    rl, sp = _arr(read_len, np.uint64), _arr(split_pos, np.uint64)
    pl, pr = _arr(passes_left, np.float64), _arr(passes_right, np.float64)
    J.read_len, J.split_pos = rl.ctypes.data_as(u64p), sp.ctypes.data_as(u64p)
    J.passes_left, J.passes_right = pl.ctypes.data_as(f64p), pr.ctypes.data_as(f64p)
    J.beyond_template_is_n = int(beyond_template_is_n)
    J.prob_dup, J.pool_reads = float(prob_dup), int(pool_reads)
    cap = int(np.sum(np.minimum(rl, lens.max())) * 2 + 200 * n + 64)
    out = C.create_string_buffer(cap)
    ln, led_n = C.c_uint64(), C.c_uint64()
    led_cap = int(np.sum(rl) * 3 + 64 * n + 64) if want_ledger else 0
    led = np.zeros(max(1, led_cap), dtype=np.uint64)
    led_cnt = np.zeros(max(1, n), dtype=np.uint64)
    plan = np.zeros(4 * max(1, n), dtype=np.uint64)
    rc = lib.orc_pacbio_generate(C.byref(J), 0, n, out, cap, C.byref(ln), plan.ctypes.data_as(u64p),
                                 led.ctypes.data_as(u64p) if want_ledger else None, led_cap, C.byref(led_n),
                                 led_cnt.ctypes.data_as(u64p))
    assert rc == 0 and ln.value <= cap and led_n.value <= max(led_cap, 0) + (0 if want_ledger else 1 << 62), (rc, ln.value, cap)
    return dict(fastq=out.raw[:ln.value], ledger=led[:led_n.value], ledger_cnt=led_cnt[:n], plan=plan[:4 * n].reshape(n, 4))


def ref_replay(names, seqs, chrom_ind, read_len, split_pos, passes_left, passes_right, script, is_dup=None, **model):
    """The unmodified reference on the same reads, its pcg64 reading `script`."""
    m = dict(DEFAULTS)
    m.update(model)
    lib = H.ref_lib(True)
    lib.jrefpb_replay.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p, f64p, f64p, C.POINTER(C.c_int32), f64p, f64p, C.c_double, C.c_double,
                                  C.c_double, C.c_double, u64p, C.c_uint64, u64p, C.c_char_p, C.c_uint64, u64p, C.c_char_p,
                                  C.c_uint64]
    g = H.RefGenomeH(names, [bytes(s) for s in seqs], replay=True)
    n = len(read_len)
    ci, rl, sp = _arr(chrom_ind, np.uint64), _arr(read_len, np.uint64), _arr(split_pos, np.uint64)
    pl, pr = _arr(passes_left, np.float64), _arr(passes_right, np.float64)
    sq, nm = _arr(m["sqrt_params"], np.float64), _arr(m["norm_params"], np.float64)
    sc = _arr(script, np.uint64)
    dup = _arr(is_dup, np.int32) if is_dup is not None else None
    consumed = np.zeros(max(1, n), dtype=np.uint64)
    cap = int(np.sum(rl) * 2 + 200 * n + 64)
    out = C.create_string_buffer(cap)
    ln = C.c_uint64()
    err = C.create_string_buffer(256)
    rc = lib.jrefpb_replay(g.h, n, ci.ctypes.data_as(u64p), rl.ctypes.data_as(u64p), sp.ctypes.data_as(u64p),
                           pl.ctypes.data_as(f64p), pr.ctypes.data_as(f64p),
                           dup.ctypes.data_as(C.POINTER(C.c_int32)) if dup is not None else None, sq.ctypes.data_as(f64p), nm.ctypes.data_as(f64p),
                           m["prob_thresh"], m["ins_prob"], m["del_prob"], m["sub_prob"], sc.ctypes.data_as(u64p), len(sc),
                           consumed.ctypes.data_as(u64p), out, cap, C.byref(ln), err, 256)
    assert rc == 0, err.value
    return out.raw[:ln.value], consumed[:n]
