"""``illumina()`` -- the user-facing call, same argument surface as the
reference's R function (/root/reference/R/hts_illumina.R:593-619), restated in
Python because this image has no R.  It validates arguments as
check_illumina_args does (R/hts_illumina.R:277-396), resolves the ART profiles,
converts fragment mean/sd to Gamma shape/scale, and calls the C ABI
(include/jlp_b200.h) where the R function calls illumina_ref_cpp /
illumina_hap_cpp.  All compute happens in CUDA kernels; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import numbers
import os

import numpy as np

from . import _lib
from .genome import Haplotypes, RefGenome
from .profiles import JackalopeError, cached_flat_profile

u8p, u32p, u64p, f64p = _lib.u8p, _lib.u32p, _lib.u64p, _lib.f64p


def _err_msg(par, *what):
    raise JackalopeError("\nFor the `illumina` function in jackalope, argument `%s` must be %s." % (par, " ".join(what)))


def _single_integer(x, lo=None, hi=None):
    ok = isinstance(x, numbers.Real) and not isinstance(x, bool) and float(x) == int(x)
    return ok and (lo is None or x >= lo) and (hi is None or x <= hi)


def _single_number(x, lo=None, hi=None):
    ok = isinstance(x, numbers.Real) and not isinstance(x, bool) and not np.isnan(x)
    return ok and (lo is None or x >= lo) and (hi is None or x <= hi)


def check_illumina_args(obj, n_reads, read_length, paired, frag_mean, frag_sd, matepair, seq_sys, profile1,
                        profile2, ins_prob1, del_prob1, ins_prob2, del_prob2, frag_len_min, frag_len_max,
                        haplotype_probs, barcodes, prob_dup, sep_files, compress, comp_method, n_threads,
                        read_pool_size, show_progress):
    """R/hts_illumina.R:277-396."""
    if not isinstance(obj, (RefGenome, Haplotypes)):
        raise JackalopeError("\nWhen providing info for the Illumina sequencer, the object providing the sequence "
                             "information should be of class \"ref_genome\" or \"haplotypes\".")
    for name, v in (("read_length", read_length), ("n_reads", n_reads), ("n_threads", n_threads),
                    ("read_pool_size", read_pool_size)):
        if not _single_integer(v, 1):
            _err_msg(name, "a single integer >= 1")
    if not isinstance(compress, bool) and not _single_integer(compress, 1, 9):
        _err_msg("compress", "a single logical or integer from 1 to 9")
    if comp_method not in ("gzip", "bgzip"):
        _err_msg("comp_method", "\"gzip\" or \"bgzip\"")
    for name, v in (("paired", paired), ("matepair", matepair), ("sep_files", sep_files),
                    ("show_progress", show_progress)):
        if not isinstance(v, (bool, np.bool_)):
            _err_msg(name, "a single logical.")
    for name, v in (("frag_mean", frag_mean), ("frag_sd", frag_sd)):
        if not _single_number(v) or v <= 0:
            _err_msg(name, "a single number > 0")
    for name, v in (("ins_prob1", ins_prob1), ("del_prob1", del_prob1), ("ins_prob2", ins_prob2),
                    ("del_prob2", del_prob2), ("prob_dup", prob_dup)):
        if not _single_number(v, 0, 1):
            _err_msg(name, "a single number in range [0,1].")
    for name, v in (("seq_sys", seq_sys), ("profile1", profile1), ("profile2", profile2)):
        if v is not None and not isinstance(v, str):
            _err_msg(name, "NULL or a single string")
    for name, v in (("frag_len_min", frag_len_min), ("frag_len_max", frag_len_max)):
        if v is not None and not _single_integer(v, 1):
            _err_msg(name, "NULL or a single integer >= 1")
    if haplotype_probs is not None:
        hp = np.asarray(haplotype_probs, dtype=np.float64)
        if hp.ndim != 1 or np.any(np.isnan(hp)) or np.any(hp < 0) or np.all(hp == 0):
            _err_msg("haplotype_probs", "NULL or a numeric/integer vector", "with no values < 0 and at least one value > 0")
    if barcodes is not None and not (isinstance(barcodes, (list, tuple)) and all(isinstance(b, str) for b in barcodes)):
        _err_msg("barcodes", "NULL or a character vector")
    if paired:
        if profile1 is not None and profile2 is None:
            raise JackalopeError("\nFor Illumina paired-end reads, if you provide a custom profile for "
                                 "read 1, you must also provide a file for read 2.")
        if profile1 is None and profile2 is not None:
            raise JackalopeError("\nFor Illumina paired-end reads, if you provide a custom profile for "
                                 "read 2, you must also provide a file for read 1.")
    elif profile2 is not None:
        raise JackalopeError("\nFor Illumina single-end reads, it makes no sense to provide a custom profile for "
                             "read 2. Terminating here in case this was a mistake.")
    is_haps = isinstance(obj, Haplotypes)
    if haplotype_probs is not None and not is_haps:
        raise JackalopeError("\nFor Illumina sequencing, it makes no sense to provide a vector of probabilities of "
                             "sequencing each haplotype if the `obj` argument is of class \"ref_genome\". "
                             "Terminating here in case this was a mistake.")
    if haplotype_probs is not None and is_haps and len(haplotype_probs) != obj.n_haps():
        _err_msg("haplotype_probs", "a vector of the same length as the number of haplotypes in the",
                 "`obj` argument, if `obj` is of class \"haplotypes\".",
                 "Use `obj$n_haps()` to see the number of haplotypes")
    if barcodes is not None:
        if not is_haps and len(barcodes) != 1:
            raise JackalopeError("\nFor Illumina sequencing, it makes no sense to provide a vector of multiple "
                                 "barcodes if the `obj` argument is of class \"ref_genome\". "
                                 "Terminating here in case this was a mistake.")
        if is_haps and len(barcodes) != obj.n_haps():
            _err_msg("barcodes", "a vector of the same length as the number of haplotypes in the",
                     "`obj` argument, if `obj` is of class \"haplotypes\".",
                     "Use `obj$n_haps()` to see the number of haplotypes")
        if any(ch not in "TCAG" for b in barcodes for ch in b):
            _err_msg("barcodes", "NULL or a character vector with only the",
                     "characters \"T\", \"C\", \"A\", and \"G\" present")


class Context:
    """A library context with a genome (and haplotypes) resident in HBM: one GPU (``device``), or several GPUs of
    this host driven by one call (``devices``: a list of device indices, or "all") -- illumina() then cuts the run
    into one contiguous piece per GPU and still writes ONE ordered set of files (jlp_ctx_create_multi)."""

    def __init__(self, device=0, devices=None):
        self.lib = _lib.lib()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.jlp_ctx_create(int(device), C.byref(h))
        elif isinstance(devices, str) and devices == "all":
            rc = self.lib.jlp_ctx_create_multi(0, None, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.jlp_ctx_create_multi(len(devices), arr, C.byref(h))
        if rc != 0:
            raise RuntimeError("jlp_ctx_create: " + self.lib.jlp_last_error(None).decode())
        self.h = h
        self.n_devices = int(self.lib.jlp_ctx_n_devices(h))
        self._genome = None
        self._haps = None
        self._profiles = [None, None]

    def close(self):
        if getattr(self, "h", None):
            self.lib.jlp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.jlp_last_error(self.h).decode(errors="replace")
            if rc == _lib.JLP_ERR_ARG:
                raise JackalopeError(msg)
            raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))

    def set_genome(self, g: RefGenome, wait=True):
        """Upload the reference.  With wait=False the copy runs chromosome by chromosome underneath
        the next illumina() call on this genome, which waits for it before returning."""
        if self._genome is g:
            return
        bases, off = g.flat()
        names = (C.c_char_p * g.n_chroms())(*[n.encode() for n in g.names])
        fn = self.lib.jlp_set_genome if wait else self.lib.jlp_set_genome_async
        self._upload_keep = (bases, off)      # the library reads `bases` until the upload is done
        self._check(fn(self.h, bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(u64p),
                       g.n_chroms(), names, g.name.encode()), "jlp_set_genome")
        self._genome, self._haps = g, None

    def set_haplotypes(self, haps: Haplotypes):
        self.set_genome(haps.reference)
        if self._haps is haps:
            return
        self._check(self.lib.jlp_clear_haplotypes(self.h), "jlp_clear_haplotypes")
        nc = haps.n_chroms()
        for h, name in enumerate(haps.hap_names):
            m = haps.muts[h]
            n_muts = np.array([x.old_pos.size for x in m], dtype=np.uint64)
            sizes = np.array([x.chrom_size for x in m], dtype=np.uint64)
            pool_len = np.array([x.pool.size for x in m], dtype=np.uint64)
            op = (u64p * nc)(*[x.old_pos.ctypes.data_as(u64p) for x in m])
            npos = (u64p * nc)(*[x.new_pos.ctypes.data_as(u64p) for x in m])
            no = (u64p * nc)(*[x.nuc_off.ctypes.data_as(u64p) for x in m])
            pools = (C.c_void_p * nc)(*[x.pool.ctypes.data for x in m])
            idx = C.c_uint64()
            self._check(self.lib.jlp_add_haplotype(self.h, name.encode(), n_muts.ctypes.data_as(u64p), op, npos, no,
                                                   pools, pool_len.ctypes.data_as(u64p), sizes.ctypes.data_as(u64p),
                                                   C.byref(idx)), "jlp_add_haplotype")
        self._haps = haps

    def haplotype_chrom(self, hap, chrom) -> bytes:
        n = int(self._haps.muts[hap][chrom].chrom_size)
        buf = C.create_string_buffer(max(n, 1))
        ln = C.c_uint64()
        self._check(self.lib.jlp_get_haplotype_chrom(self.h, hap, chrom, buf, n, C.byref(ln)), "jlp_get_haplotype_chrom")
        return buf.raw[:ln.value]

    def set_profile(self, end, flat):
        L, nq, probs, quals = flat
        old = self._profiles[end]
        if old is not None and (old is flat or (old[0] == L and all(np.array_equal(a, b) for a, b in zip(old[1:], flat[1:])))):
            return
        self._check(self.lib.jlp_set_profile(self.h, end, L, nq.ctypes.data_as(u32p), probs.ctypes.data_as(f64p),
                                             quals.ctypes.data_as(u8p)), "jlp_set_profile")
        self._profiles[end] = flat


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _prepare(obj, out_prefix, n_reads, read_length, paired, frag_mean, frag_sd, matepair, seq_sys, profile1,
             profile2, ins_prob1, del_prob1, ins_prob2, del_prob2, frag_len_min, frag_len_max, haplotype_probs,
             barcodes, prob_dup, sep_files, compress, comp_method, n_threads, read_pool_size, show_progress,
             overwrite, seed, batch_pairs, shard, check_files, comp_engine="auto"):
    """Everything illumina() does before calling into C++ (R/hts_illumina.R:621-731).
    Returns (params struct, keep-alive list, flattened profiles)."""
    if matepair:
        paired = True
    check_illumina_args(obj, n_reads, read_length, paired, frag_mean, frag_sd, matepair, seq_sys, profile1, profile2,
                        ins_prob1, del_prob1, ins_prob2, del_prob2, frag_len_min, frag_len_max, haplotype_probs,
                        barcodes, prob_dup, sep_files, compress, comp_method, n_threads, read_pool_size,
                        show_progress)
    is_haps = isinstance(obj, Haplotypes)
    out_prefix = os.path.expanduser(out_prefix) if out_prefix else ""
    if not is_haps:
        sep_files = False
    n_files = 2 if paired else 1
    if not sep_files:
        fns = ["%s_R%d.fq" % (out_prefix, i + 1) for i in range(n_files)]
    else:
        fns = ["%s_%s_R%d.fq" % (out_prefix, h, i + 1) for h in obj.hap_names for i in range(n_files)]
    if isinstance(compress, bool):
        compress = 6 if compress else 0
    if check_files:   # check_file_existence, R/util.R:59-81
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
    frag_len_shape = (frag_mean / frag_sd) ** 2
    frag_len_scale = frag_sd ** 2 / frag_mean
    if frag_len_min is None:
        frag_len_min = read_length
    if frag_len_max is None or frag_len_max > 2 ** 32 - 1:
        frag_len_max = 2 ** 32 - 1
    if frag_len_min > frag_len_max:
        raise JackalopeError("\nFragment length min can't be less than the max. "
                             "For computational reasons, both should also be < 2^32, "
                             "and if `frag_len_min` is not provided, it's automatically changed "
                             "to the read length.")
    if haplotype_probs is None and is_haps:
        haplotype_probs = [1.0] * obj.n_haps()
    if barcodes is None:
        barcodes = [""] * (obj.n_haps() if is_haps else 1)
    prof1 = cached_flat_profile(profile1, seq_sys, read_length, 1)
    prof2 = cached_flat_profile(profile2, seq_sys, read_length, 2) if paired else None

    keep = []
    p = _lib.Params()
    p.paired, p.matepair = int(paired), int(matepair)
    p.out_prefix = out_prefix.encode()
    p.sep_files, p.compress, p.comp_method = int(sep_files), int(compress), comp_method.encode()
    p.n_reads, p.prob_dup, p.n_threads = int(n_reads), float(prob_dup), int(n_threads)
    p.show_progress, p.read_pool_size = int(show_progress), int(read_pool_size)
    if is_haps:
        hp = np.ascontiguousarray(haplotype_probs, dtype=np.float64)
        keep.append(hp)
        p.haplotype_probs = hp.ctypes.data_as(f64p)
    p.frag_len_shape, p.frag_len_scale = float(frag_len_shape), float(frag_len_scale)
    p.frag_len_min, p.frag_len_max = int(frag_len_min), int(frag_len_max)
    p.ins_prob1, p.del_prob1, p.ins_prob2, p.del_prob2 = map(float, (ins_prob1, del_prob1, ins_prob2, del_prob2))
    bc = (C.c_char_p * len(barcodes))(*[b.encode() for b in barcodes])
    keep.append(bc)
    p.barcodes = bc
    p.seed = int(seed) & (2 ** 64 - 1)
    p.batch_pairs = int(batch_pairs or 0)
    p.shard_index, p.shard_count = (int(shard[0]), int(shard[1])) if shard else (0, 1)
    if comp_engine not in ("auto", "host", "device"):
        raise JackalopeError("comp_engine must be \"auto\", \"host\" or \"device\"")
    p.comp_engine = {"auto": 0, "host": 1, "device": 2}[comp_engine]
    return p, keep, (prof1, prof2), is_haps, fns


def illumina(obj, out_prefix, n_reads, read_length, paired, frag_mean=400, frag_sd=100, matepair=False,
             seq_sys=None, profile1=None, profile2=None, ins_prob1=0.00009, del_prob1=0.00011,
             ins_prob2=0.00015, del_prob2=0.00023, frag_len_min=None, frag_len_max=None, haplotype_probs=None,
             barcodes=None, prob_dup=0.02, sep_files=False, compress=False, comp_method="bgzip", n_threads=1,
             read_pool_size=1000, show_progress=False, overwrite=False, *, seed=None, device=0, ctx=None,
             batch_pairs=None, shard=None, sink="files", comp_engine="auto"):
    """Create and write Illumina reads to FASTQ file(s).

    Positional and keyword arguments up to ``overwrite`` are the reference's.
    Keyword-only additions: ``seed`` (the reference draws its seeds from R's RNG;
    here ``None`` draws one from numpy's global RNG, so ``np.random.seed`` plays
    the part of ``set.seed``), ``device``/``ctx`` (which GPU), ``batch_pairs``,
    ``shard=(index, count)``, ``comp_engine`` ("auto": files compressed at levels 1-6 come from the device
    BGZF coder, levels 7-9 from zlib on the writer threads; "host"; "device": also memory and callable sinks
    receive BGZF bytes), and ``sink``: "files" (default, returns ``None``
    like the reference), "memory" (returns ``(r1_bytes, r2_bytes, stats)``) or
    "device" (generate and discard on the GPU; returns ``stats``), or a callable
    ``sink(job, end, buffer)`` that receives every batch's FASTQ bytes from the library's
    pinned host buffers (returns ``stats``)."""
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    p, keep, (prof1, prof2), is_haps, fns = _prepare(
        obj, out_prefix, n_reads, read_length, paired, frag_mean, frag_sd, matepair, seq_sys, profile1, profile2,
        ins_prob1, del_prob1, ins_prob2, del_prob2, frag_len_min, frag_len_max, haplotype_probs, barcodes, prob_dup,
        sep_files, compress, comp_method, n_threads, read_pool_size, show_progress, overwrite, seed, batch_pairs,
        shard, check_files=(isinstance(sink, str) and sink == "files"), comp_engine=comp_engine)
    ctx = ctx or default_context(device)
    if is_haps:
        ctx.set_haplotypes(obj)
    else:
        ctx.set_genome(obj, wait=False)
    ctx.set_profile(0, prof1)
    if prof2 is not None:
        ctx.set_profile(1, prof2)
    stats = _lib.RunStats()
    lib = ctx.lib
    if isinstance(sink, str) and sink == "files":
        fn = lib.jlp_illumina_hap if is_haps else lib.jlp_illumina_ref
        ctx._check(fn(ctx.h, C.byref(p), C.byref(stats)), "illumina")
        return None
    if callable(sink):
        # sink(job, end, memoryview) for every batch, from the library's pinned host buffers
        def _cb(_user, job, end, data, n):
            try:
                sink(int(job), int(end), (C.c_char * n).from_address(data) if n else b"")
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1
        cb = _lib.CHUNK_CB(_cb)
        ctx._check(lib.jlp_illumina_stream(ctx.h, int(is_haps), C.byref(p), cb, None, C.byref(stats)), "illumina")
        return stats.as_dict()
    if sink == "device":
        ctx._check(lib.jlp_illumina_device_only(ctx.h, int(is_haps), C.byref(p), C.byref(stats)), "illumina")
        return stats.as_dict()
    if sink != "memory":
        raise JackalopeError("sink must be \"files\", \"memory\" or \"device\"")
    n_ends = 2 if p.paired else 1
    n_rec = int(n_reads) // n_ends
    if p.shard_count > 1:        # a shard holds about 1/count of every job (one job per haplotype with sep_files)
        n_jobs = obj.n_haps() if (is_haps and p.sep_files) else 1
        n_rec = n_rec // p.shard_count + n_jobs
    max_name = max(len(n) for n in (obj.reference.names if is_haps else obj.names))
    max_gn = max(len(n) for n in obj.hap_names) if is_haps else 3
    cap = n_rec * (max_name + max_gn + 32 + 2 * int(read_length) + 8) + 64
    o1 = C.create_string_buffer(cap)
    o2 = C.create_string_buffer(cap if n_ends == 2 else 1)
    l1, l2 = C.c_uint64(), C.c_uint64()
    ctx._check(lib.jlp_illumina_to_memory(ctx.h, int(is_haps), C.byref(p), o1, cap, C.byref(l1),
                                          o2 if n_ends == 2 else None, cap if n_ends == 2 else 0, C.byref(l2),
                                          C.byref(stats)), "illumina")
    return o1.raw[:l1.value], (o2.raw[:l2.value] if n_ends == 2 else b""), stats.as_dict()
