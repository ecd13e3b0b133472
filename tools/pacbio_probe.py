#!/usr/bin/env python3
"""PacBio reads (first CUDA version) next to the unmodified reference on the host cores: default pacbio() arguments
on a 100 Mb genome (10 x 10 Mb).  Prints one JSON line; kept under profiles/."""
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jackalope_b200 as J  # noqa: E402
from oracle import harness as H  # noqa: E402
from oracle.harness_pacbio import DEFAULTS  # noqa: E402

threads = min(os.cpu_count() or 1, 32)
g = J.random_genome(10, 10_000_000, seed=3)
ctx = J.Context(0)
n = 1 << 17
J.pacbio(g, "", 4096, seed=1, ctx=ctx, sink="device", n_threads=threads)
t0 = time.perf_counter()
st = J.pacbio(g, "", n, seed=2, ctx=ctx, sink="device", n_threads=threads)
t_dev = time.perf_counter() - t0
out = {"workload": "10 x 10 Mb genome, pacbio() defaults", "reads": n, "bases": st["bytes_out"][0] / 2,
       "device": {"kernel_ms": st["reads_ms"], "kernel_reads_per_s": n / (st["reads_ms"] / 1e3), "wall_s": t_dev,
                  "reads_per_s": n / t_dev, "fastq_GBps_kernel": st["bytes_out"][0] / (st["reads_ms"] / 1e3) / 1e9,
                  "host_threads": threads}}
d = tempfile.mkdtemp(dir="/dev/shm")
try:
    J.pacbio(g, os.path.join(d, "w"), 1 << 15, seed=1, ctx=ctx, n_threads=threads, overwrite=True)      # buffers
    J.pacbio(g, os.path.join(d, "w"), 1 << 15, seed=1, ctx=ctx, n_threads=threads, compress=True, overwrite=True)
    t0 = time.perf_counter()
    J.pacbio(g, os.path.join(d, "p"), n, seed=2, ctx=ctx, n_threads=threads)
    t_f = time.perf_counter() - t0
    out["files"] = {"wall_s": t_f, "reads_per_s": n / t_f, "bytes": os.path.getsize(os.path.join(d, "p_R1.fq"))}
    t0 = time.perf_counter()
    J.pacbio(g, os.path.join(d, "z"), n, seed=2, ctx=ctx, n_threads=threads, compress=True)
    t_z = time.perf_counter() - t0
    out["files_bgzf_device"] = {"wall_s": t_z, "reads_per_s": n / t_z, "bytes": os.path.getsize(os.path.join(d, "z_R1.fq.gz"))}
    if H.have_ref(False):
        lib = H.ref_lib(False)
        f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
        lib.jrefpb_pacbio_ref.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_double] * 5 + \
            [f64p, u64p, C.c_uint64, C.c_uint64, f64p, f64p, f64p, f64p] + [C.c_double] * 4 + [C.c_char_p, C.c_uint64]
        rg = H.RefGenomeH(g.names, [bytes(s) for s in g.seqs])
        arr = lambda x: np.ascontiguousarray(x, dtype=np.float64)
        cn, cs, sq, nm = arr(DEFAULTS["chi2_params_n"]), arr(DEFAULTS["chi2_params_s"]), arr(DEFAULTS["sqrt_params"]), arr(DEFAULTS["norm_params"])
        ln = DEFAULTS["lognorm_read_length"]
        err = C.create_string_buffer(256)
        for nthr, nr in ((1, 4000), (threads, 4000 * threads)):
            t0 = time.perf_counter()
            rc = lib.jrefpb_pacbio_ref(rg.h, os.path.join(d, "r%d" % nthr).encode(), nr, nthr, 100, 0.0, ln[2], ln[0], ln[1], 50.0,
                                       None, None, 0, 40, cn.ctypes.data_as(f64p), cs.ctypes.data_as(f64p), sq.ctypes.data_as(f64p),
                                       nm.ctypes.data_as(f64p), 0.2, 0.11, 0.04, 0.01, err, 256)
            assert rc == 0, err.value
            t_r = time.perf_counter() - t0
            out["reference_%d_threads" % nthr] = {"reads": nr, "wall_s": t_r, "reads_per_s": nr / t_r}
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
