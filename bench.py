#!/usr/bin/env python3
"""Benchmark of the Illumina read-generation hot path (BASELINE.json metric:
Illumina PE150 read pairs/s, HS25).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("human_pe150_hs25", BASELINE.json configs[4] / SURVEY.md section 8d config 5):
synthetic 3.1 Gb genome (24 chromosomes with human-like length spread, uniform TCAG),
PE150, HS25 profile (HiSeq2500L150R{1,2}filter), all other illumina() arguments at the
reference's defaults.  The full job is 30x coverage = 3.1e8 pairs; a STEP is one device
batch of --batch-pairs pairs (default 2^20) of that job, so K steps generate K * 2^20
pairs per GPU.  With N GPUs the job's pair-index range is sharded contiguously over the
ranks (no collective on the data path; "weak" scaling: per-GPU work is fixed).

One JSON line is printed by rank 0 (see README / DESIGN.md section 8 for the keys):
  value     pairs/s with the genome resident in HBM and the FASTQ left in HBM
            (jlp_illumina_device_only), timed with CUDA events on the library's compute
            stream, max over ranks;
  e2e       pairs/s through the C ABI with HOST buffers: the genome is uploaded from
            pinned host memory inside the timed region and every batch's FASTQ is copied to
            the library's pinned host buffers and handed to the caller (jlp_illumina_stream);
  roofline  the dominant kernel (k_reads: template gather + quality/error model + FASTQ
            records): algorithmic bytes per launch / its CUDA-event time, against
            MEASURED_PEAKS.json;
  bgzf      the same run with compress = 6 / 1 on the device (BGZF written by k_bgzf): kernel time, and end to
            end with only compressed bytes crossing PCIe;
  pacbio    a short device-resident run of pacbio() defaults on the same genome (reads/s, kernel time);
  cpu_baseline  the unmodified reference (oracle/_ref/libjlp_ref.so) or, if that is not
            built, the oracle port, on the host cores, on a bounded sample.

--impl reference times the reference's own CPU implementation (illumina_ref_cpp, all
host threads) on bounded samples of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# GRCh38 chromosome lengths (Mb, rounded) -- only their proportions are used
HUMAN_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
METRIC = "illumina_pe150_hs25_read_pairs_per_sec"
UNIT = "read pairs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload(name, genome_bases):
    """(chromosome lengths, read_length, illumina kwargs, full-job pairs)"""
    if name != "human_pe150_hs25":
        raise SystemExit("unknown workload " + name)
    w = np.array(HUMAN_MB, dtype=np.float64)
    lens = np.floor(w / w.sum() * genome_bases).astype(np.int64)
    lens[0] += genome_bases - lens.sum()
    L = 150
    full_pairs = int(genome_bases * 30 // (2 * L))
    return lens, L, dict(seq_sys="HS25"), full_pairs


def make_genome_into(buf, lens, seed):
    """Uniform TCAG bases (create_genome's default pi_tcag) written into `buf` (uint8[total])."""
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"TCAG", dtype=np.uint8)
    chunk = 1 << 26
    total = int(np.sum(lens))
    for o in range(0, total, chunk):
        n = min(chunk, total - o)
        w = rng.integers(0, 2 ** 63, size=(n + 7) // 8, dtype=np.int64).view(np.uint8)[:n]
        np.take(lut, w & 3, out=buf[o:o + n])


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every ~2 ms
    (nvidia-smi -lms cannot sample a region of a few tens of milliseconds); falls back to one
    nvidia-smi query if NVML is unavailable."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            if uuid:
                try:
                    self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
                except Exception:
                    self._h = None
            if self._h is None:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._nv is None:
            return
        self._stop.clear()
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()

    def stop(self):
        if self._nv is None:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                f = [float(x) for x in o.strip().split(",")]
                return {"sm_mhz": f[0], "sm_max_mhz": f[1], "samples": 1, "reasons": [], "how": "nvidia-smi after the region"}
            except Exception:
                return None
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons), "how": "NVML polled every 2 ms during the timed region"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------- CPU baseline ---

def cpu_reference_setup(lens, seqs_view):
    """The unmodified reference's RefGenome over the same bases (oracle/_ref), or None."""
    from oracle import harness as H
    if not H.have_ref(False):
        return None
    off = np.concatenate(([0], np.cumsum(lens)))
    return H.RefGenomeH(["chrom%d" % i for i in range(len(lens))],
                        [seqs_view[off[i]:off[i + 1]].tobytes() for i in range(len(lens))])


def cpu_reference_run(ref, n_pairs, L, prof1, prof2, threads, seed):
    """illumina_ref_cpp on `threads` host threads; returns seconds."""
    from oracle import harness as H
    import shutil
    need = n_pairs * 2 * (2 * L + 32)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need else None
    with tempfile.TemporaryDirectory(dir=shm) as d:
        if shm is None:          # no room on tmpfs: the reference writes into /dev/null instead
            for k in (1, 2):
                os.symlink("/dev/null", os.path.join(d, "r_R%d.fq" % k))
        t0 = time.perf_counter()
        H.ref_illumina_ref(ref, paired=True, matepair=False, out_prefix=os.path.join(d, "r"), n_reads=2 * n_pairs,
                           prob_dup=0.02, n_threads=threads, read_pool_size=1000, shape=16.0, scale=25.0,
                           frag_len_min=L, frag_len_max=2 ** 32 - 1, prof1=prof1, prof2=prof2,
                           ins_prob=[0.00009, 0.00015], del_prob=[0.00011, 0.00023], r_seed=seed)
        return time.perf_counter() - t0


def cpu_reference_pacbio(ref, n_reads, threads):
    """pacbio_ref_cpp (defaults of pacbio()) of the unmodified reference on `threads` host threads; returns seconds."""
    from oracle import harness as H
    from oracle.harness_pacbio import DEFAULTS as D
    import shutil
    lib = H.ref_lib(False)
    f64p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    lib.jrefpb_pacbio_ref.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_double] * 5 + \
        [f64p, u64p, C.c_uint64, C.c_uint64, f64p, f64p, f64p, f64p] + [C.c_double] * 4 + [C.c_char_p, C.c_uint64]
    arr = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    cn, cs, sq, nm = arr(D["chi2_params_n"]), arr(D["chi2_params_s"]), arr(D["sqrt_params"]), arr(D["norm_params"])
    ln = D["lognorm_read_length"]
    need = n_reads * 20000
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need else None
    err = C.create_string_buffer(256)
    with tempfile.TemporaryDirectory(dir=shm) as d:
        if shm is None:
            os.symlink("/dev/null", os.path.join(d, "p_R1.fq"))
        t0 = time.perf_counter()
        rc = lib.jrefpb_pacbio_ref(ref.h, os.path.join(d, "p").encode(), n_reads, threads, 100, 0.0, ln[2], ln[0], ln[1], 50.0, None, None,
                                   0, 40, cn.ctypes.data_as(f64p), cs.ctypes.data_as(f64p), sq.ctypes.data_as(f64p), nm.ctypes.data_as(f64p),
                                   0.2, 0.11, 0.04, 0.01, err, 256)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return time.perf_counter() - t0


def cpu_port_run(genome, n_pairs, L, kw, seed):
    """The oracle port (1 thread) on the same workload; returns seconds."""
    from oracle.compare import oracle_run
    t0 = time.perf_counter()
    oracle_run(genome, 2 * n_pairs, L, True, seed, **kw)
    return time.perf_counter() - t0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(genome, lens, flat_bases, L, kw, prof1, prof2, target_s=12.0, pacbio_out=None):
    threads = host_threads()
    ref = cpu_reference_setup(lens, flat_bases)
    if ref is not None:
        n0 = 20000 * threads
        t = cpu_reference_run(ref, n0, L, prof1, prof2, threads, 1)
        n1 = int(max(n0, min(1e7, n0 / t * target_s)))
        t1 = cpu_reference_run(ref, n1, L, prof1, prof2, threads, 2)
        if pacbio_out is not None:
            try:        # the PacBio generator of the unmodified reference on the same genome and threads: about 3 s
                tp = cpu_reference_pacbio(ref, 2000 * threads, threads)
                npb = int(max(2000 * threads, min(4e5, 2000 * threads / tp * 3.0)))
                tp = cpu_reference_pacbio(ref, npb, threads)
                pacbio_out["cpu_baseline"] = {"value": npb / tp, "unit": "reads/s", "cores": threads, "kind": "reference",
                                              "sample": "%d reads through pacbio_ref_cpp, n_threads=%d, %.1f s" % (npb, threads, tp)}
            except Exception as e:
                pacbio_out["cpu_baseline"] = {"error": str(e)[:200]}
        return {"value": n1 / t1, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "%d pairs of the same workload (3.1 Gb genome, PE150 HS25) through illumina_ref_cpp, "
                          "n_threads=%d, read_pool_size=1000, output to tmpfs (or /dev/null when tmpfs is too small), %.1f s" % (n1, threads, t1)}
    n0 = 20000
    t = cpu_port_run(genome, n0, L, kw, 1)
    n1 = int(max(n0, min(2e6, n0 / t * target_s)))
    t1 = cpu_port_run(genome, n1, L, kw, 2)
    return {"value": n1 / t1, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d pairs of the same workload through the C oracle port (1 thread), %.1f s" % (n1, t1)}


# ------------------------------------------------------------------------ main ---

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="human_pe150_hs25")
    ap.add_argument("--batch-pairs", type=int, default=1 << 20)
    ap.add_argument("--genome-bases", type=float, default=3.1e9, help="shrink for a quick functional run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=20261018)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else max(a.warmup, 1)
    rank, local_rank, world = dist_env()
    if a.impl == "reference" and rank != 0:
        return 0
    # stdout carries exactly one JSON line: everything else a library prints there (NCCL's version
    # banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    lens, L, kw, full_pairs = workload(a.workload, int(a.genome_bases))
    total = int(lens.sum())
    config = {"workload": a.workload, "genome_bases": total, "n_chroms": len(lens), "read_length": L, "paired": True,
              "seq_sys": "HS25", "profiles": "HiSeq2500L150R1filter/HiSeq2500L150R2filter", "frag_mean": 400,
              "frag_sd": 100, "prob_dup": 0.02, "full_job_pairs": full_pairs, "pairs_per_step": a.batch_pairs,
              "sharding": "contiguous pair-index ranges, one per GPU, no collective",
              "l2": "inputs (3.1 GB genome, random gather) and per-step outputs (~0.66 GB) are larger than the 126 MB L2"}

    from __graft_entry__ import build
    build()
    import jackalope_b200 as J

    if a.impl == "reference":
        return reference_arm(a, lens, L, kw, config, J, real_stdout)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: genome in PINNED host memory (the e2e leg uploads it inside its timed region)
    t0 = time.perf_counter()
    pinned = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    flat = pinned.numpy()
    make_genome_into(flat, lens, a.seed)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    genome = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
    genome.flat = lambda: (flat, off.astype(np.uint64))      # already contiguous: no copy
    log("[bench] rank %d: genome %.2f Gb built in %.1f s" % (rank, total / 1e9, time.perf_counter() - t0))

    ctx = J.Context(local_rank)
    B = a.batch_pairs

    def run(n_steps, sink, seed, **more):
        return J.illumina(genome, "", 2 * n_steps * B * world, L, True, seed=seed, ctx=ctx, sink=sink,
                          batch_pairs=B, shard=(rank, world), **kw, **more)

    # ---- device-resident leg
    for i in range(a.warmup):
        run(1, "device", a.seed + 100 + i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    st = run(a.steps, "device", a.seed)
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    assert st["pairs"] == a.steps * B and st["batches"] == a.steps, st
    run_ms = max_over_ranks(st["run_ms"])
    launches = int(sum_over_ranks(st["kernel_launches"]))
    value = a.steps * B * world / (run_ms / 1e3)
    place_ms = st["place_ms"] / st["batches"]
    reads_ms = st["reads_ms"] / st["batches"]
    if rank == 0:
        log("[bench] device leg: %.3f ms/step (events), wall %.3f s, k_place %.3f ms, k_reads %.3f ms, bytes/pair %.1f"
            % (run_ms / a.steps, wall, place_ms, reads_ms, sum(st["bytes_out"]) / st["pairs"]))

    # ---- end-to-end leg: genome H2D + every batch's FASTQ D2H into pinned host buffers
    e2e = None
    if not a.no_e2e:
        seen = [0, 0]

        def sink(job, end, buf):
            seen[end] += len(buf)

        ctx2 = ctx
        for i in range(min(a.warmup, 3)):          # also sizes the pinned buffers
            h2d_before = run(1, sink, a.seed + 200 + i)["h2d_bytes"]      # cumulative per context
        seen[0] = seen[1] = 0
        barrier()
        t0 = time.perf_counter()
        ctx2._genome = None                        # force the upload: one illumina() call = one genome H2D
        st2 = run(a.steps, sink, a.seed)
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        assert seen[0] == st2["bytes_out"][0] and seen[1] == st2["bytes_out"][1] and st2["bytes_out"] == st["bytes_out"]
        e2e = {"value": a.steps * B * world / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": (st2["h2d_bytes"] - h2d_before) / a.steps, "d2h_bytes_per_step": st2["d2h_bytes"] / a.steps,
               "ms_per_step": t_e2e / a.steps * 1e3,
               "note": "wall clock around jlp_set_genome_async + jlp_illumina_stream: the chromosomes this rank's shard reads "
                       "are copied from pinned host memory inside the region, every batch's FASTQ lands in the library's "
                       "pinned host buffers"}

    # ---- compress = TRUE on the device (BGZF members written by k_bgzf): kernel time with the output left in HBM,
    #      and end to end with only the compressed bytes crossing PCIe; level 6 (the reference's default: literals +
    #      line-prefix matches) and level 1 (literals only)
    bgzf = None
    if not a.no_e2e:
        bgzf = {"note": "compress=<level>, comp_engine=device: k_bgzf + scan + gather after k_reads; e2e hands BGZF bytes to the "
                        "caller from pinned host buffers (jlp_illumina_stream); GB/s counts FASTQ bytes read + BGZF bytes written"}
        for level in (6, 1):
            zkw = dict(compress=level, comp_engine="device")
            run(1, "device", a.seed + 400, **zkw)
            barrier()
            stz = run(a.steps, "device", a.seed, **zkw)
            barrier()
            z_ms = max_over_ranks(stz["bgzf_ms"] / stz["batches"])
            zrun_ms = max_over_ranks(stz["run_ms"])
            zseen = [0, 0]

            def zsink(job, end, buf):
                zseen[end] += len(buf)

            run(1, zsink, a.seed + 401, **zkw)
            zseen[0] = zseen[1] = 0
            barrier()
            ctx._genome = None
            t0 = time.perf_counter()
            stz2 = run(a.steps, zsink, a.seed, **zkw)
            barrier()
            t_z = max_over_ranks(time.perf_counter() - t0)
            assert stz["bytes_out"] == st["bytes_out"] and zseen[0] == stz2["z_bytes"][0] + 28, (stz, zseen)
            zin, zout = sum(stz["bytes_out"]) / stz["batches"], sum(stz["z_bytes"]) / stz["batches"]
            bgzf["level%d" % level] = {
                "device_resident": {"value": a.steps * B * world / (zrun_ms / 1e3), "unit": UNIT, "ms_per_step": zrun_ms / a.steps},
                "e2e": {"value": a.steps * B * world / t_z, "unit": UNIT, "ms_per_step": t_z / a.steps * 1e3,
                        "d2h_bytes_per_step": stz2["d2h_bytes"] / a.steps},
                "ratio": zout / zin, "k_bgzf_ms_per_step": z_ms, "k_bgzf_GBps": (zin + zout) / (z_ms / 1e3) / 1e9}
            if rank == 0:
                log("[bench] device BGZF level %d: %.3f ms/step, ratio %.3f, e2e %.2f ms/step" % (level, z_ms, zout / zin, t_z / a.steps * 1e3))

    # ---- the same through FILES (the reference-facing default sink): tmpfs, all host threads writing
    e2e_files = None
    if not a.no_e2e and rank == 0 and world == 1 and os.path.isdir("/dev/shm"):
        import shutil
        need = int(sum(st["bytes_out"]) * 1.1)
        if shutil.disk_usage("/dev/shm").free > 2 * need:
            d = tempfile.mkdtemp(dir="/dev/shm")
            try:
                nthr = min(host_threads(), 32)
                J.illumina(genome, os.path.join(d, "w"), 2 * B, L, True, seed=a.seed + 300, ctx=ctx, batch_pairs=B,
                           n_threads=nthr, overwrite=True, **kw)
                ctx._genome = None
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                J.illumina(genome, os.path.join(d, "r"), 2 * a.steps * B, L, True, seed=a.seed, ctx=ctx, batch_pairs=B,
                           n_threads=nthr, overwrite=True, **kw)
                t_files = time.perf_counter() - t0
                sz = os.path.getsize(os.path.join(d, "r_R1.fq")) + os.path.getsize(os.path.join(d, "r_R2.fq"))
                assert sz == sum(st["bytes_out"]), (sz, st["bytes_out"])
                e2e_files = {"value": a.steps * B / t_files, "unit": UNIT, "ms_per_step": t_files / a.steps * 1e3,
                             "writer_threads": nthr, "bytes_written": sz,
                             "note": "illumina(obj, out_prefix, ...) writing <prefix>_R{1,2}.fq on tmpfs; genome H2D inside"}
                # compress = TRUE (the reference's default level 6 -> the device coder)
                ctx._genome = None
                t0 = time.perf_counter()
                J.illumina(genome, os.path.join(d, "z"), 2 * a.steps * B, L, True, seed=a.seed, ctx=ctx, batch_pairs=B,
                           n_threads=nthr, compress=True, overwrite=True, **kw)
                t_z = time.perf_counter() - t0
                zsz = os.path.getsize(os.path.join(d, "z_R1.fq.gz")) + os.path.getsize(os.path.join(d, "z_R2.fq.gz"))
                e2e_files["bgzip_device"] = {"value": a.steps * B / t_z, "unit": UNIT, "steps": a.steps, "compressed_bytes": zsz,
                                             "ratio": zsz / sz, "note": "compress=TRUE, comp_method=bgzip (BGZF written by the GPU)"}
                # the same with zlib level 6 on the writer threads: 1 step
                t0 = time.perf_counter()
                J.illumina(genome, os.path.join(d, "y"), 2 * B, L, True, seed=a.seed, ctx=ctx, batch_pairs=B,
                           n_threads=nthr, compress=True, comp_engine="host", overwrite=True, **kw)
                t_z = time.perf_counter() - t0
                zsz = os.path.getsize(os.path.join(d, "y_R1.fq.gz")) + os.path.getsize(os.path.join(d, "y_R2.fq.gz"))
                e2e_files["bgzip_host_zlib6"] = {"value": B / t_z, "unit": UNIT, "steps": 1, "compressed_bytes": zsz,
                                                 "ratio": zsz / (sz / a.steps), "note": "comp_engine=host: zlib level 6 on the writer threads"}
            finally:
                shutil.rmtree(d, ignore_errors=True)

    # ---- PacBio reads, a short device-resident run of pacbio() defaults on the same genome
    pacbio = None
    if not a.no_e2e and rank == 0 and world == 1:
        try:
            nthr = min(host_threads(), 32)
            J.pacbio(genome, "", 1 << 13, seed=a.seed, ctx=ctx, sink="device", n_threads=nthr)
            runs = []
            for rep in range(3):        # three runs, the median reported (the host preparation shares the cores with whatever else runs)
                t0 = time.perf_counter()
                stp = J.pacbio(genome, "", 1 << 16, seed=a.seed + 1 + rep, ctx=ctx, sink="device", n_threads=nthr)
                runs.append(time.perf_counter() - t0)
            t_pb = sorted(runs)[1]
            pacbio = {"reads": stp["pairs"], "bases": stp["bytes_out"][0] / 2, "reads_per_s": stp["pairs"] / t_pb, "wall_s_runs": runs,
                      "kernel_ms": stp["reads_ms"], "kernel_reads_per_s": stp["pairs"] / (stp["reads_ms"] / 1e3),
                      "kernel_fastq_GBps": stp["bytes_out"][0] / (stp["reads_ms"] / 1e3) / 1e9, "host_threads": nthr,
                      "note": "pacbio() defaults, reads left on the device; reads_per_s includes the per-read host preparation"}
        except Exception as e:          # the Illumina line is the contract; never lose it to the extra leg
            pacbio = {"error": str(e)[:200]}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": run_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic", "config": config, "clocks": clk, "e2e": e2e, "e2e_files": e2e_files, "bgzf": bgzf, "pacbio": pacbio,
           "gpu_launches": launches}

    if rank == 0:
        peaks, which = None, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            which = "measured"
        except Exception:
            peaks = 6650.0
        # k_reads (fused quality/error + FASTQ): reads 2L template bases (1 B/base), writes the FASTQ records
        alg_bytes = 2 * L * B + sum(st["bytes_out"]) / st["batches"]
        achieved = alg_bytes / (reads_ms / 1e3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("k_reads")
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": "k_reads (template gather + quality/error model + FASTQ records)", "achieved": achieved, "peak": peaks,
                           "unit": "GB/s", "frac": achieved / peaks, "traffic": traffic, "peak_source": which,
                           "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": reads_ms,
                           "share_of_step": reads_ms / (run_ms / a.steps), "k_place_ms_per_launch": place_ms,
                           "whole_path_GBps": (2 * L * B + sum(st["bytes_out"]) / st["batches"]) / (run_ms / a.steps / 1e3) / 1e9}
        if not a.no_cpu_baseline and world == 1:
            prof1, prof2 = (J.flatten_profile(J.read_profile(None, "HS25", L, r)) for r in (1, 2))
            t0 = time.perf_counter()
            out["cpu_baseline"] = cpu_baseline(genome, lens, flat, L, kw, prof1, prof2,
                                               pacbio_out=pacbio if isinstance(pacbio, dict) and "error" not in pacbio else None)
            log("[bench] cpu baseline took %.1f s" % (time.perf_counter() - t0))
        print(json.dumps(out), file=real_stdout, flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm(a, lens, L, kw, config, J, real_stdout):
    """The reference's own CPU implementation on bounded samples of the workload."""
    total = int(lens.sum())
    flat = np.empty(total, dtype=np.uint8)
    make_genome_into(flat, lens, a.seed)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    genome = J.RefGenome(["chrom%d" % i for i in range(len(lens))], [flat[off[i]:off[i + 1]] for i in range(len(lens))])
    prof1, prof2 = (J.flatten_profile(J.read_profile(None, "HS25", L, r)) for r in (1, 2))
    threads = host_threads()
    ref = cpu_reference_setup(lens, flat)
    kind = "reference" if ref is not None else "port"
    cores = threads if ref is not None else 1

    def step(n, seed):
        if ref is not None:
            return cpu_reference_run(ref, n, L, prof1, prof2, threads, seed)
        return cpu_port_run(genome, n, L, kw, seed)

    n = 5000 * cores
    t = step(n, 1)
    budget = min(3.0, 150.0 / max(1, a.steps + a.warmup))        # seconds per step: the whole run ends within minutes
    n = int(max(n, min(4e6, n / t * budget)))
    for i in range(max(0, a.warmup - 1)):
        step(n, 10 + i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(n, 100 + i)
    dt = time.perf_counter() - t0
    v = a.steps * n / dt
    config = dict(config, pairs_per_step=n)
    sample = ("%d pairs per step of the same workload through %s, %d host threads, output to tmpfs (or /dev/null when tmpfs is too small)"
              % (n, "illumina_ref_cpp (unmodified reference, oracle/_ref)" if ref is not None else "the C oracle port", cores))
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), file=real_stdout, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
