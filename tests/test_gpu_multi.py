"""One call, several GPUs (jlp_ctx_create_multi): the run is cut into one contiguous piece per device, every device
writes its piece into the SAME ordered files -- plain FASTQ at offsets known from a size pass, device-compressed output at
offsets known from a dry run of the coder, host-compressed output as per-device parts joined in device order -- and the result must be the single-device run byte for byte
(/root/reference/src/hts.h:334-353 split over threads, :401-416 one set of files, :512-552 sep_files).
On a one-GPU host the devices of the multi-GPU context are device 0 listed several times (the pieces then share the
GPU: same code path); with more GPUs visible the same tests also run over all of them."""
import gzip
import os

import numpy as np
import pytest
import torch

import jackalope_b200 as J
from common import first_diff, oracle_run

pytestmark = pytest.mark.gpu


def device_lists():
    n = torch.cuda.device_count()
    out = [[0, 0, 0]]
    if n >= 2:
        out.append(list(range(n)))
    return out


def read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_reference_run_one_file_set_from_several_devices(ctx, tmp_path, devices):
    g = J.random_genome(5, [60_000, 20_000, 90_000, 5_000, 40_000], seed=201)
    kw = dict(seq_sys="HS25", prob_dup=0.3, read_pool_size=64, ins_prob1=0.002, del_prob2=0.003)
    n_reads = 50_000
    m = J.Context(devices=devices)
    try:
        assert m.n_devices == len(devices)
        one, many = str(tmp_path / "one"), str(tmp_path / "many")
        J.illumina(g, one, n_reads, 150, True, seed=5, ctx=ctx, batch_pairs=3000, **kw)
        J.illumina(g, many, n_reads, 150, True, seed=5, ctx=m, batch_pairs=3000, n_threads=4, **kw)
        for r in (1, 2):
            assert first_diff(read("%s_R%d.fq" % (many, r)), read("%s_R%d.fq" % (one, r))) is None
        o = oracle_run(g, n_reads, 150, True, 5, **kw)
        assert read(many + "_R1.fq") == o["r1"] and read(many + "_R2.fq") == o["r2"]
        # memory sink: the devices write into disjoint ranges of the caller's buffers
        r1, r2, st = J.illumina(g, "", n_reads, 150, True, seed=5, ctx=m, sink="memory", batch_pairs=3000, **kw)
        assert r1 == o["r1"] and r2 == o["r2"] and st["pairs"] == n_reads // 2
        # device-only: every pair generated exactly once
        st = J.illumina(g, "", n_reads, 150, True, seed=5, ctx=m, sink="device", **kw)
        assert st["pairs"] == n_reads // 2 and st["bytes_out"] == [len(o["r1"]), len(o["r2"])]
        # compressed on the device: sizes from a dry run, every device writes its BGZF members to their final place; one EOF block
        J.illumina(g, many, n_reads, 150, True, seed=5, ctx=m, batch_pairs=3000, compress=True, n_threads=2, **kw)
        for r in (1, 2):
            z = read("%s_R%d.fq.gz" % (many, r))
            assert z[-28:-24] == b"\x1f\x8b\x08\x04" and z.count(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00\x1b\x00\x03\x00") == 1
            assert gzip.decompress(z) == (o["r1"] if r == 1 else o["r2"])
        assert not [f for f in os.listdir(tmp_path) if ".part" in f]
        # zlib on the writer threads (levels 7-9): per-device parts, joined in device order, no parts left behind
        J.illumina(g, many + "9", n_reads, 150, True, seed=5, ctx=m, batch_pairs=3000, compress=9, n_threads=2, **kw)
        assert gzip.decompress(read(many + "9_R2.fq.gz")) == o["r2"]
        assert not [f for f in os.listdir(tmp_path) if ".part" in f]
        # the stream sink has no order over several devices
        with pytest.raises(RuntimeError, match="multi-GPU"):
            J.illumina(g, "", n_reads, 150, True, seed=5, ctx=m, sink=lambda *a: None, **kw)
    finally:
        m.close()


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_haplotypes_sep_files_from_several_devices(ctx, tmp_path, devices):
    g = J.random_genome(3, 50_000, seed=202)
    haps = J.random_haplotypes(g, 7, sub_rate=0.01, indel_rate=0.002, seed=203)
    probs = [5, 1, 1, 0, 3, 1, 8]
    n_reads = 40_000
    kw = dict(seq_sys="HS25", haplotype_probs=probs, sep_files=True, barcodes=["A", "CC", "GT", "T", "ACG", "TT", "G"])
    m = J.Context(devices=devices)
    try:
        one, many = str(tmp_path / "one"), str(tmp_path / "many")
        J.illumina(haps, one, n_reads, 100, True, seed=6, ctx=ctx, batch_pairs=2500, **kw)
        J.illumina(haps, many, n_reads, 100, True, seed=6, ctx=m, batch_pairs=2500, n_threads=3, **kw)
        total = 0
        for h in haps.hap_names:
            for r in (1, 2):
                a, b = read("%s_%s_R%d.fq" % (many, h, r)), read("%s_%s_R%d.fq" % (one, h, r))
                assert first_diff(a, b) is None, (h, r)
                total += a.count(b"\n")
        assert total == 4 * n_reads and read("%s_hap3_R1.fq" % many) == b""       # probability 0: the files exist, empty
        o = oracle_run(haps, n_reads, 100, True, 6, **kw)
        assert b"".join(read("%s_%s_R1.fq" % (many, h)) for h in haps.hap_names) == o["r1"]
        # pooled into one file pair, compressed on the device
        kw2 = dict(kw, sep_files=False)
        J.illumina(haps, many + "p", n_reads, 100, True, seed=6, ctx=m, batch_pairs=2500, compress=4, **kw2)
        o2 = oracle_run(haps, n_reads, 100, True, 6, **kw2)
        assert gzip.decompress(read(many + "p_R1.fq.gz")) == o2["r1"] and gzip.decompress(read(many + "p_R2.fq.gz")) == o2["r2"]
        r1, r2, _ = J.illumina(haps, "", n_reads, 100, True, seed=6, ctx=m, sink="memory", **kw)
        assert r1 == o["r1"] and r2 == o["r2"]
        # one compressed file pair per haplotype from several devices: the dry run sizes every job's files (a job may
        # have several writers, a device several jobs); a haplotype without reads gets a file with the EOF block only
        J.illumina(haps, many + "z", n_reads, 100, True, seed=6, ctx=m, batch_pairs=2500, compress=True, n_threads=3, **kw)
        eof = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00\x1b\x00\x03\x00\x00\x00\x00\x00\x00\x00\x00\x00"
        for h in haps.hap_names:
            for r in (1, 2):
                z = read("%sz_%s_R%d.fq.gz" % (many, h, r))
                assert z.endswith(eof) and z.count(eof[:20]) == 1, (h, r)
                assert gzip.decompress(z) == read("%s_%s_R%d.fq" % (one, h, r)), (h, r)
        assert read("%sz_hap3_R2.fq.gz" % many) == eof
        assert not [f for f in os.listdir(tmp_path) if ".part" in f]
    finally:
        m.close()


def test_callbacks_of_a_multi_device_run_come_from_the_calling_thread(tmp_path):
    import ctypes as C
    import threading
    from jackalope_b200 import _lib
    from jackalope_b200.illumina import _prepare
    g = J.random_genome(2, 80_000, seed=204)
    m = J.Context(devices=[0, 0])
    try:
        m.set_genome(g)
        a = dict(frag_mean=400, frag_sd=100, matepair=False, seq_sys="HS25", profile1=None, profile2=None, ins_prob1=0.00009,
                 del_prob1=0.00011, ins_prob2=0.00015, del_prob2=0.00023, frag_len_min=None, frag_len_max=None, haplotype_probs=None,
                 barcodes=None, prob_dup=0.02, sep_files=False, compress=False, comp_method="bgzip", n_threads=1, read_pool_size=1000,
                 show_progress=False)
        p, keep, (p1, p2), _, _ = _prepare(g, str(tmp_path / "cb"), 60_000, 100, True, *[a[k] for k in a], True, 9, 2000, None,
                                           check_files=False)
        m.set_profile(0, p1)
        m.set_profile(1, p2)
        seen, threads = [0], set()

        def progress(_u, n):
            seen[0] += n
            threads.add(threading.get_ident())

        p.progress_cb = _lib.PROGRESS_CB(progress)
        p.abort_cb = _lib.ABORT_CB(lambda _u: 0)
        st = _lib.RunStats()
        assert m.lib.jlp_illumina_ref(m.h, C.byref(p), C.byref(st)) == 0
        assert seen[0] == 60_000 and threads == {threading.get_ident()}
        # an abort request stops every device; the call reports JLP_ERR_ABORTED
        calls = [0]

        def abort(_u):
            calls[0] += 1
            return 1 if calls[0] > 1 else 0

        p.abort_cb = _lib.ABORT_CB(abort)
        p.n_reads = 40_000_000
        assert m.lib.jlp_illumina_ref(m.h, C.byref(p), C.byref(st)) == _lib.JLP_ERR_ABORTED
    finally:
        m.close()


def test_shards_of_separate_processes_write_their_own_files(ctx, tmp_path):
    """shard_count > 1 with the files sink (one process per GPU under torchrun): every shard writes
    <name>.shard<i>of<n>; concatenated in shard order they are the unsharded file (round 1 truncated one shared file)."""
    g = J.random_genome(3, 30_000, seed=205)
    kw = dict(seq_sys="HS25", prob_dup=0.2, read_pool_size=40)
    pre = str(tmp_path / "s")
    J.illumina(g, pre, 20_000, 100, True, seed=7, ctx=ctx, **kw)
    for k in range(3):
        J.illumina(g, pre, 20_000, 100, True, seed=7, ctx=ctx, shard=(k, 3), overwrite=True, **kw)
    for r in (1, 2):
        cat = b"".join(read("%s_R%d.fq.shard%dof3" % (pre, r, k)) for k in range(3))
        assert cat == read("%s_R%d.fq" % (pre, r))


def test_haplotypes_are_materialised_only_where_a_shard_reads_them():
    g = J.random_genome(2, 400_000, seed=206)
    haps = J.random_haplotypes(g, 8, sub_rate=0.01, indel_rate=0.001, seed=207)
    c = J.Context(0)
    try:
        kw = dict(seq_sys="HS25", haplotype_probs=[1] * 8)
        st = J.illumina(haps, "", 16_000, 100, True, seed=8, ctx=c, sink="device", shard=(7, 8), **kw)
        small = st["h2d_bytes"]
        st = J.illumina(haps, "", 16_000, 100, True, seed=8, ctx=c, sink="device", **kw)
        assert st["h2d_bytes"] > 4 * small                      # the other seven haplotypes' records arrive only now
        st = J.illumina(haps, "", 16_000, 100, True, seed=8, ctx=c, sink="device", **kw)
        assert st["h2d_bytes"] < small                          # everything resident
    finally:
        c.close()
