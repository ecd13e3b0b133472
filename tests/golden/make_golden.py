#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (needs /root/reference;
run `make -C oracle` first).  For every case in cases.py the reference's own read model
(oracle/_ref/libjlp_ref_replay.so: sample_indels, append_pools, fill_read_qual,
fill_fq_lines compiled from /root/reference/src) consumes the draw ledger of that case
and its FASTQ bytes are stored.  The script refuses to write a fixture unless the oracle
agrees byte for byte and the reference consumed exactly the ledger.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import jackalope_b200 as J  # noqa: E402
from golden.cases import CASES, _haps  # noqa: E402
from oracle import harness as H  # noqa: E402
from oracle.compare import oracle_run  # noqa: E402
from test_oracle_vs_reference import to_ref_haps  # noqa: E402


def main():
    H.build()
    from jackalope_b200 import build
    build.build()
    assert H.have_ref(True), "oracle/_ref/libjlp_ref_replay.so missing (no /root/reference?)"
    for name, (make, n_reads, L, paired, seed, kw) in CASES.items():
        obj = make()
        is_hap = isinstance(obj, J.Haplotypes)
        o = oracle_run(obj, n_reads, L, paired, seed, want_ledger=True, **kw)
        p = o["params"]
        nc = o["n_chroms"]
        if is_hap:
            haps, edits = _haps(return_edits=True)
            ref_obj = to_ref_haps(haps, edits, replay=True)
        else:
            ref_obj = H.RefGenomeH(obj.names, [obj.chrom(c) for c in range(nc)], replay=True)
        plan, ledger, cnt = (np.concatenate(o[k]) for k in ("plan", "ledger", "ledger_cnt"))
        nb = obj.n_haps() if is_hap else 1
        r = H.ref_replay(ref_obj, is_hap=is_hap, paired=bool(p.paired), matepair=bool(p.matepair),
                         prof1=o["profiles"][0], prof2=o["profiles"][1],
                         ins_prob=[p.ins_prob1, p.ins_prob2], del_prob=[p.del_prob1, p.del_prob2],
                         barcodes=[p.barcodes[i] for i in range(nb)], hap=plan[:, 0] // nc, chrom=plan[:, 0] % nc,
                         frag_len=plan[:, 1], frag_start=plan[:, 2], script=ledger)
        assert np.array_equal(r["consumed"], cnt), name
        assert r["r1"] == o["r1"] and r["r2"] == o["r2"], name
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(out, r1=np.frombuffer(r["r1"], np.uint8), r2=np.frombuffer(r["r2"], np.uint8),
                            counts=o["groups"].counts)
        print("%-28s %7d + %7d FASTQ bytes, %d draws" % (name, len(r["r1"]), len(r["r2"]), ledger.size))


if __name__ == "__main__":
    main()
