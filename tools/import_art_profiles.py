#!/usr/bin/env python3
"""Pack the ART empirical quality profiles jackalope ships as data
(/root/reference/inst/art_profiles/*.txt.gz) into
one container, jackalope_b200/data/art_profiles.bin, so the built-in `seq_sys`
names keep working where /root/reference does not exist (the GPU box).

Only the T/C/A/G rows are kept -- the reference drops everything else
(R/hts_illumina.R:229).  The numbers are data and are stored verbatim: quality
values and CUMULATIVE counts, exactly what the text files hold.

Container (little endian, then gzip'd):
    b"JLPP1\\n", u32 n_profiles, then per profile
    u16 len(name), name, u32 n_pos, and for nt in T,C,A,G, pos in 0..n_pos-1:
    u16 nq, nq * u8 quality, nq * u64 cumulative count
"""
import glob
import gzip
import os
import struct
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/inst/art_profiles"
DST = os.path.join(os.path.dirname(__file__), "..", "jackalope_b200", "data", "art_profiles.bin")


def parse(path):
    rows = {}
    with gzip.open(path, "rt") as fh:
        lines = [l.rstrip("\n") for l in fh if l[:1] in "TCAG"]
    assert len(lines) % 2 == 0, path
    for i in range(0, len(lines), 2):
        # R's strsplit drops one trailing empty field (some files end lines with a tab)
        a, b = lines[i].split("\t"), lines[i + 1].split("\t")
        if a[-1] == "":
            a.pop()
        if b[-1] == "":
            b.pop()
        assert a[:2] == b[:2] and len(a) == len(b), (path, i)
        rows[(a[0], int(a[1]))] = ([int(x) for x in a[2:]], [int(x) for x in b[2:]])
    return rows


def main():
    blob = bytearray(b"JLPP1\n")
    files = sorted(glob.glob(os.path.join(SRC, "*.txt.gz")))
    blob += struct.pack("<I", len(files))
    for f in files:
        name = os.path.basename(f)[: -len(".txt.gz")].encode()
        rows = parse(f)
        n_pos = 1 + max(p for (_, p) in rows)
        blob += struct.pack("<H", len(name)) + name + struct.pack("<I", n_pos)
        for nt in "TCAG":
            for pos in range(n_pos):
                q, c = rows[(nt, pos)]
                blob += struct.pack("<H", len(q)) + bytes(q) + struct.pack("<%dQ" % len(c), *c)
    with gzip.GzipFile(DST, "wb", mtime=0) as out:
        out.write(bytes(blob))
    print("wrote", os.path.normpath(DST), len(files), "profiles", os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
