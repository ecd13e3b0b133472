"""The device BGZF coder (jackalope_b200/csrc/jlp_bgzf.cu) through the C ABI.

There is no reference bit pattern to match for compressed output (htslib is not in this image and
any valid deflate stream is acceptable to its readers), so parity is the round trip: what zlib
inflates from the device's members must be the bytes that went in -- which for the generators are the
bytes of the uncompressed run, themselves checked against the oracle elsewhere -- plus the BGZF
framing htslib expects (member header with the BC field, BSIZE, CRC-32, ISIZE, EOF block), which
python's gzip module verifies member by member (CRC and size included)."""
import ctypes as C
import gzip
import struct
import zlib

import numpy as np
import pytest

import jackalope_b200 as J

pytestmark = pytest.mark.gpu


def bgzf_blocks(z: bytes):
    out, p = [], 0
    while p < len(z):
        assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 10:p + 16] == b"\x06\x00BC\x02\x00", p
        bsize = struct.unpack_from("<H", z, p + 16)[0] + 1
        isize = struct.unpack_from("<I", z, p + bsize - 4)[0]
        out.append((bsize, isize))
        p += bsize
    assert p == len(z)
    return out


def device_bgzf(ctx, data: bytes, level: int = 6) -> bytes:
    lib = ctx.lib
    n = C.c_uint64()
    cap = len(data) + (len(data) // 0xff00 + 2) * 64 + 64
    buf = C.create_string_buffer(cap)
    rc = lib.jlp_bgzf_device(ctx.h, level, data, len(data), buf, cap, C.byref(n))
    assert rc == 0, lib.jlp_last_error(ctx.h)
    return buf.raw[:n.value]


def huffman_only_size(data: bytes) -> int:
    """Bytes of an optimal (unlimited-depth) literal-only Huffman coding of every 0xff00-byte block, plus
    26 bytes of BGZF framing per block and the EOF block; the dynamic block's own header is not counted."""
    import heapq
    tot = 28
    for i in range(0, len(data), 0xff00):
        cnt = np.bincount(np.frombuffer(data[i:i + 0xff00], np.uint8), minlength=256).tolist() + [1]
        h = [c for c in cnt if c]
        heapq.heapify(h)
        bits = 0
        while len(h) > 1:
            a, b = heapq.heappop(h), heapq.heappop(h)
            bits += a + b
            heapq.heappush(h, a + b)
        # the coder keeps at most 44 KiB of block image: what does not get below that is stored (5 + 26 bytes of framing)
        tot += (bits + 7) // 8 + 26 if (bits + 7) // 8 + 200 < 44 * 1024 else len(data[i:i + 0xff00]) + 31
    return tot


def fastq_like(n, seed):
    rng = np.random.default_rng(seed)
    recs = []
    size = 0
    i = 0
    while size < n:
        seq = bytes(rng.choice(np.frombuffer(b"TCAG", np.uint8), 150))
        q = bytes(rng.choice(np.frombuffer(b"#(18=CGJ", np.uint8), 150, p=[.01, .02, .03, .04, .1, .2, .4, .2]))
        r = b"@REF-chrom%d-%d-F/1\n%s\n+\n%s\n" % (i % 24, rng.integers(0, 10 ** 8), seq, q)
        recs.append(r)
        size += len(r)
        i += 1
    return b"".join(recs)[:n]


CASES = {
    "empty": b"",
    "one_byte": b"A",
    "short_text": b"@REF-chrom0-17404-R/1\nGCGTTACAGC\n+\nCCC1GCGC8G\n",
    "one_symbol": b"G" * 70000,
    "two_symbols": b"GA" * 40000,
    "exact_block": fastq_like(0xff00, 1),
    "block_plus_one": fastq_like(0xff00 + 1, 2),
    "three_blocks_ragged": fastq_like(3 * 0xff00 + 17, 3),
    "random_bytes_stored": np.random.default_rng(4).integers(0, 256, 200000, dtype=np.uint8).tobytes(),
    "all_byte_values_skewed": bytes(np.random.default_rng(5).geometric(0.05, 150000).clip(0, 255).astype(np.uint8)),
    # Fibonacci counts (the end-of-block symbol is the first 1): an unrestricted Huffman code is 21 bits deep,
    # the 15-bit limit has to act
    "fibonacci_depth": b"".join(bytes([65 + k]) * f for k, f in enumerate(
        [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711])),
    # identical and nearly identical lines: matches up to the line's last byte, lengths up to 258, short distances
    "repeated_lines": b"".join(b"%s\n" % (b"ACGT" * (1 + i % 70)) for i in range(2400)),
    "short_lines": b"".join(b"ab%d\n" % (i % 7) for i in range(30000)),             # more line starts than the coder keeps
    "empty_lines": b"\n" * 5000 + b"x\n\n" * 3000,
    # many distinct rare bytes before each line start: literal codes of 12+ bits next to a match (the lane's bits pass 64)
    "long_codes_before_matches": b"".join(bytes(np.random.default_rng(i).integers(128, 256, 45, dtype=np.uint8)) + b"\n" + b"PREFIX-%06d" % (i * 7919 % 10**6) + b"A" * (i % 5) + b"\n" + b"qq\n" + b"rr\n" for i in range(1500)),
    # the codes of a call come from its first block: bytes that block does not hold (outside the printable range, which
    # always gets a code), and blocks the first block's code fits badly, fall back to a code of their own
    "late_binary_bytes": fastq_like(0xff00 + 5000, 8) + bytes(np.random.default_rng(9).integers(0, 32, 70000, dtype=np.uint8)) + fastq_like(30000, 10),
    "nonstationary": b"A" * 0xff00 + fastq_like(2 * 0xff00, 11) + bytes(np.random.default_rng(12).integers(0, 256, 0xff00, dtype=np.uint8)) + b"\n".join(b"line %d" % (i % 13) for i in range(9000)),
    # the coder's units are chunks of 128 and segments of 4096 bytes: newlines in the last byte of every chunk, lines that
    # drift through the chunks, identical 300-byte lines (matches of 258 bytes across chunk and segment boundaries), and
    # lines longer than a segment and than the bytes searched for earlier line starts (PacBio-like records)
    "lines_of_128": b"".join(b"%s\n" % (bytes([65 + (i * 7 + j) % 20 for j in range(120)]) + b"%07d" % (i % 97)) for i in range(1600)),
    "lines_of_129": b"".join(b"%s\n" % (b"HDR-%04d-" % (i % 50) + bytes([97 + (i + j * j) % 13 for j in range(119)])) for i in range(1500)),
    "identical_300": (b"ACGTTGCA" * 37 + b"xyz\n") * 700,
    "long_lines": b"".join(b"@m%d/%d/ccs\n%s\n+\n%s\n" % (i, 1000 + i, bytes(np.random.default_rng(30 + i).choice(np.frombuffer(b"ACGT", np.uint8), 9000 + 911 * i)),
                                                         bytes(np.random.default_rng(60 + i).choice(np.frombuffer(b"!+5?IS]", np.uint8), 9000 + 911 * i)))
                           for i in range(9)),
    "chunk_tail_127": fastq_like(0xff00 + 127, 6),
    "chunk_tail_129": fastq_like(2 * 0xff00 + 129, 7),
}


@pytest.mark.parametrize("level", [1, 6])
@pytest.mark.parametrize("name", list(CASES))
def test_device_bgzf_round_trip(ctx, name, level):
    """level 1: literals only; level 6: also the matches of line prefixes."""
    data = CASES[name]
    z = device_bgzf(ctx, data, level)
    blocks = bgzf_blocks(z)
    assert blocks[-1] == (28, 0)
    assert [i for _, i in blocks[:-1]] == [min(0xff00, len(data) - o) for o in range(0, len(data), 0xff00)]
    assert all(b <= 65536 for b, _ in blocks)
    assert gzip.decompress(z) == data            # inflates, CRC-32 and ISIZE of every member verified
    if len(data) >= 1000 and name != "random_bytes_stored":
        # a literal-only dynamic block: within 1 % (+ 150 bytes of block header per block) of the optimal Huffman coding
        assert len(z) <= 1.01 * huffman_only_size(data) + 150 * len(blocks), (len(z), huffman_only_size(data))
    if name == "random_bytes_stored":
        assert len(z) <= len(data) + (len(data) // 0xff00 + 1) * 31 + 28    # stored blocks: 5 + 26 bytes each
    if level == 6 and name in ("exact_block", "three_blocks_ragged", "repeated_lines", "long_codes_before_matches"):
        assert len(z) < len(device_bgzf(ctx, data, 1))                      # the matches pay


def small_genome(seed):
    return J.random_genome(3, 20000, seed=seed)


def test_generator_compressed_on_device(ctx, tmp_path):
    """Files, memory and stream sinks with the device coder: the inflated bytes are the plain run's."""
    g = small_genome(61)
    n_reads = 60000
    r1, r2, _ = J.illumina(g, "", n_reads, 150, True, seed=62, ctx=ctx, sink="memory")
    # files, default engine (auto -> device at level 6), several batches, several writer threads
    pre = str(tmp_path / "dz")
    J.illumina(g, pre, n_reads, 150, True, seed=62, ctx=ctx, compress=True, n_threads=4, batch_pairs=7000)
    z1, z2 = open(pre + "_R1.fq.gz", "rb").read(), open(pre + "_R2.fq.gz", "rb").read()
    assert gzip.decompress(z1) == r1 and gzip.decompress(z2) == r2
    for z in (z1, z2):
        blocks = bgzf_blocks(z)
        assert blocks[-1] == (28, 0) and all(b <= 65536 and 0 < i <= 0xff00 for b, i in blocks[:-1])
    assert len(z1) < 0.40 * len(r1)
    # level 9 with the default engine is zlib on the host: smaller
    J.illumina(g, pre + "9", n_reads, 150, True, seed=62, ctx=ctx, compress=9, n_threads=4)
    z9 = open(pre + "9_R1.fq.gz", "rb").read()
    assert gzip.decompress(z9) == r1 and len(z9) < len(z1)
    # memory sink
    m1, m2, st = J.illumina(g, "", n_reads, 150, True, seed=62, ctx=ctx, sink="memory", compress=True, comp_engine="device",
                            batch_pairs=9000)
    assert gzip.decompress(m1) == r1 and gzip.decompress(m2) == r2
    assert st["z_bytes"][0] + 28 == len(m1) and st["d2h_bytes"] == st["z_bytes"][0] + st["z_bytes"][1]
    # stream sink, single-end
    s1, _, _ = J.illumina(g, "", 5000, 100, False, seed=63, ctx=ctx, sink="memory")
    got = []
    J.illumina(g, "", 5000, 100, False, seed=63, ctx=ctx, sink=lambda job, end, buf: got.append(bytes(buf)),
               compress=1, comp_engine="device", batch_pairs=1200)
    assert gzip.decompress(b"".join(got)) == s1
    # without comp_engine="device" the memory sink stays plain
    p1, _, _ = J.illumina(g, "", 5000, 100, False, seed=63, ctx=ctx, sink="memory", compress=True)
    assert p1 == s1


def test_generator_compressed_sep_files(ctx, tmp_path):
    g = small_genome(64)
    haps = J.random_haplotypes(g, 3, seed=65)
    pre = str(tmp_path / "zh")
    J.illumina(haps, pre, 9000, 100, True, seed=66, ctx=ctx, compress=True, sep_files=True, n_threads=2, haplotype_probs=[.5, .3, .2])
    m1, m2, _ = J.illumina(haps, "", 9000, 100, True, seed=66, ctx=ctx, sep_files=True, sink="memory", haplotype_probs=[.5, .3, .2])
    assert b"".join(gzip.decompress(open("%s_%s_R1.fq.gz" % (pre, h), "rb").read()) for h in haps.hap_names) == m1
    assert b"".join(gzip.decompress(open("%s_%s_R2.fq.gz" % (pre, h), "rb").read()) for h in haps.hap_names) == m2
