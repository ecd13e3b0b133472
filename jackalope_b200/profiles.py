"""ART-style Illumina quality profiles: discovery, parsing, formatting.

Host-side mirror of the R layer of the reference for this path (R is absent
from this image, so the R functions are restated in Python with the same
names, argument meaning and error text):

* ``builtin_illumina_profiles``  R/hts_illumina.R:9-42
* ``seq_sys_by_read_length``     R/hts_illumina.R:53-72
* ``find_profile_file``          R/hts_illumina.R:85-117
* ``format_profile``             R/hts_illumina.R:133-188
* ``read_profile``               R/hts_illumina.R:210-262

The built-in profiles are the ART data files jackalope ships under
inst/art_profiles/; here they live packed in ``data/art_profiles.bin``
(made by tools/import_art_profiles.py).
"""
from __future__ import annotations

import gzip
import os
import struct
from functools import lru_cache

import numpy as np

_DATA = os.path.join(os.path.dirname(__file__), "data", "art_profiles.bin")

NTS = ("T", "C", "A", "G")  # outer order of qual_probs / quals (R/hts_illumina.R:145)


class JackalopeError(ValueError):
    """Raised where the reference calls R's stop()."""


def builtin_illumina_profiles():
    """Rows (name, read_length, read, file_name, abbrev) ordered by read length
    (stable), as R/hts_illumina.R:9-42."""
    rows = [
        ("Genome Analyzer I", 36, 1, "EmpR36R1", "GA1"),
        ("Genome Analyzer I", 36, 2, "EmpR36R2", "GA1"),
        ("Genome Analyzer I", 44, 1, "EmpR44R1", "GA1"),
        ("Genome Analyzer I", 44, 2, "EmpR44R2", "GA1"),
        ("Genome Analyzer II", 50, 1, "EmpR50R1", "GA2"),
        ("Genome Analyzer II", 50, 2, "EmpR50R2", "GA2"),
        ("MiniSeq TruSeq", 50, 1, "MiniSeqTruSeqL50", "MinS"),
        ("Genome Analyzer II", 75, 1, "EmpR75R1", "GA2"),
        ("Genome Analyzer II", 75, 2, "EmpR75R2", "GA2"),
        ("NextSeq 500 v2", 75, 1, "NextSeq500v2L75R1", "NS50"),
        ("NextSeq 500 v2", 75, 2, "NextSeq500v2L75R2", "NS50"),
        ("HiSeq 1000", 100, 1, "Emp100R1", "HS10"),
        ("HiSeq 1000", 100, 2, "Emp100R2", "HS10"),
        ("HiSeq 2000", 100, 1, "HiSeq2000L100R1", "HS20"),
        ("HiSeq 2000", 100, 2, "HiSeq2000L100R2", "HS20"),
        ("HiSeq 2500", 125, 1, "HiSeq2500L125R1", "HS25"),
        ("HiSeq 2500", 125, 2, "HiSeq2500L125R2", "HS25"),
        ("HiSeq 2500", 150, 1, "HiSeq2500L150R1filter", "HS25"),
        ("HiSeq 2500", 150, 2, "HiSeq2500L150R2filter", "HS25"),
        ("HiSeqX v2.5 PCR free", 150, 1, "HiSeqXPCRfreeL150R1", "HSXn"),
        ("HiSeqX v2.5 PCR free", 150, 2, "HiSeqXPCRfreeL150R2", "HSXn"),
        ("HiSeqX v2.5 TruSeq", 150, 1, "HiSeqXtruSeqL150R1", "HSXt"),
        ("HiSeqX v2.5 TruSeq", 150, 2, "HiSeqXtruSeqL150R2", "HSXt"),
        ("MiSeq v1", 250, 1, "EmpMiSeq250R1", "MSv1"),
        ("MiSeq v1", 250, 2, "EmpMiSeq250R2", "MSv1"),
        ("MiSeq v3", 250, 1, "MiSeqv3L250R1", "MSv3"),
        ("MiSeq v3", 250, 2, "MiSeqv3L250R2", "MSv3"),
    ]
    rows.sort(key=lambda r: r[1])  # R's order() is stable
    return [dict(name=r[0], read_length=r[1], read=r[2], file_name=r[3], abbrev=r[4]) for r in rows]


def seq_sys_by_read_length(read_length):
    if read_length <= 44:
        return "GA1"
    if read_length <= 75:
        return "GA2"
    if read_length <= 100:
        return "HS20"
    if read_length <= 150:
        return "HS25"
    if read_length <= 250:
        return "MSv1"
    raise JackalopeError("\nNo built-in Illumina profile can generate reads of length %d." % read_length)


def find_profile_file(seq_sys, read_length, read):
    """Name of the built-in profile (the reference returns its path)."""
    df = builtin_illumina_profiles()
    c1 = [r for r in df if r["name"] == seq_sys or r["abbrev"] == seq_sys]
    if not c1:
        raise JackalopeError(
            "\nThe desired Illumina platform name isn't available. "
            "See printed data frame above for names and abbreviations of those available.")
    c2 = [r for r in c1 if r["read_length"] >= read_length]
    if not c2:
        raise JackalopeError(
            "\nFor the desired Illumina platform, this package doesn't have "
            "a read length that's as long as you want. "
            "See printed values above for lengths that are available.")
    c3 = [r for r in c2 if r["read"] == read]
    if not c3:
        raise JackalopeError(
            "\nFor the desired Illumina platform and read length, "
            "this package only has a profile for read number %d." % (2 if read == 1 else 1))
    return "builtin:" + c3[0]["file_name"]


@lru_cache(maxsize=1)
def _builtin_blob():
    with gzip.open(_DATA, "rb") as fh:
        blob = fh.read()
    if blob[:6] != b"JLPP1\n":
        raise RuntimeError("corrupt " + _DATA)
    out, p = {}, 6
    (n_prof,) = struct.unpack_from("<I", blob, p)
    p += 4
    for _ in range(n_prof):
        (ln,) = struct.unpack_from("<H", blob, p)
        p += 2
        name = blob[p:p + ln].decode()
        p += ln
        (n_pos,) = struct.unpack_from("<I", blob, p)
        p += 4
        info = []
        for nt in NTS:
            for pos in range(n_pos):
                (nq,) = struct.unpack_from("<H", blob, p)
                p += 2
                quals = np.frombuffer(blob, np.uint8, nq, p).astype(np.int64)
                p += nq
                counts = np.frombuffer(blob, "<u8", nq, p).astype(np.float64)
                p += 8 * nq
                info.append((nt, pos, quals, counts))
        out[name] = info
    return out


def _info_from_counts(nt, pos, quals, cum):
    """One profile position: cumulative counts -> probabilities
    (R/hts_illumina.R:233-237)."""
    probs = np.array(cum, dtype=np.float64)
    if probs.size > 1:
        probs = probs - np.concatenate(([0.0], probs[:-1]))
    probs = probs / probs.sum()
    return dict(nt=nt, pos=int(pos), quals=np.asarray(quals, dtype=np.int64), probs=probs)


def _parse_text_profile(path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        lines = [l.rstrip("\n").rstrip("\r") for l in fh]
    lines = [l for l in lines if l[:1] in NTS]  # grepl("^T|^C|^A|^G")
    fields = []
    for l in lines:
        f = l.split("\t")
        if f and f[-1] == "":  # strsplit() drops one trailing empty field
            f.pop()
        fields.append(f)
    info = []
    for i in range(0, len(fields), 2):
        if i + 1 >= len(fields):
            raise JackalopeError("\nInput profile file does not have proper format. "
                                 "The two lines specifying quality and distances should "
                                 "always have the same values for nucleotide and position.")
        a, b = fields[i], fields[i + 1]
        if a[:2] != b[:2]:
            raise JackalopeError("\nInput profile file does not have proper format. "
                                 "The two lines specifying quality and distances should "
                                 "always have the same values for nucleotide and position.")
        if len(a) != len(b):
            raise JackalopeError("\nInput profile file does not have proper format. "
                                 "The two lines specifying quality and distances should "
                                 "always have the same number of tab-delimited columns.")
        info.append(_info_from_counts(a[0], int(a[1]), [int(x) for x in a[2:]],
                                      [float(x) for x in b[2:]]))
    return info


def format_profile(profile_info, read_length):
    """Validate and trim to ``read_length`` positions; returns
    ``dict(qual_probs=[4][L] arrays, quals=[4][L] arrays)`` in T,C,A,G order."""
    pos_all = [x["pos"] for x in profile_info]
    if min(pos_all) > 0:
        raise JackalopeError("\nMinimum profile position should be zero.")
    if max(pos_all) < read_length - 1:
        raise JackalopeError("\nMaximum profile position should be >= read_length - 1.")
    qual_probs, quals = [], []
    for nt in NTS:
        sel = [x for x in profile_info if x["nt"] == nt]
        pos_nt = [x["pos"] for x in sel]
        if pos_nt != list(range(len(pos_nt))):
            raise JackalopeError("\nFor nucleotide %s in the profile, the positions "
                                 "aren't a vector from 0 to length(positions) - 1." % nt)
        if len(sel) < read_length:
            raise JackalopeError("\nFor nucleotide %s in the profile, it doesn't provide at "
                                 "least as many positions as your desired read length." % nt)
        if any(len(x["probs"]) != len(x["quals"]) for x in sel):
            raise JackalopeError("\nFor nucleotide %s in the profile, at least "
                                 "one of the positions has a number of qualities that doesn't "
                                 "match with the number of quality probabilities." % nt)
        sel = sel[:read_length]
        qual_probs.append([x["probs"] for x in sel])
        quals.append([x["quals"] for x in sel])
    return dict(qual_probs=qual_probs, quals=quals)


def read_profile(profile_fn, seq_sys, read_length, read):
    if profile_fn is not None and seq_sys is not None:
        raise JackalopeError("\nFor Illumina sequencing, the user should never provide both a custom "
                             "profile file and a sequencing system.")
    if profile_fn is None and seq_sys is None:
        seq_sys = seq_sys_by_read_length(read_length)
    if profile_fn is None:
        profile_fn = find_profile_file(seq_sys, read_length, read)
    if profile_fn.startswith("builtin:"):
        raw = _builtin_blob()[profile_fn[len("builtin:"):]]
        info = [_info_from_counts(nt, pos, q, c) for (nt, pos, q, c) in raw]
    else:
        info = _parse_text_profile(profile_fn)
    return format_profile(info, read_length)


def flatten_profile(prof):
    """[4][L][k] nested lists -> (L, nq uint32[4L], probs float64[], quals uint8[])
    in (nt, pos) order: the layout every C entry point of this repo takes.
    Qualities are reduced mod 256 as Rcpp's conversion to std::vector<uint8> does
    (the reference's own test profile uses quality 255)."""
    L = len(prof["qual_probs"][0])
    nq = np.zeros(4 * L, dtype=np.uint32)
    probs, quals = [], []
    for nt in range(4):
        if len(prof["qual_probs"][nt]) != L or len(prof["quals"][nt]) != L:
            raise JackalopeError("In IlluminaQualityError construct, all probs' lengths not equal")
        for pos in range(L):
            p = np.asarray(prof["qual_probs"][nt][pos], dtype=np.float64)
            q = np.asarray(prof["quals"][nt][pos], dtype=np.int64)
            nq[nt * L + pos] = p.size
            probs.append(p)
            quals.append((q & 0xFF).astype(np.uint8))
    return L, nq, np.ascontiguousarray(np.concatenate(probs)), np.ascontiguousarray(np.concatenate(quals))


_flat_cache = {}


def cached_flat_profile(profile_fn, seq_sys, read_length, read):
    """flatten_profile(read_profile(...)), memoised: parsing an ART profile costs ~10 ms of
    Python per call, which a multi-million-pairs-per-second run should pay once.  Files are keyed
    by path, size and mtime."""
    key = (profile_fn, seq_sys, int(read_length), int(read))
    if profile_fn is not None and not str(profile_fn).startswith("builtin:"):
        try:
            st = os.stat(profile_fn)
            key += (st.st_size, st.st_mtime_ns)
        except OSError:
            pass
    hit = _flat_cache.get(key)
    if hit is None:
        hit = flatten_profile(read_profile(profile_fn, seq_sys, read_length, read))
        if len(_flat_cache) > 64:
            _flat_cache.clear()
        _flat_cache[key] = hit
    return hit
