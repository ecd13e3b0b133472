"""PacBio read model: the oracle's C restatement against the UNMODIFIED reference.

CPU-only.  The CUDA path for PacBio is not built yet; this pins the checker it will be compared with.  Under
replay the reference's own PacBioQualityError::sample (update_probs with its truncated normals, fill_quals, the
insertion / deletion / substitution walk) and PacBioOneGenome::append_pool consume the oracle's draw ledger and
must emit byte-identical FASTQ while consuming exactly the ledger's draw count per read.  Read length and the
pass split are injected (std::lognormal_distribution / std::chi_squared_distribution keep cached normals)."""
import numpy as np
import pytest

from oracle import harness as H
from oracle import harness_pacbio as P

needs_ref = pytest.mark.skipif(not H.have_ref(True), reason="oracle/_ref not built (no /root/reference here)")


def genome(seed, lens, with_n=True):
    rng = np.random.default_rng(seed)
    seqs = [bytes(rng.choice(np.frombuffer(b"TCAG", np.uint8), n)) for n in lens]
    if with_n and lens[0] > 400:
        s = bytearray(seqs[0])
        s[100:130] = b"N" * 30
        s[300] = ord("x")
        seqs[0] = bytes(s)
    return ["chr%d" % i for i in range(len(lens))], seqs


def reads(seed, counts, lens, mean_len, max_passes=40.0):
    rng = np.random.default_rng(seed)
    n = int(np.sum(counts))
    chrom = np.repeat(np.arange(len(counts)), counts)
    rl = np.maximum(1, rng.normal(mean_len, mean_len / 4, n).astype(np.int64))
    passes = np.minimum(max_passes, 1 + rng.chisquare(3, n) * rng.uniform(0.01, 3, n))
    passes[rng.random(n) < 0.1] = np.floor(passes[rng.random(n) < 0.1][:0].sum() + 2)      # some whole numbers
    eff = np.minimum(rl, np.asarray(lens)[chrom])
    sp, pl, pr = zip(*(P.split_passes(float(p), int(e)) for p, e in zip(passes, eff)))
    return chrom, rl, np.array(sp), np.array(pl), np.array(pr)


@needs_ref
def test_min_exp_matches_reference():
    for args in [((0.5, 0.2247), (0, 0.2), 0.2, 0.11, 0.04, 0.01), ((0.5, 0.2247), (0, 0.2), 0.5, 0.3, 0.2, 0.1),
                 ((0.3, 0.1), (0.1, 0.3), 0.05, 0.01, 0.02, 0.005), ((0.5, 0.2247), (0, 0.2), 0.16, 0.11, 0.04, 0.01)]:
        assert P.min_exp(*args) == P.min_exp(*args, ref=True)


CASES = {
    "defaults": dict(lens=[6000, 3000, 9000], counts=[5, 3, 6], mean_len=1500),
    "reads_as_long_as_chromosomes": dict(lens=[700, 400, 900], counts=[6, 6, 6], mean_len=800),
    "high_error": dict(lens=[5000, 5000], counts=[5, 5], mean_len=1200, model=dict(ins_prob=0.3, del_prob=0.25, sub_prob=0.2)),
    # the truncation point lies more than 5 sd above the mean: trunc_norm takes its tail branch (two draws per iteration)
    "trunc_norm_tail_branch": dict(lens=[4000], counts=[8], mean_len=900, model=dict(norm_params=(-10.0, 0.1))),
    "tiny_reads": dict(lens=[50, 2000], counts=[4, 4], mean_len=3),
}


@needs_ref
@pytest.mark.parametrize("name", list(CASES))
def test_reference_replays_the_oracle_ledger(name):
    c = CASES[name]
    names, seqs = genome(1, c["lens"])
    chrom, rl, sp, pl, pr = reads(2, c["counts"], c["lens"], c["mean_len"])
    model = c.get("model", {})
    o = P.generate(names, seqs, "REF", c["counts"], rl, sp, pl, pr, seed=77, **model)
    fq, consumed = P.ref_replay(names, seqs, chrom, rl, sp, pl, pr, o["ledger"], **model)
    assert np.array_equal(consumed, o["ledger_cnt"]), (consumed[:8], o["ledger_cnt"][:8])
    assert int(consumed.sum()) == len(o["ledger"])
    assert fq == o["fastq"]
    if name == "trunc_norm_tail_branch":          # 2 + 2 draws for the two truncated normals instead of 1 + 1
        plain = P.generate(names, seqs, "REF", c["counts"], rl, sp, pl, pr, seed=77)
        assert (o["ledger_cnt"] >= 4 + rl.clip(max=4000) * 0).all() and o["fastq"] != plain["fastq"]
    if name == "reads_as_long_as_chromosomes":    # read_chrom_space == chrom_len happens: read_start 0, no start draw
        assert (o["plan"][:, 3] == np.asarray(c["lens"])[o["plan"][:, 0]]).any()
    # shape of the records: 4 lines, read and quality of read_length characters, two quality values split at split_pos
    lines = fq.split(b"\n")[:-1]
    assert len(lines) == 4 * len(rl)
    for i in range(len(rl)):
        eff = min(int(rl[i]), c["lens"][chrom[i]])
        assert lines[4 * i].startswith(b"@REF-chr%d-" % chrom[i]) and lines[4 * i][-2:] in (b"-F", b"-R")
        assert len(lines[4 * i + 1]) == eff == len(lines[4 * i + 3]) and lines[4 * i + 2] == b"+"
        q = lines[4 * i + 3]
        assert len(set(q[:sp[i]])) <= 1 and len(set(q[sp[i]:])) <= 1


@needs_ref
@pytest.mark.parametrize("lens,mean_len", [([6000, 3000, 9000], 1500), ([700, 400, 900], 800)])
def test_duplicates_replay(lens, mean_len):
    """prob_dup > 0: a duplicate (re_read) keeps its chain's chromosome, read length and read_start, draws a new pass
    split, walk, strand and edits; with chromosome-long reads deletions are given up until the template fits and a
    read that still does not fit is not written."""
    names, seqs = genome(3, lens)
    counts = [12, 12, 12]
    chrom, rl, sp, pl, pr = reads(4, counts, lens, mean_len)
    prob_dup, pool = 0.45, 7
    # the chains: who duplicates whom is a function of the dup draws; a duplicate takes its leader's chromosome and length
    from oracle.harness_pacbio import _orc
    import ctypes as C
    lib = _orc()
    lib.orc_pb_draw.restype = C.c_uint64
    lib.orc_pb_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    n = len(rl)
    leader = np.arange(n)
    for j in range(1, n):
        if j % pool != 0 and (lib.orc_pb_draw(77, j - 1, 0, 2, 0) + 1) / 2.0 ** 64 < prob_dup:
            leader[j] = leader[j - 1]
    is_dup = (leader != np.arange(n)).astype(np.int32)
    assert 5 < is_dup.sum() < n - 5
    rl = rl[leader]
    eff = np.minimum(rl, np.asarray(lens)[chrom[leader]])
    sp = np.array([min(int(s), int(e)) for s, e in zip(sp, eff)])
    o = P.generate(names, seqs, "REF", counts, rl, sp, pl, pr, seed=77, prob_dup=prob_dup, pool_reads=pool)
    fq, consumed = P.ref_replay(names, seqs, chrom[leader], rl, sp, pl, pr, o["ledger"], is_dup=is_dup)
    assert np.array_equal(consumed, o["ledger_cnt"])
    assert fq == o["fastq"]
    # duplicates share their leader's start
    ids = [l for l in fq.split(b"\n")[0::4] if l]
    written = o["plan"][:, 3] != np.uint64(2 ** 64 - 1)
    assert len(ids) == written.sum()
    starts = o["plan"][:, 2]
    assert all(starts[j] == starts[leader[j]] for j in range(n))
