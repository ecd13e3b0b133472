"""Host-side genome data model handed to the C ABI.

``RefGenome`` mirrors the reference's RefGenome/RefChrom (name + nucleotide
string per chromosome, /root/reference/src/ref_classes.h:36-140) and
``Haplotypes`` its HapSet/HapGenome/HapChrom/AllMutations (sorted mutation
records ``old_pos``, ``new_pos``, ``nucleos`` per chromosome plus the mutated
size, src/hap_classes.h:100-104,287-296,540).  The objects only *hold* data;
they are flattened into plain arrays before crossing into the library.

Creating genomes and evolving haplotypes (create_genome, create_haplotypes) is
outside the hot path.  ``random_genome`` and
``random_haplotypes`` below exist so that tests and the benchmark have inputs
of the named shapes; they are input generators, not restatements.
"""
from __future__ import annotations

import numpy as np

from .profiles import JackalopeError


class RefGenome:
    """Chromosome names and sequences (bytes / uint8 arrays of ASCII)."""

    name = "REF"   # src/ref_classes.h:138

    def __init__(self, names, seqs):
        if len(names) != len(seqs) or len(names) == 0:
            raise JackalopeError("RefGenome needs one name per chromosome and at least one chromosome")
        self.names = [n.decode() if isinstance(n, bytes) else str(n) for n in names]
        self.seqs = [np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray))
                     else np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]

    def n_chroms(self):
        return len(self.seqs)

    def sizes(self):
        return np.array([s.size for s in self.seqs], dtype=np.uint64)

    def chrom(self, i) -> bytes:
        return self.seqs[i].tobytes()

    def flat(self):
        """(concatenated bases uint8[], offsets uint64[n+1])"""
        off = np.concatenate(([0], np.cumsum(self.sizes()))).astype(np.uint64)
        return np.ascontiguousarray(np.concatenate(self.seqs)), off


class HapChromMuts:
    """AllMutations of one haplotype chromosome as flat arrays."""

    __slots__ = ("old_pos", "new_pos", "nuc_off", "nuc_len", "pool", "chrom_size")

    def __init__(self, old_pos, new_pos, nuc_off, nuc_len, pool, chrom_size):
        self.old_pos = np.ascontiguousarray(old_pos, dtype=np.uint64)
        self.new_pos = np.ascontiguousarray(new_pos, dtype=np.uint64)
        self.nuc_off = np.ascontiguousarray(nuc_off, dtype=np.uint64)
        self.nuc_len = np.ascontiguousarray(nuc_len, dtype=np.uint32)
        self.pool = np.frombuffer(pool, dtype=np.uint8) if isinstance(pool, (bytes, bytearray)) \
            else np.ascontiguousarray(pool, dtype=np.uint8)
        self.chrom_size = int(chrom_size)

    @staticmethod
    def empty(ref_size):
        z = np.zeros(0, dtype=np.uint64)
        return HapChromMuts(z, z, z, np.zeros(0, dtype=np.uint32), b"", ref_size)


class Haplotypes:
    """A set of named haplotypes over one reference."""

    def __init__(self, reference: RefGenome, hap_names, muts):
        """muts[h][c] is a HapChromMuts."""
        self.reference = reference
        self.hap_names = [str(n) for n in hap_names]
        if len(muts) != len(self.hap_names) or any(len(m) != reference.n_chroms() for m in muts):
            raise JackalopeError("Haplotypes needs mutations for every haplotype and chromosome")
        self.muts = muts

    def n_haps(self):
        return len(self.hap_names)

    def n_chroms(self):
        return self.reference.n_chroms()

    def sizes(self, h):
        return np.array([m.chrom_size for m in self.muts[h]], dtype=np.uint64)


# ------------------------------------------------------------ input generators

def random_genome(n_chroms, chrom_len, seed=0, pi_tcag=(0.25, 0.25, 0.25, 0.25)):
    """Uniform random chromosomes named chrom0.. (the shape create_genome makes
    with len_sd = 0, R/create_genome.R:24-28, src/create_sequences.cpp:178-181)."""
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"TCAG", dtype=np.uint8)
    lens = chrom_len if np.ndim(chrom_len) else [int(chrom_len)] * n_chroms
    if tuple(pi_tcag) == (0.25, 0.25, 0.25, 0.25):
        seqs = [lut[rng.integers(0, 4, size=int(l), dtype=np.uint8)] for l in lens]
    else:
        seqs = [lut[rng.choice(4, size=int(l), p=np.asarray(pi_tcag) / np.sum(pi_tcag)).astype(np.uint8)] for l in lens]
    return RefGenome(["chrom%d" % i for i in range(n_chroms)], seqs)


def random_mutations(ref_seq: np.ndarray, rng, sub_rate, indel_rate, max_indel=10):
    """Sorted, non-overlapping substitutions / insertions / deletions on one
    chromosome, written directly in AllMutations form.  Indel sizes 1..max_indel
    with weights exp(-size) (the relative rates `indels()` uses, R/mevo.R:660-699).
    Returns (HapChromMuts, list of (kind, hap_pos, payload)) where the list replays
    the same edits through HapChrom::add_* in ascending order."""
    n = ref_seq.size
    n_sub = rng.binomial(n, sub_rate)
    n_ind = rng.binomial(n, indel_rate)
    m = min(n_sub + n_ind, max(0, n // (max_indel + 2)))
    if m == 0:
        return HapChromMuts.empty(n), []
    # sites spaced so that a deletion never reaches the next site
    slots = np.sort(rng.choice(n // (max_indel + 2), size=m, replace=False)) * (max_indel + 2)
    kinds = np.zeros(m, dtype=np.int8)
    ind_idx = rng.choice(m, size=min(n_ind, m), replace=False)
    kinds[ind_idx] = rng.integers(1, 3, size=ind_idx.size)      # 1 insertion, 2 deletion
    w = np.exp(-np.arange(1, max_indel + 1, dtype=np.float64))
    sizes = rng.choice(np.arange(1, max_indel + 1), size=m, p=w / w.sum())
    bases = np.frombuffer(b"TCAG", dtype=np.uint8)
    old_pos, new_pos, nuc_off, nuc_len, edits = [], [], [], [], []
    pool = bytearray()
    shift = 0
    for o, kind, sz in zip(slots.tolist(), kinds.tolist(), sizes.tolist()):
        old_pos.append(o)
        new_pos.append(o + shift)
        nuc_off.append(len(pool))
        if kind == 0:
            cur = b"TCAG".find(bytes([int(ref_seq[o])]))
            alt = bases[(cur + 1 + rng.integers(0, 3)) % 4] if cur >= 0 else bases[rng.integers(0, 4)]
            pool.append(int(alt))
            nuc_len.append(1)
            edits.append(("sub", o + shift, bytes([int(alt)])))
        elif kind == 1:
            ins = bases[rng.integers(0, 4, size=sz)].tobytes()
            pool.append(int(ref_seq[o]))
            pool += ins
            nuc_len.append(1 + sz)
            edits.append(("ins", o + shift, ins))
            shift += sz
        else:
            sz = min(sz, n - o)
            nuc_len.append(0)
            edits.append(("del", o + shift, sz))
            shift -= sz
    return HapChromMuts(old_pos, new_pos, nuc_off, nuc_len, bytes(pool), n + shift), edits


def random_haplotypes(reference: RefGenome, n_haps, sub_rate=0.01, indel_rate=0.001, seed=0, names=None,
                      return_edits=False):
    rng = np.random.default_rng(seed)
    muts, edits = [], []
    for _ in range(n_haps):
        mh, eh = [], []
        for seq in reference.seqs:
            mm, ee = random_mutations(seq, rng, sub_rate, indel_rate)
            mh.append(mm)
            eh.append(ee)
        muts.append(mh)
        edits.append(eh)
    haps = Haplotypes(reference, names or ["hap%d" % i for i in range(n_haps)], muts)
    return (haps, edits) if return_edits else haps
