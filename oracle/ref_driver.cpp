// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI driver around the UNMODIFIED
// reference implementation of the Illumina hot path.
//
// This translation unit #includes /root/reference/src/hts_illumina.cpp where
// it lies (via -I/root/reference/src) so that the member templates defined in
// that .cpp (IlluminaOneGenome<T>::re_read<U>, ...) are visible here; it is
// linked with /root/reference/src/hap_classes.cpp compiled from the same tree.
// No reference source is copied into this repository.  Built only by
// oracle/Makefile into oracle/_ref/ (git-ignored).  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
// the resulting libraries; the product (jackalope_b200/) never does.
//
// Two builds of this file exist:
//   libjlp_ref.so         real vendored pcg64  -> end-to-end reference runs
//   libjlp_ref_replay.so  scripted pcg64 (stubs_replay/) + -DJLP_REPLAY
//                         -> the reference consumes a draw stream we supply
// Both are compiled with -fno-access-control so the driver can reach
// protected members (constr_info, just_indels, append_pools) and
// AliasSampler's private tables without touching the sources.

#include "hts_illumina.cpp"  // resolved through -I/root/reference/src

#include <cstring>
#include <map>
#include <memory>

namespace {

typedef std::vector<std::vector<std::vector<double>>> ProbT;
typedef std::vector<std::vector<std::vector<uint8>>> QualT;

// Flattened profile: nq[nt*L+pos] entries per (nt,pos), values concatenated in
// (nt,pos) order.  nt order is T,C,A,G (R/hts_illumina.R:143-186).
void unflatten_profile(uint64_t L, const uint32_t* nq, const double* probs,
                       const uint8_t* quals, ProbT& P, QualT& Q) {
    P.assign(4, std::vector<std::vector<double>>(L));
    Q.assign(4, std::vector<std::vector<uint8>>(L));
    uint64_t off = 0;
    for (uint64_t nt = 0; nt < 4; nt++) {
        for (uint64_t pos = 0; pos < L; pos++) {
            uint32_t n = nq[nt * L + pos];
            P[nt][pos].assign(probs + off, probs + off + n);
            Q[nt][pos].assign(quals + off, quals + off + n);
            off += n;
        }
    }
}

int fail(char* err, size_t cap, const char* msg) {
    if (err && cap) {
        std::strncpy(err, msg, cap - 1);
        err[cap - 1] = '\0';
    }
    return -1;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- genomes ---

void* jref_genome_new(uint64_t n_chroms, const char* const* seqs,
                      const uint64_t* lens, const char* const* names) {
    RefGenome* g = new RefGenome();
    for (uint64_t i = 0; i < n_chroms; i++) {
        g->chromosomes.push_back(RefChrom(std::string(names[i]),
                                          std::string(seqs[i], lens[i])));
        g->total_size += lens[i];
    }
    return g;
}
void jref_genome_free(void* g) { delete static_cast<RefGenome*>(g); }

void* jref_hapset_new(void* ref, uint64_t n_haps, const char* const* names) {
    std::vector<std::string> nm;
    for (uint64_t i = 0; i < n_haps; i++) nm.push_back(names[i]);
    return new HapSet(*static_cast<RefGenome*>(ref), nm);
}
void jref_hapset_free(void* hs) { delete static_cast<HapSet*>(hs); }

// Editing through the reference's own HapChrom::add_* (src/hap_classes.cpp:298-509);
// 0-based haplotype coordinates, as in src/ref_hap_access.cpp:816-866.
int jref_add_sub(void* hs, uint64_t hap, uint64_t chrom, char nt, uint64_t pos) {
    try { (*static_cast<HapSet*>(hs))[hap][chrom].add_substitution(nt, pos); }
    catch (...) { return -1; }
    return 0;
}
int jref_add_ins(void* hs, uint64_t hap, uint64_t chrom, const char* nts, uint64_t pos) {
    try { (*static_cast<HapSet*>(hs))[hap][chrom].add_insertion(std::string(nts), pos); }
    catch (...) { return -1; }
    return 0;
}
int jref_add_del(void* hs, uint64_t hap, uint64_t chrom, uint64_t size, uint64_t pos) {
    try { (*static_cast<HapSet*>(hs))[hap][chrom].add_deletion(size, pos); }
    catch (...) { return -1; }
    return 0;
}
// n edits of one haplotype chromosome in one call, in the order given (kind 0 substitution: payload 1 base;
// 1 insertion: payload the inserted bases; 2 deletion of `size`); pos = 0-based haplotype coordinate at edit time
int jref_add_edits(void* hs, uint64_t hap, uint64_t chrom, uint64_t n, const uint8_t* kind, const uint64_t* pos,
                   const uint64_t* size, const uint64_t* pay_off, const char* payload) {
    try {
        HapChrom& hc = (*static_cast<HapSet*>(hs))[hap][chrom];
        for (uint64_t i = 0; i < n; i++) {
            if (kind[i] == 0) hc.add_substitution(payload[pay_off[i]], pos[i]);
            else if (kind[i] == 1) hc.add_insertion(std::string(payload + pay_off[i], size[i]), pos[i]);
            else hc.add_deletion(size[i], pos[i]);
        }
    } catch (...) { return -1; }
    return 0;
}
uint64_t jref_hap_chrom_size(void* hs, uint64_t hap, uint64_t chrom) {
    return (*static_cast<HapSet*>(hs))[hap][chrom].size();
}
// HapChrom::get_chrom_full (src/hap_classes.cpp:80-116)
int jref_hap_chrom_full(void* hs, uint64_t hap, uint64_t chrom, char* out, uint64_t cap) {
    std::string s = (*static_cast<HapSet*>(hs))[hap][chrom].get_chrom_full();
    if (s.size() > cap) return -1;
    std::memcpy(out, s.data(), s.size());
    return 0;
}
uint64_t jref_hap_n_muts(void* hs, uint64_t hap, uint64_t chrom) {
    return (*static_cast<HapSet*>(hs))[hap][chrom].mutations.size();
}
uint64_t jref_hap_nuc_bytes(void* hs, uint64_t hap, uint64_t chrom) {
    const AllMutations& m = (*static_cast<HapSet*>(hs))[hap][chrom].mutations;
    uint64_t n = 0;
    for (uint64_t i = 0; i < m.size(); i++)
        if (m.nucleos[i] != nullptr) n += std::strlen(m.nucleos[i]);
    return n;
}
// Export the AllMutations SoA (src/hap_classes.h:100-104) as flat arrays.
int jref_hap_get_muts(void* hs, uint64_t hap, uint64_t chrom, uint64_t* old_pos,
                      uint64_t* new_pos, uint64_t* nuc_off, uint32_t* nuc_len,
                      char* pool) {
    const AllMutations& m = (*static_cast<HapSet*>(hs))[hap][chrom].mutations;
    uint64_t off = 0;
    for (uint64_t i = 0; i < m.size(); i++) {
        old_pos[i] = m.old_pos[i];
        new_pos[i] = m.new_pos[i];
        nuc_off[i] = off;
        uint32_t n = m.nucleos[i] ? static_cast<uint32_t>(std::strlen(m.nucleos[i])) : 0;
        nuc_len[i] = n;
        if (n) std::memcpy(pool + off, m.nucleos[i], n);
        off += n;
    }
    return 0;
}

// ------------------------------------------------------------ primitives ---

// AliasSampler::construct (src/alias_sampler.h:68-106), tables read back
// through -fno-access-control.
int jref_alias_build(const double* probs, uint64_t n, double* Prob, uint64_t* Alias) {
    AliasSampler s(std::vector<double>(probs, probs + n));
    for (uint64_t i = 0; i < n; i++) {
        Prob[i] = s.Prob[i];
        Alias[i] = s.Alias[i];
    }
    return 0;
}

// qual_prob_map of IlluminaQualityError (src/hts_illumina.h:182-187)
int jref_qual_prob_map(uint64_t L, const uint32_t* nq, const double* probs,
                       const uint8_t* quals, double* out, uint64_t cap) {
    ProbT P; QualT Q;
    unflatten_profile(L, nq, probs, quals, P, Q);
    IlluminaQualityError qe(P, Q);
    if (qe.qual_prob_map.size() > cap) return -1;
    for (size_t i = 0; i < qe.qual_prob_map.size(); i++) out[i] = qe.qual_prob_map[i];
    return static_cast<int>(qe.qual_prob_map.size());
}

void jref_rev_comp(char* s, uint64_t n) {
    std::string x(s, n);
    rev_comp(x);  // src/str_manip.h:214-229
    std::memcpy(s, x.data(), n);
}

// reads_per_group (src/hts.h:58-103); engine seeded from the stub R RNG.
int jref_reads_per_group(uint64_t n_reads, const double* probs, uint64_t n, uint64_t* out) {
    std::vector<uint64> r = reads_per_group(n_reads, std::vector<double>(probs, probs + n));
    for (uint64_t i = 0; i < n; i++) out[i] = r[i];
    return 0;
}

void jref_set_r_seed(uint64_t s) { jlp_stub::r_rng_state() = s; }

// The literal long-double expressions of src/pcg.h:99-101 and their call sites,
// exposed so the oracle's integer restatements can be checked against them.
//   kind 0: (uint64)(runif_01 * n)         long double product, n integral
//           (src/alias_sampler.h:55, src/hts_illumina.h:216,254)
//   kind 1: double u = runif_01; u < p     (src/alias_sampler.h:57-58, src/hts_illumina.h:251-252, src/hts.h:265-266)
//   kind 2: double u = runif_01; u > p     (src/hts_illumina.cpp:132-133)
//   kind 3: runif_01 < p (long double cmp) (src/hts_illumina.cpp:352)
//   kind 4: (uint8)(runif_01 * 10 + 33)    (src/hts_illumina.h:238)
//   kind 5: double u = runif_01; (uint64)(u * n)   (src/hts_illumina.cpp:215-216)
uint64_t jref_unif_expr(int kind, uint64_t x, double p, uint64_t n) {
    const long double max64 = static_cast<long double>(~0ULL);
    long double ul = (static_cast<long double>(x) + 1) / (max64 + 2);
    switch (kind) {
    case 0: return static_cast<uint64>(ul * n);
    case 1: { double u = ul; return u < p; }
    case 2: { double u = ul; return u > p; }
    case 3: return ul < p;
    case 4: { uint8 q = ul * 10 + static_cast<uint8>('!'); return q; }
    case 5: { double u = ul; return static_cast<uint64>(u * n); }
    }
    return 0;
}

#ifndef JLP_REPLAY
// ------------------------------------------------- end-to-end reference ---

int jref_illumina_ref(void* ref, int paired, int matepair, const char* out_prefix,
                      uint64_t n_reads, double prob_dup, uint64_t n_threads,
                      uint64_t read_pool_size, double shape, double scale,
                      uint64_t frag_len_min, uint64_t frag_len_max,
                      uint64_t L, const uint32_t* nq1, const double* probs1,
                      const uint8_t* quals1, double ins1, double del1,
                      const uint32_t* nq2, const double* probs2,
                      const uint8_t* quals2, double ins2, double del2,
                      const char* barcode, char* err, uint64_t errcap) {
    try {
        ProbT P1, P2; QualT Q1, Q2;
        unflatten_profile(L, nq1, probs1, quals1, P1, Q1);
        if (paired) unflatten_profile(L, nq2, probs2, quals2, P2, Q2);
        else { P2 = ProbT(1, std::vector<std::vector<double>>(1)); Q2 = QualT(1, std::vector<std::vector<uint8>>(1)); }
        std::vector<std::string> barcodes(1, std::string(barcode ? barcode : ""));
        illumina_ref_cpp(ref, paired, matepair, out_prefix, 0, "bgzip", n_reads, prob_dup,
                         n_threads, false, read_pool_size, shape, scale,
                         frag_len_min, frag_len_max, P1, Q1, ins1, del1,
                         P2, Q2, ins2, del2, barcodes);
    } catch (std::exception& e) {
        return fail(err, errcap, e.what());
    }
    return 0;
}

int jref_illumina_hap(void* hs, int paired, int matepair, const char* out_prefix,
                      int sep_files, uint64_t n_reads, double prob_dup,
                      uint64_t n_threads, uint64_t read_pool_size,
                      const double* hap_probs, double shape, double scale,
                      uint64_t frag_len_min, uint64_t frag_len_max,
                      uint64_t L, const uint32_t* nq1, const double* probs1,
                      const uint8_t* quals1, double ins1, double del1,
                      const uint32_t* nq2, const double* probs2,
                      const uint8_t* quals2, double ins2, double del2,
                      const char* const* barcodes_, char* err, uint64_t errcap) {
    try {
        HapSet* h = static_cast<HapSet*>(hs);
        ProbT P1, P2; QualT Q1, Q2;
        unflatten_profile(L, nq1, probs1, quals1, P1, Q1);
        if (paired) unflatten_profile(L, nq2, probs2, quals2, P2, Q2);
        else { P2 = ProbT(1, std::vector<std::vector<double>>(1)); Q2 = QualT(1, std::vector<std::vector<uint8>>(1)); }
        std::vector<std::string> barcodes;
        for (uint64_t i = 0; i < h->size(); i++)
            barcodes.push_back(barcodes_ && barcodes_[i] ? barcodes_[i] : "");
        std::vector<double> hp(hap_probs, hap_probs + h->size());
        illumina_hap_cpp(hs, paired, matepair, out_prefix, sep_files, 0, "bgzip", n_reads,
                         prob_dup, n_threads, false, read_pool_size, hp, shape, scale,
                         frag_len_min, frag_len_max, P1, Q1, ins1, del1,
                         P2, Q2, ins2, del2, barcodes);
    } catch (std::exception& e) {
        return fail(err, errcap, e.what());
    }
    return 0;
}

int jref_is_replay() { return 0; }

#else  // JLP_REPLAY
// ----------------------------------------------------------------- replay ---
//
// Drive the unmodified read model on fragments chosen by the harness, with
// every pcg64 output taken from `script`.  For each instance i the driver sets
// (chrom_ind, frag_len, frag_start) -- the three values chrom_indels_frag /
// indels_frag (src/hts_illumina.cpp:191-265) would have produced -- and then
// runs exactly what re_read does (src/hts_illumina.cpp:80-110):
//   just_indels(eng)  = sample_indels + adjust_chrom_spaces
//   append_pools(...) = strand, fill_read/rev_comp/barcode, fill_read_qual, fill_fq_lines
// which is also everything one_read does after the fragment is known.
// consumed[i] returns how many draws instance i took from the script.

int jref_replay(void* obj, int is_hap, int paired, int matepair,
                uint64_t L, const uint32_t* nq1, const double* probs1,
                const uint8_t* quals1, double ins1, double del1,
                const uint32_t* nq2, const double* probs2,
                const uint8_t* quals2, double ins2, double del2,
                const char* const* barcodes_, uint64_t n_inst,
                const uint64_t* hap, const uint64_t* chrom,
                const uint64_t* frag_len, const uint64_t* frag_start,
                const uint64_t* script, uint64_t script_len, uint64_t* consumed,
                char* out1, uint64_t cap1, uint64_t* len1,
                char* out2, uint64_t cap2, uint64_t* len2,
                char* err, uint64_t errcap) {
    try {
        ProbT P1, P2; QualT Q1, Q2;
        unflatten_profile(L, nq1, probs1, quals1, P1, Q1);
        if (paired) unflatten_profile(L, nq2, probs2, quals2, P2, Q2);

        jlp_replay::Script& sc = jlp_replay::script();
        sc.data = script; sc.len = script_len; sc.pos = 0; sc.underruns = 0;
        pcg64 eng;
        std::vector<std::vector<char>> pools(paired ? 2 : 1);
        bool finished = false;
        const uint64_t huge = ~0ULL >> 1;

        if (!is_hap) {
            RefGenome* g = static_cast<RefGenome*>(obj);
            std::string bc(barcodes_ && barcodes_[0] ? barcodes_[0] : "");
            std::unique_ptr<IlluminaReference> rd;
            if (paired) rd.reset(new IlluminaReference(*g, matepair, 16.0, 25.0, 1, huge,
                                                       P1, Q1, ins1, del1, P2, Q2, ins2, del2, bc));
            else rd.reset(new IlluminaReference(*g, 16.0, 25.0, 1, huge, P1, Q1, ins1, del1, bc));
            rd->chrom_reads.assign(g->size(), huge);
            for (uint64_t i = 0; i < n_inst; i++) {
                size_t before = sc.pos;
                rd->constr_info.chrom_ind = chrom[i];
                rd->constr_info.frag_len = frag_len[i];
                rd->constr_info.frag_start = frag_start[i];
                rd->re_read<std::vector<char>>(pools, finished, eng);
                consumed[i] = sc.pos - before;
            }
        } else {
            HapSet* h = static_cast<HapSet*>(obj);
            std::map<uint64_t, std::unique_ptr<IlluminaOneHaplotype>> makers;
            std::map<std::pair<uint64_t, uint64_t>, std::string> seqs;
            for (uint64_t i = 0; i < n_inst; i++) {
                size_t before = sc.pos;
                uint64_t hi = hap[i], ci = chrom[i];
                if (!makers.count(hi)) {
                    std::string bc(barcodes_ && barcodes_[hi] ? barcodes_[hi] : "");
                    if (paired) makers[hi].reset(new IlluminaOneHaplotype(
                        (*h)[hi], matepair, 16.0, 25.0, 1, huge,
                        P1, Q1, ins1, del1, P2, Q2, ins2, del2, bc));
                    else makers[hi].reset(new IlluminaOneHaplotype(
                        (*h)[hi], 16.0, 25.0, 1, huge, P1, Q1, ins1, del1, bc));
                }
                std::pair<uint64_t, uint64_t> key(hi, ci);
                // IlluminaHaplotypes::one_read materialises the same way (src/hts_illumina.cpp:527)
                if (!seqs.count(key)) seqs[key] = (*h)[hi][ci].get_chrom_full();
                IlluminaOneHaplotype& rd = *makers[hi];
                rd.constr_info.frag_len = frag_len[i];
                rd.constr_info.frag_start = frag_start[i];
                rd.re_read<std::vector<char>>(seqs[key], ci, pools, eng);
                consumed[i] = sc.pos - before;
            }
        }
        if (sc.underruns) return fail(err, errcap, "replay script underrun");
        *len1 = pools[0].size();
        if (pools[0].size() > cap1) return fail(err, errcap, "out1 too small");
        std::memcpy(out1, pools[0].data(), pools[0].size());
        if (paired) {
            *len2 = pools[1].size();
            if (pools[1].size() > cap2) return fail(err, errcap, "out2 too small");
            std::memcpy(out2, pools[1].data(), pools[1].size());
        } else if (len2) *len2 = 0;
    } catch (std::exception& e) {
        return fail(err, errcap, e.what());
    }
    return 0;
}

// The inner loop of create_chromosomes_ (src/create_sequences.cpp:129-132) on a scripted engine:
// chrom.push_back(bases_[sampler.sample(engine)]) with the reference's own AliasSampler.
int jref_create_chrom_replay(const double* pi_tcag, uint64_t len, const uint64_t* script, uint64_t script_len,
                             char* out, uint64_t* consumed) {
    jlp_replay::Script& sc = jlp_replay::script();
    sc.data = script; sc.len = script_len; sc.pos = 0; sc.underruns = 0;
    pcg64 engine;
    const AliasSampler sampler(std::vector<double>(pi_tcag, pi_tcag + 4));
    std::string bases_ = jlp::bases;
    std::string chrom;
    chrom.reserve(len);
    for (uint64_t j = 0; j < len; j++) {
        uint64 k = sampler.sample(engine);
        chrom.push_back(bases_[k]);
    }
    std::memcpy(out, chrom.data(), len);
    if (consumed) *consumed = sc.pos;
    return sc.underruns ? -1 : 0;
}

int jref_is_replay() { return 1; }
#endif

}  // extern "C"
