// TEST INFRASTRUCTURE ONLY (oracle/): link-level test harness of the Rcpp glue.
//
// oracle/Makefile builds oracle/_ref/libjlp_glue.so from
//   * integration/hts_illumina_b200.cpp, integration/hts_pacbio_b200.cpp   (the glue a maintainer adds)
//   * the STOCK generated wrappers _jackalope_{illumina,pacbio}_{ref,hap}_cpp, extracted at build time from
//     /root/reference/src/RcppExports.cpp:109-243 into oracle/_ref/ (git-ignored; never copied into the repository)
//   * /root/reference/src/hap_classes.cpp (HapSet editing, to build inputs through the reference's own add_*)
//   * this driver
// against the package's own headers and the stub Rcpp headers of oracle/stubs, linked with
// jackalope_b200/libjlp_b200.so under -Wl,-z,defs: a glue export whose C++ prototype differs from the one the
// stock wrapper declares (src/RcppExports.cpp:110,143,178,211) is an undefined symbol and the build fails.
// The functions below call the wrappers the way R's .Call does -- one SEXP per argument (here: a pointer to the
// C++ value, see stubs/Rcpp.h) -- so tests/test_gpu_glue.py exercises wrapper -> glue -> C ABI -> CUDA -> files.
#include <Rcpp.h>

#include <cstring>
#include <string>
#include <vector>

#include "jackalope_types.h"
#include "ref_classes.h"
#include "hap_classes.h"

typedef std::vector<std::vector<std::vector<double>>> ProbT;
typedef std::vector<std::vector<std::vector<uint8>>> QualT;

// the stock wrappers (src/RcppExports.cpp:111,144,179,212): 24 / 26 / 24 / 26 SEXPs
extern "C" {
SEXP _jackalope_illumina_ref_cpp(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                                 SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP _jackalope_illumina_hap_cpp(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                                 SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP _jackalope_pacbio_ref_cpp(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                               SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
}

namespace {

void unflatten_profile(uint64_t L, const uint32_t* nq, const double* probs, const uint8_t* quals, ProbT& P, QualT& Q) {
    P.assign(4, std::vector<std::vector<double>>(L));
    Q.assign(4, std::vector<std::vector<uint8>>(L));
    uint64_t off = 0;
    for (uint64_t nt = 0; nt < 4; nt++)
        for (uint64_t pos = 0; pos < L; pos++) {
            const uint32_t n = nq[nt * L + pos];
            P[nt][pos].assign(probs + off, probs + off + n);
            Q[nt][pos].assign(quals + off, quals + off + n);
            off += n;
        }
}

int finish(SEXP r, char* err, uint64_t cap) {
    if (r != jlp_stub::r_error_token()) return 0;
    if (err && cap) {
        std::strncpy(err, jlp_stub::last_r_error().c_str(), cap - 1);
        err[cap - 1] = '\0';
    }
    return -1;
}

template <typename T> SEXP sx(const T& v) { return const_cast<void*>(static_cast<const void*>(&v)); }

}  // namespace

extern "C" {

void* jglue_genome_new(uint64_t n_chroms, const char* const* seqs, const uint64_t* lens, const char* const* names) {
    RefGenome* g = new RefGenome();
    for (uint64_t i = 0; i < n_chroms; i++) {
        g->chromosomes.push_back(RefChrom(std::string(names[i]), std::string(seqs[i], lens[i])));
        g->total_size += lens[i];
    }
    return g;
}
void jglue_genome_free(void* g) { delete static_cast<RefGenome*>(g); }

void* jglue_hapset_new(void* ref, uint64_t n_haps, const char* const* names) {
    std::vector<std::string> nm;
    for (uint64_t i = 0; i < n_haps; i++) nm.push_back(names[i]);
    return new HapSet(*static_cast<RefGenome*>(ref), nm);
}
void jglue_hapset_free(void* hs) { delete static_cast<HapSet*>(hs); }

// n edits of one haplotype chromosome through HapChrom::add_* (src/hap_classes.cpp:298-509), in the order given:
// kind 0 substitution (payload: 1 base), 1 insertion (payload: the inserted bases), 2 deletion (size);
// pos = 0-based haplotype coordinate at the time of the edit
int jglue_add_edits(void* hs, uint64_t hap, uint64_t chrom, uint64_t n, const uint8_t* kind, const uint64_t* pos,
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
uint64_t jglue_hap_chrom_size(void* hs, uint64_t hap, uint64_t chrom) { return (*static_cast<HapSet*>(hs))[hap][chrom].size(); }

void jglue_set_r_seed(uint64_t s) { jlp_stub::r_rng_state() = s; }
// the run seed the glue's seed_from_r() derives right after jglue_set_r_seed(s) (two Rcpp::runif(0, 2^32) values)
uint64_t jglue_seed_after(uint64_t s) {
    const uint64_t keep = jlp_stub::r_rng_state();
    jlp_stub::r_rng_state() = s;
    std::vector<uint64> v = Rcpp::as<std::vector<uint64>>(Rcpp::runif(2, 0, 4294967296.0));
    jlp_stub::r_rng_state() = keep;
    return (static_cast<uint64_t>(v[0]) << 32) | static_cast<uint64_t>(v[1]);
}

int jglue_illumina_ref(void* ref, int paired_, int matepair_, const char* out_prefix_, int compress_, const char* comp_method_,
                       uint64_t n_reads_, double prob_dup, uint64_t n_threads_, uint64_t read_pool_size_, double shape, double scale,
                       uint64_t frag_len_min_, uint64_t frag_len_max_, uint64_t L, const uint32_t* nq1, const double* probs1,
                       const uint8_t* quals1, double ins1, double del1, const uint32_t* nq2, const double* probs2, const uint8_t* quals2,
                       double ins2, double del2, const char* barcode, char* err, uint64_t errcap) {
    ProbT P1, P2; QualT Q1, Q2;
    unflatten_profile(L, nq1, probs1, quals1, P1, Q1);
    if (paired_) unflatten_profile(L, nq2, probs2, quals2, P2, Q2);
    else { P2 = ProbT(1, std::vector<std::vector<double>>(1)); Q2 = QualT(1, std::vector<std::vector<uint8>>(1)); }   // R/hts_illumina.R:689-690
    const bool paired = paired_, matepair = matepair_, show_progress = false;
    const std::string out_prefix(out_prefix_), comp_method(comp_method_);
    const int compress = compress_;
    const uint64 n_reads = n_reads_, n_threads = n_threads_, read_pool_size = read_pool_size_, frag_len_min = frag_len_min_,
                 frag_len_max = frag_len_max_;
    const std::vector<std::string> barcodes(1, std::string(barcode ? barcode : ""));
    SEXP r = _jackalope_illumina_ref_cpp(ref, sx(paired), sx(matepair), sx(out_prefix), sx(compress), sx(comp_method), sx(n_reads),
                                         sx(prob_dup), sx(n_threads), sx(show_progress), sx(read_pool_size), sx(shape), sx(scale),
                                         sx(frag_len_min), sx(frag_len_max), sx(P1), sx(Q1), sx(ins1), sx(del1), sx(P2), sx(Q2),
                                         sx(ins2), sx(del2), sx(barcodes));
    return finish(r, err, errcap);
}

int jglue_illumina_hap(void* hs, int paired_, int matepair_, const char* out_prefix_, int sep_files_, int compress_,
                       const char* comp_method_, uint64_t n_reads_, double prob_dup, uint64_t n_threads_, uint64_t read_pool_size_,
                       const double* hap_probs, double shape, double scale, uint64_t frag_len_min_, uint64_t frag_len_max_, uint64_t L,
                       const uint32_t* nq1, const double* probs1, const uint8_t* quals1, double ins1, double del1, const uint32_t* nq2,
                       const double* probs2, const uint8_t* quals2, double ins2, double del2, const char* const* barcodes_, char* err,
                       uint64_t errcap) {
    HapSet* h = static_cast<HapSet*>(hs);
    ProbT P1, P2; QualT Q1, Q2;
    unflatten_profile(L, nq1, probs1, quals1, P1, Q1);
    if (paired_) unflatten_profile(L, nq2, probs2, quals2, P2, Q2);
    else { P2 = ProbT(1, std::vector<std::vector<double>>(1)); Q2 = QualT(1, std::vector<std::vector<uint8>>(1)); }
    const bool paired = paired_, matepair = matepair_, show_progress = false, sep_files = sep_files_;
    const std::string out_prefix(out_prefix_), comp_method(comp_method_);
    const int compress = compress_;
    const uint64 n_reads = n_reads_, n_threads = n_threads_, read_pool_size = read_pool_size_, frag_len_min = frag_len_min_,
                 frag_len_max = frag_len_max_;
    std::vector<std::string> barcodes;
    for (uint64_t i = 0; i < h->size(); i++) barcodes.push_back(barcodes_ && barcodes_[i] ? barcodes_[i] : "");
    const std::vector<double> hp(hap_probs, hap_probs + h->size());
    SEXP r = _jackalope_illumina_hap_cpp(hs, sx(paired), sx(matepair), sx(out_prefix), sx(sep_files), sx(compress), sx(comp_method),
                                         sx(n_reads), sx(prob_dup), sx(n_threads), sx(show_progress), sx(read_pool_size), sx(hp),
                                         sx(shape), sx(scale), sx(frag_len_min), sx(frag_len_max), sx(P1), sx(Q1), sx(ins1), sx(del1),
                                         sx(P2), sx(Q2), sx(ins2), sx(del2), sx(barcodes));
    return finish(r, err, errcap);
}

// pacbio_ref_cpp through its stock wrapper with the defaults of pacbio() (R/hts_pacbio.R) except what is passed
int jglue_pacbio_ref(void* ref, const char* out_prefix_, uint64_t n_reads_, uint64_t n_threads_, uint64_t read_pool_size_, double prob_dup,
                     double scale, double sigma, double loc, double min_read_len, uint64_t max_passes_, const double* chi2_n,
                     const double* chi2_s, const double* sqrt_p, const double* norm_p, double prob_thresh, double prob_ins,
                     double prob_del, double prob_subst, char* err, uint64_t errcap) {
    const std::string out_prefix(out_prefix_), comp_method("bgzip");
    const int compress = 0;
    const bool show_progress = false;
    const uint64 n_reads = n_reads_, n_threads = n_threads_, read_pool_size = read_pool_size_, max_passes = max_passes_;
    const std::vector<double> read_probs, cn(chi2_n, chi2_n + 3), cs(chi2_s, chi2_s + 5), sq(sqrt_p, sqrt_p + 2), nm(norm_p, norm_p + 2);
    const std::vector<uint64> read_lens;
    SEXP r = _jackalope_pacbio_ref_cpp(ref, sx(out_prefix), sx(compress), sx(comp_method), sx(n_reads), sx(n_threads), sx(show_progress),
                                       sx(read_pool_size), sx(prob_dup), sx(scale), sx(sigma), sx(loc), sx(min_read_len), sx(read_probs),
                                       sx(read_lens), sx(max_passes), sx(cn), sx(cs), sx(sq), sx(nm), sx(prob_thresh), sx(prob_ins),
                                       sx(prob_del), sx(prob_subst));
    return finish(r, err, errcap);
}

}  // extern "C"
