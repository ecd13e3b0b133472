// Rcpp glue that makes jackalope's illumina() run on libjlp_b200.so.
//
// Drop this file into the package's src/ in place of the bodies of
// illumina_ref_cpp / illumina_hap_cpp (/root/reference/src/hts_illumina.cpp:589-739),
// add `-ljlp_b200 -lcudart` to PKG_LIBS in src/Makevars and put include/jlp_b200.h on the
// include path.  The two exported signatures -- and therefore R/RcppExports.R:96-107,
// src/RcppExports.cpp:111-175 and the R function illumina() -- stay exactly as they are.
// R, Rcpp and the package's build chain are absent from the image this repository was
// built in.  What is checked there: this file is compiled TOGETHER with the stock wrappers of
// src/RcppExports.cpp:109-243 (extracted at build time, never copied into the repository), the package's own
// ref_classes.h / hap_classes.h and stub Rcpp headers into oracle/_ref/libjlp_glue.so, linked with
// libjlp_b200.so (oracle/Makefile, oracle/glue_driver.cpp); tests/test_gpu_glue.py then calls
// _jackalope_illumina_{ref,hap}_cpp the way .Call does and byte-compares the files with the oracle.
//
// Only jackalope's own types are read, through const access, as the original code does
// (XPtr<RefGenome> / XPtr<HapSet> are borrowed, never freed: src/hts_illumina.h:307,513).

#include <RcppArmadillo.h>
#include <progress.hpp>

#include <string>
#include <vector>

#include "jlp_b200_glue.h"

using namespace Rcpp;
using namespace jlp_glue;

namespace {

typedef std::vector<std::vector<std::vector<double>>> QualProbs;
typedef std::vector<std::vector<std::vector<uint8>>> Quals;

// nested [nt in T,C,A,G][pos][k] -> flat arrays (R/hts_illumina.R:143-186 builds the nesting)
void set_profile(Ctx& ctx, int end, const QualProbs& probs, const Quals& quals) {
    if (probs.size() != 4 || quals.size() != 4)
        stop("In IlluminaQualityError construct, probs and quals must be of length 4.");   // src/hts_illumina.h:161-164
    std::vector<uint32_t> nq;
    std::vector<double> p;
    std::vector<uint8_t> q;
    for (size_t nt = 0; nt < 4; nt++)
        for (size_t pos = 0; pos < probs[nt].size(); pos++) {
            nq.push_back(static_cast<uint32_t>(probs[nt][pos].size()));
            p.insert(p.end(), probs[nt][pos].begin(), probs[nt][pos].end());
            q.insert(q.end(), quals[nt][pos].begin(), quals[nt][pos].end());
        }
    ctx.check(jlp_set_profile(ctx.p, end, probs[0].size(), nq.data(), p.data(), q.data()));
}

int abort_cb(void* prog) { return static_cast<Progress*>(prog)->check_abort() ? 1 : 0; }      // src/hts.h:396-399
void progress_cb(void* prog, uint64_t reads) { static_cast<Progress*>(prog)->increment(reads); }  // src/hts.h:414

}  // namespace

//[[Rcpp::export]]
void illumina_ref_cpp(SEXP ref_genome_ptr, const bool& paired, const bool& matepair, const std::string& out_prefix,
                      const int& compress, const std::string& comp_method, const uint64& n_reads, const double& prob_dup,
                      const uint64& n_threads, const bool& show_progress, const uint64& read_pool_size,
                      const double& frag_len_shape, const double& frag_len_scale, const uint64& frag_len_min,
                      const uint64& frag_len_max, const QualProbs& qual_probs1, const Quals& quals1,
                      const double& ins_prob1, const double& del_prob1, const QualProbs& qual_probs2, const Quals& quals2,
                      const double& ins_prob2, const double& del_prob2, const std::vector<std::string>& barcodes) {
    XPtr<RefGenome> ref_genome(ref_genome_ptr);
    Ctx ctx;
    set_genome(ctx, *ref_genome);
    set_profile(ctx, 0, qual_probs1, quals1);
    if (paired) set_profile(ctx, 1, qual_probs2, quals2);

    uint64 n_bar = n_reads;                                  // src/hts_illumina.cpp:639-640
    if (compress > 0 && n_threads > 1) n_bar += n_reads / 2;
    Progress prog_bar(n_bar, show_progress);
    const char* bc[1] = {barcodes.empty() ? "" : barcodes[0].c_str()};
    jlp_illumina_params P = {};
    P.paired = paired; P.matepair = matepair; P.out_prefix = out_prefix.c_str();   // expand_path: do it in R (path.expand)
    P.compress = compress; P.comp_method = comp_method.c_str();
    P.comp_engine = JLP_COMP_AUTO;   // levels 1-6: BGZF written by the GPU; 7-9: zlib on the writer threads
    P.n_reads = n_reads; P.prob_dup = prob_dup; P.n_threads = n_threads; P.read_pool_size = read_pool_size;
    P.frag_len_shape = frag_len_shape; P.frag_len_scale = frag_len_scale;
    P.frag_len_min = frag_len_min; P.frag_len_max = frag_len_max;
    P.ins_prob1 = ins_prob1; P.del_prob1 = del_prob1; P.ins_prob2 = ins_prob2; P.del_prob2 = del_prob2;
    P.barcodes = bc;
    P.seed = seed_from_r();
    P.abort_cb = abort_cb; P.progress_cb = progress_cb; P.cb_user = &prog_bar;
    int rc = jlp_illumina_ref(ctx.p, &P, nullptr);
    if (rc != JLP_OK && rc != JLP_ERR_ABORTED) ctx.check(rc);       // an interrupt just ends the run, files closed (src/hts.h:396-399)
}

//[[Rcpp::export]]
void illumina_hap_cpp(SEXP hap_set_ptr, const bool& paired, const bool& matepair, const std::string& out_prefix,
                      const bool& sep_files, const int& compress, const std::string& comp_method, const uint64& n_reads,
                      const double& prob_dup, const uint64& n_threads, const bool& show_progress,
                      const uint64& read_pool_size, const std::vector<double>& haplotype_probs,
                      const double& frag_len_shape, const double& frag_len_scale, const uint64& frag_len_min,
                      const uint64& frag_len_max, const QualProbs& qual_probs1, const Quals& quals1,
                      const double& ins_prob1, const double& del_prob1, const QualProbs& qual_probs2, const Quals& quals2,
                      const double& ins_prob2, const double& del_prob2, const std::vector<std::string>& barcodes) {
    XPtr<HapSet> hap_set(hap_set_ptr);
    Ctx ctx;
    set_genome(ctx, *(hap_set->reference));
    for (uint64 h = 0; h < hap_set->size(); h++) add_haplotype(ctx, (*hap_set)[h]);
    set_profile(ctx, 0, qual_probs1, quals1);
    if (paired) set_profile(ctx, 1, qual_probs2, quals2);

    uint64 n_bar = n_reads;
    if (compress > 0 && n_threads > 1) n_bar += n_reads / 2;
    Progress prog_bar(n_bar, show_progress);
    std::vector<const char*> bc;
    for (const std::string& b : barcodes) bc.push_back(b.c_str());
    jlp_illumina_params P = {};
    P.paired = paired; P.matepair = matepair; P.out_prefix = out_prefix.c_str(); P.sep_files = sep_files;
    P.compress = compress; P.comp_method = comp_method.c_str();
    P.comp_engine = JLP_COMP_AUTO;   // levels 1-6: BGZF written by the GPU; 7-9: zlib on the writer threads
    P.n_reads = n_reads; P.prob_dup = prob_dup; P.n_threads = n_threads; P.read_pool_size = read_pool_size;
    P.haplotype_probs = haplotype_probs.data();
    P.frag_len_shape = frag_len_shape; P.frag_len_scale = frag_len_scale;
    P.frag_len_min = frag_len_min; P.frag_len_max = frag_len_max;
    P.ins_prob1 = ins_prob1; P.del_prob1 = del_prob1; P.ins_prob2 = ins_prob2; P.del_prob2 = del_prob2;
    P.barcodes = bc.data();
    P.seed = seed_from_r();
    P.abort_cb = abort_cb; P.progress_cb = progress_cb; P.cb_user = &prog_bar;
    int rc = jlp_illumina_hap(ctx.p, &P, nullptr);
    if (rc != JLP_OK && rc != JLP_ERR_ABORTED) ctx.check(rc);
}
