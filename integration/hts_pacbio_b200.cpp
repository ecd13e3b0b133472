// Rcpp glue that makes jackalope's pacbio() run on libjlp_b200.so: the bodies of pacbio_ref_cpp / pacbio_hap_cpp
// (/root/reference/src/hts_pacbio.cpp) on the C ABI of include/jlp_b200.h; the exported signatures -- and with them
// R/RcppExports.R, src/RcppExports.cpp and the R function pacbio() (R/hts_pacbio.R) -- stay as they are.  Build notes
// as in hts_illumina_b200.cpp; type-checked only (no R in the image this repository was built in).

#include <RcppArmadillo.h>
#include <progress.hpp>

#include <string>
#include <vector>

#include "jlp_b200_glue.h"

using namespace Rcpp;
using namespace jlp_glue;

namespace {

int abort_cb(void* prog) { return static_cast<Progress*>(prog)->check_abort() ? 1 : 0; }      // src/hts.h:396-399
void progress_cb(void* prog, uint64_t reads) { static_cast<Progress*>(prog)->increment(reads); }  // src/hts.h:414

void fill_params(jlp_pacbio_params& P, const std::string& out_prefix, const int& compress, const std::string& comp_method,
                 const uint64& n_reads, const uint64& n_threads, const uint64& read_pool_size, const double& prob_dup,
                 const double& scale, const double& sigma, const double& loc, const double& min_read_len,
                 const std::vector<double>& read_probs, const std::vector<uint64_t>& read_lens, const uint64& max_passes,
                 const std::vector<double>& chi2_params_n, const std::vector<double>& chi2_params_s,
                 const std::vector<double>& sqrt_params, const std::vector<double>& norm_params, const double& prob_thresh,
                 const double& prob_ins, const double& prob_del, const double& prob_subst) {
    if (chi2_params_n.size() != 3 || chi2_params_s.size() != 5 || sqrt_params.size() != 2 || norm_params.size() != 2)
        stop("PacBio parameter vectors have the wrong lengths.");
    P.out_prefix = out_prefix.c_str();
    P.compress = compress; P.comp_method = comp_method.c_str(); P.comp_engine = JLP_COMP_AUTO;
    P.n_reads = n_reads; P.n_threads = n_threads; P.read_pool_size = read_pool_size; P.prob_dup = prob_dup;
    P.scale = scale; P.sigma = sigma; P.loc = loc; P.min_read_len = min_read_len;
    P.read_probs = read_probs.empty() ? nullptr : read_probs.data();
    P.read_lens = read_lens.empty() ? nullptr : read_lens.data();
    P.n_custom = read_probs.size();
    P.max_passes = max_passes;
    for (int i = 0; i < 3; i++) P.chi2_params_n[i] = chi2_params_n[i];
    for (int i = 0; i < 5; i++) P.chi2_params_s[i] = chi2_params_s[i];
    for (int i = 0; i < 2; i++) { P.sqrt_params[i] = sqrt_params[i]; P.norm_params[i] = norm_params[i]; }
    P.prob_thresh = prob_thresh; P.prob_ins = prob_ins; P.prob_del = prob_del; P.prob_subst = prob_subst;
    P.seed = seed_from_r();
}

}  // namespace

//[[Rcpp::export]]
void pacbio_ref_cpp(SEXP ref_genome_ptr, const std::string& out_prefix, const int& compress, const std::string& comp_method,
                    const uint64& n_reads, const uint64& n_threads, const bool& show_progress, const uint64& read_pool_size,
                    const double& prob_dup, const double& scale, const double& sigma, const double& loc,
                    const double& min_read_len, const std::vector<double>& read_probs, const std::vector<uint64>& read_lens,
                    const uint64& max_passes, const std::vector<double>& chi2_params_n, const std::vector<double>& chi2_params_s,
                    const std::vector<double>& sqrt_params, const std::vector<double>& norm_params, const double& prob_thresh,
                    const double& prob_ins, const double& prob_del, const double& prob_subst) {
    XPtr<RefGenome> ref_genome(ref_genome_ptr);
    Ctx ctx;
    set_genome(ctx, *ref_genome);
    Progress prog_bar(n_reads, show_progress);
    const std::vector<uint64_t> lens(read_lens.begin(), read_lens.end());
    jlp_pacbio_params P = {};
    fill_params(P, out_prefix, compress, comp_method, n_reads, n_threads, read_pool_size, prob_dup, scale, sigma, loc,
                min_read_len, read_probs, lens, max_passes, chi2_params_n, chi2_params_s, sqrt_params, norm_params, prob_thresh,
                prob_ins, prob_del, prob_subst);
    P.abort_cb = abort_cb; P.progress_cb = progress_cb; P.cb_user = &prog_bar;
    const int rc = jlp_pacbio(ctx.p, 0, &P, nullptr);
    if (rc != JLP_OK && rc != JLP_ERR_ABORTED) ctx.check(rc);       // an interrupt just ends the run (src/hts.h:396-399)
}

//[[Rcpp::export]]
void pacbio_hap_cpp(SEXP hap_set_ptr, const std::string& out_prefix, const bool& sep_files, const int& compress,
                    const std::string& comp_method, const uint64& n_reads, const uint64& n_threads, const bool& show_progress,
                    const uint64& read_pool_size, const std::vector<double>& haplotype_probs, const double& prob_dup,
                    const double& scale, const double& sigma, const double& loc, const double& min_read_len,
                    const std::vector<double>& read_probs, const std::vector<uint64>& read_lens, const uint64& max_passes,
                    const std::vector<double>& chi2_params_n, const std::vector<double>& chi2_params_s,
                    const std::vector<double>& sqrt_params, const std::vector<double>& norm_params, const double& prob_thresh,
                    const double& prob_ins, const double& prob_del, const double& prob_subst) {
    XPtr<HapSet> hap_set(hap_set_ptr);
    Ctx ctx;
    set_genome(ctx, *(hap_set->reference));
    for (uint64 h = 0; h < hap_set->size(); h++) add_haplotype(ctx, (*hap_set)[h]);
    Progress prog_bar(n_reads, show_progress);
    const std::vector<uint64_t> lens(read_lens.begin(), read_lens.end());
    jlp_pacbio_params P = {};
    fill_params(P, out_prefix, compress, comp_method, n_reads, n_threads, read_pool_size, prob_dup, scale, sigma, loc,
                min_read_len, read_probs, lens, max_passes, chi2_params_n, chi2_params_s, sqrt_params, norm_params, prob_thresh,
                prob_ins, prob_del, prob_subst);
    P.sep_files = sep_files;
    P.haplotype_probs = haplotype_probs.data();
    P.abort_cb = abort_cb; P.progress_cb = progress_cb; P.cb_user = &prog_bar;
    const int rc = jlp_pacbio(ctx.p, 1, &P, nullptr);
    if (rc != JLP_OK && rc != JLP_ERR_ABORTED) ctx.check(rc);
}
