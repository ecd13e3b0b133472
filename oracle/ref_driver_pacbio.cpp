// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI driver around the UNMODIFIED reference implementation of the PacBio
// read generator, compiled into the same two libraries as ref_driver.cpp
// (oracle/_ref/libjlp_ref.so with the real pcg64, libjlp_ref_replay.so with the scripted one).
//
// This translation unit #includes /root/reference/src/hts_pacbio.cpp where it lies (-I/root/reference/src); no
// reference source is copied.  -fno-access-control lets the driver set the per-read state of PacBioOneGenome and call
// its private append_pool.  R::pnorm5 / R::qnorm5 / R::qchisq come from oracle/rmath_standin.h (R is not in this
// image); nothing in the product loads these libraries.

#include "hts_pacbio.cpp"  // resolved through -I/root/reference/src

#include <cstring>

namespace {

int pb_fail(char* err, size_t cap, const char* msg) {
    if (err && cap) {
        std::strncpy(err, msg, cap - 1);
        err[cap - 1] = '\0';
    }
    return -1;
}

}  // namespace

extern "C" {

// PacBioQualityError::calc_min_exp (src/hts_pacbio.cpp), through the constructor
double jrefpb_min_exp(const double* sqrt_params, const double* norm_params, double prob_thresh, double prob_ins,
                      double prob_del, double prob_subst) {
    PacBioQualityError qe(std::vector<double>(sqrt_params, sqrt_params + 2), std::vector<double>(norm_params, norm_params + 2),
                          prob_thresh, prob_ins, prob_del, prob_subst);
    return qe.min_exp;
}

#ifndef JLP_REPLAY
// ------------------------------------------------- end-to-end reference ---

int jrefpb_pacbio_ref(void* ref, const char* out_prefix, uint64_t n_reads, uint64_t n_threads, uint64_t read_pool_size,
                      double prob_dup, double scale, double sigma, double loc, double min_read_len,
                      const double* read_probs, const uint64_t* read_lens, uint64_t n_custom, uint64_t max_passes,
                      const double* chi2_params_n, const double* chi2_params_s, const double* sqrt_params,
                      const double* norm_params, double prob_thresh, double prob_ins, double prob_del, double prob_subst,
                      char* err, uint64_t errcap) {
    try {
        pacbio_ref_cpp(ref, out_prefix, 0, "bgzip", n_reads, n_threads, false, read_pool_size, prob_dup, scale, sigma, loc,
                       min_read_len, std::vector<double>(read_probs, read_probs + n_custom),
                       std::vector<uint64>(read_lens, read_lens + n_custom), max_passes,
                       std::vector<double>(chi2_params_n, chi2_params_n + 3), std::vector<double>(chi2_params_s, chi2_params_s + 5),
                       std::vector<double>(sqrt_params, sqrt_params + 2), std::vector<double>(norm_params, norm_params + 2),
                       prob_thresh, prob_ins, prob_del, prob_subst);
    } catch (std::exception& e) {
        return pb_fail(err, errcap, e.what());
    }
    return 0;
}

// read lengths and (split_pos, passes_left, passes_right) as the reference's own samplers draw them, n of each,
// for the statistical comparison of the product's samplers (PacBioReadLenSampler::sample, PacBioPassSampler::sample)
int jrefpb_sample_lengths_passes(uint64_t n, uint64_t seed, double scale, double sigma, double loc, double min_read_len,
                                 uint64_t max_passes, const double* chi2_params_n, const double* chi2_params_s,
                                 uint64_t* read_len, uint64_t* split_pos, double* passes_left, double* passes_right) {
    pcg64 eng(seed, seed ^ 0x9E3779B97F4A7C15ULL);
    PacBioReadLenSampler ls(scale, sigma, loc, min_read_len);
    PacBioPassSampler ps(max_passes, std::vector<double>(chi2_params_n, chi2_params_n + 3),
                         std::vector<double>(chi2_params_s, chi2_params_s + 5));
    for (uint64_t i = 0; i < n; i++) {
        read_len[i] = ls.sample(eng);
        uint64 sp; double pl, pr;
        ps.sample(sp, pl, pr, eng, static_cast<double>(read_len[i]));
        split_pos[i] = sp; passes_left[i] = pl; passes_right[i] = pr;
    }
    return 0;
}

#else  // JLP_REPLAY
// ----------------------------------------------------------------- replay ---
//
// Drive the unmodified read model on reads chosen by the harness, every pcg64 output taken from `script`.  For each
// read i the driver sets what the two stateful samplers would have produced -- chrom_ind, read_length
// (PacBioReadLenSampler::sample: std::lognormal_distribution keeps a cached normal, so its draw count depends on the
// history) and split_pos / passes_left / passes_right (PacBioPassSampler::sample: std::chi_squared_distribution, the
// same) -- and then runs exactly what PacBioOneGenome::one_read does after them (src/hts_pacbio.cpp):
//   qe_sampler.sample(...)   update_probs (two truncated normals), fill_quals, the insertion/deletion/substitution walk
//   read_chrom_space, read_start   (for a duplicate: re_read's rule instead, see below)
//   append_pool(...)         strand, ID line, fill_read / rev_comp, the edits, the two-valued quality line
// consumed[i] returns how many draws read i took from the script.
int jrefpb_replay(void* ref, uint64_t n_reads, const uint64_t* chrom_ind, const uint64_t* read_len, const uint64_t* split_pos,
                  const double* passes_left, const double* passes_right, const int32_t* is_dup, const double* sqrt_params,
                  const double* norm_params,
                  double prob_thresh, double prob_ins, double prob_del, double prob_subst, const uint64_t* script,
                  uint64_t script_len, uint64_t* consumed, char* out, uint64_t cap, uint64_t* len, char* err, uint64_t errcap) {
    try {
        RefGenome* g = static_cast<RefGenome*>(ref);
        jlp_replay::Script& sc = jlp_replay::script();
        sc.data = script; sc.len = script_len; sc.pos = 0; sc.underruns = 0;
        pcg64 eng;
        const double chi2_n[3] = {0.00189237136, 2.53944970, 5500}, chi2_s[5] = {0.01214, -5.12, 675, 48303.0732881, 1.4691051212330266};
        PacBioReference rd(*g, 17922.611306, 0.200110276521, -10075.4363813, 50.0, 40,
                           std::vector<double>(chi2_n, chi2_n + 3), std::vector<double>(chi2_s, chi2_s + 5),
                           std::vector<double>(sqrt_params, sqrt_params + 2), std::vector<double>(norm_params, norm_params + 2),
                           prob_thresh, prob_ins, prob_del, prob_subst);
        std::vector<char> pool;
        for (uint64_t i = 0; i < n_reads; i++) {
            const size_t before = sc.pos;
            rd.chrom_ind = chrom_ind[i];
            const uint64 chrom_len = (*g)[rd.chrom_ind].size();
            rd.read_length = read_len[i];
            if (rd.read_length >= chrom_len) rd.read_length = chrom_len;       // one_read, after len_sampler.sample
            rd.split_pos = split_pos[i]; rd.passes_left = passes_left[i]; rd.passes_right = passes_right[i];
            rd.qe_sampler.sample(eng, rd.qual_left, rd.qual_right, rd.insertions, rd.deletions, rd.substitutions, chrom_len,
                                 rd.read_length, rd.split_pos, rd.passes_left, rd.passes_right);
            rd.read_chrom_space = rd.read_length + rd.deletions.size() - rd.insertions.size();
            if (is_dup && is_dup[i]) {
                // re_read (src/hts_pacbio.cpp): chrom_ind, read_length and read_start are the chain's (the harness passes the
                // leader's chromosome and length; rd.read_start still holds the leader's); deletions are given up until the
                // template fits, a read that does not fit is not written
                while ((rd.read_chrom_space + rd.read_start) > chrom_len) {
                    if (rd.deletions.empty()) break;
                    rd.deletions.pop_back();
                    rd.read_chrom_space--;
                }
                if ((rd.read_chrom_space + rd.read_start) > chrom_len) { consumed[i] = sc.pos - before; continue; }
            } else if (rd.read_chrom_space < chrom_len) {
                double u = runif_01(eng);
                rd.read_start = static_cast<uint64>(u * (chrom_len - rd.read_chrom_space + 1));
            } else if (rd.read_chrom_space == chrom_len) {
                rd.read_start = 0;
            } else {
                return pb_fail(err, errcap, "read_chrom_space should never exceed the chromosome length.");
            }
            rd.append_pool<std::vector<char>>(pool, eng);
            consumed[i] = sc.pos - before;
        }
        if (sc.underruns) return pb_fail(err, errcap, "the script ran out of draws");
        *len = pool.size();
        if (pool.size() <= cap) std::memcpy(out, pool.data(), pool.size());
    } catch (std::exception& e) {
        return pb_fail(err, errcap, e.what());
    }
    return 0;
}
#endif

}  // extern "C"
