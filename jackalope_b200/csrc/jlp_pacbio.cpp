// Host side of the PacBio read generator: see jlp_pacbio.h.
#include "jlp_pacbio.h"

#include <atomic>
#include <cmath>
#include <stdexcept>

#include "jlp_draws.h"
#include "jlp_host.h"

namespace jlp {

namespace {

// Draw planes of the PacBio generator: counter = (j_lo, j_hi, block, PL_PB | sub << 8), two 64-bit draws per block.
// sub 0 read-level (block 0: truncated normals left / right; block 1: start, strand; blocks 16.. the tail branch's
// rejection loop), sub 1 the walk (position p: block p >> 1, half p & 1), sub 2 the inserted / substituted bases,
// sub 3 the samplers of this file.
constexpr uint32_t PL_PB = 6;
uint64_t pb_draw(uint64_t seed, uint64_t j, uint32_t sub, uint32_t block, uint32_t half) {
    const U4 w = draw_block(seed, j, block, PL_PB, sub);
    return half ? hi64(w) : lo64(w);
}

// runif_01 (src/pcg.h:99-101) in the reference's own type
long double runif_01(uint64_t x) { return ((long double)x + 1) / ((long double)UINT64_MAX + 2); }

// Standard normal distribution function and quantile.  The reference calls R::pnorm5 / R::qnorm5 here
// (src/hts_pacbio.h:349-352); these agree with Rmath to a few ulp: the lower tail through erfc, the quantile by
// Newton steps on it from Abramowitz and Stegun 26.2.23, the upper half by symmetry.
double pnorm(double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); }
double qnorm(double p) {
    if (!(p > 0.0)) return -INFINITY;
    if (!(p < 1.0)) return INFINITY;
    if (p > 0.5) return -qnorm(1.0 - p);
    const double t = std::sqrt(-2.0 * std::log(p));
    double x = -(t - (2.515517 + 0.802853 * t + 0.010328 * t * t) / (1.0 + 1.432788 * t + 0.189269 * t * t + 0.001308 * t * t * t));
    for (int i = 0; i < 5; i++) {
        const double d = 0.39894228040143267794 * std::exp(-0.5 * x * x);
        if (!(d > 1e-300)) break;
        x -= (pnorm(x) - p) / d;
    }
    return x;
}

// quantile of the chi-squared distribution (R::qchisq, src/hts_pacbio.h:178) by bisection on gamma_p
double qchisq(double p, double df) {
    const long double a = 0.5L * df;
    double lo = 0.0, hi = df + 10.0 * std::sqrt(2.0 * df) + 50.0;
    while ((double)gamma_p(a, 0.5L * hi) < p) hi *= 2.0;
    for (int i = 0; i < 60; i++) {
        const double mid = 0.5 * (lo + hi);
        if ((double)gamma_p(a, 0.5L * mid) < p) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}

double total_prob(const PbModel& m, double e) { return std::pow(m.prob_ins, e) + std::pow(m.prob_del, e) + std::pow(m.prob_subst, e); }

double sigmoid(double x) { return 1 / (1 + std::pow(2, (-2.5 / 3 * x + 6.5 / 3))); }

// a stream of uniforms in (0, 1) for the samplers of one read
struct SampleStream {
    uint64_t seed, j;
    uint32_t k = 0;
    double next() {
        const uint64_t x = pb_draw(seed, j, 3, k >> 1, k & 1);
        k++;
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
};

// Gamma(a, 1), Marsaglia and Tsang (2000); a < 1 through Gamma(a + 1) U^(1/a)
double sample_gamma(double a, SampleStream& S) {
    if (a < 1.0) return sample_gamma(a + 1.0, S) * std::pow(S.next(), 1.0 / a);
    const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
        const double z = qnorm(S.next());
        const double v0 = 1.0 + c * z;
        if (v0 <= 0) continue;
        const double v = v0 * v0 * v0, u = S.next();
        if (std::log(u) < 0.5 * z * z + d - d * v + d * std::log(v)) return d * v;
    }
}

}  // namespace

void pb_prepare(PbModel& m) {
    if (!(m.prob_ins >= 0 && m.prob_del >= 0 && m.prob_subst >= 0) || m.prob_ins + m.prob_del + m.prob_subst > 1)
        throw std::runtime_error("\nWhen providing info for the PacBio sequencer, the insertion, deletion, and substitution "
                                 "probabilities cannot sum to > 1.");
    if (m.read_probs.size() != m.read_lens.size())
        throw std::runtime_error("Probability and read lengths vector should be the same length.");   // src/hts_pacbio.h:72-74
    // PacBioQualityError::calc_min_exp
    double min_exp = 1, total = total_prob(m, min_exp), left, right;
    if (total < m.prob_thresh) {
        while (total < m.prob_thresh) { min_exp /= 2; total = total_prob(m, min_exp); }
        left = min_exp; right = min_exp * 2;
    } else {
        while (total > m.prob_thresh) { min_exp *= 2; total = total_prob(m, min_exp); }
        left = min_exp / 2; right = min_exp;
    }
    for (int i = 0; i < 15; i++) {
        const double mid = (left + right) / 2;
        total = total_prob(m, mid);
        if (total == m.prob_thresh) { min_exp = mid; break; }
        else if (total > m.prob_thresh) { left = mid; min_exp = (mid + right) / 2; }
        else { right = mid; min_exp = (left + mid) / 2; }
    }
    m.min_exp = min_exp;
    m.qchisq_n = (size_t)std::max(1.0, std::floor(m.chi2_n[2])) + 2;
    m.qchisq_cache.reset(new std::atomic<double>[m.qchisq_n]);
    for (size_t i = 0; i < m.qchisq_n; i++) m.qchisq_cache[i].store(-1.0, std::memory_order_relaxed);
    if (!m.read_probs.empty()) {
        m.len_prob.resize(m.read_probs.size());
        m.len_alias.resize(m.read_probs.size());
        alias_build(m.read_probs.data(), m.read_probs.size(), m.len_prob.data(), m.len_alias.data());
    }
}

namespace {
// PacBioPassSampler::sample for r.read_length
void pb_passes(const PbModel& m, SampleStream& S, PbSample& r) {
    const double read_length = (double)r.read_length;
    const double lcap = std::min(read_length, m.chi2_n[2]);
    double n = m.chi2_n[0] * lcap + m.chi2_n[1];
    if (n < 0.001) n = 0.001;
    double s;
    if (read_length <= m.chi2_s[2]) {
        s = m.chi2_s[0] * read_length - m.chi2_s[1];
        if (s < 0.001) s = 0.001;
    } else {
        s = m.chi2_s[3] / std::pow(read_length, m.chi2_s[4]);
    }
    // the outlier threshold is a pure function of lcap (an integer): filled on first use; concurrent callers may compute
    // the same value twice
    std::atomic<double>& slot = m.qchisq_cache[(size_t)std::floor(lcap)];
    double thr = slot.load(std::memory_order_relaxed);
    if (thr < 0) { thr = qchisq(0.9925, n); slot.store(thr, std::memory_order_relaxed); }
    double passes = 2.0 * sample_gamma(0.5 * n, S);
    while (passes > thr) passes = 2.0 * sample_gamma(0.5 * n, S);
    passes *= s;
    passes += 1;
    if (passes > (double)m.max_passes) passes = (double)m.max_passes;
    double wholes;
    const double fraction = std::modf(passes, &wholes);
    if (((uint64_t)wholes & 1ull) == 0ull) {
        r.split_pos = (uint64_t)std::round(read_length * fraction);
        r.passes_left = std::ceil(passes);
        r.passes_right = std::floor(passes);
    } else {
        r.split_pos = (uint64_t)std::round(read_length * (1 - fraction));
        r.passes_left = std::floor(passes);
        r.passes_right = std::ceil(passes);
    }
}
}  // namespace

PbSample pb_sample(const PbModel& m, uint64_t seed, uint64_t j, uint64_t chrom_len) {
    SampleStream S{seed, j};
    PbSample r;
    // ---- PacBioReadLenSampler::sample
    if (m.read_probs.empty()) {
        const double mu = std::log(m.scale);
        double min_len = std::ceil(m.min_read_len);
        if (min_len < 1) min_len = 1;
        double rnd = std::exp(mu + m.sigma * qnorm(S.next())) + m.loc;
        for (int it = 0; rnd < min_len && it < 10; it++) rnd = std::exp(mu + m.sigma * qnorm(S.next())) + m.loc;
        if (rnd < min_len) rnd = min_len;
        r.read_length = (uint64_t)rnd;
    } else {
        const uint64_t n = m.read_lens.size();
        uint64_t i = (uint64_t)(S.next() * (double)n);
        if (i >= n) i = n - 1;
        if (!(S.next() < m.len_prob[i])) i = m.len_alias[i];
        r.read_length = m.read_lens[i];
    }
    if (r.read_length >= chrom_len) r.read_length = chrom_len;          // one_read, src/hts_pacbio.cpp
    pb_passes(m, S, r);
    return r;
}

PbSample pb_sample_passes(const PbModel& m, uint64_t seed, uint64_t j, uint64_t read_length) {
    SampleStream S{seed, j};
    PbSample r;
    r.read_length = read_length;
    pb_passes(m, S, r);
    return r;
}

uint64_t pb_dup_draw(uint64_t seed, uint64_t j) { return pb_draw(seed, j, 0, 2, 0); }

namespace {

// PacBioQualityError::trunc_norm (src/hts_pacbio.h:340-370); side 0 = left, 1 = right
double trunc_norm(const PbModel& m, double lower_thresh, uint64_t seed, uint64_t j, uint32_t side) {
    const double a_bar = (lower_thresh - m.norm_params[0]) / m.norm_params[1];
    if (lower_thresh < (m.norm_params[0] + 5 * m.norm_params[1])) {
        const double p = pnorm(a_bar);
        const long double u = (long double)p + runif_01(pb_draw(seed, j, 0, 0, side)) * ((long double)1 - (long double)p);   // runif_ab
        return qnorm((double)u) * m.norm_params[1] + m.norm_params[0];
    }
    double x_bar;
    for (uint32_t it = 0;; it++) {
        const double u = (double)runif_01(pb_draw(seed, j, 0, 16 + 2 * it, side));
        x_bar = std::sqrt(a_bar * a_bar - 2 * std::log(1 - u));
        const double v = (double)runif_01(pb_draw(seed, j, 0, 17 + 2 * it, side));
        if (!(v > (x_bar / a_bar))) break;
    }
    return m.norm_params[1] * x_bar + m.norm_params[0];
}

}  // namespace

void pb_read_model(const PbModel& m, uint64_t seed, uint64_t j, const PbSample& s, PbRead& out) {
    // ---- update_probs
    const double pass[2] = {s.passes_left, s.passes_right};
    double cum[2][3];
    for (uint32_t side = 0; side < 2; side++) {
        const double root = std::sqrt(pass[side] + m.sqrt_params[0]) - m.sqrt_params[1];
        const double thresh = (m.min_exp - root) / sigmoid(pass[side]);
        const double incr = trunc_norm(m, thresh, seed, j, side);
        double e = incr * sigmoid(pass[side]) + std::sqrt(pass[side] + m.sqrt_params[0]) - m.sqrt_params[1];
        if (e < 0.6) e = 0.6;
        cum[side][0] = std::pow(m.prob_ins, e);
        cum[side][1] = std::pow(m.prob_del, e) + cum[side][0];
        cum[side][2] = std::pow(m.prob_subst, e) + cum[side][1];
    }
    // ---- fill_quals
    uint32_t flags = 0;
    for (uint32_t side = 0; side < 2; side++) {
        uint64_t q = (uint64_t)std::round(-10.0 * std::log10(cum[side][2]));
        if (q > 93) q = 93;
        flags |= (uint32_t)(q + '!') << (8 + 8 * side);
        // ---- the walk's comparisons on `double u` (src/hts_pacbio.h:300-316): u > cum[2] no event; u < cum[0] insertion;
        //      u < cum[1] deletion; else substitution
        const Thr t[3] = {thr_double_lt(cum[side][0]), thr_double_lt(cum[side][1]), thr_double_le(cum[side][2])};
        for (uint32_t k = 0; k < 3; k++) {
            out.thr[side][k] = t[k].thr;
            if (t[k].all) flags |= 1u << (3 * side + k);
        }
    }
    out.read_length = (uint32_t)s.read_length;
    out.split_pos = (uint32_t)s.split_pos;
    out.flags = flags;
}

}  // namespace jlp
