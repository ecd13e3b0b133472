/* TEST INFRASTRUCTURE ONLY (oracle/): stand-ins for the three Rmath functions the reference's PacBio path calls
 * (R::pnorm5, R::qnorm5 in PacBioQualityError::trunc_norm, /root/reference/src/hts_pacbio.h:349-352; R::qchisq in
 * PacBioPassSampler::sample, :178).  R is not in this image.  The same functions are used by the stub header the
 * unmodified reference is compiled against (oracle/stubs/RcppArmadillo.h) and by the C restatement
 * (oracle/jlp_oracle.c), so that both sides of a replay comparison compute identical doubles; against real Rmath
 * they agree to a few ulp (pnorm through erfc, qnorm by Wichura's AS 241 -- R's own algorithm --, qchisq by bisection on
 * the regularised incomplete gamma function). */
#ifndef JLP_ORACLE_RMATH_STANDIN_H
#define JLP_ORACLE_RMATH_STANDIN_H

#include <math.h>

/* lower tail of the standard normal */
static inline double jlp_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* quantile of the standard normal */
static inline double jlp_qnorm(double p) {
    /* Wichura (1988), Algorithm AS 241, routine PPND16 -- the algorithm R's qnorm5 itself uses: rational approximations
     * in q = p - 1/2 (central part) and in r = sqrt(-log(min(p, 1 - p))) (the tails), about 16 digits */
    if (!(p > 0.0)) return -INFINITY;
    if (!(p < 1.0)) return INFINITY;
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * (((((((2.5090809287301226727e+3 * r + 3.3430575583588128105e+4) * r + 6.7265770927008700853e+4) * r +
                        4.5921953931549871457e+4) * r + 1.3731693765509461125e+4) * r + 1.9715909503065514427e+3) * r +
                      1.3314166789178437745e+2) * r + 3.3871328727963666080) /
               (((((((5.2264952788528545610e+3 * r + 2.8729085735721942674e+4) * r + 3.9307895800092710610e+4) * r +
                    2.1213794301586595867e+4) * r + 5.3941960214247511077e+3) * r + 6.8718700749205790830e+2) * r +
                  4.2313330701600911252e+1) * r + 1.0);
    }
    double r = sqrt(-log(q < 0 ? p : 1.0 - p));
    double v;
    if (r <= 5.0) {
        r -= 1.6;
        v = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
                 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.76949722146069140550) * r +
               4.63033784615654529590) * r + 1.42343711074968357734) /
            (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
                 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940) * r +
               2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        v = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
                 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580) * r +
               5.46378491116411436990) * r + 6.65790464350110377720) /
            (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
                 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
               5.99832206555887937690e-1) * r + 1.0);
    }
    return q < 0 ? -v : v;
}

/* regularised lower incomplete gamma P(a, x): series below a + 1, continued fraction above */
static inline double jlp_pgamma(double a, double x) {
    if (x <= 0.0) return 0.0;
    const double lg = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 10000; n++) {
            ap += 1.0; del *= x / ap; sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * exp(-x + a * log(x) - lg);
    }
    double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; i++) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(-x + a * log(x) - lg) * h;
}

/* quantile of the chi-squared distribution with df degrees of freedom (lower tail) */
static inline double jlp_qchisq(double p, double df) {
    const double a = 0.5 * df;
    double lo = 0.0, hi = df + 10.0 * sqrt(2.0 * df) + 50.0;
    while (jlp_pgamma(a, 0.5 * hi) < p) hi *= 2.0;
    for (int i = 0; i < 200; i++) {
        const double mid = 0.5 * (lo + hi);
        if (jlp_pgamma(a, 0.5 * mid) < p) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}

#endif
