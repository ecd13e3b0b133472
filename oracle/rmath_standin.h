/* TEST INFRASTRUCTURE ONLY (oracle/): stand-ins for the three Rmath functions the reference's PacBio path calls
 * (R::pnorm5, R::qnorm5 in PacBioQualityError::trunc_norm, /root/reference/src/hts_pacbio.h:349-352; R::qchisq in
 * PacBioPassSampler::sample, :178).  R is not in this image.  The same functions are used by the stub header the
 * unmodified reference is compiled against (oracle/stubs/RcppArmadillo.h) and by the C restatement
 * (oracle/jlp_oracle.c), so that both sides of a replay comparison compute identical doubles; against real Rmath
 * they agree to a few ulp (pnorm through erfc, qnorm by Newton steps on it, qchisq by bisection on the
 * regularised incomplete gamma function). */
#ifndef JLP_ORACLE_RMATH_STANDIN_H
#define JLP_ORACLE_RMATH_STANDIN_H

#include <math.h>

/* lower tail of the standard normal */
static inline double jlp_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* quantile: Abramowitz and Stegun 26.2.23 as the starting point, five Newton steps on the lower tail (relative error
 * below 1e-15 against scipy over 1e-300 .. 1 - 1e-15), the upper half by symmetry */
static inline double jlp_qnorm(double p) {
    if (!(p > 0.0)) return -INFINITY;
    if (!(p < 1.0)) return INFINITY;
    if (p > 0.5) return -jlp_qnorm(1.0 - p);      /* 1 - p is exact there; the lower tail keeps its relative precision */
    const double t = sqrt(-2.0 * log(p));
    double x = -(t - (2.515517 + 0.802853 * t + 0.010328 * t * t) / (1.0 + 1.432788 * t + 0.189269 * t * t + 0.001308 * t * t * t));
    for (int i = 0; i < 5; i++) {
        const double d = 0.39894228040143267794 * exp(-0.5 * x * x);
        if (!(d > 1e-300)) break;
        x -= (jlp_pnorm(x) - p) / d;
    }
    return x;
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
