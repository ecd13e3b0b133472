// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <RcppArmadillo.h>.
//
// Purpose: let the UNMODIFIED reference sources under /root/reference/src
// (hts_illumina.cpp, hap_classes.cpp and the headers they include) compile
// with plain g++ in a container that has no R / Rcpp / Armadillo.  Only the
// handful of symbols those files touch are provided.  Nothing in the product
// (jackalope_b200/) includes or links this file.
#ifndef JLP_ORACLE_STUB_RCPPARMADILLO_H
#define JLP_ORACLE_STUB_RCPPARMADILLO_H

// (the real header transitively provides these standard headers)
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <deque>
#include <fstream>
#include <random>
#include <sstream>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

typedef void* SEXP;

namespace jlp_stub {
// Deterministic stand-in for R's RNG (the reference seeds pcg64 from
// Rcpp::runif, src/pcg.h:42,65).  splitmix64; state settable by the driver.
inline uint64_t& r_rng_state() {
    static uint64_t s = 0x9E3779B97F4A7C15ULL;
    return s;
}
inline uint64_t r_rng_next() {
    uint64_t z = (r_rng_state() += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
}  // namespace jlp_stub

// Rmath: the three functions the PacBio path calls (src/hts_pacbio.h:178,349,352); see rmath_standin.h
#include "../rmath_standin.h"
namespace R {
inline double pnorm5(double x, double mu, double sigma, int lower, int log_p) {
    (void)log_p;
    const double p = jlp_pnorm((x - mu) / sigma);
    return lower ? p : 1.0 - p;
}
inline double qnorm5(double p, double mu, double sigma, int lower, int log_p) {
    (void)log_p;
    return mu + sigma * jlp_qnorm(lower ? p : 1.0 - p);
}
inline double qchisq(double p, double df, int lower, int log_p) {
    (void)log_p;
    return jlp_qchisq(lower ? p : 1.0 - p, df);
}
}  // namespace R

namespace Rcpp {

struct exception : public std::runtime_error {
    exception(const char* msg, bool = false) : std::runtime_error(msg) {}
};

inline void stop(const std::string& msg) { throw exception(msg.c_str()); }
inline void stop(const char* msg) { throw exception(msg); }
inline void warning(const std::string& msg) { std::cerr << "Warning: " << msg << std::endl; }
inline void warning(const char* msg) { std::cerr << "Warning: " << msg << std::endl; }

static std::ostream& Rcout = std::cout;
inline void Rprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vprintf(fmt, ap);
    va_end(ap);
}

// Rcpp::runif(n, a, b): uniform doubles in (a, b)
inline std::vector<double> runif(int n, double a, double b) {
    std::vector<double> out(n);
    for (int i = 0; i < n; i++) {
        double u = (static_cast<double>(jlp_stub::r_rng_next() >> 11) + 0.5) / 9007199254740992.0;
        out[i] = a + u * (b - a);
    }
    return out;
}

template <typename T> struct as_impl;
template <> struct as_impl<std::vector<uint_fast64_t>> {
    static std::vector<uint_fast64_t> get(const std::vector<double>& x) {
        std::vector<uint_fast64_t> out(x.size());
        for (size_t i = 0; i < x.size(); i++) out[i] = static_cast<uint_fast64_t>(x[i]);
        return out;
    }
};
template <typename T> inline T as(const std::vector<double>& x) { return as_impl<T>::get(x); }
// as<std::string>(SEXP): our Function stub hands back a heap std::string*
template <typename T> inline T as(SEXP x) {
    T* p = static_cast<T*>(x);
    T out(*p);
    delete p;
    return out;
}

// XPtr<T>: the entry points only borrow the pointee (src/hts_illumina.cpp:614,689)
template <typename T> struct XPtr {
    T* ptr;
    XPtr(SEXP p) : ptr(static_cast<T*>(p)) {}
    XPtr(T* p, bool = true) : ptr(p) {}
    T& operator*() const { return *ptr; }
    T* operator->() const { return ptr; }
    T* get() const { return ptr; }
};

// Environment / Function / List: only path.expand (src/io.h:37-46, identity
// here) and options("width") (src/util.h:37-46, 80 here) are ever called.
struct ListElem {
    operator int() const { return 80; }
};
struct List {
    ListElem operator[](const char*) const { return ListElem(); }
};
struct FunctionResult {
    std::string arg;
    operator SEXP() const { return static_cast<SEXP>(new std::string(arg)); }
    operator List() const { return List(); }
};
struct Function {
    FunctionResult operator()(const std::string& x) const { return FunctionResult{x}; }
    FunctionResult operator()(const char* x) const { return FunctionResult{std::string(x)}; }
};
struct Environment {
    Environment(const char*) {}
    Function operator[](const char*) const { return Function(); }
};

}  // namespace Rcpp

namespace arma {
// Only what AliasSampler::construct uses (src/alias_sampler.h:41-106).
struct rowvec {
    std::vector<double> v;
    unsigned long long n_elem;
    rowvec() : v(), n_elem(0) {}
    rowvec(const std::vector<double>& x) : v(x), n_elem(x.size()) {}
    double& operator()(unsigned long long i) { return v[i]; }
    const double& operator()(unsigned long long i) const { return v[i]; }
    rowvec& operator/=(double d) { for (double& x : v) x /= d; return *this; }
    rowvec& operator*=(double d) { for (double& x : v) x *= d; return *this; }
};
inline double accu(const rowvec& p) { return std::accumulate(p.v.begin(), p.v.end(), 0.0); }
}  // namespace arma

#endif
