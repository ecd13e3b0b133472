// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <Rcpp.h> as the generated wrappers of
// /root/reference/src/RcppExports.cpp use it (BEGIN_RCPP / END_RCPP, RcppExport, Rcpp::RNGScope,
// Rcpp::traits::input_parameter<T>::type, R_NilValue).  With these the STOCK wrappers
// _jackalope_illumina_{ref,hap}_cpp / _jackalope_pacbio_{ref,hap}_cpp compile unmodified, so the
// link test of the Rcpp glue (oracle/glue_driver.cpp) goes through exactly the prototypes the
// package's .Call table binds.  A "SEXP" here is a plain pointer to the C++ value of the
// parameter's type (the driver owns the values).  Nothing in the product includes this file.
#ifndef JLP_ORACLE_STUB_RCPP_H
#define JLP_ORACLE_STUB_RCPP_H

#include "RcppArmadillo.h"

#include <type_traits>

#define RcppExport extern "C"
#define R_NilValue nullptr

namespace jlp_stub {
// what END_RCPP turns into an R error: the message is kept, the wrapper returns this token
inline std::string& last_r_error() {
    static thread_local std::string s;
    return s;
}
inline SEXP r_error_token() {
    static int token;
    return &token;
}
}  // namespace jlp_stub

#define BEGIN_RCPP try {
#define END_RCPP                                                    \
    }                                                               \
    catch (std::exception & ex__) {                                 \
        jlp_stub::last_r_error() = ex__.what();                     \
        return jlp_stub::r_error_token();                           \
    }                                                               \
    catch (...) {                                                   \
        jlp_stub::last_r_error() = "c++ exception (unknown reason)"; \
        return jlp_stub::r_error_token();                           \
    }

namespace Rcpp {
struct RNGScope {};
namespace traits {
template <typename T> struct input_parameter {
    typedef typename std::remove_const<typename std::remove_reference<T>::type>::type V;
    struct type {
        const V* p;
        type(SEXP x) : p(static_cast<const V*>(x)) {}
        operator const V&() const { return *p; }
    };
};
template <> struct input_parameter<SEXP> {
    struct type {
        SEXP p;
        type(SEXP x) : p(x) {}
        operator SEXP() const { return p; }
    };
};
}  // namespace traits
}  // namespace Rcpp

#endif
