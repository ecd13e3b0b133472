// TEST INFRASTRUCTURE ONLY (oracle/): inert stand-in for htslib's BGZF writer.
// The oracle never exercises compressed output through the reference.
#ifndef JLP_ORACLE_STUB_BGZF_H
#define JLP_ORACLE_STUB_BGZF_H
#include <cstddef>
#include <sys/types.h>
struct BGZF { int errcode; };
inline BGZF* bgzf_open(const char*, const char*) { return nullptr; }
inline int bgzf_mt(BGZF*, int, int) { return 0; }
inline ssize_t bgzf_write(BGZF*, const void*, size_t n) { return static_cast<ssize_t>(n); }
inline int bgzf_close(BGZF*) { return 0; }
#endif
