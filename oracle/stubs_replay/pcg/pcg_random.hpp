// TEST INFRASTRUCTURE ONLY (oracle/): scripted stand-in for `pcg64`.
//
// In the replay build this header shadows /root/reference/inst/include/pcg/
// pcg_random.hpp, so that the UNMODIFIED reference classes consume a stream of
// 64-bit values supplied by the test harness instead of a real PRNG.  Every
// `pcg64` object reads from the same global script (the reference uses one
// engine per thread and the replay driver is single-threaded).
#ifndef JLP_ORACLE_REPLAY_PCG_RANDOM_HPP
#define JLP_ORACLE_REPLAY_PCG_RANDOM_HPP

#include <cstddef>
#include <cstdint>
#include "pcg/pcg_extras.hpp"

namespace jlp_replay {
struct Script {
    const uint64_t* data = nullptr;
    size_t len = 0;
    size_t pos = 0;
    size_t underruns = 0;
};
inline Script& script() {
    static Script s;
    return s;
}
}  // namespace jlp_replay

class pcg64 {
public:
    typedef uint64_t result_type;
    pcg64() {}
    pcg64(pcg_extras::pcg128_t, pcg_extras::pcg128_t) {}
    static constexpr result_type min() { return 0ULL; }
    static constexpr result_type max() { return ~0ULL; }
    result_type operator()() {
        jlp_replay::Script& s = jlp_replay::script();
        if (s.pos >= s.len) {
            s.underruns++;
            return 0x8000000000000000ULL;
        }
        return s.data[s.pos++];
    }
};

#endif
