// TEST INFRASTRUCTURE ONLY (oracle/): replay build's stand-in for the vendored
// pcg_extras.hpp -- only the 128-bit integer typedef the reference uses
// (src/jackalope_types.h:31).
#ifndef JLP_ORACLE_REPLAY_PCG_EXTRAS_HPP
#define JLP_ORACLE_REPLAY_PCG_EXTRAS_HPP
namespace pcg_extras {
typedef __uint128_t pcg128_t;
}
#endif
