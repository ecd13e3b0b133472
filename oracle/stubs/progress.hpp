// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for RcppProgress' <progress.hpp>.
#ifndef JLP_ORACLE_STUB_PROGRESS_HPP
#define JLP_ORACLE_STUB_PROGRESS_HPP
struct Progress {
    Progress(unsigned long, bool) {}
    bool check_abort() { return false; }
    bool is_aborted() { return false; }
    void increment(unsigned long = 1) {}
    void cleanup() {}
};
#endif
