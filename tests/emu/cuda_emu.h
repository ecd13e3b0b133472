// A tiny CUDA execution model on host threads, for the CPU test suite (tests/test_bgzf_emu.py): one OS thread per
// CUDA thread of ONE block at a time, __syncthreads() and the warp collectives on pthread barriers, shared-memory
// atomics as host atomics.  It exists so that the logic of a kernel (jackalope_b200/csrc/jlp_bgzf.cu is compiled
// unchanged, as C++, with JLP_CPU_EMU defined) can be exercised where there is no GPU.  Test infrastructure only:
// nothing under jackalope_b200/ includes or links it, and nothing here is ever timed.
//
// Rules it relies on (the same ones CUDA states): every *_sync collective is reached by all 32 lanes of the warp,
// __syncthreads() by all threads of the block (an early return must be block-uniform).  Shared memory is filled
// with 0xcd bytes before each block so that reads of what was never written show up.
#ifndef JLP_CUDA_EMU_H
#define JLP_CUDA_EMU_H

#include <pthread.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) (std::memcpy((void*)&(sym), (src), (n)), cudaSuccess)

namespace jlp_emu {
struct Warp {
    pthread_barrier_t bar;
    uint64_t vals[32];
};
struct Block {
    pthread_barrier_t bar;
    std::vector<Warp> warps;
    uint8_t* smem;
};
extern thread_local Block* tl_block;
extern thread_local uint3 tl_tid, tl_bid, tl_gdim;
inline uint8_t* dyn_smem() { return tl_block->smem; }
inline Warp& my_warp() { return tl_block->warps[tl_tid.x >> 5]; }

template <class F>
void launch(unsigned grid, unsigned block, size_t smem_bytes, F body) {
    for (unsigned b = 0; b < grid; b++) {
        Block B;
        pthread_barrier_init(&B.bar, nullptr, block);
        B.warps.resize((block + 31) / 32);
        for (unsigned w = 0; w < B.warps.size(); w++) pthread_barrier_init(&B.warps[w].bar, nullptr, std::min(32u, block - 32 * w));
        std::vector<uint8_t> sm(smem_bytes + 64, 0xcd);
        B.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm.data()) + 15) & ~(uintptr_t)15);
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t]() {
                tl_block = &B; tl_tid = uint3{t, 0, 0}; tl_bid = uint3{b, 0, 0}; tl_gdim = uint3{grid, 1, 1};
                body();
            });
        for (auto& x : th) x.join();
        pthread_barrier_destroy(&B.bar);
        for (auto& w : B.warps) pthread_barrier_destroy(&w.bar);
    }
}
}  // namespace jlp_emu

#define threadIdx (jlp_emu::tl_tid)
#define blockIdx (jlp_emu::tl_bid)
#define gridDim (jlp_emu::tl_gdim)

static inline void __syncthreads() { pthread_barrier_wait(&jlp_emu::tl_block->bar); }
static inline void __syncwarp() { pthread_barrier_wait(&jlp_emu::my_warp().bar); }

template <class T, class Pick>
static inline T emu_exchange(T v, Pick pick) {
    jlp_emu::Warp& W = jlp_emu::my_warp();
    const unsigned lane = threadIdx.x & 31;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    W.vals[lane] = raw;
    pthread_barrier_wait(&W.bar);
    const int src = pick(lane);
    uint64_t got = src >= 0 && src < 32 ? W.vals[src] : raw;
    pthread_barrier_wait(&W.bar);
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, unsigned src) { return emu_exchange(v, [&](unsigned) { return (int)(src & 31u); }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) { return emu_exchange(v, [&](unsigned l) { return (int)l - (int)d; }); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) { return emu_exchange(v, [&](unsigned l) { return (int)(l + d); }); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, unsigned m) { return emu_exchange(v, [&](unsigned l) { return (int)(l ^ m); }); }
static inline unsigned __ballot_sync(unsigned, bool p) {
    jlp_emu::Warp& W = jlp_emu::my_warp();
    W.vals[threadIdx.x & 31] = p ? 1 : 0;
    pthread_barrier_wait(&W.bar);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(W.vals[i] & 1) << i;
    pthread_barrier_wait(&W.bar);
    return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicSub(unsigned* p, unsigned v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
using std::max;
using std::min;

#endif
