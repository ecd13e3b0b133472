// BGZF on the device: the FASTQ of a batch is compressed where it was generated, so that
// only the compressed bytes cross PCIe (the file format is what
// FileBGZF / bgzip_file write through htslib, /root/reference/src/io.h:58-135,
// src/hts.h:140-180: gzip members (RFC 1952) of at most 0xff00 input bytes carrying the
// "BC" extra field with the member's size, each holding one RFC 1951 deflate stream).
//
// One CTA per BGZF block, 512 threads, thread t owning bytes [128 t, 128 t + 128) of the block
// in registers from the single global read:
//   pass 1  per-warp byte histograms (shared atomics) and the CRC-32 of each thread's chunk;
//           chunk CRCs are combined by a tree of "advance by 2^j zero bytes" operators
//   build   one dynamic-Huffman block of literals (no matches: on FASTQ of random genomes an
//           order-0 code gives the ratio of zlib level 1, 0.37): code lengths by
//           two-queue Huffman over the rank-sorted symbols, limited to 15 / 7 bits by the
//           usual count fix-up, canonical codes, run-length coded code lengths
//   pass 2  every thread sizes its chunk, a block scan gives its bit offset, and it packs
//           its codes into the block image in shared memory
//   a block that would not shrink is emitted as a stored block through the same packing
//   code with the identity table.
// The image leaves with 128-bit stores into a 64 KiB slot; k_bgzf_scan / k_bgzf_gather
// then make the file contiguous.
#include "jlp_kernels.cuh"

#include <cstring>

namespace jlp {

namespace {

constexpr int kZT = 512;                 // threads per CTA
constexpr int kZW = kZT / 32;
constexpr uint32_t kChunk = 128;         // input bytes per thread
constexpr uint32_t kNLit = 286;          // literal/length alphabet
constexpr uint32_t kHdr = 18;            // BGZF member header bytes
constexpr uint32_t kStageWords = kBgzfSlot / 4;
constexpr uint32_t kHistWord0 = kStageWords - kZW * 256;   // the histograms alias the end of the image

__constant__ uint32_t c_crc_tab[256];
__constant__ uint32_t c_crc_adv[17][32];   // operator "advance the CRC register over 2^j zero bytes", by bit image
__constant__ uint32_t c_crc_init_full;     // the initial register 0xffffffff advanced over a full block
// the same operators for 2^7 .. 2^15 zero bytes (the sizes a full block's CRC tree combines) as 4 x 256-entry tables:
// advance(v) = T[0][v & 255] ^ T[1][v >> 8 & 255] ^ T[2][v >> 16 & 255] ^ T[3][v >> 24]
__device__ uint32_t g_crc_lvl[9][4][256];

struct ZShared {
    uint32_t stage[kStageWords];          // the block image; its last 16 KiB hold the per-warp literal counts until the packing starts
    uint32_t cnt[288];                    // literal/length counts; [256] = end of block
    uint32_t crc_tab[256];
    uint32_t ctab[256];                   // code (bit-reversed, LSB first) | length << 16 per literal
    uint32_t sw[288];                     // Huffman: weights of the leaves in sorted order
    uint32_t w_int[288];                  // ... of the internal nodes in creation order
    uint32_t scan_tmp[kZW + 1];
    uint32_t crc_w[kZW];
    uint32_t bl[16];                      // codes per length
    uint32_t cl_cnt[19];                  // counts of the code-length alphabet
    uint16_t sorted[288];                 // symbols with count > 0, ascending count
    uint16_t par_leaf[288], par_int[288]; // parent (internal node index) of leaves / internal nodes
    uint16_t code[288];
    uint16_t cl_code[19];
    alignas(16) uint8_t len[320];         // code lengths: [0, 257) literals and end of block, [257, 259) the two distance codes
    uint8_t d_int[288];
    uint8_t cl_len[19];
    uint32_t maxd, hdr_fixed_bits, hdr_bits, eob, total_bits, stored, crc_out;
};

__device__ __forceinline__ uint32_t crc_apply(const uint32_t* m, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) r ^= ((v >> i) & 1u) ? m[i] : 0u;
    return r;
}
__device__ __forceinline__ uint32_t crc_advance_lvl(uint32_t j, uint32_t v) {
    const uint32_t (*T)[256] = g_crc_lvl[j];
    return __ldg(&T[0][v & 0xffu]) ^ __ldg(&T[1][(v >> 8) & 0xffu]) ^ __ldg(&T[2][(v >> 16) & 0xffu]) ^ __ldg(&T[3][v >> 24]);
}
// CRC register advanced over n zero bytes
__device__ uint32_t crc_advance(uint32_t v, uint32_t n) {
    for (int j = 0; n; j++, n >>= 1)
        if (n & 1u) v = crc_apply(c_crc_adv[j], v);
    return v;
}

// bits OR-ed into the zeroed block image at a bit position; threads write disjoint bits of shared words
__device__ __forceinline__ void put_bits(uint32_t* w, uint32_t& pos, uint32_t v, uint32_t n) {
    if (!n) return;
    const uint32_t i = pos >> 5, s = pos & 31;
    atomicOr(&w[i], v << s);
    if (s + n > 32) atomicOr(&w[i + 1], v >> (32 - s));
    pos += n;
}

// Two-queue Huffman over the leaves S.sw[0 .. n) (ascending): parents into par_leaf / par_int, the root is
// internal node n - 2.  One thread; the heads of both queues are kept in registers.
__device__ void huff_merge(ZShared& S, uint32_t n) {
    uint32_t i = 0, k = 0, m = 0;
    uint32_t lw = S.sw[0], iw = 0xffffffffu;
    while (m + 1 < n) {
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (i < n && (k >= m || lw <= iw)) {
                w += lw; S.par_leaf[i] = (uint16_t)m; i++;
                lw = i < n ? S.sw[i] : 0xffffffffu;
            } else {
                w += iw; S.par_int[k] = (uint16_t)m; k++;
                iw = k < m ? S.w_int[k] : 0xffffffffu;
            }
        }
        S.w_int[m] = w;
        if (k == m) iw = w;
        m++;
    }
}

// Code lengths limited to max_bits from the tree huff_merge left (one thread): depths top-down, clamped, then the
// count fix-up for over-long codes, then lengths handed out in sorted order (longest code to the rarest symbol).
// Leaves S.bl[] = codes per length.
__device__ void huff_limit(ZShared& S, uint32_t n, uint32_t max_bits, uint8_t* len) {
    for (uint32_t b = 0; b < 16; b++) S.bl[b] = 0;
    int overflow = 0;
    S.d_int[n - 2] = 0;
    for (int q = (int)n - 3; q >= 0; q--) {
        uint32_t d = S.d_int[S.par_int[q]] + 1u;
        if (d > max_bits) { d = max_bits; overflow++; }
        S.d_int[q] = (uint8_t)d;
    }
    for (uint32_t q = 0; q < n; q++) {
        uint32_t d = S.d_int[S.par_leaf[q]] + 1u;
        if (d > max_bits) { d = max_bits; overflow++; }
        S.bl[d]++;
    }
    while (overflow > 0) {
        uint32_t b = max_bits - 1;
        while (S.bl[b] == 0) b--;
        S.bl[b]--; S.bl[b + 1] += 2; S.bl[max_bits]--;
        overflow -= 2;
    }
    // Kraft check; a code that does not add up is replaced by a flat one (never seen, kept as a guard)
    uint32_t kraft = 0;
    for (uint32_t b = 1; b <= max_bits; b++) kraft += S.bl[b] << (max_bits - b);
    if (kraft != (1u << max_bits)) {
        uint32_t flat = 1;
        while ((1u << flat) < n) flat++;
        for (uint32_t b = 0; b < 16; b++) S.bl[b] = 0;
        S.bl[flat] = n;
        for (uint32_t q = 0; q < n; q++) len[S.sorted[q]] = (uint8_t)flat;
        return;
    }
    uint32_t q = 0;
    for (uint32_t b = max_bits; b >= 1; b--)
        for (uint32_t c = S.bl[b]; c; c--) len[S.sorted[q++]] = (uint8_t)b;
}

__device__ __forceinline__ uint32_t rev_bits(uint32_t code, uint32_t n) { return __brev(code) >> (32 - n); }

// first canonical code of length l, from the codes per length
__device__ __forceinline__ uint32_t first_code(const uint32_t* bl, uint32_t l) {
    uint32_t c = 0;
    for (uint32_t q = 1; q <= l; q++) c = (c + (q > 1 ? bl[q - 1] : 0u)) << 1;
    return c;
}

// exclusive prefix of v over the CTA and the total (two barriers)
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t* tmp, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t x = lane < kZW ? tmp[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= (uint32_t)o) xi += y; }
        if (lane < kZW) tmp[lane] = xi - x;
        if (lane == kZW - 1) tmp[kZW] = xi;
    }
    __syncthreads();
    total = tmp[kZW];
    return tmp[warp] + incl - v;
}

// pass 1 of a thread: its 128 contiguous bytes, 16 at a time from global memory -> warp histogram, chunk CRC
template <bool FULL>
__device__ __forceinline__ uint32_t pass1(const uint8_t* chunk, uint32_t my_len, uint32_t* hist, const uint32_t* tab) {
    uint32_t reg = 0;
    uint4 d[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        d[k] = (FULL || 16u * k < my_len) ? __ldg(reinterpret_cast<const uint4*>(chunk) + k) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (FULL || (uint32_t)(k * 16 + q * 4 + r) < my_len) {
                    const uint32_t b = (w[q] >> (8 * r)) & 0xffu;
                    atomicAdd(&hist[b], 1u);
                    reg = tab[(reg ^ b) & 0xffu] ^ (reg >> 8);
                }
    }
    return reg;
}

// pass 2 of a warp: its segment (4096 bytes, fewer at the end of the file) again, this time 4 bytes per lane and
// 128 contiguous bytes per round; the codes of a lane's 4 bytes are joined (<= 60 bits), a warp scan of the
// lengths places them, and they are OR-ed into the block image.
template <bool FULL>
__device__ __forceinline__ void pack_warp(const uint8_t* seg, uint32_t seg_len, const uint32_t* ctab, uint32_t* stage,
                                          uint32_t base, uint32_t lane) {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(seg) + lane;
    for (uint32_t r0 = 0; r0 < 32; r0 += 4) {
        if (!FULL && r0 * 128u >= seg_len) break;
        uint32_t w[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; u++)
            w[u] = (FULL || (r0 + u) * 128u + lane * 4u < seg_len) ? __ldg(g + (r0 + u) * 32u) : 0u;
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t off = (r0 + u) * 128u + lane * 4u;
            const uint32_t nv = FULL ? 4u : (off >= seg_len ? 0u : min(4u, seg_len - off));
            uint32_t e0 = ctab[w[u] & 0xffu], e1 = ctab[(w[u] >> 8) & 0xffu], e2 = ctab[(w[u] >> 16) & 0xffu], e3 = ctab[w[u] >> 24];
            if (!FULL) { if (nv < 1) e0 = 0; if (nv < 2) e1 = 0; if (nv < 3) e2 = 0; if (nv < 4) e3 = 0; }
            const uint32_t l0 = e0 >> 16, l1 = e1 >> 16, l2 = e2 >> 16, l3 = e3 >> 16;
            const uint32_t c01 = (e0 & 0xffffu) | (e1 & 0xffffu) << l0, l01 = l0 + l1;
            const uint32_t c23 = (e2 & 0xffffu) | (e3 & 0xffffu) << l2, l23 = l2 + l3;
            const uint64_t c = (uint64_t)c01 | (uint64_t)c23 << l01;
            const uint32_t l = l01 + l23;
            uint32_t incl = l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
            const uint32_t pos = base + incl - l;
            base += __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t i = pos >> 5, sh = pos & 31;
            const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
            if (l) atomicOr(&stage[i], lo << sh);
            if (sh + l > 32) atomicOr(&stage[i + 1], __funnelshift_l(lo, hi, sh));
            if (sh + l > 64) atomicOr(&stage[i + 2], __funnelshift_l(hi, 0u, sh));
        }
    }
}

// The run of equal code lengths starting at a position, as the tokens of RFC 1951 section 3.2.7:
// zeros: 18 (11..138 zeros, 7 extra bits), 17 (3..10, 3 bits), else literal 0s; a length v: v once, then
// 16 (repeat 3..6 times, 2 bits), else literal v's.
struct RunTok {
    uint32_t v, n_big, big_extra_last, n_mid, mid_extra, n_lit;   // zeros: big = 18, mid = 17; else big = 16, mid unused
};
__device__ __forceinline__ RunTok run_tokens(uint32_t v, uint32_t run) {
    RunTok r{v, 0, 0, 0, 0, 0};
    if (v == 0) {
        r.n_big = run / 138; uint32_t rem = run % 138;
        if (rem >= 11) { r.n_big++; r.big_extra_last = rem - 11; rem = 0; } else r.big_extra_last = 127;
        if (rem >= 3) { r.n_mid = 1; r.mid_extra = rem - 3; rem = 0; }
        r.n_lit = rem;
    } else {
        uint32_t rem = run - 1;
        r.n_big = rem / 6; rem %= 6;
        if (rem >= 3) { r.n_big++; r.big_extra_last = rem - 3; rem = 0; } else r.big_extra_last = 3;
        r.n_lit = 1 + rem;
    }
    return r;
}

__global__ void __launch_bounds__(kZT, 3)
k_bgzf(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, const uint64_t* __restrict__ totals, uint32_t nblk_max,
       uint8_t* __restrict__ slots0, uint8_t* __restrict__ slots1, uint32_t* __restrict__ zlen0, uint32_t* __restrict__ zlen1) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ZShared& S = *reinterpret_cast<ZShared*>(smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // the grid is sized for the largest batch; the FASTQ byte counts of this one are on the device
    const bool second = blockIdx.x >= nblk_max;
    const uint32_t b = second ? blockIdx.x - nblk_max : blockIdx.x;
    const uint64_t n_all = totals[second ? 1 : 0];
    if ((uint64_t)b * kBgzfIn >= n_all) return;
    const uint8_t* in = (second ? in1 : in0) + (uint64_t)b * kBgzfIn;
    const uint32_t len = (uint32_t)min((uint64_t)kBgzfIn, n_all - (uint64_t)b * kBgzfIn);
    uint8_t* slot = (second ? slots1 : slots0) + (uint64_t)b * kBgzfSlot;
    uint32_t* zlen = (second ? zlen1 : zlen0) + b;

    const uint32_t my_off = t * kChunk;
    const uint32_t my_len = my_off >= len ? 0u : min(kChunk, len - my_off);

    // ---- clear the image and the histograms
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.stage);
        for (uint32_t i = t; i < kStageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        if (t < 256) S.crc_tab[t] = c_crc_tab[t];
        if (t < 320) S.len[t] = (t == 257 || t == 258) ? 1 : 0;   // two unused distance codes of one bit, as zlib sends for a block of literals
        if (t < 16) S.bl[t] = 0;
        if (t < 19) { S.cl_cnt[t] = 0; S.cl_len[t] = 0; }
        if (t == 0) S.maxd = 0;
    }
    __syncthreads();

    // ---- pass 1: histogram + chunk CRC
    uint32_t* hist = S.stage + kHistWord0 + warp * 256;
    uint32_t crc = my_len == kChunk ? pass1<true>(in + my_off, my_len, hist, S.crc_tab) : pass1<false>(in + my_off, my_len, hist, S.crc_tab);
    // CRC tree inside the warp: the node at lane covers chunks [t, t + 2s); its right half has right_len bytes
#pragma unroll
    for (uint32_t j = 0; j < 5; j++) {
        const uint32_t s = 1u << j;
        const uint32_t other = __shfl_down_sync(0xffffffffu, crc, s);
        if ((lane & (2 * s - 1)) == 0) {
            const uint32_t r0 = (t + s) * kChunk;
            const uint32_t right_len = r0 >= len ? 0u : min(s * kChunk, len - r0);
            crc = (right_len == s * kChunk ? crc_advance_lvl(j, crc) : crc_advance(crc, right_len)) ^ other;
        }
    }
    if (lane == 0) S.crc_w[warp] = crc;
    __syncthreads();
    uint32_t my_cnt = 0;
    if (t <= 256) {
        if (t < 256) for (int w = 0; w < kZW; w++) my_cnt += S.stage[kHistWord0 + w * 256 + t];
        else my_cnt = 1;     // end of block
        S.cnt[t] = my_cnt;
    }
    // the active symbols, compacted in symbol order: count << 9 | symbol orders them by (count, symbol)
    const uint32_t act = __ballot_sync(0xffffffffu, my_cnt != 0);
    if (lane == 0) S.scan_tmp[warp] = __popc(act);
    __syncthreads();
    uint32_t n_active = 0, before = 0;
#pragma unroll
    for (uint32_t w = 0; w < 9; w++) { const uint32_t c = S.scan_tmp[w]; n_active += c; before += w < warp ? c : 0u; }
    if (my_cnt) S.w_int[before + __popc(act & ((1u << lane) - 1u))] = my_cnt << 9 | t;
    __syncthreads();
    // rank sort (ascending count, ties by symbol)
    if (my_cnt) {
        const uint32_t key = my_cnt << 9 | t;
        uint32_t rank = 0;
        for (uint32_t q = 0; q < n_active; q++) rank += S.w_int[q] < key ? 1u : 0u;
        S.sorted[rank] = (uint16_t)t;
        S.sw[rank] = my_cnt;
    }
    if (warp == kZW - 1) {   // meanwhile the last warp (no symbols of its own) finishes the CRC over the warps' results
        crc = lane < kZW ? S.crc_w[lane] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t s = 1u << j;
            const uint32_t other = __shfl_down_sync(0xffffffffu, crc, s);
            if ((lane & (2 * s - 1)) == 0 && lane < kZW) {
                const uint32_t r0 = (lane + s) * 32u * kChunk;
                const uint32_t right_len = r0 >= len ? 0u : min(s * 32u * kChunk, len - r0);
                crc = (right_len == s * 32u * kChunk ? crc_advance_lvl(5 + j, crc) : crc_advance(crc, right_len)) ^ other;
            }
        }
        if (lane == 0) S.crc_out = crc ^ (len == kBgzfIn ? c_crc_init_full : crc_advance(0xffffffffu, len)) ^ 0xffffffffu;
    }
    __syncthreads();

    // ---- literal code lengths: the tree by one thread, the depths by one thread per leaf
    if (t == 0) huff_merge(S, n_active);       // n_active >= 2: a literal and the end of block
    __syncthreads();
    if (t < n_active) {
        uint32_t p = S.par_leaf[t], dep = 1;
        const uint32_t root = n_active - 2;
        while (p != root) { p = S.par_int[p]; dep++; }
        S.len[S.sorted[t]] = (uint8_t)min(dep, 15u);
        atomicAdd(&S.bl[min(dep, 15u)], 1u);
        if (dep > 15) atomicMax(&S.maxd, dep);
    }
    __syncthreads();
    if (S.maxd > 15) {                         // rare: the 15-bit limit has to act
        if (t == 0) huff_limit(S, n_active, 15, S.len);
        __syncthreads();
    }
    // canonical codes: first code of the length + the symbols of the same length before this one
    if (t <= 256) {
        const uint32_t l = S.len[t];
        if (l) {
            const uint32_t pat = l * 0x01010101u;
            const uint32_t* lw = reinterpret_cast<const uint32_t*>(S.len);
            uint32_t same = 0;
            for (uint32_t j = 0; j < t / 4; j++) same += __popc(__vcmpeq4(lw[j], pat));
            same += __popc(__vcmpeq4(lw[t / 4], pat) & ((1u << (8 * (t & 3))) - 1u));
            S.code[t] = (uint16_t)rev_bits(first_code(S.bl, l) + same / 8, l);
        }
        if (t < 256) S.ctab[t] = l ? (uint32_t)S.code[t] | l << 16 : 0u;
    }
    // the runs of the code-length sequence [0, 259): one thread per run counts its tokens
    RunTok rt{0, 0, 0, 0, 0, 0};
    bool run_start = false;
    if (t < 259) {
        const uint32_t v = S.len[t];
        run_start = t == 0 || S.len[t - 1] != v;
        if (run_start) {
            uint32_t run = 1;
            while (t + run < 259 && S.len[t + run] == v) run++;
            rt = run_tokens(v, run);
            if (v == 0) {
                if (rt.n_big) atomicAdd(&S.cl_cnt[18], rt.n_big);
                if (rt.n_mid) atomicAdd(&S.cl_cnt[17], 1u);
                if (rt.n_lit) atomicAdd(&S.cl_cnt[0], rt.n_lit);
            } else {
                if (rt.n_big) atomicAdd(&S.cl_cnt[16], rt.n_big);
                atomicAdd(&S.cl_cnt[v], rt.n_lit);
            }
        }
    }
    __syncthreads();

    // ---- the code-length code and the fixed part of the block header (warp 0) while the others size their segments
    if (warp == 0) {
        const uint32_t c = lane < 19 ? S.cl_cnt[lane] : 0u;
        const uint32_t n_cl = __popc(__ballot_sync(0xffffffffu, c != 0));
        const uint32_t key = c << 5 | lane;
        uint32_t rank = 0;
#pragma unroll
        for (uint32_t q = 0; q < 19; q++) {
            const uint32_t k2 = __shfl_sync(0xffffffffu, key, q);
            rank += (k2 >> 5) != 0 && k2 < key ? 1u : 0u;
        }
        if (c) { S.sorted[rank] = (uint16_t)lane; S.sw[rank] = c; }
        if (lane < 16) S.bl[lane] = 0;
        __syncwarp();
        if (n_cl == 1) {                       // a complete code needs two symbols
            if (lane == 0) { S.cl_len[S.sorted[0]] = 1; S.cl_len[S.sorted[0] ? 0 : 1] = 1; S.bl[1] = 2; }
        } else {
            if (lane == 0) huff_merge(S, n_cl);
            __syncwarp();
            uint32_t dep = 0;
            if (lane < n_cl) {
                uint32_t p = S.par_leaf[lane];
                dep = 1;
                while (p != n_cl - 2) { p = S.par_int[p]; dep++; }
                S.cl_len[S.sorted[lane]] = (uint8_t)min(dep, 7u);
                atomicAdd(&S.bl[min(dep, 7u)], 1u);
            }
            if (__any_sync(0xffffffffu, dep > 7)) {   // the 7-bit limit has to act
                __syncwarp();
                if (lane == 0) huff_limit(S, n_cl, 7, S.cl_len);
            }
        }
        __syncwarp();
        const uint32_t l = lane < 19 ? S.cl_len[lane] : 0u;
        uint32_t same = 0;
#pragma unroll
        for (uint32_t q = 0; q < 19; q++) {
            const uint32_t l2 = __shfl_sync(0xffffffffu, l, q);
            same += (l2 == l && q < lane) ? 1u : 0u;
        }
        if (lane < 19) S.cl_code[lane] = (uint16_t)(l ? rev_bits(first_code(S.bl, l) + same, l) : 0u);
        const uint32_t order_lane = lane < 19 ? (uint32_t)"\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f"[lane] : 0u;
        // HCLEN: the lengths are sent in the order 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 up to the last one used
        const uint32_t l_ord = __shfl_sync(0xffffffffu, l, order_lane);
        const uint32_t used = __ballot_sync(0xffffffffu, lane < 19 && l_ord != 0);
        const uint32_t hclen = max(4u, 32u - (uint32_t)__clz(used));
        if (lane == 0) {
            uint32_t pos = kHdr * 8;
            put_bits(S.stage, pos, 1, 1); put_bits(S.stage, pos, 2, 2);                   // BFINAL, BTYPE = dynamic Huffman
            put_bits(S.stage, pos, 0, 5); put_bits(S.stage, pos, 1, 5); put_bits(S.stage, pos, hclen - 4, 4);   // HLIT = 257, HDIST = 2
            S.hdr_fixed_bits = pos + 3 * hclen;
            S.eob = (uint32_t)S.code[256] | (uint32_t)S.len[256] << 16;
        }
        if (lane < hclen) { uint32_t pos = kHdr * 8 + 17 + 3 * lane; put_bits(S.stage, pos, l_ord, 3); }
    }
    // bits of this warp's segment: its histogram times the code lengths
    uint32_t bits = 0;
#pragma unroll
    for (uint32_t q = 0; q < 8; q++) bits += hist[lane + 32 * q] * (S.ctab[lane + 32 * q] >> 16);
#pragma unroll
    for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if (lane) bits = 0;
    __syncthreads();

    // ---- one scan for both the header tokens of the runs and the chunks: header bits << 20 | chunk bits
    uint32_t hb = 0;
    uint32_t l_big = 0, c_big = 0, l_mid = 0, c_mid = 0, l_lit = 0, c_lit = 0;
    if (run_start) {
        const uint32_t big = rt.v == 0 ? 18u : 16u, xb = rt.v == 0 ? 7u : 2u;
        l_big = S.cl_len[big]; c_big = S.cl_code[big];
        l_mid = S.cl_len[17]; c_mid = S.cl_code[17];
        l_lit = S.cl_len[rt.v]; c_lit = S.cl_code[rt.v];
        hb = rt.n_big * (l_big + xb) + rt.n_mid * (l_mid + 3u) + rt.n_lit * l_lit;
    }
    uint32_t total;
    const uint32_t excl = block_scan(hb << 20 | bits, S.scan_tmp, total);
    const uint32_t hdr_bits = S.hdr_fixed_bits + (total >> 20);
    const uint32_t eob = S.eob;
    const uint32_t end_bits = hdr_bits + (total & 0xfffffu) + (eob >> 16);
    const bool stored = ((end_bits + 7) / 8 - kHdr) >= len + 5;     // uniform
    uint32_t bitpos, total_bits;
    if ((stored ? kHdr + 5 + len : (end_bits + 7) / 8) + 8 > kHistWord0 * 4) {
        // the image reaches into the histograms (every warp has read its own before the scan's barriers)
        for (uint32_t i = kHistWord0 + t; i < kStageWords; i += kZT) S.stage[i] = 0;
        __syncthreads();
    }
    if (!stored) {
        if (run_start) {
            uint32_t pos = S.hdr_fixed_bits + (excl >> 20);
            if (rt.v == 0) {
                for (uint32_t q = 0; q < rt.n_big; q++) { put_bits(S.stage, pos, c_big, l_big); put_bits(S.stage, pos, q + 1 == rt.n_big ? rt.big_extra_last : 127u, 7); }
                if (rt.n_mid) { put_bits(S.stage, pos, c_mid, l_mid); put_bits(S.stage, pos, rt.mid_extra, 3); }
                for (uint32_t q = 0; q < rt.n_lit; q++) put_bits(S.stage, pos, c_lit, l_lit);
            } else {
                put_bits(S.stage, pos, c_lit, l_lit);
                for (uint32_t q = 0; q < rt.n_big; q++) { put_bits(S.stage, pos, c_big, l_big); put_bits(S.stage, pos, q + 1 == rt.n_big ? rt.big_extra_last : 3u, 2); }
                for (uint32_t q = 1; q < rt.n_lit; q++) put_bits(S.stage, pos, c_lit, l_lit);
            }
        }
        bitpos = hdr_bits + (__shfl_sync(0xffffffffu, excl, 0) & 0xfffffu);
        total_bits = end_bits;
    } else {
        // does not shrink: stored block (BTYPE 00) through the same packing code with the identity table
        const uint32_t clear_words = (S.hdr_fixed_bits + 31) / 32 + 1;
        __syncthreads();
        for (uint32_t i = kHdr / 4 + t; i < clear_words; i += kZT) S.stage[i] = 0;
        if (t < 256) S.ctab[t] = t | 8u << 16;
        __syncthreads();
        if (t == 0) {
            uint32_t pos = kHdr * 8;
            put_bits(S.stage, pos, 1, 8);                              // BFINAL = 1, BTYPE = 00, padding to the byte
            put_bits(S.stage, pos, len, 16); put_bits(S.stage, pos, len ^ 0xffffu, 16);
        }
        bitpos = (kHdr + 5 + warp * 32u * kChunk) * 8;
        total_bits = (kHdr + 5 + len) * 8;
    }
    {
        const uint32_t seg_off = warp * 32u * kChunk;
        const uint32_t seg_len = seg_off >= len ? 0u : min(32u * kChunk, len - seg_off);
        if (seg_len == 32u * kChunk) pack_warp<true>(in + seg_off, seg_len, S.ctab, S.stage, bitpos, lane);
        else if (seg_len) pack_warp<false>(in + seg_off, seg_len, S.ctab, S.stage, bitpos, lane);
    }
    const uint32_t body_end = (total_bits + 7) / 8;
    const uint32_t total_bytes = body_end + 8;
    if (t == 0) {
        uint32_t pos = total_bits - (eob >> 16);
        if (!stored) put_bits(S.stage, pos, eob & 0xffffu, eob >> 16);
        pos = body_end * 8;
        put_bits(S.stage, pos, S.crc_out, 32); put_bits(S.stage, pos, len, 32);
        // member header: ID1 ID2 CM FLG(FEXTRA) MTIME XFL OS(255) XLEN=6 'B' 'C' SLEN=2 BSIZE = total - 1
        S.stage[0] = 0x04088b1fu; S.stage[1] = 0; S.stage[2] = 0x0006ff00u; S.stage[3] = 0x00024342u;
        atomicOr(&S.stage[4], (total_bytes - 1) & 0xffffu);
        *zlen = total_bytes;
    }
    __syncthreads();
    const uint4* s4 = reinterpret_cast<const uint4*>(S.stage);
    uint4* o4 = reinterpret_cast<uint4*>(slot);
    for (uint32_t i = t; i < (total_bytes + 15) / 16; i += kZT) o4[i] = s4[i];
}

// exclusive prefix of the member sizes of one file (one CTA per file); the file's compressed size into totals[2 + file]
__global__ void __launch_bounds__(1024)
k_bgzf_scan(const uint32_t* __restrict__ zlen0, const uint32_t* __restrict__ zlen1, uint64_t* __restrict__ zoff0,
            uint64_t* __restrict__ zoff1, uint64_t* __restrict__ totals) {
    __shared__ uint64_t wsum[32];
    const uint32_t* zlen = blockIdx.x ? zlen1 : zlen0;
    uint64_t* zoff = blockIdx.x ? zoff1 : zoff0;
    const uint32_t n = (uint32_t)((totals[blockIdx.x] + kBgzfIn - 1) / kBgzfIn);
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    uint64_t mine = 0;
    for (uint32_t i = lo; i < hi; i++) mine += zlen[i];
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint64_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (t == 0) {
        uint64_t run = 0;
        for (int w = 0; w < 32; w++) { const uint64_t v = wsum[w]; wsum[w] = run; run += v; }
        totals[2 + blockIdx.x] = run;
    }
    __syncthreads();
    uint64_t run = wsum[warp] + incl - mine;
    for (uint32_t i = lo; i < hi; i++) { zoff[i] = run; run += zlen[i]; }
}

// member b of a file moves from its slot to its place in the contiguous file image
__global__ void __launch_bounds__(256)
k_bgzf_gather(const uint8_t* __restrict__ slots0, const uint8_t* __restrict__ slots1, const uint32_t* __restrict__ zlen0,
              const uint32_t* __restrict__ zlen1, const uint64_t* __restrict__ zoff0, const uint64_t* __restrict__ zoff1,
              const uint64_t* __restrict__ totals, uint32_t nblk_max, uint8_t* __restrict__ out0, uint8_t* __restrict__ out1) {
    const bool second = blockIdx.x >= nblk_max;
    const uint32_t b = second ? blockIdx.x - nblk_max : blockIdx.x;
    if ((uint64_t)b * kBgzfIn >= totals[second ? 1 : 0]) return;
    const uint8_t* src = (second ? slots1 : slots0) + (uint64_t)b * kBgzfSlot;
    const uint32_t n = (second ? zlen1 : zlen0)[b];
    uint8_t* dst = (second ? out1 : out0) + (second ? zoff1 : zoff0)[b];
    const uint32_t t = threadIdx.x;
    // destination-aligned 16-byte stores; the source words are funnel-shifted into place
    const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    if (t < head) dst[t] = src[t];
    const uint32_t body = (n - head) / 16;
    const uint32_t sh = (head & 3u) * 8u;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = t; i < body; i += 256) {
        const uint32_t* p = s32 + 4 * i;
        const uint32_t a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3], a4 = sh ? p[4] : 0u;
        d4[i] = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh),
                           __funnelshift_r(a3, a4, sh));
    }
    const uint32_t done = head + body * 16;
    if (t < n - done) dst[done + t] = src[done + t];
}

}  // namespace

cudaError_t bgzf_init() {
    uint32_t tab[256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        tab[i] = c;
    }
    uint32_t adv[17][32];
    for (int i = 0; i < 32; i++) { const uint32_t v = 1u << i; adv[0][i] = tab[v & 0xffu] ^ (v >> 8); }
    for (int j = 1; j < 17; j++)
        for (int i = 0; i < 32; i++) {
            uint32_t v = adv[j - 1][i], r = 0;
            for (int q = 0; q < 32; q++) if ((v >> q) & 1u) r ^= adv[j - 1][q];
            adv[j][i] = r;
        }
    cudaError_t e = cudaMemcpyToSymbol(c_crc_tab, tab, sizeof tab);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_crc_adv, adv, sizeof adv);
    if (e != cudaSuccess) return e;
    uint32_t init_full = 0xffffffffu;
    for (int j = 0; j < 17; j++)
        if ((kBgzfIn >> j) & 1u) {
            uint32_t r = 0;
            for (int q = 0; q < 32; q++) if ((init_full >> q) & 1u) r ^= adv[j][q];
            init_full = r;
        }
    e = cudaMemcpyToSymbol(c_crc_init_full, &init_full, sizeof init_full);
    if (e != cudaSuccess) return e;
    static uint32_t lvl[9][4][256];
    for (int j = 0; j < 9; j++)
        for (int k = 0; k < 4; k++)
            for (uint32_t x = 0; x < 256; x++) {
                const uint32_t v = x << (8 * k);
                uint32_t r = 0;
                for (int q = 0; q < 32; q++) if ((v >> q) & 1u) r ^= adv[7 + j][q];
                lvl[j][k][x] = r;
            }
    e = cudaMemcpyToSymbol(g_crc_lvl, lvl, sizeof lvl);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_bgzf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZShared));
}

cudaError_t launch_bgzf(const uint8_t* in0, const uint8_t* in1, uint64_t* totals, uint32_t nblk_max, uint8_t* slots0,
                        uint8_t* slots1, uint32_t* zlen0, uint32_t* zlen1, uint64_t* zoff0, uint64_t* zoff1, uint8_t* out0,
                        uint8_t* out1, cudaStream_t s) {
    if (nblk_max) k_bgzf<<<2 * nblk_max, kZT, sizeof(ZShared), s>>>(in0, in1, totals, nblk_max, slots0, slots1, zlen0, zlen1);
    k_bgzf_scan<<<2, 1024, 0, s>>>(zlen0, zlen1, zoff0, zoff1, totals);
    if (nblk_max)
        k_bgzf_gather<<<2 * nblk_max, 256, 0, s>>>(slots0, slots1, zlen0, zlen1, zoff0, zoff1, totals, nblk_max, out0, out1);
    return cudaGetLastError();
}

}  // namespace jlp
