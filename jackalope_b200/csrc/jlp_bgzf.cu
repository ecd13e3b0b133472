// BGZF on the device: the FASTQ of a batch is compressed where it was generated, so that
// only the compressed bytes cross PCIe (the file format is what
// FileBGZF / bgzip_file write through htslib, /root/reference/src/io.h:58-135,
// src/hts.h:140-180: gzip members (RFC 1952) of at most 0xff00 input bytes carrying the
// "BC" extra field with the member's size, each holding one RFC 1951 deflate stream).
//
// FASTQ is stationary -- every block of a batch has the statistics of every other -- so the Huffman codes are built
// ONCE per file and batch, from the file's first block, and every block is coded with them (each member still
// carries the code in its own dynamic-block header, as the format wants):
//
// k_bgzf_code   one CTA per file.  Block 0: literal histograms (shared atomics), the matches of its lines, and
//               from the counts -- those of the printable bytes, the lengths and the distances raised by one, so
//               that what the sample does not hold still has a code -- the literal/length and distance codes
//               (two-queue Huffman over the rank-sorted symbols, limited to 15 bits, canonical), the code-length
//               code, and the bits of the member + block header, which are the same for every block.  Left in a
//               ZCode in device memory.
// k_bgzf        one CTA of 512 threads per BGZF block, 3 CTAs per SM; thread t owns bytes [128 t, 128 t + 128) in
//               both passes, warp w the segment [4096 w, 4096 w + 4096); between the first barrier and the one at
//               which the sizes meet a warp synchronises with nobody but itself.
//   pass 1  the CRC-32 of the chunk (slicing by four; the warp's tree of "advance over 2^j zero bytes" operators
//           gives the segment's, which is advanced over the bytes behind the segment: the block's is the XOR of the
//           sixteen), its newlines as a bit mask, and the bits its bytes take as literals
//   match   (levels 4-6) the line starts of the segment by a warp scan of the newline counts, the four before the
//           segment from the newlines of the 1024 bytes in front of it; every thread compares the line starts of its
//           chunk with the line start 4 lines earlier -- in FASTQ the same line of the previous record -- and a
//           common prefix of >= 4 bytes becomes one length/distance pair (it ends with its line and with the
//           segment; two per chunk are kept): a bit mask marks the bytes so covered, the chunk sizes lose their
//           literals and gain the pair
//   sizes   a warp scan of the chunk sizes, the segments' sums through shared memory: every chunk's bit position
//   pass 2  every thread codes its chunk serially into a 64-bit window that is emptied, a word at a time, into the
//           member's image in shared memory at the chunk's own bit position; the image leaves with 128-bit stores
//   a block that holds a byte without a code, or would not shrink, is left to
// k_bgzf_own    the same block with a Huffman code of its own (histograms in pass 1, the code construction of
//               k_bgzf_code between the passes, warp-cooperative pass 2); stores the block when it does not shrink.
// The member lands in a 64 KiB slot; k_bgzf_scan / k_bgzf_gather then make the file contiguous.
#ifdef JLP_CPU_EMU      // tests/emu: the kernels of this file on host threads (CPU test suite), see tests/emu/cuda_emu.h
#include "cuda_emu.h"
namespace jlp {
constexpr uint32_t kBgzfIn = 0xff00;
constexpr uint32_t kBgzfSlot = 65536;
constexpr uint32_t kBgzfCodeBytes = 4096;
}
#define JLP_DYN_SMEM(name) uint8_t* name = jlp_emu::dyn_smem()
#define JLP_LAUNCH(k, grid, block, smem, stream, ...) jlp_emu::launch(grid, block, smem, [&]() { k(__VA_ARGS__); })
#else
#include "jlp_kernels.cuh"
#define JLP_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define JLP_LAUNCH(k, grid, block, smem, stream, ...) k<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <cstddef>
#include <cstring>

namespace jlp {


namespace {

constexpr int kZT = 512;                 // threads per CTA
constexpr int kZW = kZT / 32;
constexpr uint32_t kChunk = 128;         // input bytes per thread
constexpr uint32_t kSeg = 32 * kChunk;   // input bytes per warp
constexpr uint32_t kNLit = 286;          // literal/length alphabet
constexpr uint32_t kNDist = 30;
constexpr uint32_t kHdr = 18;            // BGZF member header bytes
constexpr uint32_t kStageBytes = 44 * 1024;                // the image of a compressed block; larger ones are stored
constexpr uint32_t kStageWords = kStageBytes / 4;
constexpr uint32_t kHistWord0 = kStageWords - kZW * 256;   // the histograms alias the end of the image
constexpr uint32_t kLsCap = 1024;        // line starts kept per block (two per thread)
constexpr uint32_t kMCap = 1024;         // matches kept per block
constexpr uint32_t kSlots = 2;           // k_bgzf: matches kept per chunk (kSlots * kZT = kMCap)
constexpr uint32_t kWls = 96;            // k_bgzf: line starts kept per segment, the kBack before it included
constexpr uint32_t kLookBack = 1024;     // k_bgzf: bytes before a segment searched for those
constexpr uint32_t kNoLine = 0xffffu;
constexpr uint32_t kBack = 4;            // a line is compared with the line this many lines earlier
constexpr uint32_t kMaskWords = kBgzfIn / 32 + 2;
constexpr uint32_t kImageBytes = 48 * 1024;   // k_bgzf: the image of a member in shared memory
constexpr uint32_t kImageWords = kImageBytes / 4;
constexpr uint32_t kHdrWords = 160;      // 18 bytes of member header + at most 4498 bits of block header
constexpr uint32_t kOwnCode = 0xffffffffu;   // in zlen[]: k_bgzf leaves this block to k_bgzf_own
constexpr uint32_t kNoCode = 4096u << 16;    // more bits than a chunk of coded bytes can take

// what k_bgzf_code leaves for k_bgzf, one per file
struct ZCode {
    uint32_t ctab[256];        // literal: code (bit-reversed, LSB first) | length << 16; kNoCode: the code has no word for it
    uint32_t lcode[32];        // length symbol 257 + i, likewise
    uint32_t dcode[32];        // distance symbol i
    uint32_t eob;              // end of block
    uint32_t hdr_bits;         // bits in use of hdr[]
    uint32_t pad[2];
    uint32_t hdr[kHdrWords];   // the member header (BSIZE left zero) and the dynamic block's header
};
static_assert(2 * sizeof(ZCode) <= kBgzfCodeBytes, "launch_bgzf's code scratch");
static_assert(kHdrWords % 4 == 0 && kHdrWords < kHistWord0, "the header words are copied, the rest of the image is cleared as uint4");

__constant__ uint32_t c_crc_tab[4][256];  // slicing by four: [k][b] = the register after byte b and k zero bytes
__constant__ uint32_t c_crc_adv[17][32];   // operator "advance the CRC register over 2^j zero bytes", by bit image
__constant__ uint32_t c_crc_init_full;     // the initial register 0xffffffff advanced over a full block
// the same operators for 2^7 .. 2^15 zero bytes (the sizes a full block's CRC tree combines) as 4 x 256-entry tables:
// advance(v) = T[0][v & 255] ^ T[1][v >> 8 & 255] ^ T[2][v >> 16 & 255] ^ T[3][v >> 24]
__device__ uint32_t g_crc_lvl[10][4][256];     // [9]: 3840 zero bytes, the last segment of a full block

// scratch of one Huffman construction
struct HuffScratch {
    uint32_t* sw;        // weights of the leaves in sorted order
    uint32_t* w_int;     // ... of the internal nodes in creation order
    uint16_t* sorted;    // symbols with count > 0, ascending count
    uint16_t* par_leaf;  // parent (internal node index) of leaves
    uint16_t* par_int;   // ... of internal nodes
    uint8_t* d_int;
    uint32_t* bl;        // codes per length
};

// shared memory of k_bgzf_code and k_bgzf_own
struct ZShared {
    uint32_t stage[kStageWords];          // the block image; its last 16 KiB hold the per-warp literal counts until the packing starts
    uint32_t mask[kMaskWords];            // bit q + 1: byte q of the block is covered by a match
    uint64_t mbits[kMCap];                // the matches in block order: first their tokens (match_token), then their bits | bit count << 56
    uint32_t cnt[288];                    // literal/length counts; [256] = end of block, [257, 286) lengths
    uint32_t crc_tab[4][256];
    uint32_t ctab[256];                   // code (bit-reversed, LSB first) | length << 16 per literal
    uint32_t sw[288];                     // Huffman: weights of the leaves in sorted order
    uint32_t w_int[288];                  // ... of the internal nodes in creation order
    uint32_t scan_tmp[kZW + 1];
    uint32_t crc_w[kZW];
    uint32_t wm[kZW];                     // matches that start in each warp's segment
    uint32_t run_bal[kZW];                // bit per position of the code-length sequence: a run starts here
    uint32_t bl[16];                      // codes per length
    uint32_t dcnt[32];                    // counts of the distance alphabet
    uint32_t cl_cnt[19];                  // counts of the code-length alphabet
    uint16_t ls[kLsCap + 1];              // line starts, ascending; ls[0] = 0
    uint16_t sorted[288];                 // symbols with count > 0, ascending count
    uint16_t par_leaf[288], par_int[288]; // parent (internal node index) of leaves / internal nodes
    uint16_t code[288];
    uint16_t dcode[32];
    uint16_t cl_code[19];
    alignas(16) uint8_t len[288];         // code lengths of literals, end of block and lengths
    uint8_t dlen[32];
    uint8_t d_int[288];
    uint8_t cl_len[19];
    // a second, small scratch: the distance code is built by warp 1 while thread 0 builds the literal/length tree
    uint32_t sw2[32], w_int2[32], bl2[16];
    uint16_t sorted2[32], par_leaf2[32], par_int2[32];
    uint8_t d_int2[32];
    uint32_t maxd, hlit, hdist, hdr_fixed_bits, eob, crc_out;
};

static_assert(sizeof(ZShared) <= 75 * 1024, "three CTAs per SM");
static_assert(offsetof(ZShared, mbits) % 8 == 0, "mbits alignment");

// The literal/length code limited to 15 bits by the whole CTA (uniform call).  S.bl[] holds the leaves per depth, the
// ones deeper than 15 counted at 15, so the Kraft sum exceeds one by E units of 2^-15.  zlib's step (gen_bitlen) --
// a leaf of the deepest level above 15 moves one down and takes a leaf of level 15 up beside it -- lowers the sum by
// exactly one unit, so E steps make the code complete again; then the lengths are handed out in sorted order, the
// longest code to the rarest symbol, by one thread per symbol.
__device__ __forceinline__ void limit_15(ZShared& S, uint32_t n, uint32_t t) {
    if (t == 0) {
        uint32_t kraft = 0;
        for (uint32_t b = 1; b <= 15; b++) kraft += S.bl[b] << (15 - b);
        for (int32_t e = (int32_t)kraft - (1 << 15); e > 0; e--) {
            uint32_t b = 14;
            while (S.bl[b] == 0) b--;
            S.bl[b]--; S.bl[b + 1] += 2; S.bl[15]--;
        }
    }
    __syncthreads();
    if (t < n) {
        uint32_t b = 15, cum = S.bl[15];
        while (t >= cum && b > 1) cum += S.bl[--b];
        S.len[S.sorted[t]] = (uint8_t)b;
    }
    __syncthreads();
}

// shared memory of k_bgzf
struct ZMain {
    uint32_t image[kImageWords];          // the member: the header words, then the bits of every chunk at their final place
    uint64_t mbits[kMCap];                // the matches that start in chunk t, in order, from kSlots * t: bits | bit count << 56
    uint32_t mask[kMaskWords];            // bit q + 1: byte q of the block is covered by a match
    uint32_t crc_tab[4][256];
    uint32_t ctab[256];
    uint32_t lcode[32], dcode[32];
    uint32_t cbits[kZT];                  // per chunk: the bits it sends
    uint32_t wtot[kZW];                   // per segment: the bits it sends
    uint32_t crc_w[kZW];                  // per segment: its CRC register advanced over the bytes behind the segment
    uint32_t eob, hdr_bits, miss;
    uint16_t ls[kZW][kWls];               // per segment: the last kBack line starts before it (kNoLine: none), then its own, ascending
};
static_assert(sizeof(ZMain) <= 75 * 1024, "three CTAs per SM");
static_assert(offsetof(ZMain, mbits) % 8 == 0, "mbits alignment");

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
__device__ void huff_merge(const HuffScratch& S, uint32_t n) {
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
__device__ void huff_limit(const HuffScratch& S, uint32_t n, uint32_t max_bits, uint8_t* len) {
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

// A Huffman code for an alphabet of at most 32 symbols by one warp: lane s holds the count c of symbol s.
// Lengths (<= max_bits) into len[], bit-reversed canonical codes into code[]; an alphabet with fewer than two
// used symbols gets a second, unused one-bit code (a complete code, as zlib sends).  Returns the number of
// used symbols, dummies included, through the highest one + 1.  Uses the CTA's Huffman scratch.
__device__ uint32_t warp_huffman(const HuffScratch& S, uint32_t c, uint32_t n_sym, uint32_t max_bits, uint8_t* len, uint16_t* code,
                                 uint32_t lane) {
    const uint32_t n = __popc(__ballot_sync(0xffffffffu, c != 0));
    const uint32_t key = c << 5 | lane;
    uint32_t rank = 0;
    for (uint32_t q = 0; q < n_sym; q++) {
        const uint32_t k2 = __shfl_sync(0xffffffffu, key, q);
        rank += (k2 >> 5) != 0 && k2 < key ? 1u : 0u;
    }
    if (c) { S.sorted[rank] = (uint16_t)lane; S.sw[rank] = c; }
    if (lane < 16) S.bl[lane] = 0;
    if (lane < n_sym) len[lane] = 0;
    __syncwarp();
    if (n < 2) {
        if (lane == 0) {
            const uint32_t s0 = n ? S.sorted[0] : 0u;
            len[s0] = 1; len[s0 ? 0 : 1] = 1; S.bl[1] = 2;
        }
    } else {
        if (lane == 0) huff_merge(S, n);
        __syncwarp();
        uint32_t dep = 0;
        if (lane < n) {
            uint32_t p = S.par_leaf[lane];
            dep = 1;
            while (p != n - 2) { p = S.par_int[p]; dep++; }
            len[S.sorted[lane]] = (uint8_t)min(dep, max_bits);
            atomicAdd(&S.bl[min(dep, max_bits)], 1u);
        }
        if (__any_sync(0xffffffffu, dep > max_bits)) {   // the length limit has to act
            __syncwarp();
            if (lane == 0) huff_limit(S, n, max_bits, len);
        }
    }
    __syncwarp();
    const uint32_t l = lane < n_sym ? len[lane] : 0u;
    uint32_t same = 0;
    for (uint32_t q = 0; q < n_sym; q++) {
        const uint32_t l2 = __shfl_sync(0xffffffffu, l, q);
        same += (l2 == l && q < lane) ? 1u : 0u;
    }
    if (lane < n_sym) code[lane] = (uint16_t)(l ? rev_bits(first_code(S.bl, l) + same, l) : 0u);
    const uint32_t used = __ballot_sync(0xffffffffu, l != 0);
    __syncwarp();
    return 32u - (uint32_t)__clz(used);
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

// A match as one word: length symbol - 257 (5 bits) | its extra bits' value (5) | distance symbol (5) |
// its extra bits' value (13) | bit 31.  RFC 1951 section 3.2.5.
__device__ __forceinline__ uint32_t match_token(uint32_t m, uint32_t d) {
    uint32_t li, lx;
    const uint32_t x = m - 3;
    if (x < 8) { li = x; lx = 0; }
    else if (m == 258) { li = 28; lx = 0; }
    else { const uint32_t e = 29u - (uint32_t)__clz(x); li = 4 * e + (x >> e); lx = x & ((1u << e) - 1u); }
    uint32_t ds, dx;
    const uint32_t y = d - 1;
    if (y < 4) { ds = y; dx = 0; }
    else { const uint32_t lg = 31u - (uint32_t)__clz(y), e = lg - 1; ds = 2 * lg + ((y >> e) & 1u); dx = y & ((1u << e) - 1u); }
    return li | lx << 5 | ds << 10 | dx << 15 | 0x80000000u;
}
__device__ __forceinline__ uint32_t len_extra_bits(uint32_t li) { return li < 8 || li == 28 ? 0u : (li - 4) >> 2; }
__device__ __forceinline__ uint32_t dist_extra_bits(uint32_t ds) { return ds < 4 ? 0u : (ds - 2) >> 1; }
// the bits of a match (length code, its extra bits, distance code, its extra bits): at most 48
__device__ __forceinline__ uint64_t match_bits(const ZShared& S, uint32_t tok, uint32_t& n) {
    const uint32_t li = tok & 31u, lx = (tok >> 5) & 31u, ds = (tok >> 10) & 31u, dx = (tok >> 15) & 0x1fffu;
    const uint32_t ll = S.len[257 + li], le = len_extra_bits(li), dl = S.dlen[ds], de = dist_extra_bits(ds);
    uint64_t v = S.code[257 + li];
    v |= (uint64_t)lx << ll;
    v |= (uint64_t)S.dcode[ds] << (ll + le);
    v |= (uint64_t)dx << (ll + le + dl);
    n = ll + le + dl + de;
    return v;
}
__device__ __forceinline__ uint64_t match_bits(const ZMain& S, uint32_t tok, uint32_t& n) {
    const uint32_t li = tok & 31u, lx = (tok >> 5) & 31u, ds = (tok >> 10) & 31u, dx = (tok >> 15) & 0x1fffu;
    const uint32_t lc = S.lcode[li], dc = S.dcode[ds];
    const uint32_t ll = lc >> 16, le = len_extra_bits(li), dl = dc >> 16, de = dist_extra_bits(ds);
    uint64_t v = lc & 0xffffu;
    v |= (uint64_t)lx << ll;
    v |= (uint64_t)(dc & 0xffffu) << (ll + le);
    v |= (uint64_t)dx << (ll + le + dl);
    n = ll + le + dl + de;
    return v;
}

// 16 bytes starting at an arbitrary address, from five aligned word loads (all in flight together)
__device__ __forceinline__ void load16(const uint8_t* a, uint32_t (&x)[4]) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)u & 3u) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
    x[0] = __funnelshift_r(w0, w1, sh); x[1] = __funnelshift_r(w1, w2, sh);
    x[2] = __funnelshift_r(w2, w3, sh); x[3] = __funnelshift_r(w3, w4, sh);
}

// pass 1 of a thread: its 128 contiguous bytes, 16 at a time from global memory -> chunk CRC, the chunk's newlines as
// one bit per byte; HIST (k_bgzf_code, k_bgzf_own): the warp histogram of all bytes (the bytes a match covers are
// taken out again when the match is found); SIZE (k_bgzf): the bits the chunk's bytes take as literals (a byte without
// a code counts kNoCode >> 16 of them)
template <bool FULL, bool HIST, bool SIZE = false>
__device__ __forceinline__ uint32_t pass1(const uint8_t* chunk, uint32_t my_len, uint32_t* hist, const uint32_t* tab,
                                          uint32_t (&nlm)[4], const uint32_t* ctab = nullptr, uint32_t* lit_bits = nullptr) {
    uint32_t reg = 0, bits = 0;
    uint4 d[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        d[k] = (FULL || 16u * k < my_len) ? __ldg(reinterpret_cast<const uint4*>(chunk) + k) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // the four "is a newline" bits of the word gathered into a nibble by one multiplication
            const uint32_t eq = __vcmpeq4(w[q], 0x0a0a0a0au) & 0x01010101u;
            nlm[k >> 1] |= ((eq * 0x01020408u) >> 24) << (((k & 1) * 16) + q * 4);
            // the CRC register takes a whole word per step (slicing by four: one round of table look-ups instead of four
            // dependent ones); the ragged end of a short chunk goes byte by byte
            const bool whole = FULL || (uint32_t)(k * 16 + q * 4 + 3) < my_len;
            if (whole) {
                const uint32_t x = reg ^ w[q];
                reg = tab[768 + (x & 0xffu)] ^ tab[512 + ((x >> 8) & 0xffu)] ^ tab[256 + ((x >> 16) & 0xffu)] ^ tab[x >> 24];
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int i = k * 16 + q * 4 + r;
                if (FULL || (uint32_t)i < my_len) {
                    const uint32_t b = (w[q] >> (8 * r)) & 0xffu;
                    if (HIST) atomicAdd(&hist[b], 1u);
                    if (SIZE) bits += ctab[b] >> 16;
                    if (!whole) reg = tab[(reg ^ b) & 0xffu] ^ (reg >> 8);
                }
            }
        }
    }
    if (SIZE) *lit_bits = bits;
    if (!FULL) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            nlm[j] &= my_len <= 32u * j ? 0u : (my_len - 32u * j >= 32u ? 0xffffffffu : (1u << (my_len - 32u * j)) - 1u);
    }
    return reg;
}

// pass 2 of a warp of k_bgzf_own: its segment again, 4 bytes per lane and 128 contiguous bytes per round.  A byte a
// match covers sends nothing, except the first one, which sends the match.
template <bool FULL, bool LZ>
__device__ __forceinline__ void pack_warp(const ZShared& S, const uint8_t* seg, uint32_t seg_off, uint32_t seg_len, uint32_t* stage,
                                          uint32_t base, uint32_t cursor, uint32_t lane) {
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
            // bit 0: the byte before this word is covered; bits 1..4: this word's bytes
            const uint32_t q = seg_off + off;
            const uint32_t mb = LZ && (FULL || nv) ? __funnelshift_r(S.mask[q >> 5], S.mask[(q >> 5) + 1], q & 31u) & 0x1fu : 0u;
            uint32_t e0 = S.ctab[w[u] & 0xffu], e1 = S.ctab[(w[u] >> 8) & 0xffu], e2 = S.ctab[(w[u] >> 16) & 0xffu], e3 = S.ctab[w[u] >> 24];
            if (!FULL) { if (nv < 1) e0 = 0; if (nv < 2) e1 = 0; if (nv < 3) e2 = 0; if (nv < 4) e3 = 0; }
            if (LZ && (mb & 0x1eu)) { if (mb & 2u) e0 = 0; if (mb & 4u) e1 = 0; if (mb & 8u) e2 = 0; if (mb & 16u) e3 = 0; }
            const uint32_t l0 = e0 >> 16, l1 = e1 >> 16, l2 = e2 >> 16, l3 = e3 >> 16;
            const uint32_t c01 = (e0 & 0xffffu) | (e1 & 0xffffu) << l0, l01 = l0 + l1;
            const uint32_t c23 = (e2 & 0xffffu) | (e3 & 0xffffu) << l2, l23 = l2 + l3;
            uint64_t c = (uint64_t)c01 | (uint64_t)c23 << l01;
            uint32_t l = l01 + l23, top = 0;
            // a match starts at a covered byte whose predecessor is not covered (the byte before a line start is a '\n',
            // which no match covers); everything after it in this word is covered, so its bits come last
            // the matches are kept in block order: the cursor counts the ones before this round, the ballot the ones before this lane
            const uint32_t heads = (mb >> 1) & ~mb & 0xfu;
            const uint32_t hbal = LZ ? __ballot_sync(0xffffffffu, heads != 0) : 0u;
            if (LZ && heads) {
                const uint64_t mbv = S.mbits[cursor + __popc(hbal & ((1u << lane) - 1u))];
                const uint32_t hn = (uint32_t)(mbv >> 56);
                const uint64_t hv = mbv & 0x00ffffffffffffffull;
                if (l + hn > 64) top = (uint32_t)(hv >> (64 - l));
                c |= hv << l;
                l += hn;
            }
            cursor += __popc(hbal);
            uint32_t incl = l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
            const uint32_t at = base + incl - l;
            base += __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t i = at >> 5, sh = at & 31;
            const uint32_t x0 = (uint32_t)c, x1 = (uint32_t)(c >> 32);
            if (l) atomicOr(&stage[i], x0 << sh);
            if (sh + l > 32) atomicOr(&stage[i + 1], __funnelshift_l(x0, x1, sh));
            if (sh + l > 64) atomicOr(&stage[i + 2], __funnelshift_l(x1, top, sh));
            if (LZ && sh + l > 96) atomicOr(&stage[i + 3], __funnelshift_l(top, 0u, sh));
        }
    }
}

// pass 2 of a thread of k_bgzf: its chunk again, byte after byte into a 64-bit window that is emptied, a word at a time,
// into the member's image at the chunk's own bit position (the first and the last word are shared with the
// neighbours: OR-ed).  A byte a match covers sends nothing, except the first one, which sends the match.
template <bool FULL, bool LZ>
__device__ __forceinline__ void encode_chunk(const ZMain& S, uint32_t* image, const uint8_t* chunk, uint32_t my_off, uint32_t my_len,
                                             uint32_t bitpos, uint32_t cursor) {
    uint64_t acc = 0;
    uint32_t fill = bitpos & 31u;
    uint32_t* wp = image + (bitpos >> 5);
    // n <= 32 bits behind the fill < 32 bits waiting in the window
    auto put = [&](uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << fill;
        fill += n;
        if (fill >= 32) { atomicOr(wp, (uint32_t)acc); wp++; acc >>= 32; fill -= 32; }
    };
    uint4 nx = __ldg(reinterpret_cast<const uint4*>(chunk));
#pragma unroll 1
    for (uint32_t k = 0; k < 8; k++) {
        if (!FULL && 16u * k >= my_len) break;
        const uint4 d = nx;
        if (k + 1 < 8 && (FULL || 16u * (k + 1) < my_len)) nx = __ldg(reinterpret_cast<const uint4*>(chunk) + k + 1);   // in flight while d is coded
        const uint32_t w[4] = {d.x, d.y, d.z, d.w};
        // bit 0: the byte before these sixteen is covered; bits 1..16: the sixteen
        const uint32_t q0 = my_off + 16u * k;
        const uint32_t mb16 = LZ ? __funnelshift_r(S.mask[q0 >> 5], S.mask[(q0 >> 5) + 1], q0 & 31u) : 0u;
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) {
            const uint32_t off = 16u * k + 4u * q;
            const uint32_t nv = FULL ? 4u : (off >= my_len ? 0u : min(4u, my_len - off));
            const uint32_t mb = (mb16 >> (4u * q)) & 0x1fu;
            uint32_t e0 = S.ctab[w[q] & 0xffu], e1 = S.ctab[(w[q] >> 8) & 0xffu], e2 = S.ctab[(w[q] >> 16) & 0xffu], e3 = S.ctab[w[q] >> 24];
            if (!FULL) { if (nv < 1) e0 = 0; if (nv < 2) e1 = 0; if (nv < 3) e2 = 0; if (nv < 4) e3 = 0; }
            if (LZ && (mb & 0x1eu)) { if (mb & 2u) e0 = 0; if (mb & 4u) e1 = 0; if (mb & 8u) e2 = 0; if (mb & 16u) e3 = 0; }
            const uint32_t l0 = e0 >> 16, l1 = e1 >> 16, l2 = e2 >> 16, l3 = e3 >> 16;
            const uint32_t c01 = (e0 & 0xffffu) | (e1 & 0xffffu) << l0, l01 = l0 + l1;
            const uint32_t c23 = (e2 & 0xffffu) | (e3 & 0xffffu) << l2, l23 = l2 + l3;
            if (l01 + l23 <= 32) put(c01 | c23 << l01, l01 + l23);        // l01 <= 30
            else { put(c01, l01); put(c23, l23); }
            // a match starts at a covered byte whose predecessor is not covered (the byte before a line start is a '\n',
            // which no match covers); everything after it in this word is covered, so its bits come last
            if (LZ && ((mb >> 1) & ~mb & 0xfu) && (FULL || nv)) {
                const uint64_t mbv = S.mbits[cursor++];
                const uint32_t hn = (uint32_t)(mbv >> 56);
                put((uint32_t)mbv, min(hn, 32u));
                if (hn > 32) put((uint32_t)(mbv >> 32) & 0x00ffffffu, hn - 32);
            }
        }
    }
    if (fill) atomicOr(wp, (uint32_t)acc);
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

// The match of line start k (ls[k]) with line start k - kBack, never beyond the line's last byte before its '\n':
// position, length and token; tok = 0: none.
__device__ __forceinline__ void find_match(const uint8_t* in, uint32_t len, const uint16_t* ls, uint32_t n_ls, uint32_t k, uint32_t& at,
                                           uint32_t& mlen, uint32_t& tok) {
    if (k < kBack || k >= n_ls) return;
    const uint32_t i = ls[k];
    if (i >= len) return;
    const uint32_t p = ls[k - kBack];
    const uint32_t next = k + 1 < n_ls ? (uint32_t)ls[k + 1] - 1u : len;   // the '\n' that ends this line, or the end
    const uint32_t maxm = min(258u, next - i);
    // common prefix, 16 bytes per round trip (reads may run up to 19 bytes past the block: the buffers have slack)
    uint32_t m = 0;
    while (m < maxm) {
        uint32_t a[4], c[4];
        load16(in + i + m, a);
        load16(in + p + m, c);
        uint32_t same = 16;
#pragma unroll
        for (int j = 3; j >= 0; j--) { const uint32_t x = a[j] ^ c[j]; if (x) same = 4 * j + (((uint32_t)__ffs(x) - 1u) >> 3); }
        m += same;
        if (same < 16) break;
    }
    m = min(m, maxm);
    if (m >= 4 && i - p <= 32768u) { at = i; mlen = m; tok = match_token(m, i - p); }
}

// The codes of one file of a batch, from its first block (one CTA per file).
__global__ void __launch_bounds__(kZT, 1)
k_bgzf_code(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, const uint64_t* __restrict__ totals, uint32_t lz,
            ZCode* __restrict__ codes) {
    JLP_DYN_SMEM(smem_raw);
    ZShared& S = *reinterpret_cast<ZShared*>(smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool second = blockIdx.x != 0;
    const uint64_t n_all = totals[second ? 1 : 0];
    if (n_all == 0) return;
    const uint8_t* in = second ? in1 : in0;
    const uint32_t len = (uint32_t)min((uint64_t)kBgzfIn, n_all);
    ZCode& Z = codes[second ? 1 : 0];

    const uint32_t my_off = t * kChunk;
    const uint32_t my_len = my_off >= len ? 0u : min(kChunk, len - my_off);
    const HuffScratch H1{S.sw, S.w_int, S.sorted, S.par_leaf, S.par_int, S.d_int, S.bl};
    const HuffScratch H2{S.sw2, S.w_int2, S.sorted2, S.par_leaf2, S.par_int2, S.d_int2, S.bl2};

    // ---- clear the header image (and with it the histograms) and the counters
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.stage);
        for (uint32_t i = t; i < kStageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        if (t < 256) { S.crc_tab[0][t] = c_crc_tab[0][t]; S.crc_tab[1][t] = c_crc_tab[1][t]; S.crc_tab[2][t] = c_crc_tab[2][t]; S.crc_tab[3][t] = c_crc_tab[3][t]; }
        if (t < 288) { S.len[t] = 0; S.cnt[t] = 0; }
        if (t < 32) { S.dcnt[t] = 0; S.dlen[t] = 0; }
        if (t < 16) S.bl[t] = 0;
        if (t < 19) S.cl_cnt[t] = 0;
        if (t == 0) { S.maxd = 0; S.ls[0] = 0; }
    }
    __syncthreads();
    // ---- pass 1: histogram of all bytes + the chunk's newlines (the CRC comes along unused)
    uint32_t* hist = S.stage + kHistWord0 + warp * 256;
    uint32_t nlm[4] = {0, 0, 0, 0};
    if (my_len == kChunk) pass1<true, true>(in + my_off, my_len, hist, &S.crc_tab[0][0], nlm);
    else pass1<false, true>(in + my_off, my_len, hist, &S.crc_tab[0][0], nlm);

    if (lz) {       // uniform
        // ---- line starts and matches as k_bgzf finds them; here only their symbols are counted
        const uint32_t my_nl = __popc(nlm[0]) + __popc(nlm[1]) + __popc(nlm[2]) + __popc(nlm[3]);
        uint32_t total_nl;
        const uint32_t nl_before = block_scan(my_nl, S.scan_tmp, total_nl);
        {
            uint32_t idx = nl_before + 1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t m = nlm[j];
                while (m) {
                    const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1u;
                    if (idx <= kLsCap) S.ls[idx] = (uint16_t)(my_off + 32u * j + bit + 1u);
                    idx++;
                }
            }
        }
        __syncthreads();
        const uint32_t n_ls = min(min(total_nl, kLsCap) + 1, 2u * kZT);
        uint32_t m_at[2] = {0, 0}, m_len[2] = {0, 0}, m_tok[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) find_match(in, len, S.ls, n_ls, 2 * t + h, m_at[h], m_len[h], m_tok[h]);
        uint32_t total_m;
        uint32_t mj = block_scan((m_tok[0] ? 1u : 0u) + (m_tok[1] ? 1u : 0u), S.scan_tmp, total_m);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (m_tok[h] && mj < kMCap) {
                const uint32_t tok = m_tok[h], i = m_at[h];
                atomicAdd(&S.cnt[257 + (tok & 31u)], 1u);
                atomicAdd(&S.dcnt[(tok >> 10) & 31u], 1u);
                // its bytes leave the literal histograms (of the warps whose segments hold them)
                for (uint32_t q0 = 0; q0 < m_len[h]; q0 += 16) {
                    uint32_t a[4];
                    load16(in + i + q0, a);
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (q0 + j < m_len[h]) atomicSub(&S.stage[kHistWord0 + ((i + q0 + j) / kSeg) * 256 + ((a[j >> 2] >> (8 * (j & 3))) & 0xffu)], 1u);
                }
            }
            if (m_tok[h]) mj++;
        }
    }
    __syncthreads();

    // ---- the counts, each raised by one: what the sample does not hold can still be coded in the other blocks
    //      (level 1 sends no matches: its length and distance alphabets stay empty)
    const uint32_t hlit = lz ? kNLit : 257u;
    uint32_t my_cnt = 0;
    if (t < hlit) {
        if (t < 256) {
            for (int w = 0; w < kZW; w++) my_cnt += S.stage[kHistWord0 + w * 256 + t];
            if (t == '\n' || (t >= 0x20 && t < 0x7f)) my_cnt += 1u;
        } else my_cnt = t == 256 ? 1u : S.cnt[t] + 1u;     // [256]: the end of block, once per block
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
#pragma unroll 4
        for (uint32_t q = 0; q < n_active; q++) rank += S.w_int[q] < key ? 1u : 0u;
        S.sorted[rank] = (uint16_t)t;
        S.sw[rank] = my_cnt;
    }
    __syncthreads();

    // ---- literal/length code lengths: the tree by one thread, the depths by one thread per leaf
    if (t == 0) huff_merge(H1, n_active);      // n_active >= 257
    if (warp == 1) {                           // meanwhile the distance code
        const uint32_t hd = warp_huffman(H2, lane < kNDist ? S.dcnt[lane] + (lz ? 1u : 0u) : 0u, kNDist, 15, S.dlen, S.dcode, lane);
        if (lane == 0) S.hdist = hd;
    }
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
    if (S.maxd > 15) limit_15(S, n_active, t);   // the 15-bit limit has to act (with the raised counts: nearly always)
    // canonical codes: first code of the length + the symbols of the same length before this one
    if (t < kNLit) {
        const uint32_t l = S.len[t];
        if (l) {
            const uint32_t pat = l * 0x01010101u;
            const uint32_t* lw = reinterpret_cast<const uint32_t*>(S.len);
            uint32_t same = 0;
#pragma unroll 4
            for (uint32_t j = 0; j < t / 4; j++) same += __popc(__vcmpeq4(lw[j], pat));
            same += __popc(__vcmpeq4(lw[t / 4], pat) & ((1u << (8 * (t & 3))) - 1u));
            S.code[t] = (uint16_t)rev_bits(first_code(S.bl, l) + same / 8, l);
        } else S.code[t] = 0;
    }
    __syncthreads();
    // ---- the runs of the code-length sequence (HLIT literal/length lengths, then HDIST distance lengths): one thread
    //      per run counts its tokens
    const uint32_t n_seq = hlit + S.hdist;
    auto seq = [&](uint32_t i) -> uint32_t { return i < hlit ? S.len[i] : S.dlen[i - hlit]; };
    RunTok rt{0, 0, 0, 0, 0, 0};
    // where the runs start, one bit per position (n_seq <= 316 < kZT): a run ends where the next one starts
    const uint32_t v_seq = t < n_seq ? seq(t) : 0u;
    const bool run_start = t < n_seq && (t == 0 || seq(t - 1) != v_seq);
    {
        const uint32_t bal = __ballot_sync(0xffffffffu, run_start);
        if (lane == 0) S.run_bal[warp] = bal;
    }
    __syncthreads();
    if (run_start) {
        uint32_t next = n_seq;
        uint32_t m = lane == 31 ? 0u : S.run_bal[warp] & (0xffffffffu << (lane + 1u));
        for (uint32_t w = warp; ; ) {
            if (m) { next = min(n_seq, w * 32u + (uint32_t)__ffs(m) - 1u); break; }
            if (++w >= (uint32_t)kZW) break;
            m = S.run_bal[w];
        }
        rt = run_tokens(v_seq, next - t);
        if (v_seq == 0) {
            if (rt.n_big) atomicAdd(&S.cl_cnt[18], rt.n_big);
            if (rt.n_mid) atomicAdd(&S.cl_cnt[17], 1u);
            if (rt.n_lit) atomicAdd(&S.cl_cnt[0], rt.n_lit);
        } else {
            if (rt.n_big) atomicAdd(&S.cl_cnt[16], rt.n_big);
            atomicAdd(&S.cl_cnt[v_seq], rt.n_lit);
        }
    }
    __syncthreads();

    // ---- the code-length code and the fixed part of the headers (warp 0)
    if (warp == 0) {
        warp_huffman(H2, lane < 19 ? S.cl_cnt[lane] : 0u, 19, 7, S.cl_len, S.cl_code, lane);
        const uint32_t l = lane < 19 ? S.cl_len[lane] : 0u;
        const uint32_t order_lane = lane < 19 ? (uint32_t)"\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f"[lane] : 0u;
        // HCLEN: the lengths are sent in the order 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 up to the last one used
        const uint32_t l_ord = __shfl_sync(0xffffffffu, l, order_lane);
        const uint32_t used = __ballot_sync(0xffffffffu, lane < 19 && l_ord != 0);
        const uint32_t hclen = max(4u, 32u - (uint32_t)__clz(used));
        if (lane == 0) {
            // member header: ID1 ID2 CM FLG(FEXTRA) MTIME XFL OS(255) XLEN=6 'B' 'C' SLEN=2 BSIZE (= total - 1, left to k_bgzf)
            S.stage[0] = 0x04088b1fu; S.stage[1] = 0; S.stage[2] = 0x0006ff00u; S.stage[3] = 0x00024342u;
            uint32_t pos = kHdr * 8;
            put_bits(S.stage, pos, 1, 1); put_bits(S.stage, pos, 2, 2);                   // BFINAL, BTYPE = dynamic Huffman
            put_bits(S.stage, pos, hlit - 257, 5); put_bits(S.stage, pos, S.hdist - 1, 5); put_bits(S.stage, pos, hclen - 4, 4);
            S.hdr_fixed_bits = pos + 3 * hclen;
        }
        if (lane < hclen) { uint32_t pos = kHdr * 8 + 17 + 3 * lane; put_bits(S.stage, pos, l_ord, 3); }
    }
    __syncthreads();
    // ---- the tokens of the runs, placed by a scan of their bits
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
    const uint32_t excl = block_scan(hb, S.scan_tmp, total);
    if (run_start) {
        uint32_t pos = S.hdr_fixed_bits + excl;
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
    __syncthreads();
    // ---- hand over
    const uint32_t hdr_bits = S.hdr_fixed_bits + total;
    if (t < 256) Z.ctab[t] = S.len[t] ? (uint32_t)S.code[t] | (uint32_t)S.len[t] << 16 : kNoCode;
    if (t < 32) {
        Z.lcode[t] = t < kNLit - 257 && S.len[257 + t] ? (uint32_t)S.code[257 + t] | (uint32_t)S.len[257 + t] << 16 : 0u;
        Z.dcode[t] = t < kNDist && S.dlen[t] ? (uint32_t)S.dcode[t] | (uint32_t)S.dlen[t] << 16 : 0u;
    }
    if (t < kHdrWords) Z.hdr[t] = t * 32u < hdr_bits ? S.stage[t] : 0u;
    if (t == 0) { Z.eob = (uint32_t)S.code[256] | (uint32_t)S.len[256] << 16; Z.hdr_bits = hdr_bits; Z.pad[0] = 0; Z.pad[1] = 0; }
}

// The common prefix of the lines at i and p < i, at most maxm bytes (16 per round trip; reads may run up to 19 bytes
// past the block: the buffers have slack)
__device__ __forceinline__ uint32_t common_prefix(const uint8_t* in, uint32_t i, uint32_t p, uint32_t maxm) {
    uint32_t m = 0;
    while (m < maxm) {
        uint32_t a[4], c[4];
        load16(in + i + m, a);
        load16(in + p + m, c);
        uint32_t same = 16;
#pragma unroll
        for (int j = 3; j >= 0; j--) { const uint32_t x = a[j] ^ c[j]; if (x) same = 4 * j + (((uint32_t)__ffs(x) - 1u) >> 3); }
        m += same;
        if (same < 16) break;
    }
    return min(m, maxm);
}

// One BGZF block per CTA, with the codes k_bgzf_code left.  Between the first barrier and the one at which the
// segments' sizes meet, every warp works on its own segment of 4096 bytes with warp-level synchronisation only.
__global__ void __launch_bounds__(kZT, 3)
k_bgzf(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, const uint64_t* __restrict__ totals, uint32_t nblk_max,
       uint32_t lz, const ZCode* __restrict__ codes, uint8_t* __restrict__ slots0, uint8_t* __restrict__ slots1,
       uint32_t* __restrict__ zlen0, uint32_t* __restrict__ zlen1) {
    JLP_DYN_SMEM(smem_raw);
    ZMain& S = *reinterpret_cast<ZMain*>(smem_raw);
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
    const ZCode& Z = codes[second ? 1 : 0];

    const uint32_t my_off = t * kChunk;
    const uint32_t my_len = my_off >= len ? 0u : min(kChunk, len - my_off);
    const uint32_t seg_off = warp * kSeg, seg_end = min(len, seg_off + kSeg);

    // ---- the image: the header words (the same for every block), zeros behind them; the match mask; the codes
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.image);
        for (uint32_t i = kHdrWords / 4 + t; i < kImageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        if (t < kHdrWords) S.image[t] = __ldg(&Z.hdr[t]);
        if (lz) for (uint32_t i = t; i < kMaskWords; i += kZT) S.mask[i] = 0;
        if (t < 256) {
            S.crc_tab[0][t] = c_crc_tab[0][t]; S.crc_tab[1][t] = c_crc_tab[1][t]; S.crc_tab[2][t] = c_crc_tab[2][t]; S.crc_tab[3][t] = c_crc_tab[3][t];
            S.ctab[t] = __ldg(&Z.ctab[t]);
        }
        if (t < 32) { S.lcode[t] = __ldg(&Z.lcode[t]); S.dcode[t] = __ldg(&Z.dcode[t]); }
        if (t == 0) { S.eob = __ldg(&Z.eob); S.hdr_bits = __ldg(&Z.hdr_bits); S.miss = 0; }
    }
    __syncthreads();
    // ---- pass 1: chunk CRC, the chunk's newlines, the bits of its bytes as literals
    uint32_t nlm[4] = {0, 0, 0, 0};
    uint32_t lit_bits = 0;
    uint32_t crc = my_len == kChunk ? pass1<true, false, true>(in + my_off, my_len, nullptr, &S.crc_tab[0][0], nlm, S.ctab, &lit_bits)
                                    : pass1<false, false, true>(in + my_off, my_len, nullptr, &S.crc_tab[0][0], nlm, S.ctab, &lit_bits);
    if (lit_bits >= (kNoCode >> 16)) { S.miss = 1; lit_bits = 0; }       // a byte the code has no word for
    S.cbits[t] = lit_bits;
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
    if (lane == 0) {
        // the segment's register advanced over the bytes behind the segment: the block's is the XOR of the sixteen
        if (len == kBgzfIn) {
            if (warp + 1 < (uint32_t)kZW) {
                crc = crc_advance_lvl(9, crc);                                     // the last segment's 3840 bytes
                const uint32_t n = kZW - 2 - warp;                                  // whole segments in between
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) if ((n >> q) & 1u) crc = crc_advance_lvl(5 + q, crc);
            }
        } else crc = crc_advance(crc, len - seg_end);
        S.crc_w[warp] = crc;
    }

    if (lz && seg_off < len) {       // uniform in the warp
        uint16_t* wls = S.ls[warp];
        // ---- the last kBack line starts before the segment, from the newlines of the kLookBack bytes before it
        //      (32 per lane); one in the byte just before the segment starts a line of the segment itself
        if (lane < kBack) wls[lane] = kNoLine;
        __syncwarp();
        if (warp) {
            const uint4* lb = reinterpret_cast<const uint4*>(in + seg_off - kLookBack + 32u * lane);
            const uint4 d0 = __ldg(lb), d1 = __ldg(lb + 1);
            const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) m |= (((__vcmpeq4(w[q], 0x0a0a0a0au) & 0x01010101u) * 0x01020408u) >> 24) << (4 * q);
            if (lane == 31) m &= 0x7fffffffu;
            // newlines behind this lane's, by a suffix sum
            uint32_t after = __popc(m);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_down_sync(0xffffffffu, after, o); if (lane + o < 32) after += x; }
            after -= __popc(m);
            while (m && after < kBack) {            // from the last one down
                const uint32_t bit = 31u - (uint32_t)__clz(m);
                m &= ~(1u << bit);
                wls[kBack - 1 - after] = (uint16_t)(seg_off - kLookBack + 32u * lane + bit + 1u);
                after++;
            }
        }
        // ---- the line starts inside the segment: behind every newline but one in the chunk's last byte, and at the
        //      chunk's first byte when the byte before is a newline (or the block starts there)
        const bool prev_nl = my_len && (t == 0 || __ldg(in + my_off - 1) == '\n');
        const bool last_nl = my_len && __ldg(in + my_off + my_len - 1) == '\n';
        const uint32_t my_nl = __popc(nlm[0]) + __popc(nlm[1]) + __popc(nlm[2]) + __popc(nlm[3]);
        const uint32_t my_ls = my_nl + (prev_nl ? 1u : 0u) - (last_nl ? 1u : 0u);
        uint32_t incl = my_ls;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
        const uint32_t n_ls = min(kBack + __shfl_sync(0xffffffffu, incl, 31), kWls);     // entries of wls in use
        const uint32_t k_lo = kBack + incl - my_ls;
        {
            uint32_t idx = k_lo;
            if (prev_nl) { if (idx < kWls) wls[idx] = (uint16_t)my_off; idx++; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t m = nlm[j];
                while (m) {
                    const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1u;
                    const uint32_t pos = 32u * j + bit + 1u;
                    if (pos < my_len) { if (idx < kWls) wls[idx] = (uint16_t)(my_off + pos); idx++; }
                }
            }
        }
        __syncwarp();
        // ---- matches: every thread looks at the line starts of its chunk -- the thread that will code them -- and keeps
        //      up to kSlots matches, in order, in slots of its own; a match ends with its line and with the segment
        //      (no match covers a newline: the coder finds a match's first byte by its uncovered predecessor)
        const uint32_t last_end = seg_end - (__ldg(in + seg_end - 1) == '\n' ? 1u : 0u);
        uint32_t slot_n = 0;
        for (uint32_t k = k_lo; k < min(k_lo + my_ls, n_ls); k++) {
            const uint32_t p = wls[k - kBack];
            if (p == kNoLine || (k + 1 == kWls)) continue;     // the list's last entry when it is full: its line's end is not known
            const uint32_t i = wls[k];
            const uint32_t next = k + 1 < n_ls ? (uint32_t)wls[k + 1] - 1u : last_end;   // the '\n' that ends this line, or the segment's end
            const uint32_t m = common_prefix(in, i, p, min(258u, next - i));
            if (m >= 4 && slot_n < kSlots) {
                uint32_t hn;
                const uint64_t hv = match_bits(S, match_token(m, i - p), hn);
                S.mbits[kSlots * t + slot_n++] = hv | (uint64_t)hn << 56;
                atomicAdd(&S.cbits[t], hn);
                // its bytes no longer send their literals ...
                for (uint32_t q0 = 0; q0 < m; q0 += 16) {
                    uint32_t a[4];
                    load16(in + i + q0, a);
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (q0 + j < m) atomicSub(&S.cbits[(i + q0 + j) / kChunk], S.ctab[(a[j >> 2] >> (8 * (j & 3))) & 0xffu] >> 16);
                }
                // ... and are marked: bits [i + 1, i + m + 1) of the mask
                uint32_t lo = i + 1, hi = i + m + 1;
                while (lo < hi) {
                    const uint32_t wi = lo >> 5, b0 = lo & 31u, n = min(32u - b0, hi - lo);
                    atomicOr(&S.mask[wi], (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << b0);
                    lo += n;
                }
            }
        }
        __syncwarp();
    }

    // ---- where every chunk's bits go: a scan in the warp, the segments' sums through shared memory
    const uint32_t my_bits = S.cbits[t];
    uint32_t excl = my_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, excl, o); if (lane >= (uint32_t)o) excl += x; }
    if (lane == 31) S.wtot[warp] = excl;
    excl -= my_bits;
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (uint32_t w = 0; w < (uint32_t)kZW; w++) { const uint32_t x = S.wtot[w]; total += x; excl += w < warp ? x : 0u; }
    const uint32_t hdr_bits = S.hdr_bits, eob = S.eob;
    const uint32_t end_bits = hdr_bits + total + (eob >> 16);
    const uint32_t z_bytes = (end_bits + 7) / 8 + 8;                 // the member with a dynamic block
    if (S.miss || z_bytes - kHdr - 8 >= len + 5 || z_bytes > kImageBytes - 16) {     // uniform
        // a byte without a code, or the block does not shrink with this code: k_bgzf_own takes it
        if (t == 0) *zlen = kOwnCode;
        return;
    }
    // ---- pass 2: every thread codes its chunk to its place
    if (my_len) {
        const uint32_t bitpos = hdr_bits + excl, cursor = kSlots * t;
        if (lz) {
            if (my_len == kChunk) encode_chunk<true, true>(S, S.image, in + my_off, my_off, my_len, bitpos, cursor);
            else encode_chunk<false, true>(S, S.image, in + my_off, my_off, my_len, bitpos, cursor);
        } else {
            if (my_len == kChunk) encode_chunk<true, false>(S, S.image, in + my_off, my_off, my_len, bitpos, cursor);
            else encode_chunk<false, false>(S, S.image, in + my_off, my_off, my_len, bitpos, cursor);
        }
    }
    if (t == 0) {
        uint32_t pos = end_bits - (eob >> 16);
        put_bits(S.image, pos, eob & 0xffffu, eob >> 16);
        atomicOr(&S.image[4], (z_bytes - 1) & 0xffffu);              // BSIZE
        pos = (z_bytes - 8) * 8;
        uint32_t crc_all = 0;
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)kZW; w++) crc_all ^= S.crc_w[w];
        put_bits(S.image, pos, crc_all ^ (len == kBgzfIn ? c_crc_init_full : crc_advance(0xffffffffu, len)) ^ 0xffffffffu, 32);
        put_bits(S.image, pos, len, 32);
        *zlen = z_bytes;
    }
    __syncthreads();
    const uint4* s4 = reinterpret_cast<const uint4*>(S.image);
    uint4* o4 = reinterpret_cast<uint4*>(slot);
    for (uint32_t i = t; i < (z_bytes + 15) / 16; i += kZT) o4[i] = s4[i];
}

// A block with a Huffman code of its own, for what k_bgzf left (zlen = kOwnCode): warp-cooperative passes with
// per-warp literal histograms in pass 1, the code construction of k_bgzf_code between the passes, one block image.
__device__ __forceinline__ void own_block(ZShared& S, uint32_t blk, const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1,
                                          const uint64_t* __restrict__ totals, uint32_t nblk_max, uint32_t lz, uint8_t* __restrict__ slots0,
                                          uint8_t* __restrict__ slots1, uint32_t* __restrict__ zlen0, uint32_t* __restrict__ zlen1) {
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool second = blk >= nblk_max;
    const uint32_t b = second ? blk - nblk_max : blk;
    const uint64_t n_all = totals[second ? 1 : 0];
    if ((uint64_t)b * kBgzfIn >= n_all) return;
    const uint8_t* in = (second ? in1 : in0) + (uint64_t)b * kBgzfIn;
    const uint32_t len = (uint32_t)min((uint64_t)kBgzfIn, n_all - (uint64_t)b * kBgzfIn);
    uint8_t* slot = (second ? slots1 : slots0) + (uint64_t)b * kBgzfSlot;
    uint32_t* zlen = (second ? zlen1 : zlen0) + b;
    if (*zlen != kOwnCode) return;      // uniform

    const uint32_t my_off = t * kChunk;
    const uint32_t my_len = my_off >= len ? 0u : min(kChunk, len - my_off);
    const HuffScratch H1{S.sw, S.w_int, S.sorted, S.par_leaf, S.par_int, S.d_int, S.bl};
    const HuffScratch H2{S.sw2, S.w_int2, S.sorted2, S.par_leaf2, S.par_int2, S.d_int2, S.bl2};

    // ---- clear the image (and with it the histograms), the match mask and the counters
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.stage);
        for (uint32_t i = t; i < kStageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t i = t; i < kMaskWords; i += kZT) S.mask[i] = 0;
        if (t < 256) { S.crc_tab[0][t] = c_crc_tab[0][t]; S.crc_tab[1][t] = c_crc_tab[1][t]; S.crc_tab[2][t] = c_crc_tab[2][t]; S.crc_tab[3][t] = c_crc_tab[3][t]; }
        if (t < 288) { S.len[t] = 0; S.cnt[t] = 0; }
        if (t < 32) { S.dcnt[t] = 0; S.dlen[t] = 0; }
        if (t < 16) S.bl[t] = 0;
        if (t < 19) S.cl_cnt[t] = 0;
        if (t < kZW) S.wm[t] = 0;
        if (t == 0) { S.maxd = 0; S.hlit = 257; S.ls[0] = 0; }
    }

    __syncthreads();
    // ---- pass 1: histogram of all bytes + chunk CRC + the chunk's newlines
    uint32_t* hist = S.stage + kHistWord0 + warp * 256;
    uint32_t nlm[4] = {0, 0, 0, 0};
    uint32_t crc = my_len == kChunk ? pass1<true, true>(in + my_off, my_len, hist, &S.crc_tab[0][0], nlm) : pass1<false, true>(in + my_off, my_len, hist, &S.crc_tab[0][0], nlm);
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

    uint32_t n_m = 0;
    if (lz) {       // uniform
    // ---- line starts: the byte after every newline, numbered by a block scan of the newline counts
    const uint32_t my_nl = __popc(nlm[0]) + __popc(nlm[1]) + __popc(nlm[2]) + __popc(nlm[3]);
    uint32_t total_nl;
    const uint32_t nl_before = block_scan(my_nl, S.scan_tmp, total_nl);
    {
        uint32_t idx = nl_before + 1;       // line start idx follows the idx-th newline
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t m = nlm[j];
            while (m) {
                const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                m &= m - 1u;
                if (idx <= kLsCap) S.ls[idx] = (uint16_t)(my_off + 32u * j + bit + 1u);
                idx++;
            }
        }
    }
    __syncthreads();
    const uint32_t n_ls = min(min(total_nl, kLsCap) + 1, 2u * kZT);   // entries of S.ls in use (a last one may equal len: no line there)
    // ---- matches: line start k against line start k - kBack, never beyond the line's last byte before its '\n';
    //      thread t looks at line starts 2t and 2t + 1, a block scan numbers the matches in block order
    uint32_t m_at[2] = {0, 0}, m_len[2] = {0, 0}, m_tok[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) find_match(in, len, S.ls, n_ls, 2 * t + h, m_at[h], m_len[h], m_tok[h]);
    uint32_t total_m;
    uint32_t mj = block_scan((m_tok[0] ? 1u : 0u) + (m_tok[1] ? 1u : 0u), S.scan_tmp, total_m);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (m_tok[h] && mj < kMCap) {
            const uint32_t tok = m_tok[h], i = m_at[h];
            S.mbits[mj] = tok;
            atomicAdd(&S.cnt[257 + (tok & 31u)], 1u);
            atomicAdd(&S.dcnt[(tok >> 10) & 31u], 1u);
            atomicMax(&S.hlit, 258 + (tok & 31u));
            atomicAdd(&S.wm[i / kSeg], 1u);
            // its bytes leave the literal histograms (of the warps whose segments hold them) ...
            for (uint32_t q0 = 0; q0 < m_len[h]; q0 += 16) {
                uint32_t a[4];
                load16(in + i + q0, a);
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (q0 + j < m_len[h]) atomicSub(&S.stage[kHistWord0 + ((i + q0 + j) / kSeg) * 256 + ((a[j >> 2] >> (8 * (j & 3))) & 0xffu)], 1u);
            }
            // ... and are marked: bits [i + 1, i + m + 1) of the mask
            uint32_t lo = i + 1, hi = i + m_len[h] + 1;
            while (lo < hi) {
                const uint32_t wi = lo >> 5, b0 = lo & 31u, n = min(32u - b0, hi - lo);
                atomicOr(&S.mask[wi], (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << b0);
                lo += n;
            }
        }
        if (m_tok[h]) mj++;
    }
    n_m = min(total_m, kMCap);
    }
    __syncthreads();

    uint32_t my_cnt = 0;
    if (t < kNLit) {
        if (t < 256) for (int w = 0; w < kZW; w++) my_cnt += S.stage[kHistWord0 + w * 256 + t];
        else my_cnt = t == 256 ? 1u : S.cnt[t];     // end of block; lengths
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
#pragma unroll 4
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
                const uint32_t r0 = (lane + s) * kSeg;
                const uint32_t right_len = r0 >= len ? 0u : min(s * kSeg, len - r0);
                crc = (right_len == s * kSeg ? crc_advance_lvl(5 + j, crc) : crc_advance(crc, right_len)) ^ other;
            }
        }
        if (lane == 0) S.crc_out = crc ^ (len == kBgzfIn ? c_crc_init_full : crc_advance(0xffffffffu, len)) ^ 0xffffffffu;
    }
    __syncthreads();

    // ---- literal/length code lengths: the tree by one thread, the depths by one thread per leaf
    if (t == 0) huff_merge(H1, n_active);      // n_active >= 2: a literal or a match, and the end of block
    if (warp == 1) {                           // meanwhile the distance code
        const uint32_t hd = warp_huffman(H2, lane < kNDist ? S.dcnt[lane] : 0u, kNDist, 15, S.dlen, S.dcode, lane);
        if (lane == 0) S.hdist = hd;
    }
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
    if (S.maxd > 15) limit_15(S, n_active, t);   // rare: the 15-bit limit has to act
    // canonical codes: first code of the length + the symbols of the same length before this one
    if (t < kNLit) {
        const uint32_t l = S.len[t];
        if (l) {
            const uint32_t pat = l * 0x01010101u;
            const uint32_t* lw = reinterpret_cast<const uint32_t*>(S.len);
            uint32_t same = 0;
#pragma unroll 4
            for (uint32_t j = 0; j < t / 4; j++) same += __popc(__vcmpeq4(lw[j], pat));
            same += __popc(__vcmpeq4(lw[t / 4], pat) & ((1u << (8 * (t & 3))) - 1u));
            S.code[t] = (uint16_t)rev_bits(first_code(S.bl, l) + same / 8, l);
        }
        if (t < 256) S.ctab[t] = l ? (uint32_t)S.code[t] | l << 16 : 0u;
    }
    __syncthreads();
    // the runs of the code-length sequence (HLIT literal/length lengths, then HDIST distance lengths): one thread
    // per run counts its tokens
    // the matches' tokens become their bits now that both codes are known
    for (uint32_t j = t; j < n_m; j += kZT) {
        uint32_t hn;
        const uint64_t hv = match_bits(S, (uint32_t)S.mbits[j], hn);
        S.mbits[j] = hv | (uint64_t)hn << 56;
    }
    const uint32_t hlit = S.hlit, n_seq = hlit + S.hdist;
    auto seq = [&](uint32_t i) -> uint32_t { return i < hlit ? S.len[i] : S.dlen[i - hlit]; };
    RunTok rt{0, 0, 0, 0, 0, 0};
    // where the runs start, one bit per position (n_seq <= 316 < kZT): a run ends where the next one starts, so no thread
    // has to walk its run (the zero lengths of the bytes FASTQ never uses make runs of a hundred and more)
    const uint32_t v_seq = t < n_seq ? seq(t) : 0u;
    const bool run_start = t < n_seq && (t == 0 || seq(t - 1) != v_seq);
    {
        const uint32_t bal = __ballot_sync(0xffffffffu, run_start);
        if (lane == 0) S.run_bal[warp] = bal;
    }
    __syncthreads();
    if (t < n_seq) {
        const uint32_t v = v_seq;
        if (run_start) {
            uint32_t next = n_seq;
            uint32_t m = lane == 31 ? 0u : S.run_bal[warp] & (0xffffffffu << (lane + 1u));
            for (uint32_t w = warp; ; ) {
                if (m) { next = min(n_seq, w * 32u + (uint32_t)__ffs(m) - 1u); break; }
                if (++w >= kZW) break;
                m = S.run_bal[w];
            }
            const uint32_t run = next - t;
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
        warp_huffman(H2, lane < 19 ? S.cl_cnt[lane] : 0u, 19, 7, S.cl_len, S.cl_code, lane);
        const uint32_t l = lane < 19 ? S.cl_len[lane] : 0u;
        const uint32_t order_lane = lane < 19 ? (uint32_t)"\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f"[lane] : 0u;
        // HCLEN: the lengths are sent in the order 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 up to the last one used
        const uint32_t l_ord = __shfl_sync(0xffffffffu, l, order_lane);
        const uint32_t used = __ballot_sync(0xffffffffu, lane < 19 && l_ord != 0);
        const uint32_t hclen = max(4u, 32u - (uint32_t)__clz(used));
        if (lane == 0) {
            uint32_t pos = kHdr * 8;
            put_bits(S.stage, pos, 1, 1); put_bits(S.stage, pos, 2, 2);                   // BFINAL, BTYPE = dynamic Huffman
            put_bits(S.stage, pos, hlit - 257, 5); put_bits(S.stage, pos, S.hdist - 1, 5); put_bits(S.stage, pos, hclen - 4, 4);
            S.hdr_fixed_bits = pos + 3 * hclen;
            S.eob = (uint32_t)S.code[256] | (uint32_t)S.len[256] << 16;
        }
        if (lane < hclen) { uint32_t pos = kHdr * 8 + 17 + 3 * lane; put_bits(S.stage, pos, l_ord, 3); }
    }
    // bits of this warp's segment: its histogram times the code lengths, plus its matches
    const uint32_t seg_off = warp * kSeg;
    uint32_t m_first = 0;                       // matches before this warp's segment
    for (uint32_t w = 0; w < warp; w++) m_first += S.wm[w];
    const uint32_t m_end = m_first + S.wm[warp];
    uint32_t bits = 0;
#pragma unroll
    for (uint32_t q = 0; q < 8; q++) bits += hist[lane + 32 * q] * (S.ctab[lane + 32 * q] >> 16);
    for (uint32_t j = m_first + lane; j < m_end; j += 32) bits += (uint32_t)(S.mbits[j] >> 56);
#pragma unroll
    for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    if (lane) bits = 0;
    __syncthreads();

    // ---- one scan for both the header tokens of the runs and the segments: header bits << 20 | segment bits
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
    const uint32_t z_bytes = (end_bits + 7) / 8 + 8;                 // the member with a dynamic block
    const bool stored = z_bytes - kHdr - 8 >= len + 5 || z_bytes > kStageBytes - 16;     // uniform
    if (stored) {
        // does not shrink (or not enough for the image): stored block (BTYPE 00) straight from the input
        const uint32_t total_bytes = kHdr + 5 + len + 8;
        uint32_t* s32 = reinterpret_cast<uint32_t*>(slot);
        const uint32_t* i32 = reinterpret_cast<const uint32_t*>(in);
        // data byte j lands on slot byte 23 + j: slot word k >= 6 is input bytes [4 (k - 6) + 1, 4 (k - 6) + 5)
        const uint32_t n_words = (total_bytes + 3) / 4;
        for (uint32_t k = 6 + t; k < n_words; k += kZT) s32[k] = __funnelshift_r(__ldg(i32 + (k - 6)), __ldg(i32 + (k - 5)), 8);
        __syncthreads();
        if (t == 0) {
            s32[0] = 0x04088b1fu; s32[1] = 0; s32[2] = 0x0006ff00u; s32[3] = 0x00024342u;
            s32[4] = ((total_bytes - 1) & 0xffffu) | 0x01u << 16 | (len & 0xffu) << 24;      // BSIZE, BFINAL = 1 / BTYPE = 00, LEN
            s32[5] = (len >> 8) | ((len ^ 0xffffu) & 0xffffu) << 8 | (uint32_t)__ldg(in) << 24;   // LEN, NLEN, first byte
            const uint32_t crc_v = S.crc_out;
            for (uint32_t q = 0; q < 4; q++) { slot[23 + len + q] = (uint8_t)(crc_v >> (8 * q)); slot[27 + len + q] = (uint8_t)(len >> (8 * q)); }
            *zlen = total_bytes;
        }
        return;
    }
    if (z_bytes > kHistWord0 * 4) {
        // the image reaches into the histograms (every warp has read its own before the scan's barriers)
        for (uint32_t i = kHistWord0 + t; i < kStageWords; i += kZT) S.stage[i] = 0;
        __syncthreads();
    }
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
    {
        const uint32_t bitpos = hdr_bits + (__shfl_sync(0xffffffffu, excl, 0) & 0xfffffu);
        const uint32_t seg_len = seg_off >= len ? 0u : min(kSeg, len - seg_off);
        if (lz) {
            if (seg_len == kSeg) pack_warp<true, true>(S, in + seg_off, seg_off, seg_len, S.stage, bitpos, m_first, lane);
            else if (seg_len) pack_warp<false, true>(S, in + seg_off, seg_off, seg_len, S.stage, bitpos, m_first, lane);
        } else {
            if (seg_len == kSeg) pack_warp<true, false>(S, in + seg_off, seg_off, seg_len, S.stage, bitpos, m_first, lane);
            else if (seg_len) pack_warp<false, false>(S, in + seg_off, seg_off, seg_len, S.stage, bitpos, m_first, lane);
        }
    }
    if (t == 0) {
        uint32_t pos = end_bits - (eob >> 16);
        put_bits(S.stage, pos, eob & 0xffffu, eob >> 16);
        pos = (z_bytes - 8) * 8;
        put_bits(S.stage, pos, S.crc_out, 32); put_bits(S.stage, pos, len, 32);
        // member header: ID1 ID2 CM FLG(FEXTRA) MTIME XFL OS(255) XLEN=6 'B' 'C' SLEN=2 BSIZE = total - 1
        S.stage[0] = 0x04088b1fu; S.stage[1] = 0; S.stage[2] = 0x0006ff00u; S.stage[3] = 0x00024342u;
        atomicOr(&S.stage[4], (z_bytes - 1) & 0xffffu);
        *zlen = z_bytes;
    }
    __syncthreads();
    const uint4* s4 = reinterpret_cast<const uint4*>(S.stage);
    uint4* o4 = reinterpret_cast<uint4*>(slot);
    for (uint32_t i = t; i < (z_bytes + 15) / 16; i += kZT) o4[i] = s4[i];
}
// The blocks k_bgzf left (usually none), one CTA per 32 blocks of the launch.
__global__ void __launch_bounds__(kZT, 3)
k_bgzf_own(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, const uint64_t* __restrict__ totals, uint32_t nblk_max,
           uint32_t lz, uint8_t* __restrict__ slots0, uint8_t* __restrict__ slots1, uint32_t* __restrict__ zlen0, uint32_t* __restrict__ zlen1) {
    JLP_DYN_SMEM(smem_raw);
    ZShared& S = *reinterpret_cast<ZShared*>(smem_raw);
    __shared__ uint32_t todo;
    // CTA c looks after blocks [32 c, 32 c + 32) of the launch: one load per lane finds the ones left to it
    if (threadIdx.x < 32) {
        const uint32_t blk = 32 * blockIdx.x + threadIdx.x;
        const bool second = blk >= nblk_max;
        const uint32_t b = second ? blk - nblk_max : blk;
        const bool mine = blk < 2 * nblk_max && (uint64_t)b * kBgzfIn < totals[second ? 1 : 0] && (second ? zlen1 : zlen0)[b] == kOwnCode;
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (threadIdx.x == 0) todo = bal;
    }
    __syncthreads();
    for (uint32_t m = todo; m; m &= m - 1u) {
        own_block(S, 32 * blockIdx.x + (uint32_t)__ffs(m) - 1u, in0, in1, totals, nblk_max, lz, slots0, slots1, zlen0, zlen1);
        __syncthreads();       // the next block reuses the shared memory
    }
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
    uint32_t tab4[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        tab4[0][i] = tab[i];
        for (int k = 1; k < 4; k++) tab4[k][i] = (tab4[k - 1][i] >> 8) ^ tab[tab4[k - 1][i] & 0xffu];
    }
    cudaError_t e = cudaMemcpyToSymbol(c_crc_tab, tab4, sizeof tab4);
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
    static uint32_t lvl[10][4][256];
    for (int j = 0; j < 10; j++)
        for (int k = 0; k < 4; k++)
            for (uint32_t x = 0; x < 256; x++) {
                uint32_t v = x << (8 * k);
                // j < 9: one operator; 9: kBgzfIn % kSeg = 3840 = 2^11 + 2^10 + 2^9 + 2^8 bytes, four of them in turn
                for (int a = (j < 9 ? 7 + j : 8); a <= (j < 9 ? 7 + j : 11); a++) {
                    uint32_t r = 0;
                    for (int q = 0; q < 32; q++) if ((v >> q) & 1u) r ^= adv[a][q];
                    v = r;
                }
                lvl[j][k][x] = v;
            }
    e = cudaMemcpyToSymbol(g_crc_lvl, lvl, sizeof lvl);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bgzf_code, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZShared));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bgzf_own, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZShared));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_bgzf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZMain));
}

cudaError_t launch_bgzf(const uint8_t* in0, const uint8_t* in1, uint64_t* totals, uint32_t nblk_max, bool matches, uint8_t* slots0,
                        uint8_t* slots1, uint32_t* zlen0, uint32_t* zlen1, uint64_t* zoff0, uint64_t* zoff1, uint8_t* out0,
                        uint8_t* out1, uint8_t* codes, cudaStream_t s) {
    ZCode* zc = reinterpret_cast<ZCode*>(codes);
    if (nblk_max) {
        JLP_LAUNCH(k_bgzf_code, 2, kZT, sizeof(ZShared), s, in0, in1, totals, matches ? 1u : 0u, zc);
        JLP_LAUNCH(k_bgzf, 2 * nblk_max, kZT, sizeof(ZMain), s, in0, in1, totals, nblk_max, matches ? 1u : 0u, zc, slots0, slots1, zlen0, zlen1);
        JLP_LAUNCH(k_bgzf_own, (2 * nblk_max + 31) / 32, kZT, sizeof(ZShared), s, in0, in1, totals, nblk_max, matches ? 1u : 0u, slots0, slots1, zlen0, zlen1);
    }
    JLP_LAUNCH(k_bgzf_scan, 2, 1024, 0, s, zlen0, zlen1, zoff0, zoff1, totals);
    if (nblk_max)
        JLP_LAUNCH(k_bgzf_gather, 2 * nblk_max, 256, 0, s, slots0, slots1, zlen0, zlen1, zoff0, zoff1, totals, nblk_max, out0, out1);
    return cudaGetLastError();
}

}  // namespace jlp
