// BGZF on the device: the FASTQ of a batch is compressed where it was generated, so that
// only the compressed bytes cross PCIe (the file format is what
// FileBGZF / bgzip_file write through htslib, /root/reference/src/io.h:58-135,
// src/hts.h:140-180: gzip members (RFC 1952) of at most 0xff00 input bytes carrying the
// "BC" extra field with the member's size, each holding one RFC 1951 deflate stream).
//
// One CTA per BGZF block, 512 threads, 3 CTAs per SM.  Thread t owns bytes [128 t, 128 t + 128)
// of the block in the passes that need contiguous bytes, warp w the segment [4096 w, 4096 w + 4096).
//   pass 0  line starts (the byte after every '\n', by a block scan of the newline counts); every
//           line start is compared with the line start 4 lines earlier -- in FASTQ the same line of the
//           previous record -- and a common prefix of >= 4 bytes (the "@<genome>-<chrom>-" of the
//           ID lines, mostly) becomes one length/distance pair; a bit mask marks the bytes so covered
//   pass 1  per-warp histograms of the remaining literals (shared atomics) and the CRC-32 of each
//           thread's chunk; chunk CRCs are combined by a tree of "advance by 2^j zero bytes" operators
//   build   one dynamic-Huffman block: code lengths by two-queue Huffman over the rank-sorted
//           symbols (the tree by one thread, the depths by one thread per leaf), limited to 15 / 7
//           bits by the usual count fix-up, canonical codes; the distance code and the code-length
//           code by one warp; the run-length coded code lengths by one thread per run
//   pass 2  every warp packs its segment, 4 bytes per lane and 128 contiguous bytes per round: the
//           codes of a lane's bytes are joined, a warp scan of the lengths places them, and they are
//           OR-ed into the block image in shared memory
//   a block that would not shrink below 48 KiB is emitted as a stored block straight from the input.
// The image leaves with 128-bit stores into a 64 KiB slot; k_bgzf_scan / k_bgzf_gather
// then make the file contiguous.
#include "jlp_kernels.cuh"

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
constexpr uint32_t kBack = 4;            // a line is compared with the line this many lines earlier
constexpr uint32_t kMaskWords = kBgzfIn / 32 + 2;

__constant__ uint32_t c_crc_tab[256];
__constant__ uint32_t c_crc_adv[17][32];   // operator "advance the CRC register over 2^j zero bytes", by bit image
__constant__ uint32_t c_crc_init_full;     // the initial register 0xffffffff advanced over a full block
// the same operators for 2^7 .. 2^15 zero bytes (the sizes a full block's CRC tree combines) as 4 x 256-entry tables:
// advance(v) = T[0][v & 255] ^ T[1][v >> 8 & 255] ^ T[2][v >> 16 & 255] ^ T[3][v >> 24]
__device__ uint32_t g_crc_lvl[9][4][256];

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

struct ZShared {
    uint32_t stage[kStageWords];          // the block image; its last 16 KiB hold the per-warp literal counts until the packing starts
    uint32_t mask[kMaskWords];            // bit q + 1: byte q of the block is covered by a match
    uint64_t mbits[kMCap];                // the matches in block order: first their tokens (match_token), then their bits | bit count << 56
    uint32_t cnt[288];                    // literal/length counts; [256] = end of block, [257, 286) lengths
    uint32_t crc_tab[256];
    uint32_t ctab[256];                   // code (bit-reversed, LSB first) | length << 16 per literal
    uint32_t sw[288];                     // Huffman: weights of the leaves in sorted order
    uint32_t w_int[288];                  // ... of the internal nodes in creation order
    uint32_t scan_tmp[kZW + 1];
    uint32_t crc_w[kZW];
    uint32_t wm[kZW];                     // matches that start in each warp's segment
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

// the mask bits of bytes [q, q + 32) of the block (bit i: byte q + i), q a multiple of 4
__device__ __forceinline__ uint32_t mask_at(const uint32_t* mask, uint32_t q) {
    const uint32_t g = q + 1;
    return __funnelshift_r(mask[g >> 5], mask[(g >> 5) + 1], g & 31u);
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

// pass 1 of a thread: its 128 contiguous bytes, 16 at a time from global memory -> warp histogram of all bytes
// (the bytes a match covers are taken out again when the match is found), chunk CRC, and the chunk's newlines as
// one bit per byte
template <bool FULL>
__device__ __forceinline__ uint32_t pass1(const uint8_t* chunk, uint32_t my_len, uint32_t* hist, const uint32_t* tab,
                                          uint32_t (&nlm)[4]) {
    uint32_t reg = 0;
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
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int i = k * 16 + q * 4 + r;
                if (FULL || (uint32_t)i < my_len) {
                    const uint32_t b = (w[q] >> (8 * r)) & 0xffu;
                    atomicAdd(&hist[b], 1u);
                    reg = tab[(reg ^ b) & 0xffu] ^ (reg >> 8);
                }
            }
        }
    }
    if (!FULL) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            nlm[j] &= my_len <= 32u * j ? 0u : (my_len - 32u * j >= 32u ? 0xffffffffu : (1u << (my_len - 32u * j)) - 1u);
    }
    return reg;
}

// pass 2 of a warp: its segment again, 4 bytes per lane and 128 contiguous bytes per round.  A byte a match
// covers sends nothing, except the first one, which sends the match.
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
       uint32_t lz, uint8_t* __restrict__ slots0, uint8_t* __restrict__ slots1, uint32_t* __restrict__ zlen0, uint32_t* __restrict__ zlen1) {
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
    const HuffScratch H1{S.sw, S.w_int, S.sorted, S.par_leaf, S.par_int, S.d_int, S.bl};
    const HuffScratch H2{S.sw2, S.w_int2, S.sorted2, S.par_leaf2, S.par_int2, S.d_int2, S.bl2};

    // ---- clear the image (and with it the histograms), the match mask and the counters
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.stage);
        for (uint32_t i = t; i < kStageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t i = t; i < kMaskWords; i += kZT) S.mask[i] = 0;
        if (t < 256) S.crc_tab[t] = c_crc_tab[t];
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
    uint32_t crc = my_len == kChunk ? pass1<true>(in + my_off, my_len, hist, S.crc_tab, nlm) : pass1<false>(in + my_off, my_len, hist, S.crc_tab, nlm);
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
    for (int h = 0; h < 2; h++) {
        const uint32_t k = 2 * t + h;
        if (k >= kBack && k < n_ls) {
            const uint32_t i = S.ls[k];
            if (i < len) {
                const uint32_t p = S.ls[k - kBack];
                const uint32_t next = k + 1 < n_ls ? (uint32_t)S.ls[k + 1] - 1u : len;   // the '\n' that ends this line, or the end
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
                if (m >= 4 && i - p <= 32768u) { m_at[h] = i; m_len[h] = m; m_tok[h] = match_token(m, i - p); }
            }
        }
    }
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
    if (S.maxd > 15) {                         // rare: the 15-bit limit has to act
        if (t == 0) huff_limit(H1, n_active, 15, S.len);
        __syncthreads();
    }
    // canonical codes: first code of the length + the symbols of the same length before this one
    if (t < kNLit) {
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
    bool run_start = false;
    if (t < n_seq) {
        const uint32_t v = seq(t);
        run_start = t == 0 || seq(t - 1) != v;
        if (run_start) {
            uint32_t run = 1;
            while (t + run < n_seq && seq(t + run) == v) run++;
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

cudaError_t launch_bgzf(const uint8_t* in0, const uint8_t* in1, uint64_t* totals, uint32_t nblk_max, bool matches, uint8_t* slots0,
                        uint8_t* slots1, uint32_t* zlen0, uint32_t* zlen1, uint64_t* zoff0, uint64_t* zoff1, uint8_t* out0,
                        uint8_t* out1, cudaStream_t s) {
    if (nblk_max) k_bgzf<<<2 * nblk_max, kZT, sizeof(ZShared), s>>>(in0, in1, totals, nblk_max, matches ? 1u : 0u, slots0, slots1, zlen0, zlen1);
    k_bgzf_scan<<<2, 1024, 0, s>>>(zlen0, zlen1, zoff0, zoff1, totals);
    if (nblk_max)
        k_bgzf_gather<<<2 * nblk_max, 256, 0, s>>>(slots0, slots1, zlen0, zlen1, zoff0, zoff1, totals, nblk_max, out0, out1);
    return cudaGetLastError();
}

}  // namespace jlp
