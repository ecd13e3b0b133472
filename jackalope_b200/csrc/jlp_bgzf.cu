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

__constant__ uint32_t c_crc_tab[256];
__constant__ uint32_t c_crc_adv[17][32];   // operator "advance the CRC register over 2^j zero bytes", by bit image

struct ZShared {
    uint32_t stage[kStageWords];          // the block image
    uint32_t hist[kZW][288];              // per-warp literal counts
    uint32_t cnt[288];                    // literal/length counts; [256] = end of block
    uint32_t crc_tab[256];
    uint32_t crc[kZT];
    uint32_t ctab[256];                   // code (bit-reversed, LSB first) | length << 16 per literal
    uint32_t warp_sum[kZW];
    uint16_t sorted[288];                 // symbols with count > 0, ascending count
    uint8_t len[320];                     // code lengths: [0, 286) literal/length, [286, 316) distance
    uint16_t code[288];
    uint32_t w_int[288];                  // Huffman scratch: internal node weights
    uint16_t par_leaf[288], par_int[288];
    uint8_t d_int[288];
    uint16_t tok[336];                    // run-length tokens of the code lengths: symbol | extra << 5
    uint32_t next_code[16];
    uint32_t n_active, n_tok, hdr_bits, eob, total_bits, stored, crc_out;
};

__device__ __forceinline__ uint32_t crc_apply(const uint32_t* m, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) r ^= ((v >> i) & 1u) ? m[i] : 0u;
    return r;
}
// CRC register advanced over n zero bytes
__device__ uint32_t crc_advance(uint32_t v, uint32_t n) {
    for (int j = 0; n; j++, n >>= 1)
        if (n & 1u) v = crc_apply(c_crc_adv[j], v);
    return v;
}

// serial bit writer into the zeroed block image (one thread)
struct BitW {
    uint32_t* w; uint32_t pos;
    __device__ void put(uint32_t v, uint32_t n) {
        if (!n) return;
        const uint32_t i = pos >> 5, s = pos & 31;
        w[i] |= v << s;
        if (s + n > 32) w[i + 1] |= v >> (32 - s);
        pos += n;
    }
};

// Code lengths (<= max_bits) of the symbols sorted[0 .. n) (ascending count) into len[sym].
// Two-queue Huffman, then the count fix-up for over-long codes, then lengths handed out in
// sorted order (longest code to the rarest symbol).  One thread.
__device__ void huff_lengths(ZShared& S, const uint32_t* cnt, uint32_t n, uint32_t max_bits, uint8_t* len) {
    if (n == 0) return;
    if (n == 1) { len[S.sorted[0]] = 1; return; }
    uint32_t i = 0, k = 0, m = 0;
    while (m + 1 < n) {
        uint32_t w = 0;
        for (int q = 0; q < 2; q++) {
            const bool leaf = i < n && (k >= m || cnt[S.sorted[i]] <= S.w_int[k]);
            if (leaf) { w += cnt[S.sorted[i]]; S.par_leaf[i++] = (uint16_t)m; }
            else { w += S.w_int[k]; S.par_int[k++] = (uint16_t)m; }
        }
        S.w_int[m++] = w;
    }
    uint32_t bl[16];
    for (uint32_t b = 0; b < 16; b++) bl[b] = 0;
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
        bl[d]++;
    }
    while (overflow > 0) {
        uint32_t b = max_bits - 1;
        while (bl[b] == 0) b--;
        bl[b]--; bl[b + 1] += 2; bl[max_bits]--;
        overflow -= 2;
    }
    // Kraft check; a code that does not add up is replaced by a flat one (never seen, kept as a guard)
    uint32_t kraft = 0;
    for (uint32_t b = 1; b <= max_bits; b++) kraft += bl[b] << (max_bits - b);
    if (kraft != (1u << max_bits)) {
        uint32_t flat = 1;
        while ((1u << flat) < n) flat++;
        for (uint32_t q = 0; q < n; q++) len[S.sorted[q]] = (uint8_t)flat;
        return;
    }
    uint32_t q = 0;
    for (uint32_t b = max_bits; b >= 1; b--)
        for (uint32_t c = bl[b]; c; c--) len[S.sorted[q++]] = (uint8_t)b;
}

__device__ __forceinline__ uint32_t rev_bits(uint32_t code, uint32_t n) { return __brev(code) >> (32 - n); }

// canonical codes of a small alphabet (one thread); code[] bit-reversed for LSB-first packing
__device__ void small_codes(const uint8_t* len, uint32_t n_sym, uint32_t max_bits, uint16_t* code) {
    uint32_t bl[16], next[16];
    for (uint32_t b = 0; b < 16; b++) bl[b] = 0;
    for (uint32_t s = 0; s < n_sym; s++) bl[len[s]]++;
    bl[0] = 0;
    uint32_t c = 0;
    for (uint32_t b = 1; b <= max_bits; b++) { c = (c + bl[b - 1]) << 1; next[b] = c; }
    for (uint32_t s = 0; s < n_sym; s++)
        if (len[s]) code[s] = (uint16_t)rev_bits(next[len[s]]++, len[s]);
}

template <bool FULL, typename F>
__device__ __forceinline__ void for_bytes(const uint4 (&d)[8], uint32_t my_len, F f) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t w[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (FULL || (uint32_t)(k * 16 + q * 4 + r) < my_len) f((w[q] >> (8 * r)) & 0xffu);
    }
}

template <bool FULL>
__device__ __forceinline__ void pass1(const uint4 (&d)[8], uint32_t my_len, uint32_t* hist, const uint32_t* tab, uint32_t& crc) {
    uint32_t reg = 0;
    for_bytes<FULL>(d, my_len, [&](uint32_t b) {
        atomicAdd(&hist[b], 1u);
        reg = tab[(reg ^ b) & 0xffu] ^ (reg >> 8);
    });
    crc = reg;
}

template <bool FULL>
__device__ __forceinline__ uint32_t size_bits(const uint4 (&d)[8], uint32_t my_len, const uint32_t* ctab) {
    uint32_t n = 0;
    for_bytes<FULL>(d, my_len, [&](uint32_t b) { n += ctab[b] >> 16; });
    return n;
}

template <bool FULL>
__device__ __forceinline__ void pack(const uint4 (&d)[8], uint32_t my_len, const uint32_t* ctab, uint32_t* stage, uint32_t bitpos) {
    uint32_t wi = bitpos >> 5, nb = bitpos & 31;
    uint64_t acc = 0;
    for_bytes<FULL>(d, my_len, [&](uint32_t b) {
        const uint32_t e = ctab[b];
        acc |= (uint64_t)(e & 0xffffu) << nb;
        nb += e >> 16;
        if (nb >= 32) { atomicOr(&stage[wi++], (uint32_t)acc); acc >>= 32; nb -= 32; }
    });
    if (nb) atomicOr(&stage[wi], (uint32_t)acc);
}

__global__ void __launch_bounds__(kZT, 1)
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

    // ---- the chunk of this thread, once from global memory
    const uint32_t my_off = t * kChunk;
    const uint32_t my_len = my_off >= len ? 0u : min(kChunk, len - my_off);
    uint4 d[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        d[k] = (my_off + 16u * k < len) ? __ldg(reinterpret_cast<const uint4*>(in + my_off) + k) : make_uint4(0, 0, 0, 0);

    // ---- clear the image and the histograms
    {
        uint4* s4 = reinterpret_cast<uint4*>(S.stage);
        for (uint32_t i = t; i < kStageWords / 4; i += kZT) s4[i] = make_uint4(0, 0, 0, 0);
        uint32_t* h = &S.hist[0][0];
        for (uint32_t i = t; i < kZW * 288; i += kZT) h[i] = 0;
        if (t < 256) S.crc_tab[t] = c_crc_tab[t];
        if (t < 320) S.len[t] = 0;
    }
    __syncthreads();

    // ---- pass 1: histogram + chunk CRC
    uint32_t crc;
    if (my_len == kChunk) pass1<true>(d, my_len, S.hist[warp], S.crc_tab, crc);
    else pass1<false>(d, my_len, S.hist[warp], S.crc_tab, crc);
    S.crc[t] = crc;
    __syncthreads();
    uint32_t my_cnt = 0;
    if (t < 288) {
        if (t < 256) for (int w = 0; w < kZW; w++) my_cnt += S.hist[w][t];
        if (t == 256) my_cnt = 1;     // end of block
        S.cnt[t] = my_cnt;
    }
    // CRC tree: the node at t covers chunks [t, t + 2s); the right half has right_len bytes
    for (uint32_t s = 1; s < kZT; s <<= 1) {
        __syncthreads();
        if ((t & (2 * s - 1)) == 0) {
            const uint32_t r0 = (t + s) * kChunk;
            const uint32_t right_len = r0 >= len ? 0u : min(s * kChunk, len - r0);
            S.crc[t] = crc_advance(S.crc[t], right_len) ^ S.crc[t + s];
        }
    }
    const uint32_t n_active = __syncthreads_count(my_cnt != 0);
    // rank sort of the active symbols (ascending count, ties by symbol)
    if (my_cnt) {
        uint32_t rank = 0;
        for (uint32_t s = 0; s <= 256; s++) {
            const uint32_t c = S.cnt[s];
            rank += (c != 0 && (c < my_cnt || (c == my_cnt && s < t))) ? 1u : 0u;
        }
        S.sorted[rank] = (uint16_t)t;
    }
    if (t == 32) S.crc_out = S.crc[0] ^ crc_advance(0xffffffffu, len) ^ 0xffffffffu;
    __syncthreads();

    // ---- literal code lengths (one thread), canonical codes (all)
    if (t == 0) {
        huff_lengths(S, S.cnt, n_active, 15, S.len);
        uint32_t bl[16];
        for (int q = 0; q < 16; q++) bl[q] = 0;
        for (uint32_t s = 0; s <= 256; s++) bl[S.len[s]]++;
        bl[0] = 0;
        uint32_t c = 0;
        for (uint32_t q = 1; q < 16; q++) { c = (c + bl[q - 1]) << 1; S.next_code[q] = c; }
    }
    __syncthreads();
    if (t < 288) {
        const uint32_t l = t < kNLit ? S.len[t] : 0;
        if (l) {
            // canonical order inside a length is by symbol
            uint32_t same = 0;
            for (uint32_t s = 0; s < t; s++) same += S.len[s] == l ? 1u : 0u;
            S.code[t] = (uint16_t)rev_bits(S.next_code[l] + same, l);
        }
    }
    __syncthreads();
    if (t < 256) S.ctab[t] = S.len[t] ? (uint32_t)S.code[t] | (uint32_t)S.len[t] << 16 : 0u;

    // ---- header of the dynamic block: run-length coded code lengths and their own code (one thread)
    if (t == 0) {
        // no length symbols; two unused distance codes of one bit each, as zlib sends for a block of
        // literals ("at least one distance code exists and at least one bit is sent", trees.c)
        const uint32_t hlit = 257, hdist = 2;
        const uint32_t N = hlit + hdist;
        uint8_t* L = S.len;                            // [0, 257) literal/length lengths, then the distance code lengths
        L[257] = 1; L[258] = 1;
        uint32_t nt = 0;
        uint32_t clc[19];
        for (int q = 0; q < 19; q++) clc[q] = 0;
        auto tok = [&](uint32_t sym, uint32_t extra) { S.tok[nt++] = (uint16_t)(sym | extra << 5); clc[sym]++; };
        for (uint32_t i = 0; i < N;) {
            const uint32_t v = L[i];
            uint32_t run = 1;
            while (i + run < N && L[i + run] == v) run++;
            uint32_t r = run;
            if (v == 0) {
                while (r >= 11) { const uint32_t k = r < 138 ? r : 138; tok(18, k - 11); r -= k; }
                if (r >= 3) { tok(17, r - 3); r = 0; }
                while (r) { tok(0, 0); r--; }
            } else {
                tok(v, 0); r--;
                while (r >= 3) { const uint32_t k = r < 6 ? r : 6; tok(16, k - 3); r -= k; }
                while (r) { tok(v, 0); r--; }
            }
            i += run;
        }
        L[257] = 0; L[258] = 0;
        S.n_tok = nt;
        // code-length code: sort its used symbols, lengths <= 7, codes
        uint32_t n_cl = 0;
        for (uint32_t s = 0; s < 19; s++)
            if (clc[s]) {
                uint32_t p = n_cl++;
                while (p > 0 && (clc[S.sorted[p - 1]] > clc[s])) { S.sorted[p] = S.sorted[p - 1]; p--; }
                S.sorted[p] = (uint16_t)s;
            }
        uint8_t cl_len[19];
        uint16_t cl_code[19];
        for (int q = 0; q < 19; q++) { cl_len[q] = 0; cl_code[q] = 0; }
        huff_lengths(S, clc, n_cl, 7, cl_len);
        small_codes(cl_len, 19, 7, cl_code);
        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint32_t hclen = 19;
        while (hclen > 4 && cl_len[order[hclen - 1]] == 0) hclen--;
        BitW bw{S.stage, kHdr * 8};
        bw.put(1, 1); bw.put(2, 2);                   // BFINAL, BTYPE = dynamic Huffman
        bw.put(hlit - 257, 5); bw.put(hdist - 1, 5); bw.put(hclen - 4, 4);
        for (uint32_t q = 0; q < hclen; q++) bw.put(cl_len[order[q]], 3);
        for (uint32_t q = 0; q < nt; q++) {
            const uint32_t sym = S.tok[q] & 31u, extra = S.tok[q] >> 5;
            bw.put(cl_code[sym], cl_len[sym]);
            if (sym == 16) bw.put(extra, 2);
            else if (sym == 17) bw.put(extra, 3);
            else if (sym == 18) bw.put(extra, 7);
        }
        S.hdr_bits = bw.pos;
        S.eob = (uint32_t)S.code[256] | (uint32_t)S.len[256] << 16;
    }
    __syncthreads();

    // ---- pass 2: size, scan, pack
    uint32_t bits = my_len == kChunk ? size_bits<true>(d, my_len, S.ctab) : size_bits<false>(d, my_len, S.ctab);
    uint32_t incl = bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
    if (lane == 31) S.warp_sum[warp] = incl;
    __syncthreads();
    if (t == 0) {
        uint32_t run = S.hdr_bits;
        for (int w = 0; w < kZW; w++) { const uint32_t v = S.warp_sum[w]; S.warp_sum[w] = run; run += v; }
        const uint32_t end_bits = run + (S.eob >> 16);
        S.total_bits = end_bits;
        S.stored = ((end_bits + 7) / 8 - kHdr) >= len + 5 ? 1u : 0u;
    }
    __syncthreads();
    uint32_t bitpos = S.warp_sum[warp] + incl - bits;
    const bool stored = S.stored != 0;
    if (stored) {
        // does not shrink: stored block (BTYPE 00) through the same packing code with the identity table
        const uint32_t clear_words = (S.hdr_bits + 31) / 32 + 1;
        __syncthreads();
        for (uint32_t i = t; i < clear_words; i += kZT) S.stage[i] = 0;
        if (t < 256) S.ctab[t] = t | 8u << 16;
        __syncthreads();
        if (t == 0) {
            BitW bw{S.stage, kHdr * 8};
            bw.put(1, 8);                              // BFINAL = 1, BTYPE = 00, padding to the byte
            bw.put(len, 16); bw.put(len ^ 0xffffu, 16);
            S.total_bits = bw.pos + len * 8;
        }
        bitpos = (kHdr + 5 + my_off) * 8;
        __syncthreads();
    }
    if (my_len == kChunk) pack<true>(d, my_len, S.ctab, S.stage, bitpos);
    else if (my_len) pack<false>(d, my_len, S.ctab, S.stage, bitpos);
    __syncthreads();
    if (t == 0) {
        uint32_t pos = S.total_bits;
        if (!stored) {
            BitW bw{S.stage, pos - (S.eob >> 16)};
            bw.put(S.eob & 0xffffu, S.eob >> 16);
        }
        const uint32_t body_end = (pos + 7) / 8;       // bytes so far
        const uint32_t total = body_end + 8;
        BitW bw{S.stage, body_end * 8};
        bw.put(S.crc_out, 32); bw.put(len, 32);
        // member header: ID1 ID2 CM FLG(FEXTRA) MTIME XFL OS(255) XLEN=6 'B' 'C' SLEN=2 BSIZE = total - 1
        S.stage[0] = 0x04088b1fu; S.stage[1] = 0; S.stage[2] = 0x0006ff00u; S.stage[3] = 0x00024342u;
        S.stage[4] |= (total - 1) & 0xffffu;
        S.total_bits = total;
        *zlen = total;
    }
    __syncthreads();
    const uint32_t total = S.total_bits;
    const uint4* s4 = reinterpret_cast<const uint4*>(S.stage);
    uint4* o4 = reinterpret_cast<uint4*>(slot);
    for (uint32_t i = t; i < (total + 15) / 16; i += kZT) o4[i] = s4[i];
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
