// CUDA kernels of the Illumina read generator (sm_100a).
//
//   k_materialize  haplotype chromosome from sorted mutation records
//                  (HapChrom::get_chrom_full, /root/reference/src/hap_classes.cpp:80-116)
//   k_place        one thread per read end: duplicate leader, (haplotype, chromosome),
//                  fragment length/start, strand, indel counts -> span and read length,
//                  FASTQ record length (chrom_indels_frag / sample_indels /
//                  adjust_chrom_spaces, src/hts_illumina.cpp:116-265)
//   k_scan_*       exclusive scan of FASTQ record lengths -> output offsets
//   k_reads        one warp per run of consecutive pairs: template gather, ART quality/error model, ID line and
//                  FASTQ record assembly in shared memory at file alignment, 128-bit stores
//                  (append_pools / fill_read_qual / fill_fq_lines, src/hts_illumina.cpp:285-482,
//                  src/hts_illumina.h:202-259)
//
// None of this is a dense contraction, so no tensor-core path exists here; the
// kernels are integer/byte work bounded by HBM traffic and integer issue rate
// (DESIGN.md section 5).
#include "jlp_kernels.cuh"
#include "jlp_draws.h"

#include <algorithm>

namespace jlp {

// ------------------------------------------------------------ materialise ---

// One thread per 16 output bases: binary search for the last mutation record
// with new_pos <= p (records may share a new_pos: a deletion has zero width in
// haplotype coordinates), then walk, applying
// HapChrom::get_char_ (src/hap_classes.h:439-455).
__global__ void __launch_bounds__(256)
k_materialize(const uint8_t* __restrict__ ref, uint64_t ref_size, uint64_t n_muts,
              const uint64_t* __restrict__ old_pos, const uint64_t* __restrict__ new_pos,
              const int64_t* __restrict__ size_mod, const uint64_t* __restrict__ nuc_off,
              const uint8_t* __restrict__ pool, uint64_t chrom_size, uint8_t* __restrict__ out) {
    const uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    // The 32 lanes of a warp cover 512 consecutive output bases, so their records lie next to each
    // other: one cooperative 32-ary search per warp (5 rounds of one probe per lane for 2^21 records,
    // instead of 21 dependent loads per thread) finds the last record at or before the warp's first
    // base; each lane then steps forward to its own.
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t pw = p0 - 16ull * lane;                 // the warp's first base
    int64_t lo = -1, hi = (int64_t)n_muts;                 // new_pos[lo] <= pw < new_pos[hi]
    while (hi - lo > 1) {
        const int64_t step = (hi - lo - 1 + 31) / 32;
        const int64_t idx = lo + (int64_t)(lane + 1u) * step;
        const bool ok = idx < hi && new_pos[idx] <= pw;
        const int c = __popc(__ballot_sync(0xffffffffu, ok)); // sorted: the lanes that answer yes form a prefix
        const int64_t nlo = lo + (int64_t)c * step, nhi = lo + (int64_t)(c + 1) * step;
        lo = c ? nlo : lo;
        hi = nhi < hi ? nhi : hi;
    }
    if (p0 >= chrom_size) return;
    int64_t i = lo;
    while (i + 1 < (int64_t)n_muts && new_pos[i + 1] <= p0) i++;
    uint64_t next = (i + 1 < (int64_t)n_muts) ? new_pos[i + 1] : ~0ull;
    uint64_t np = 0, op = 0, no = 0;
    int64_t sm = 0;
    if (i >= 0) { np = new_pos[i]; op = old_pos[i]; sm = size_mod[i]; no = nuc_off[i]; }
    uint32_t w[4] = {0, 0, 0, 0};
    int n = (chrom_size - p0 < 16) ? (int)(chrom_size - p0) : 16;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        if (q < n) {
            uint64_t p = p0 + q;
            while (p >= next) {
                i++;
                np = next; op = old_pos[i]; sm = size_mod[i]; no = nuc_off[i];
                next = (i + 1 < (int64_t)n_muts) ? new_pos[i + 1] : ~0ull;
            }
            uint8_t c;
            if (i < 0) c = ref[p];
            else {
                uint64_t ind = p - np;
                if ((int64_t)ind > sm) c = ref[ind + op - (uint64_t)sm];
                else c = pool[no + ind];
            }
            w[q >> 2] |= (uint32_t)c << (8 * (q & 3));
        }
    }
    if (n == 16) *reinterpret_cast<uint4*>(out + p0) = make_uint4(w[0], w[1], w[2], w[3]);
    else for (int q = 0; q < n; q++) out[p0 + q] = (uint8_t)(w[q >> 2] >> (8 * (q & 3)));
}

cudaError_t launch_materialize(const uint8_t* ref, uint64_t ref_size, uint64_t n_muts,
                               const uint64_t* old_pos, const uint64_t* new_pos,
                               const int64_t* size_mod, const uint64_t* nuc_off,
                               const uint8_t* pool, uint64_t chrom_size, uint8_t* out,
                               cudaStream_t s) {
    if (chrom_size == 0) return cudaSuccess;
    uint64_t n16 = (chrom_size + 15) / 16;
    uint64_t blocks = (n16 + 255) / 256;
    k_materialize<<<(unsigned)blocks, 256, 0, s>>>(ref, ref_size, n_muts, old_pos, new_pos, size_mod,
                                                   nuc_off, pool, chrom_size, out);
    return cudaGetLastError();
}

// --------------------------------------------------------- create_genome ---

// One thread per 16 bases = 4 Philox blocks.  Per base the reference's AliasSampler::sample
// (src/alias_sampler.h:53-60, called from create_chromosomes_, src/create_sequences.cpp:129-132):
// slot i = (uint64)(u1 * 4), then i or Alias[i] depending on u2 < Prob[i].  With four slots the die
// roll is the top two bits of its draw; the coin is decided on 16 bits; the rare ties fetch the rest.
__global__ void __launch_bounds__(256)
k_create_chrom(uint8_t* __restrict__ out, uint64_t len, uint32_t chrom, uint64_t seed, const __grid_constant__ GenomeTables t) {
    // each thread owns one 16-byte-aligned chunk of the OUTPUT (chromosomes start at arbitrary byte offsets of
    // the genome buffer), i.e. bases [16 t - sh, 16 t - sh + 16) of the chromosome
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    const int64_t bs = (int64_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16) - (int64_t)sh;
    if (bs >= (int64_t)len) return;
    uint32_t o[4] = {0, 0, 0, 0};
    U4 w4 = {0, 0, 0, 0};
    uint64_t have = ~0ull;
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) {
        const int64_t bb = bs + (int64_t)k;
        if (bb < 0 || bb >= (int64_t)len) continue;
        const uint64_t b = (uint64_t)bb;
        if ((b >> 2) != have) { have = b >> 2; w4 = genome_block(seed, chrom, have); }
        const uint32_t w = genome_word(w4, (uint32_t)b & 3u);
        uint32_t i = w >> 30;
        if ((w & 0x3fffffffu) == 0x3fffffffu) {                // the low bits could carry: (x + 1) >> 62 on the full draw
            const uint64_t x = genome_draw(seed, chrom, b, 0);
            i = x == ~0ull ? 3u : (uint32_t)((x + 1) >> 62);
        }
        const uint32_t H = (w >> 14) & 0xffffu, th = t.thr16[i];
        bool self = H < th;
        if (H == th) self = genome_draw(seed, chrom, b, 1) < t.thr[i];
        const uint32_t kk = self ? i : t.alias[i];
        o[k >> 2] |= (__byte_perm(0x47414354u, 0u, kk) & 0xffu) << (8u * (k & 3u));
    }
    uint8_t* dst = out + bs;                                   // 16-byte aligned
    if (bs >= 0 && bs + 16 <= (int64_t)len) *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    else
        for (uint32_t k = 0; k < 16; k++)
            if (bs + (int64_t)k >= 0 && bs + (int64_t)k < (int64_t)len) dst[k] = (uint8_t)(o[k >> 2] >> (8u * (k & 3u)));
}

cudaError_t launch_create_chrom(uint8_t* out, uint64_t len, uint32_t chrom, uint64_t seed, const GenomeTables& t,
                                cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const uint64_t n16 = (len + 15 + 15) / 16;                 // + the shift of the first chunk
    k_create_chrom<<<(unsigned)((n16 + 255) / 256), 256, 0, s>>>(out, len, chrom, seed, t);
    return cudaGetLastError();
}

// ------------------------------------------------------------ primitives ---

// nt_map (src/hts.h:36-44) as codes: T,C,A,G -> 0..3, anything else 4.
__device__ __forceinline__ uint32_t nt_code(uint32_t c) {
    // (c >> 1) & 3 sends A,C,T,G to 0,1,2,3; the byte-permute turns that into T0 C1 A2 G3,
    // and a second one rebuilds the ASCII letter to reject every other byte
    uint32_t code = __byte_perm(0x03000102u, 0u, (c >> 1) & 3u) & 0xffu;
    uint32_t asc = __byte_perm(0x47414354u, 0u, code) & 0xffu;
    return asc == c ? code : 4u;
}
// "TCAGN"[code]
__device__ __forceinline__ uint32_t code_ascii(uint32_t code) {
    return __byte_perm(0x47414354u, 0x0000004Eu, code | 0x5550u);   // bytes 1..3 select a zero byte
}

// classification of one template position (sample_indels, src/hts_illumina.cpp:131-146):
// 0 plain base, 1 deletion, 2 insertion
__device__ __forceinline__ int indel_class(const GenParams& p, uint32_t e, uint64_t j, uint32_t t, uint32_t H) {
    const EndDev& E = p.end[e];
    if (H > E.hA) return 0;
    uint64_t x = full_draw(H, p.seed, j, e, PU_INDEL, t);
    if (!E.tA_all && x >= E.tA) return 0;
    if (!E.tI_all && x >= E.tI) return 1;
    return 2;
}

__device__ __forceinline__ uint32_t field16(const U4& w, uint32_t f) {
    uint32_t v = (f >> 1) == 0 ? w.w0 : (f >> 1) == 1 ? w.w1 : (f >> 1) == 2 ? w.w2 : w.w3;
    return (f & 1) ? (v >> 16) : (v & 0xffffu);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}

// -------------------------------------------------------------- placement ---

// One thread per read end: everything scalar about the read.
//   duplicate-chain leader      ReadWriterOneThread::create_reads, src/hts.h:254-280
//   (haplotype, chromosome)     src/hts_illumina.cpp:199-201, :505-528
//   fragment length and start   chrom_indels_frag / indels_frag, src/hts_illumina.cpp:203-217, :245-258
//   strand                      append_pools, src/hts_illumina.cpp:352
//   indel counts -> span, len   sample_indels :116-150, adjust_chrom_spaces :153-184
//   record length, start digits fill_fq_lines :285-326
__global__ void __launch_bounds__(256)
k_place(const __grid_constant__ GenParams p) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.batch_pairs * p.n_ends) return;
    const uint32_t e = (p.n_ends == 2) ? (r & 1u) : 0u;
    const uint64_t j = p.batch_lo + (p.n_ends == 2 ? (r >> 1) : r);
    const uint32_t L = p.L;

    uint64_t k = j;
    if (!p.dup_never) {
        while (k > p.job_lo && ((k - p.job_lo) % p.pool_pairs) != 0) {
            uint64_t xd = hi64(draw_block(p.seed, k - 1, 1, PL_PAIR, 0));
            if (!(xd < p.c_dup)) break;
            k--;
        }
    }
    uint32_t g;
    {
        uint32_t lo = 0, hi = p.n_groups;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.group_off[mid] <= k) lo = mid; else hi = mid;
        }
        g = lo;
    }
    const GroupDev G = p.groups[g];
    const uint64_t chrom_len = G.len;

    uint64_t frag_len, frag_start;
    {
        U4 w = draw_block(p.seed, k, 0, PL_PAIR, 0);
        uint64_t xf = lo64(w);
        uint32_t lo = 0, hi = p.frag_n;                // first i with cdf[i] > xf
        if (hi) { lo = p.frag_guide[xf >> 56]; hi = p.frag_guide[(xf >> 56) + 1]; }
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.frag_cdf[mid] <= xf) lo = mid + 1; else hi = mid;
        }
        frag_len = p.frag_min + lo;
        if (frag_len >= chrom_len) { frag_len = chrom_len; frag_start = 0; }
        else {
            // double u = runif_01(eng); frag_start = u * (chrom_len - frag_len + 1)   (src/hts_illumina.cpp:215-216)
            uint64_t xs = hi64(w);
            double u = (xs == ~0ull) ? 1.0 : __ull2double_rn(xs + 1) * 5.421010862427522170037e-20;
            frag_start = __double2ull_rz(__dmul_rn(u, __ull2double_rn(chrom_len - frag_len + 1)));
        }
    }
    bool reverse = lo64(draw_block(p.seed, j, 1, PL_PAIR, 0)) < p.c_rev;
    if (e) reverse = !reverse;

    uint32_t frag_pos = 0, length_now = 0, n_ev = 0, ev0 = 0xffffffffu, ev1 = 0xffffffffu;
    const uint32_t frag_cap = frag_len > 0xffffffffull ? 0xffffffffu : (uint32_t)frag_len;
    const uint32_t hA = p.end[e].hA;
    // all eight fields of a block above the gate, tested two at a time: with hA + 1 <= 0x8000 a field f is above it iff
    // f has its top bit set or (f | 0x8000) - (hA + 1) keeps it (no borrow crosses the halves: each is >= 0x8000 > hA)
    const bool swar = hA < 0x8000u;
    const uint32_t gate2 = (hA + 1u) * 0x00010001u;
    while (length_now < L && frag_pos < frag_cap) {
        // 8 template positions per Philox block, 16 bits each: a field above the gate is a plain base.
        // (An 8-bit gate would halve the Philox work but make a candidate -- and with it the slow path of
        // the whole warp -- 256 times more likely; measured: 0.46 ms instead of 0.25 ms.)
        U4 w = draw_block(p.seed, j, frag_pos >> 3, PL_INDEL, e);
        bool plain;
        if (swar) {
            const uint32_t a0 = w.w0 | ((w.w0 | 0x80008000u) - gate2), a1 = w.w1 | ((w.w1 | 0x80008000u) - gate2);
            const uint32_t a2 = w.w2 | ((w.w2 | 0x80008000u) - gate2), a3 = w.w3 | ((w.w3 | 0x80008000u) - gate2);
            plain = (a0 & a1 & a2 & a3 & 0x80008000u) == 0x80008000u;
        } else {
            plain = ((w.w0 & 0xffffu) > hA) & ((w.w0 >> 16) > hA) & ((w.w1 & 0xffffu) > hA) & ((w.w1 >> 16) > hA) &
                    ((w.w2 & 0xffffu) > hA) & ((w.w2 >> 16) > hA) & ((w.w3 & 0xffffu) > hA) & ((w.w3 >> 16) > hA);
        }
        if (plain && (frag_pos & 7u) == 0 && length_now + 8 <= L && frag_pos + 8 <= frag_cap) {
            length_now += 8; frag_pos += 8;
            continue;
        }
        uint32_t f = frag_pos & 7u;
        for (; f < 8 && length_now < L && frag_pos < frag_cap; f++, frag_pos++) {
            int cls = indel_class(p, e, j, frag_pos, field16(w, f));
            if (cls == 2 && length_now == L - 1) cls = 0;  // src/hts_illumina.cpp:138-139: treated as a plain base
            if (cls == 0) { length_now++; continue; }
            length_now += cls == 2 ? 2u : 0u;
            const uint32_t ev = frag_pos | ((uint32_t)cls << 16);
            if (n_ev == 0) ev0 = ev; else if (n_ev == 1) ev1 = ev;
            n_ev++;
        }
    }
    uint32_t S = frag_pos, len = length_now;
    const uint32_t b = G.bc_len;
    if (S <= b) {                                       // barcode covers the whole template: undefined in the reference
        atomicOr(p.status, 1u);
        S = b; len = 0; n_ev = 0;
    }
    const uint32_t space = S - b;
    const uint64_t start = (p.matepair != 0) == reverse ? frag_start : frag_start + frag_len - space;

    // ---- length of the ID line "@<genome>-<chrom>-<start>-<F|R>[/<1|2>]\n" (fill_fq_lines, src/hts_illumina.cpp:296-312);
    //      its text is written by the read kernel
    uint32_t nd = 1;
    for (uint64_t v = start; v >= 10; v /= 10) nd++;
    const uint32_t idlen = G.prefix_len + nd + 3u + (p.n_ends == 2 ? 2u : 0u);
    uint32_t flags = (reverse ? kPlanReverse : 0u) | (n_ev ? kPlanIndels : 0u) | (b ? kPlanBarcode : 0u) | (n_ev > 2 ? kPlanManyEv : 0u);
    if (idlen > 255u) flags |= kPlanLongId;
    const uint32_t rec_len = idlen + 2u * len + 4u;   // ID line | read '\n' '+' '\n' qual '\n'
    const uint64_t sega = reinterpret_cast<uint64_t>(G.seq + start);
    uint4* dst = reinterpret_cast<uint4*>(p.plan + r);
    dst[0] = make_uint4((uint32_t)sega, (uint32_t)(sega >> 32), S, len | (flags << 16) | ((idlen & 0xffu) << 24));
    dst[1] = make_uint4(rec_len, g, ev0, ev1);
    p.rec_len[r] = rec_len;
}

cudaError_t launch_place(const GenParams& p, cudaStream_t s) {
    uint32_t n = p.batch_pairs * p.n_ends;
    if (n == 0) return cudaSuccess;
    k_place<<<(n + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------- quality / errors / FASTQ ---

// Shared memory is addressed through 32-bit shared-window addresses and explicit
// ld.shared / st.shared, so that no access goes through a generic pointer.
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128s(uint32_t a) { return lds128(a); }
// tables: written once before the CTA-wide barrier, read-only afterwards (no "memory" clobber,
// so the loads can be scheduled freely)
__device__ __forceinline__ uint32_t lds32_ro(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64_ro(uint32_t a) { uint2 v; asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }

__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ U4 qual_block(const GenParams& p, uint64_t j, uint32_t blk, uint32_t e) {
    return philox4x32_rk<kQualRounds>((uint32_t)j, (uint32_t)(j >> 32), blk, PL_QUAL + e * 256u, p.rk);
}

// One base evaluated exactly (full 64-bit draws wherever the 16 high bits do not decide);
// every rare case of fill_read_qual (src/hts_illumina.h:230-256) lands here: non-TCAG bases,
// ambiguous high bits, mismatches.  Returns ascii | qualchar << 8.
template <bool SMEM>
__device__ __noinline__ uint32_t base_rare(const GenParams& p, uint32_t meta_a, uint32_t ent_a, uint32_t e, uint64_t j,
                                           uint32_t pos, uint32_t code, uint32_t die, uint32_t cm) {
    const EndDev& E = p.end[e];
    const uint32_t Hcoin = cm >> 16, Hmis = cm & 0xffffu;
    if (code > 3) {
        // 'N' with a quality below 10 (src/hts_illumina.h:237-242)
        const uint32_t dn = die & 0xffffff00u;
        uint32_t qc = __umulhi(dn, 10u) + 33u;
        if (dn * 10u + 2560u < 2560u) qc = nqual_x87(full_draw(die >> 8, p.seed, j, e, PU_DIE, pos));
        return 0x4Eu | ((qc & 0xffu) << 8);
    }
    uint32_t m;
    if (SMEM) m = lds32_ro(meta_a + (code * p.L + pos) * 4u);
    else m = __ldg(E.meta + code * p.L + pos);
    const uint32_t n = m & 0xffu, off = m >> 8;
    const uint32_t dc = die & 0xffffff00u;      // the high 24 bits of X_die (the low byte of the word belongs to X_sub)
    uint32_t i = __umulhi(dc, n);
    if (dc * n + (n << 8) < (n << 8)) {         // the draw's low 40 bits can carry into the slot index
        uint64_t ii = mul_floor_x87(full_draw(die >> 8, p.seed, j, e, PU_DIE, pos), n);
        i = ii >= n ? n - 1u : (uint32_t)ii;
    }
    uint2 ent;
    if (SMEM) ent = lds64_ro(ent_a + (off + i) * 8u);
    else ent = __ldg(reinterpret_cast<const uint2*>(E.entry64) + off + i);
    const uint32_t thr = ent.x >> 16;
    bool self = Hcoin < thr;
    if (Hcoin == thr) self = full_draw(Hcoin, p.seed, j, e, PU_COIN, pos) < E.coin[off + i];
    const uint32_t qc = self ? ((ent.x >> 8) & 0xffu) : (ent.x & 0xffu);          // quality character
    const uint32_t mt = self ? (ent.y & 0xffffu) : (ent.y >> 16);
    bool mism = Hmis < mt;
    if (Hmis == mt) mism = full_draw(Hmis, p.seed, j, e, PU_MIS, pos) < E.mis[(qc - 33u) & 0xffu];
    if (mism) {
        // mm_nucleos[nt][(uint64)(u * 3)] (src/hts.h:46): the si-th code other than `code`
        const uint32_t Hsub = die & 0xffu, p3 = Hsub * 3u;
        uint32_t si = p3 >> 8;
        if ((p3 & 0xffu) + 3u > 0xffu) {
            uint64_t s3 = mul_floor_x87(full_draw(Hsub, p.seed, j, e, PU_SUB, pos), 3);
            si = s3 > 2 ? 2u : (uint32_t)s3;
        }
        code = si + (si >= code ? 1u : 0u);
    }
    return (code_ascii(code) & 0xffu) | (qc << 8);
}

// The common case of one base, decided on the 16 high bits of its three draws: quality by the
// alias method (IllQualPos::sample, src/hts_illumina.h:128-132; AliasSampler::sample,
// src/alias_sampler.h:53-60) and the mismatch test (src/hts_illumina.h:251-252).
// ct: base code clamped to 0..3 (keeps the table walk in bounds for 'N').  Outputs the table
// entry's low word, which of its two qualities was drawn, whether the base is a mismatch, and
// whether base_rare() must decide instead (ambiguous high bits).
template <bool SMEM>
__device__ __forceinline__ void base_fast(const GenParams& p, uint32_t meta_a, uint32_t ent_a, uint32_t e, uint32_t pos,
                                          uint32_t ct, uint32_t die, uint32_t cm, uint32_t& entx, bool& self, bool& rare, bool& mism) {
    uint32_t m;
    if (SMEM) m = lds32_ro(meta_a + (ct * p.L + pos) * 4u);
    else m = __ldg(p.end[e].meta + ct * p.L + pos);
    const uint32_t n = m & 0xffu, nsh = __byte_perm(m, 0u, 0x4404u);   // n and n << 8
    // die * n / 2^32 is the alias slot.  The word's low byte belongs to X_sub, not to X_die: it is masked off, so the
    // product's fraction lacks at most n * 2^8 and the slot is decided unless adding that much carries.
    const uint32_t dc = die & 0xffffff00u;
    const uint32_t slot = (m >> 8) + __umulhi(dc, n);
    uint2 ent;
    if (SMEM) ent = lds64_ro(ent_a + slot * 8u);
    else ent = __ldg(reinterpret_cast<const uint2*>(p.end[e].entry64) + slot);
    // coin: high half of cm against the threshold in the high half of the entry, compared in place
    const uint32_t thr_hi = ent.x & 0xffff0000u;
    self = cm < thr_hi;
    const uint32_t mt = __byte_perm(ent.y, 0u, self ? 0x4410u : 0x4432u);      // high 16 bits of the quality's mismatch threshold
    // coin and mismatch are undecided when their 16 high bits equal the threshold's
    const uint32_t mh = cm & 0xffffu;
    mism = mh < mt;
    rare = (dc * n + nsh < nsh) || (cm - thr_hi < 0x10000u) || (mh == mt);
    entx = ent.x;
}

// Base code at template position t of a read end (fill_read / rev_comp / barcode,
// src/hts_illumina.cpp:369-391): barcode first, then the template, reverse-complemented
// (cmp_map: T<->A, C<->G, everything else becomes 'N' later) on the reverse strand.
__device__ __forceinline__ uint32_t template_code(const uint8_t* __restrict__ seg, const uint8_t* __restrict__ bc,
                                                  uint32_t t, uint32_t b, uint32_t space, bool reverse) {
    if (t < b) return nt_code(bc[t]);
    uint32_t i = t - b;
    uint32_t code = nt_code(reverse ? seg[space - 1u - i] : seg[i]);
    if (reverse && code < 4u) code ^= 2u;
    return code;
}

__device__ __noinline__ uint32_t codes4_fix(uint32_t code, uint32_t bad) {
#pragma unroll
    for (uint32_t k = 0; k < 4; k++)
        if ((bad >> (8u * k)) & 0xffu) code = (code & ~(0xffu << (8u * k))) | (4u << (8u * k));
    return code;
}

// Four ASCII bases in one word -> four base codes (T0 C1 A2 G3, anything else 4),
// complemented on the reverse strand.
__device__ __forceinline__ uint32_t codes4(uint32_t x, bool reverse) {
    uint32_t x1 = (x >> 1) & 0x03030303u;                       // A0 C1 T2 G3
    uint32_t code = x1 ^ 0x02020202u ^ ((x1 << 1) & 0x02020202u);  // T0 C1 A2 G3
    uint32_t t = code | (code >> 4);
    uint32_t sel = __byte_perm(t, 0u, 0x4420u);                 // one selector nibble per base
    uint32_t bad = __byte_perm(0x47414354u, 0u, sel) ^ x;       // non-zero byte: not T/C/A/G
    if (reverse) code ^= 0x02020202u;
    if (bad) code = codes4_fix(code, bad);
    return code;
}

// 8 consecutive bytes starting at an arbitrary address, from three aligned word loads
__device__ __forceinline__ void load8(const uint8_t* a, uint32_t& lo, uint32_t& hi) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)u & 3u) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

// Template codes of an end WITH insertions / deletions: every template position gets its
// class again (same draws as the placement kernel); a warp prefix sum gives each surviving
// base its place in the read (fill_read_qual applies the same edits from the back,
// src/hts_illumina.h:213-225).
__device__ __noinline__ void gather_indels(const GenParams& p, uint32_t e, uint64_t j, uint32_t w, const uint8_t* seg,
                                           const uint8_t* bc, uint32_t S, uint32_t b, uint32_t ln, bool reverse) {
    const uint32_t lane = threadIdx.x & 31u, L = p.L, space = S - b;
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < S; t0 += 256) {
        const uint32_t tb = t0 + 8u * lane;
        uint32_t cls = 0, wsum = 0;
        if (tb < S) {
            U4 dw = draw_block(p.seed, j, tb >> 3, PL_INDEL, e);
#pragma unroll
            for (uint32_t f = 0; f < 8; f++) {
                if (tb + f < S) {
                    uint32_t c = (uint32_t)indel_class(p, e, j, tb + f, field16(dw, f));
                    cls |= c << (2u * f);
                    wsum += c == 0 ? 1u : c == 1 ? 0u : 2u;
                }
            }
        }
        uint32_t incl = warp_incl_scan(wsum);
        uint32_t o2 = carry + incl - wsum;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        if (tb < S) {
#pragma unroll 1
            for (uint32_t f = 0; f < 8; f++) {
                const uint32_t t = tb + f;
                if (t < S) {
                    uint32_t c = (cls >> (2u * f)) & 3u;
                    if (c == 2u && o2 == L - 1u) c = 0;     // src/hts_illumina.cpp:138-139
                    if (c != 1u && o2 < ln) {
                        sts8(w + o2++, template_code(seg, bc, t, b, space, reverse));
                        if (c == 2u && o2 < ln) {
                            // bases[(uint64)(u * 4)], src/hts_illumina.h:216; index 4 reads the
                            // string terminator, which nt_map then turns into 'N'
                            sts8(w + o2++, ins_base_index(slow64(p.seed, j, e, PU_INS, t)));
                        }
                    }
                }
            }
        }
    }
}

// Template codes of an end with a barcode and no indels (8 positions per lane, byte-wise)
__device__ __noinline__ void gather_barcode(uint32_t w, const uint8_t* seg, const uint8_t* bc, uint32_t S, uint32_t b,
                                            uint32_t ln, bool reverse) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t t = lane; t < ln; t += 32) sts8(w + t, template_code(seg, bc, t, b, S - b, reverse));
}

#ifndef JLP_READS_THREADS
// 20 warps, one CTA per SM sharing one copy of the tables: 96 registers per thread, nothing spilled.  Measured
// (before the last changes to phase B): 896 threads (72 registers, spills) 1.56 ms, 768 1.54, 640 1.46, 512 1.50.
#define JLP_READS_THREADS 640
#define JLP_READS_CTAS 1
#endif
constexpr int kReadsThreads = JLP_READS_THREADS;
constexpr uint32_t kTplSlots = 2;       // staged template windows per warp: pair k is read while pair k+1 is on its way

// bytes of shared memory one warp of k_reads owns (host and device agree through this):
// 1 mbarrier (16 bytes) | 64 bytes of scratch | 2 chunks of plans | kTplSlots template windows per end |
// one base-code line per end | one output buffer per end
__host__ __device__ inline uint32_t reads_warp_bytes(uint32_t n_ends, uint32_t rec_buf, uint32_t tpl_buf, uint32_t cod_buf) {
    return 80u + 2u * kPlanChunk * n_ends * 32u + kTplSlots * n_ends * tpl_buf + n_ends * cod_buf + n_ends * rec_buf;
}
constexpr uint32_t kReadsCtaBytes = 80;  // per CTA, in front of everything: ceil(2^64 / 10^k), k = 0..9

// ---- bulk asynchronous copies (cp.async.bulk: SASS UBLKCP) and their mbarriers
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(a), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier; both addresses and the size are multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
// generic-proxy accesses to shared memory before, asynchronous-proxy accesses after
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ID line of a record with a very long genome / chromosome name, or a coordinate >= 2^32: one lane writes it byte by byte.
__device__ __noinline__ void slow_idline(const GenParams& p, uint32_t dst, const GroupDev* Gp, uint64_t start, bool reverse, uint32_t e) {
    const uint8_t* pre = p.strpool + Gp->prefix_off;
    const uint32_t n = Gp->prefix_len;
    for (uint32_t t = 0; t < n; t++) sts8(dst + t, pre[t]);
    dst += n;
    uint8_t dg[20];
    uint32_t nd = 0;
    do { dg[nd++] = (uint8_t)('0' + (uint32_t)(start % 10)); start /= 10; } while (start);
    for (uint32_t t = 0; t < nd; t++) sts8(dst + t, dg[nd - 1u - t]);
    dst += nd;
    sts8(dst++, '-');
    sts8(dst++, reverse ? 'R' : 'F');
    if (p.n_ends == 2) { sts8(dst++, '/'); sts8(dst++, '1' + e); }
    sts8(dst, '\n');
}

// The one or two insertions / deletions of a read applied to its line of template codes (code[t] = template position t,
// S of them), last event first: a deletion at t drops code[t] (the tail moves one to the left), an insertion at t puts
// bases[(uint64)(u * 4)] behind code[t] (the tail moves one to the right).  The whole warp works on one end.
__device__ __noinline__ void apply_events(const GenParams& p, uint32_t e, uint64_t j, uint32_t C, uint32_t S, uint32_t ev0, uint32_t ev1) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t top = S;                                         // codes in the line
#pragma unroll 1
    for (int k = 1; k >= 0; k--) {
        const uint32_t ev = k ? ev1 : ev0;
        if (ev == 0xffffffffu) continue;
        const uint32_t t = ev & 0xffffu;
        if ((ev >> 16) == 1u) {                               // deletion: code[x] = code[x + 1] for x >= t, front to back
            for (uint32_t x0 = t; x0 + 1u < top; x0 += 32u) {
                const uint32_t x = x0 + lane;
                const uint32_t v = x + 1u < top ? lds8(C + x + 1u) : 0u;
                __syncwarp();
                if (x + 1u < top) sts8(C + x, v);
                __syncwarp();
            }
            top--;
        } else {                                              // insertion: code[x + 1] = code[x] for x > t, back to front
            for (uint32_t hi = top; hi > t + 1u; hi = hi > t + 33u ? hi - 32u : t + 1u) {
                const uint32_t lo = hi > t + 33u ? hi - 32u : t + 1u, x = lo + lane;
                const uint32_t v = x < hi ? lds8(C + x) : 0u;
                __syncwarp();
                if (x < hi) sts8(C + x + 1u, v);
                __syncwarp();
            }
            if (lane == 0) sts8(C + t + 1u, ins_base_index(slow64(p.seed, j, e, PU_INS, t)));
            __syncwarp();
            top++;
        }
    }
}

// ceil(2^64 / 10^k), k = 1..9: floor(n / 10^k) = umul64hi(n, c_m10[k]) for every n < 2^32
__constant__ uint64_t c_m10[10] = {0ull, 1844674407370955162ull, 184467440737095517ull, 18446744073709552ull, 1844674407370956ull,
                                   184467440737096ull, 18446744073710ull, 1844674407371ull, 184467440738ull, 18446744074ull};

// Four ASCII bases in one word -> four base codes (T0 C1 A2 G3), complemented on the reverse strand; `bad` collects
// the bytes that are not T/C/A/G (non-zero: codes4_fix sets those codes to 4).
__device__ __forceinline__ uint32_t codes4_acc(uint32_t x, bool reverse, uint32_t& bad) {
    uint32_t x1 = (x >> 1) & 0x03030303u;                       // A0 C1 T2 G3
    uint32_t code = x1 ^ 0x02020202u ^ ((x1 << 1) & 0x02020202u);  // T0 C1 A2 G3
    uint32_t t = code | (code >> 4);
    uint32_t sel = __byte_perm(t, 0u, 0x4420u);                 // one selector nibble per base
    bad = __byte_perm(0x47414354u, 0u, sel) ^ x;
    return reverse ? code ^ 0x02020202u : code;
}

// Phase B of k_reads for one pair: both ends in one index space, two bases per lane and Philox block -- quality by the
// alias method, mismatch test, substitution, written straight into the records (sq0 / sq1: shared addresses of the two
// sequence lines; cd: the code lines, codw bytes apart).  Undecided draws branch to base_rare().
// inlined: as a function of its own (registers allocated apart from the bookkeeping around it) it was slower, 1.63 ms against 1.46
#ifdef JLP_PHASEB_NOINLINE
#define JLP_PHASEB_ATTR __noinline__
#else
#define JLP_PHASEB_ATTR __forceinline__
#endif
template <bool SMEM, uint32_t NE, bool EVEN>
__device__ __forceinline__ void phase_b_loop(const GenParams& p, uint64_t j, uint32_t lane, uint32_t sq0, uint32_t sq1, uint32_t len0, uint32_t len1,
                                     uint32_t cd, uint32_t codw, uint32_t mA0, uint32_t mA1, uint32_t eA0, uint32_t eA1) {
    // One half-warp per end (single-end: the whole warp): which end a lane works on never changes, so everything that
    // depends on the end is set up once per pair, outside the loop.  Same number of passes as one index space over both
    // ends (ceil(n / 16) = ceil(2n / 32) for equally long reads).
    const uint32_t e = NE == 2 ? lane >> 4 : 0u;
    const uint32_t W = NE == 2 ? 16u : 32u;
    const uint32_t ln = e ? len1 : len0;
    const uint32_t nb = (ln + 1u) >> 1;
    const uint32_t sq = e ? sq1 : sq0, meta_a = e ? mA1 : mA0, ent_a = e ? eA1 : eA0, ca = cd + e * codw;
#pragma unroll 1
    for (uint32_t blk = NE == 2 ? lane & 15u : lane; blk < nb; blk += W) {
        const uint32_t pos = 2u * blk;
        const uint32_t s0 = sq + pos;
        const uint32_t q0 = s0 + ln + 3u;
        const bool two = EVEN || pos + 1u < ln;        // EVEN: both reads have an even length, every block holds two bases
        const U4 w = qual_block(p, j, blk, e);
        const uint32_t cl = lds16(ca + pos);
        const uint32_t cc = two ? cl : cl & 0xffu;          // the block's two base codes, one per byte
        const uint32_t c0 = cc & 0xffu, c1 = cc >> 8, pos1 = two ? pos + 1u : pos;
        const uint32_t ct0 = min(c0, 3u), ct1 = min(c1, 3u);
        uint32_t x0, x1;
        bool self0, self1, rare0, rare1, mis0, mis1;
        base_fast<SMEM>(p, meta_a, ent_a, e, pos, ct0, w.w0, w.w1, x0, self0, rare0, mis0);
        base_fast<SMEM>(p, meta_a, ent_a, e, pos1, ct1, w.w2, w.w3, x1, self1, rare1, mis1);
        // both quality characters with one byte permute; both letters with another whose selector is the pair of codes
        // itself (nibbles c0, 0, c1, 0: T C A G, and N for code 4): letters in bytes 0 and 2
        uint32_t qq = __byte_perm(x0, x1, (self0 ? 1u : 0u) | (self1 ? 0x50u : 0x40u));
        uint32_t asc = __byte_perm(0x47414354u, 0x4e4e4e4eu, cc);
        if (mis0 | mis1) {
            // mm_nucleos[nt][(uint64)(u * 3)] (src/hts.h:46): the si-th code other than the base's, si from the 8 high
            // bits of X_sub unless the low bits could change it (3 of 256 values: the exact path decides)
            if (mis0 && !rare0 && c0 <= 3u) {
                const uint32_t p0 = (w.w0 & 0xffu) * 3u;
                if ((p0 & 0xffu) >= 253u) {
                    const uint32_t r = base_rare<SMEM>(p, meta_a, ent_a, e, j, pos, c0, w.w0, w.w1);
                    asc = (asc & 0xff0000u) | (r & 0xffu);
                    qq = (qq & 0xff00u) | (r >> 8);
                } else {
                    const uint32_t si = p0 >> 8, k0 = si + (si >= ct0 ? 1u : 0u);
                    asc = (asc & 0xff0000u) | (__byte_perm(0x47414354u, 0u, k0) & 0xffu);
                }
            }
            if (mis1 && !rare1 && c1 <= 3u && two) {
                const uint32_t p1 = (w.w2 & 0xffu) * 3u;
                if ((p1 & 0xffu) >= 253u) {
                    const uint32_t r = base_rare<SMEM>(p, meta_a, ent_a, e, j, pos1, c1, w.w2, w.w3);
                    asc = (asc & 0xffu) | ((r & 0xffu) << 16);
                    qq = (qq & 0xffu) | (r & 0xff00u);
                } else {
                    const uint32_t si = p1 >> 8, k1 = si + (si >= ct1 ? 1u : 0u);
                    asc = (asc & 0xffu) | ((__byte_perm(0x47414354u, 0u, k1) & 0xffu) << 16);
                }
            }
        }
        if (rare0 || rare1 || (cc & 0x0404u) != 0u) {          // undecided draws; a base that is not T, C, A or G (code 4)
            if (rare0 || c0 > 3u) {
                uint32_t r = base_rare<SMEM>(p, meta_a, ent_a, e, j, pos, c0, w.w0, w.w1);
                asc = (asc & 0xff0000u) | (r & 0xffu);
                qq = (qq & 0xff00u) | (r >> 8);
            }
            if (two && (rare1 || c1 > 3u)) {
                uint32_t r = base_rare<SMEM>(p, meta_a, ent_a, e, j, pos1, c1, w.w2, w.w3);
                asc = (asc & 0xffu) | ((r & 0xffu) << 16);
                qq = (qq & 0xffu) | (r & 0xff00u);
            }
        }
        sts8(s0, asc);
        sts8(q0, qq);
        if (two) {
            sts8(s0 + 1u, asc >> 16);
            sts8(q0 + 1u, qq >> 8);
        }
    }
}

template <bool SMEM, uint32_t NE>
__device__ JLP_PHASEB_ATTR void phase_b(const GenParams& p, uint64_t j, uint32_t lane, uint32_t sq0, uint32_t sq1, uint32_t len0, uint32_t len1,
                                     uint32_t cd, uint32_t codw, uint32_t mA0, uint32_t mA1, uint32_t eA0, uint32_t eA1) {
    if (((len0 | len1) & 1u) == 0u) phase_b_loop<SMEM, NE, true>(p, j, lane, sq0, sq1, len0, len1, cd, codw, mA0, mA1, eA0, eA1);
    else phase_b_loop<SMEM, NE, false>(p, j, lane, sq0, sq1, len0, len1, cd, codw, mA0, mA1, eA0, eA1);
}

// One warp per run of consecutive read pairs: the R1 records of a run are one contiguous span of file 1, its R2
// records one of file 2, so a record is assembled in shared memory AT ITS FILE ALIGNMENT: the 16-byte chunks it
// completes go to the file with one 128-bit load and one 128-bit store per lane and chunk (no shifting), the
// unfinished last chunk moves to the front of the buffer and the next record continues behind it.  Only the two
// ragged ends of a run are written byte by byte.
//   staging   the plans of 8 pairs per bulk copy (cp.async.bulk global -> shared on an mbarrier, SASS UBLKCP); the
//             bytes around both templates of pair k+1 by cp.async (16 bytes per lane) while pair k is processed
//   phase A   template bytes -> base codes (T0 C1 A2 G3, other 4), reverse-complemented on the reverse strand,
//             16 positions per lane from aligned shared words; ends with indels / a barcode take the whole warp
//   ID line   written in place, one byte per lane: the group's prefix from a per-warp cache, one decimal digit per lane
//   phase B   (both ends in one index space, two bases per lane and Philox block) quality by the alias method,
//             mismatch test, substitution, straight into the record; undecided draws branch to base_rare()
//   flush     all lanes: chunk c of the two records' finished chunks, shared -> global
// (Measured: the same flush as cp.async.bulk shared -> global copies, one per record, costs about twice the
//  instructions -- operands through uniform registers, proxy fence, commit / wait, a second buffer -- DESIGN.md.)
// What changes once per pair lives in the warp's scratch, not in registers: per end where the next record starts in
// its buffer (apos), the bytes of the run's first chunk that belong to the previous run (hole), and the offset of
// byte 0 of the buffer in the file's batch buffer (goff).  The template slot and the mbarrier parity follow from the
// pair's index in the run.  NE: read ends per fragment (compile-time: paired / single).
template <bool SMEM, uint32_t NE>
__global__ void __launch_bounds__(kReadsThreads, JLP_READS_CTAS)
k_reads(const __grid_constant__ GenParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t L = p.L;
    // table placement is a function of the launch parameters only
    const uint32_t ent0 = kReadsCtaBytes;
    const uint32_t meta0 = ent0 + p.end[0].entry_n * 8u;
    const uint32_t ent1 = (meta0 + 16u * L + 15u) & ~15u;
    const uint32_t meta1 = ent1 + (NE == 2 ? p.end[1].entry_n * 8u : 0u);
    const uint32_t tab_bytes = !SMEM ? kReadsCtaBytes : NE == 2 ? ((meta1 + 16u * L + 15u) & ~15u) : ent1;
    if (threadIdx.x < 10) reinterpret_cast<uint64_t*>(smem)[threadIdx.x] = c_m10[threadIdx.x];
    if (SMEM) {
        for (uint32_t e = 0; e < NE; e++) {
            const EndDev& E = p.end[e];
            uint64_t* en = reinterpret_cast<uint64_t*>(smem + (e ? ent1 : ent0));
            for (uint32_t i = threadIdx.x; i < E.entry_n; i += blockDim.x) en[i] = E.entry64[i];
            uint32_t* m = reinterpret_cast<uint32_t*>(smem + (e ? meta1 : meta0));
            for (uint32_t i = threadIdx.x; i < 4 * L; i += blockDim.x) m[i] = E.meta[i];
        }
    }
    __syncthreads();
    uint32_t lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const uint32_t warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint32_t tplw = p.tpl_buf, codw = p.cod_buf, obw = p.rec_buf;
    const uint32_t W0 = sbase + tab_bytes + warp * reads_warp_bytes(NE, obw, tplw, codw);
    constexpr uint32_t plan_pair = NE * 32u;                  // bytes of plan per pair
#define MB0 (W0)
#define SC0 (W0 + 16u)                                        /* 32 bytes of the group's ID-line prefix; +32 group, +36 prefix length, +40 its chromosome's address */
#define ST0 (W0 + 64u)                                        /* per end 8 bytes: apos | hole << 8, goff */
#define PL0 (W0 + 80u)
#define TP0 (PL0 + 2u * kPlanChunk * plan_pair)
#define CD0 (TP0 + kTplSlots * NE * tplw)
#define OB0 (CD0 + NE * codw)

    // ---- this warp's run of pairs
    const uint32_t n_warps = gridDim.x * wpc;
    const uint32_t per = (p.batch_pairs + n_warps - 1u) / n_warps;
    const uint32_t r0 = min((blockIdx.x * wpc + warp) * per, p.batch_pairs);
    const uint32_t n_run = min(per, p.batch_pairs - r0);
    if (n_run == 0) return;

    const uint32_t he = lane >> 4, hl = lane & 15u;           // phase A and ID line: one half-warp per end
    const bool mine = NE == 2 || he == 0;
    if (lane < NE) {
        // where the run's first record of end `lane` goes: its offset in the file's batch buffer
        const uint64_t fo = p.block_base[(size_t)lane * p.n_scan_blocks + r0 / kScanBlock] + p.rec_local[r0 * NE + lane];
        sts32(ST0 + 8u * lane, ((uint32_t)fo & 15u) * 0x101u);    // apos = hole = fo % 16
        sts32(ST0 + 8u * lane + 4u, (uint32_t)fo & ~15u);
    }
    if (lane == 0) {
        mbar_init(MB0, 1);
        sts32(SC0 + 32u, 0xffffffffu);                        // no group cached
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // plans: chunk c (pairs 8c .. of the run) -> buffer c & 1; ONE mbarrier, whose phase c is chunk c: chunk c + 2 is
    // issued (after pair 8c + 7) only when chunk c + 1 has been waited for (during pair 8c + 7 at the latest)
    const uint32_t n_chunks = (n_run + kPlanChunk - 1u) / kPlanChunk;
    auto stage_chunk = [&](uint32_t c) {
        if (lane == 0) {
            const uint32_t np = min(kPlanChunk, n_run - c * kPlanChunk), bytes = np * plan_pair;
            fence_async_smem();                               // the buffer's last readers (generic proxy) came before this __syncwarp
            mbar_expect_tx(MB0, bytes);
            bulk_g2s(PL0 + (c & 1u) * kPlanChunk * plan_pair, reinterpret_cast<const uint8_t*>(p.plan) + (size_t)(r0 + c * kPlanChunk) * plan_pair, bytes, MB0);
        }
    };
    auto plan_addr = [&](uint32_t k) { return PL0 + (k & 15u) * plan_pair; };      // buffer (k >> 3) & 1, entry k & 7
    // the bytes around the templates of pair k -> slot k & 1: lane (he, hl) copies 16-byte chunk hl (+ 16, ...) of end he's window
    auto stage_tpl = [&](uint32_t k) {
        if (mine) {
            const uint4 pa = lds128(plan_addr(k) + he * 32u);
            const uint64_t sa = ((uint64_t)pa.y << 32) | pa.x;
            const uint64_t ws = (sa - 16u) & ~(uint64_t)15;
            const uint32_t need = (pa.x & 15u) + 16u + pa.z + 19u;          // bytes of the window the gather can touch
            const uint32_t dst = TP0 + ((k & 1u) * NE + he) * tplw;
            const uint32_t o = 16u * hl, top = min(need, tplw);
            if (o < top) cp_async16(dst + o, reinterpret_cast<const uint8_t*>(ws) + o);
            if (top > 256u)                                               // read lengths above 190
                for (uint32_t o2 = o + 256u; o2 < top; o2 += 256u) cp_async16(dst + o2, reinterpret_cast<const uint8_t*>(ws) + o2);
        }
        cp_async_commit();
    };
    stage_chunk(0);
    mbar_wait(MB0, 0);
    if (n_chunks > 1) stage_chunk(1);
    stage_tpl(0);
    // what a lane contributes to an ID line does not change: its power of ten (lanes 1-9), its byte of the tail
    // "-F/1\n" / "-F\n" (lanes 10-14; the strand letter is filled in per read), and -- per group -- two bytes of the prefix
    const uint64_t m10r = c_m10[min(hl, 9u)];
    const uint32_t tail0 = NE == 2 ? (hl == 10u ? '-' : hl == 12u ? '/' : hl == 13u ? '1' + he : '\n') : (hl == 10u ? '-' : '\n');
    uint32_t pfxa = 0, pfxb = 0;

#pragma unroll 1
    for (uint32_t k = 0; k < n_run; k++) {
        const uint64_t j = p.batch_lo + r0 + k;
        const uint32_t PLk = plan_addr(k);
        // ---- staging: the next pair's templates start their way; this pair's have landed
        if (k + 1u < n_run) {
            if (((k + 1u) & 7u) == 0u) mbar_wait(MB0, ((k + 1u) >> 3) & 1u);       // the chunk of pair k + 1 has landed
            stage_tpl(k + 1u);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            cp_async_wait_all();
        }
        __syncwarp();
        uint32_t sq, ln_e;                                    // this end's sequence line (shared address) and read length
        {
            // ---- this lane's end: plan fields
            uint4 pa = make_uint4(0, 0, 0, 0);
            uint32_t rec = 0, grp_e = 0, apos = 0;
            if (mine) {
                pa = lds128(PLk + he * 32u);
                const uint2 pb = lds64(PLk + he * 32u + 16u);
                rec = pb.x; grp_e = pb.y;
                apos = lds32(ST0 + 8u * he) & 0xffu;
            }
            const uint32_t S = pa.z, ln = pa.w & 0xffffu, flags = (pa.w >> 16) & 0xffu;
            uint32_t idlen = pa.w >> 24;
            if (flags & kPlanLongId) idlen = rec - 2u * ln - 4u;
            const uint32_t rs = OB0 + he * obw + apos;        // the record's first byte
            sq = rs + idlen;
            ln_e = ln;
            const uint32_t f_any = NE == 2 ? __shfl_sync(0xffffffffu, flags, 0) | __shfl_sync(0xffffffffu, flags, 16) : __shfl_sync(0xffffffffu, flags, 0);
            const uint32_t grp = __shfl_sync(0xffffffffu, grp_e, 0);
            if (grp != lds32(SC0 + 32u)) {                    // a new (haplotype, chromosome) group: its prefix into the cache
                __syncwarp();
                const GroupDev* Gp = p.groups + grp;
                const uint32_t n = Gp->prefix_len;
                pfxa = hl < n ? p.strpool[Gp->prefix_off + hl] : 0u;
                pfxb = hl + 16u < n ? p.strpool[Gp->prefix_off + hl + 16u] : 0u;
                if (lane == 0) {
                    const uint64_t gs = reinterpret_cast<uint64_t>(Gp->seq);
                    sts32(SC0 + 32u, grp); sts32(SC0 + 36u, n); sts64(SC0 + 40u, (uint32_t)gs, (uint32_t)(gs >> 32));
                }
                __syncwarp();
            }
            // ---- phase A: template base codes into the end's code line, 16 positions per lane from the staged window
            if (mine && !(flags & (kPlanManyEv | kPlanBarcode)))
            for (uint32_t tb = 16u * hl; tb < max(ln, S); tb += 256u) {                 // one round up to 256 positions
                const bool reverse = flags & kPlanReverse;
                const uint32_t d0 = (pa.x & 15u) + 16u;                               // seg's place in the window: 16 .. 31
                // forward: seg[tb .. tb+16); reverse: seg[S-1-tb-15 .. S-1-tb] read backwards and complemented
                // (positions past the read's end hold garbage nobody reads)
                const uint32_t bo = reverse ? d0 + S - 16u - tb : d0 + tb;
                const uint32_t wa = TP0 + ((k & 1u) * NE + he) * tplw + (bo & ~3u), sh = (bo & 3u) * 8u;
                const uint32_t g0 = lds32(wa), g1 = lds32(wa + 4u), g2 = lds32(wa + 8u), g3 = lds32(wa + 12u), g4 = lds32(wa + 16u);
                uint32_t x0 = __funnelshift_r(g0, g1, sh), x1 = __funnelshift_r(g1, g2, sh), x2 = __funnelshift_r(g2, g3, sh),
                         x3 = __funnelshift_r(g3, g4, sh);
                if (reverse) {
                    const uint32_t t0 = __byte_perm(x3, 0u, 0x0123u), t1 = __byte_perm(x2, 0u, 0x0123u);
                    x3 = __byte_perm(x0, 0u, 0x0123u); x2 = __byte_perm(x1, 0u, 0x0123u);
                    x0 = t0; x1 = t1;
                }
                uint32_t b0, b1, b2, b3;
                uint4 c = make_uint4(codes4_acc(x0, reverse, b0), codes4_acc(x1, reverse, b1), codes4_acc(x2, reverse, b2),
                                     codes4_acc(x3, reverse, b3));
                if (b0 | b1 | b2 | b3) { c.x = codes4_fix(c.x, b0); c.y = codes4_fix(c.y, b1); c.z = codes4_fix(c.z, b2); c.w = codes4_fix(c.w, b3); }
                sts128(CD0 + he * codw + tb, c);
            }
            // ends with indels or a barcode take the whole warp, one end after the other
            if (f_any & (kPlanIndels | kPlanBarcode)) {
                __syncwarp();
#pragma unroll 1
                for (uint32_t e = 0; e < NE; e++) {
                    const uint4 px = lds128(PLk + e * 32u);
                    const uint32_t fe = (px.w >> 16) & 0xffu;
                    if (!(fe & (kPlanIndels | kPlanBarcode))) continue;
                    if (fe & (kPlanManyEv | kPlanBarcode)) {
                        const uint8_t* seg = reinterpret_cast<const uint8_t*>(((uint64_t)px.y << 32) | px.x);
                        const GroupDev* Gp = p.groups + grp;
                        const uint8_t* bc = p.strpool + Gp->bc_off;
                        if (fe & kPlanIndels) gather_indels(p, e, j, CD0 + e * codw, seg, bc, px.z, Gp->bc_len, px.w & 0xffffu, fe & kPlanReverse);
                        else gather_barcode(CD0 + e * codw, seg, bc, px.z, Gp->bc_len, px.w & 0xffffu, fe & kPlanReverse);
                    } else {
                        // one or two insertions / deletions: phase A has laid down the codes of all S template positions;
                        // the edits shift the line's tail, from the back as fill_read_qual applies them (src/hts_illumina.h:213-225)
                        const uint2 evs = lds64(PLk + e * 32u + 24u);
                        apply_events(p, e, j, CD0 + e * codw, px.z, evs.x, evs.y);
                    }
                }
            }
            // ---- ID line "@<genome>-<chrom>-<start>-<F|R>[/<1|2>]\n" (fill_fq_lines, src/hts_illumina.cpp:296-312) and the
            //      separators, in place: prefix bytes hl and hl + 16; lanes 0-9 one decimal digit each, lanes 10-14 the
            //      tail "-F/1\n" ("-F\n" single-end), lane 15 the four separator bytes of the record
            if (mine) {
                const uint32_t pfx_len = lds32(SC0 + 36u);
                uint2 sv = lds64(SC0 + 40u);                                           // start coordinate = template - chromosome
                sv.y = pa.y - sv.y - (pa.x < sv.x ? 1u : 0u);
                sv.x = pa.x - sv.x;
                if (pfx_len <= 32u && sv.y == 0u) {
                    if (hl < pfx_len) sts8(rs + hl, pfxa);
                    if (hl + 16u < pfx_len) sts8(rs + hl + 16u, pfxb);
                    constexpr uint32_t nt = NE == 2 ? 5u : 3u;
                    const uint32_t nd = idlen - pfx_len - nt;
                    const uint32_t q = hl == 0 ? sv.x : (uint32_t)__umul64hi((uint64_t)sv.x, m10r);   // start / 10^hl
                    const uint32_t dig = q - 10u * (__umulhi(q, 0xCCCCCCCDu) >> 3);
                    const bool is_dig = hl < 10u;
                    const uint32_t ch = is_dig ? '0' + dig : hl == 11u ? ((flags & kPlanReverse) ? 'R' : 'F') : tail0;
                    const uint32_t at = rs + pfx_len + (is_dig ? nd - 1u - hl : nd + hl - 10u);
                    if (is_dig ? hl < nd : hl < 10u + nt) sts8(at, ch);
                } else if (hl == 0) {
                    slow_idline(p, rs, p.groups + grp, ((uint64_t)sv.y << 32) | sv.x, flags & kPlanReverse, he);
                }
                if (hl == 15u) {
                    sts8(sq + ln, '\n'); sts8(sq + ln + 1, '+'); sts8(sq + ln + 2, '\n');
                    sts8(sq + 2 * ln + 3, '\n');
                }
            }
        }
        const uint32_t sq0 = __shfl_sync(0xffffffffu, sq, 0), sq1 = NE == 2 ? __shfl_sync(0xffffffffu, sq, 16) : 0u;
        const uint32_t len0 = __shfl_sync(0xffffffffu, ln_e, 0), len1 = NE == 2 ? __shfl_sync(0xffffffffu, ln_e, 16) : 0u;
        __syncwarp();

        // ---- phase B
        phase_b<SMEM, NE>(p, j, lane, sq0, sq1, len0, len1, CD0, codw, sbase + meta0, sbase + meta1, sbase + ent0, sbase + ent1);
        __syncwarp();
        // ---- flush: the 16-byte chunks a record completes, one 128-bit load and store per lane and chunk; one half-warp per end
        if (mine) {
            const uint32_t sh = lds32(ST0 + 8u * he), goff = lds32(ST0 + 8u * he + 4u);
            const uint32_t tot = (sh & 0xffu) + lds32(PLk + he * 32u + 16u), nfl = tot & ~15u;   // bytes of the buffer in use, whole chunks of them
            const uint32_t buf = OB0 + he * obw;
            uint8_t* dst = (he ? p.out[1] : p.out[0]) + goff;
            for (uint32_t o = 16u * hl; o < nfl; o += 256u) {
                if (o == 0 && sh >= 0x100u) {                     // the run's first chunk: only our bytes of it
                    for (uint32_t b = sh >> 8; b < 16u; b++) dst[b] = (uint8_t)lds8(buf + b);
                } else {
                    *reinterpret_cast<uint4*>(dst + o) = lds128(buf + o);
                }
            }
            uint4 carry = make_uint4(0, 0, 0, 0);
            if (hl == 0 && nfl) carry = lds128(buf + nfl);
            __syncwarp(NE == 2 ? 0xffffffffu : 0x0000ffffu);
            // the unfinished chunk moves to the front of the buffer; the next record continues right behind it
            if (hl == 0) {
                if (nfl) sts128(buf, carry);
                sts32(ST0 + 8u * he, (tot & 15u) | (nfl ? 0u : sh & 0xff00u));
                sts32(ST0 + 8u * he + 4u, goff + nfl);
            }
        }
        __syncwarp();
        // chunk c's buffer is free once its last pair is done: chunk c + 2 takes it
        if ((k & 7u) == 7u && (k >> 3) + 2u < n_chunks) stage_chunk((k >> 3) + 2u);
    }
    // ---- the run's last, unfinished chunk: byte by byte
    if (mine) {
        const uint32_t sh = lds32(ST0 + 8u * he), apos = sh & 0xffu, hole = sh >> 8;
        if (hl >= hole && hl < apos) (he ? p.out[1] : p.out[0])[lds32(ST0 + 8u * he + 4u) + hl] = (uint8_t)lds8(OB0 + he * obw + hl);
    }
#undef MB0
#undef SC0
#undef ST0
#undef PL0
#undef TP0
#undef CD0
#undef OB0
}

static size_t reads_table_bytes(const GenParams& p) {
    // same layout as in k_reads: entry64[0], meta[0], (entry64[1], meta[1]), each end 16-byte aligned
    size_t b = kReadsCtaBytes + (size_t)p.end[0].entry_n * 8 + 16 * (size_t)p.L;
    b = (b + 15) & ~(size_t)15;
    if (p.n_ends == 2) {
        b += (size_t)p.end[1].entry_n * 8 + 16 * (size_t)p.L;
        b = (b + 15) & ~(size_t)15;
    }
    return b;
}

// Threads per CTA of k_reads for this run: the most warps (at most kReadsThreads / 32) whose record and staging
// buffers fit the shared memory of one SM -- together with the profile tables when those fit too (at least 8 warps
// then), else with the tables read from global memory.  0: not even one warp's buffers fit (a custom profile with
// a read length of several thousand).
static int reads_threads(const GenParams& p, bool& use_smem) {
    const size_t per_warp = reads_warp_bytes(p.n_ends, p.rec_buf, p.tpl_buf, p.cod_buf), tab = reads_table_bytes(p);
    const size_t with_tab = 200 * 1024, without = 226 * 1024;
    const int max_w = kReadsThreads / 32;
    int w = (int)std::min<size_t>(max_w, tab < with_tab ? (with_tab - tab) / per_warp : 0);
    if (w >= 8) { use_smem = true; return 32 * w; }
    use_smem = false;
    w = (int)std::min<size_t>(max_w, without / per_warp);
    return 32 * w;
}
bool reads_fits(const GenParams& p) {
    bool u;
    return reads_threads(p, u) > 0;
}

template <bool SMEM, uint32_t NE>
static cudaError_t launch_reads_as(const GenParams& p, int n_sm, int threads, size_t smem_bytes, cudaStream_t s) {
    cudaError_t err = cudaFuncSetAttribute(k_reads<SMEM, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    // persistent: one CTA per SM, every warp takes one contiguous run of the batch's pairs
    const int wpc = threads / 32;
    uint32_t blocks = (p.batch_pairs + wpc - 1) / wpc;
    const uint32_t cap = (uint32_t)(n_sm > 0 ? n_sm : 148);
    if (blocks > cap) blocks = cap;
    k_reads<SMEM, NE><<<blocks, threads, smem_bytes, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_reads(const GenParams& p, int n_sm, cudaStream_t s) {
    if (p.batch_pairs == 0) return cudaSuccess;
    bool use_smem = false;
    const int threads = reads_threads(p, use_smem), wpc = threads / 32;
    if (threads <= 0) return cudaErrorInvalidConfiguration;
    const size_t rec_bytes = (size_t)wpc * reads_warp_bytes(p.n_ends, p.rec_buf, p.tpl_buf, p.cod_buf);
    const size_t smem_bytes = rec_bytes + (use_smem ? reads_table_bytes(p) : kReadsCtaBytes);
    if (p.n_ends == 2) return use_smem ? launch_reads_as<true, 2>(p, n_sm, threads, smem_bytes, s) : launch_reads_as<false, 2>(p, n_sm, threads, smem_bytes, s);
    return use_smem ? launch_reads_as<true, 1>(p, n_sm, threads, smem_bytes, s) : launch_reads_as<false, 1>(p, n_sm, threads, smem_bytes, s);
}

// ------------------------------------------------------------------ scan ---

// block (file e, scan block b): exclusive prefix of up to kScanBlock record lengths
__global__ void __launch_bounds__(1024)
k_scan_local(const uint32_t* __restrict__ rec_len, uint32_t n_records, uint32_t n_ends, uint32_t n_blocks,
             uint32_t* __restrict__ rec_local, uint64_t* __restrict__ block_tot) {
    __shared__ uint32_t wsum[32];
    const uint32_t e = blockIdx.x / n_blocks, b = blockIdx.x % n_blocks;
    const uint32_t i = b * kScanBlock + threadIdx.x;
    const uint32_t r = i * n_ends + e;
    uint32_t v = (r < n_records) ? rec_len[r] : 0u;
    uint32_t inc = warp_incl_scan(v);
    if ((threadIdx.x & 31u) == 31u) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = wsum[threadIdx.x];
        uint32_t si = warp_incl_scan(s);
        wsum[threadIdx.x] = si - s;
        if (threadIdx.x == 31) block_tot[blockIdx.x] = si;
    }
    __syncthreads();
    if (r < n_records) rec_local[r] = inc - v + wsum[threadIdx.x >> 5];
}

// one block per file: exclusive prefix of the block totals
__global__ void __launch_bounds__(1024)
k_scan_tops(const uint64_t* __restrict__ block_tot, uint32_t n_blocks, uint64_t* __restrict__ block_base,
            uint64_t* __restrict__ totals) {
    __shared__ uint64_t part[1024];
    const uint32_t e = blockIdx.x;
    const uint64_t* in = block_tot + (size_t)e * n_blocks;
    uint64_t* out = block_base + (size_t)e * n_blocks;
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n_blocks);
    uint64_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += in[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t acc = 0;
        for (uint32_t t = 0; t < 1024; t++) { uint64_t v = part[t]; part[t] = acc; acc += v; }
        totals[e] = acc;
    }
    __syncthreads();
    uint64_t acc = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { out[i] = acc; acc += in[i]; }
}

cudaError_t launch_scan(const uint32_t* rec_len, uint32_t n_records, uint32_t n_ends,
                        uint32_t* rec_local, uint64_t* block_tot,
                        uint64_t* block_base, uint64_t* totals_out, cudaStream_t s) {
    uint32_t per_file = n_records / n_ends;
    uint32_t n_blocks = (per_file + kScanBlock - 1) / kScanBlock;
    if (n_blocks == 0) return cudaSuccess;
    k_scan_local<<<n_blocks * n_ends, 1024, 0, s>>>(rec_len, n_records, n_ends, n_blocks, rec_local, block_tot);
    k_scan_tops<<<n_ends, 1024, 0, s>>>(block_tot, n_blocks, block_base, totals_out);
    return cudaGetLastError();
}

}  // namespace jlp
