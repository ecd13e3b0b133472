// CUDA kernels of the Illumina read generator (sm_100a).
//
//   k_materialize  haplotype chromosome from sorted mutation records
//                  (HapChrom::get_chrom_full, /root/reference/src/hap_classes.cpp:80-116)
//   k_gen          fragment placement + indels + ART quality/error model, one
//                  thread per read end (chrom_indels_frag / sample_indels /
//                  append_pools / fill_read_qual, src/hts_illumina.cpp:116-482,
//                  src/hts_illumina.h:202-259)
//   k_scan_*       exclusive scan of FASTQ record lengths -> output offsets
//   k_fmt          FASTQ record assembly (fill_fq_lines, src/hts_illumina.cpp:285-326)
//
// None of this is a dense contraction, so no tensor-core path exists here; the
// kernels are integer/byte work bounded by HBM traffic and integer issue rate
// (DESIGN.md section 6).
#include "jlp_kernels.cuh"
#include "jlp_draws.h"

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
    uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= chrom_size) return;
    int64_t lo = -1, hi = (int64_t)n_muts;
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        if (new_pos[mid] <= p0) lo = mid; else hi = mid;
    }
    int64_t i = lo;
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

// ------------------------------------------------------- quality / errors ---

// nt_map (src/hts.h:36-44) as codes: T,C,A,G -> 0..3, anything else 4.
__device__ __forceinline__ uint32_t nt_code(uint32_t c) {
    return c == 'T' ? 0u : c == 'C' ? 1u : c == 'A' ? 2u : c == 'G' ? 3u : 4u;
}
// "TCAGN"[code]
__device__ __forceinline__ uint32_t code_ascii(uint32_t code) {
    return __byte_perm(0x47414354u, 0x0000004Eu, code | 0x5550u);   // bytes 1..3 select a zero byte
}

struct Tables {
    const uint32_t* meta[2];
    const uint32_t* entry[2];
    const uint16_t* mis16[2];
};

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

// Exact (full 64-bit draws) evaluation of one base; used when the 16 high bits
// of any draw do not decide it.  fill_read_qual, src/hts_illumina.h:230-256.
__device__ __noinline__ void base_slow(const GenParams& p, uint32_t e, uint64_t j, uint32_t pos,
                                       uint32_t code, uint32_t Hdie, uint32_t Hcoin, uint32_t Hmis,
                                       uint32_t& q_out, bool& mism_out) {
    const EndDev& E = p.end[e];
    if (code > 3) {
        q_out = nqual_x87(full_draw(Hdie, p.seed, j, e, PU_DIE, pos)) - 33u;
        mism_out = false;
        return;
    }
    uint32_t m = E.meta[code * p.L + pos];
    uint32_t n = m & 0xffu, off = m >> 8;
    uint64_t i = mul_floor_x87(full_draw(Hdie, p.seed, j, e, PU_DIE, pos), n);
    if (i >= n) i = n - 1;
    uint32_t ent = E.entry[off + i];
    bool self = full_draw(Hcoin, p.seed, j, e, PU_COIN, pos) < E.coin[off + i];
    uint32_t q = self ? ((ent >> 16) & 0xffu) : (ent >> 24);
    q_out = q;
    mism_out = full_draw(Hmis, p.seed, j, e, PU_MIS, pos) < E.mis[q];
}

__device__ __noinline__ uint32_t sub_slow(const GenParams& p, uint32_t e, uint64_t j, uint32_t pos, uint32_t Hsub) {
    uint64_t si = mul_floor_x87(full_draw(Hsub, p.seed, j, e, PU_SUB, pos), 3);
    return si > 2 ? 2u : (uint32_t)si;
}

// One base: quality draw, mismatch draw, substitution.  Returns ascii | qualchar << 8.
__device__ __forceinline__ uint32_t do_base(const GenParams& p, const Tables& T, uint32_t e, uint64_t j,
                                            uint32_t pos, uint32_t code, uint32_t wa, uint32_t wb) {
    uint32_t Hdie = wa & 0xffffu, Hcoin = wa >> 16, Hmis = wb & 0xffffu, Hsub = wb >> 16;
    uint32_t q;
    bool mism;
    if (code > 3) {
        // non-TCAG: 'N' with a quality below 10 (src/hts_illumina.h:237-242)
        uint32_t prod = Hdie * 10u;
        q = prod >> 16;
        mism = false;
        if ((prod & 0xffffu) + 10u > 0xffffu) base_slow(p, e, j, pos, code, Hdie, Hcoin, Hmis, q, mism);
        return 0x4Eu | (((q + 33u) & 0xffu) << 8);
    }
    uint32_t m = T.meta[e][code * p.L + pos];
    uint32_t n = m & 0xffu, off = m >> 8;
    uint32_t prod = Hdie * n;
    uint32_t ent = T.entry[e][off + (prod >> 16)];
    uint32_t thr = ent & 0xffffu;
    q = Hcoin < thr ? ((ent >> 16) & 0xffu) : (ent >> 24);
    uint32_t mt = T.mis16[e][q];
    mism = Hmis < mt;
    bool amb = ((prod & 0xffffu) + n > 0xffffu) | (Hcoin == thr) | (Hmis == mt);
    if (amb) base_slow(p, e, j, pos, code, Hdie, Hcoin, Hmis, q, mism);
    if (mism) {
        // mm_nucleos[nt][(uint64)(u * 3)] (src/hts.h:46): the si-th code other than `code`
        uint32_t p3 = Hsub * 3u;
        uint32_t si = p3 >> 16;
        if ((p3 & 0xffffu) + 3u > 0xffffu) si = sub_slow(p, e, j, pos, Hsub);
        code = si + (si >= code ? 1u : 0u);
    }
    return code_ascii(code) | (((q + 33u) & 0xffu) << 8);
}

template <bool SMEM_TABLES>
__global__ void __launch_bounds__(256)
k_gen(const __grid_constant__ GenParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    Tables T;
    if (SMEM_TABLES) {
        uint32_t o = 0;
        for (uint32_t e = 0; e < p.n_ends; e++) {
            const EndDev& E = p.end[e];
            uint32_t* m = smem + o;            o += 4 * p.L;
            uint32_t* en = smem + o;           o += E.entry_n;
            uint32_t* mi = smem + o;           o += 128;
            for (uint32_t i = threadIdx.x; i < 4 * p.L; i += blockDim.x) m[i] = E.meta[i];
            for (uint32_t i = threadIdx.x; i < E.entry_n; i += blockDim.x) en[i] = E.entry[i];
            for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) mi[i] = reinterpret_cast<const uint32_t*>(E.mis16)[i];
            T.meta[e] = m; T.entry[e] = en; T.mis16[e] = reinterpret_cast<const uint16_t*>(mi);
        }
        __syncthreads();
    } else {
        for (uint32_t e = 0; e < p.n_ends; e++) {
            T.meta[e] = p.end[e].meta; T.entry[e] = p.end[e].entry; T.mis16[e] = p.end[e].mis16;
        }
    }

    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.batch_pairs * p.n_ends) return;
    const uint32_t e = (p.n_ends == 2) ? (r & 1u) : 0u;
    const uint64_t j = p.batch_lo + (p.n_ends == 2 ? (r >> 1) : r);
    const uint32_t L = p.L;

    // ---- duplicate-chain leader (ReadWriterOneThread::create_reads, src/hts.h:254-280)
    uint64_t k = j;
    if (!p.dup_never) {
        while (k > p.job_lo && ((k - p.job_lo) % p.pool_pairs) != 0) {
            uint64_t xd = hi64(draw_block(p.seed, k - 1, 1, PL_PAIR, 0));
            if (!(xd < p.c_dup)) break;
            k--;
        }
    }
    // ---- (haplotype, chromosome) group of the leader
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

    // ---- fragment length and start (chrom_indels_frag, src/hts_illumina.cpp:203-217)
    uint64_t frag_len, frag_start;
    {
        U4 w = draw_block(p.seed, k, 0, PL_PAIR, 0);
        uint64_t xf = lo64(w);
        uint32_t lo = 0, hi = p.frag_n;                // first i with cdf[i] > xf
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.frag_cdf[mid] <= xf) lo = mid + 1; else hi = mid;
        }
        frag_len = p.frag_min + lo;
        if (frag_len >= chrom_len) { frag_len = chrom_len; frag_start = 0; }
        else {
            uint64_t xs = hi64(w);
            double u = (xs == ~0ull) ? 1.0 : __ull2double_rn(xs + 1) * 5.421010862427522170037e-20;
            frag_start = __double2ull_rz(__dmul_rn(u, __ull2double_rn(chrom_len - frag_len + 1)));
        }
    }
    // ---- strand (append_pools, src/hts_illumina.cpp:352); the second end is the opposite
    bool reverse = lo64(draw_block(p.seed, j, 1, PL_PAIR, 0)) < p.c_rev;
    if (e) reverse = !reverse;

    // ---- indels, pass 1: counts only (sample_indels, src/hts_illumina.cpp:116-150)
    uint32_t frag_pos = 0, length_now = 0, n_ev = 0;
    const uint32_t frag_cap = frag_len > 0xffffffffull ? 0xffffffffu : (uint32_t)frag_len;
    while (length_now < L && frag_pos < frag_cap) {
        U4 w = draw_block(p.seed, j, frag_pos >> 3, PL_INDEL, e);
        const uint32_t hA = p.end[e].hA;
        bool plain = ((w.w0 & 0xffffu) > hA) & ((w.w0 >> 16) > hA) & ((w.w1 & 0xffffu) > hA) & ((w.w1 >> 16) > hA) &
                     ((w.w2 & 0xffffu) > hA) & ((w.w2 >> 16) > hA) & ((w.w3 & 0xffffu) > hA) & ((w.w3 >> 16) > hA);
        if (plain && (frag_pos & 7u) == 0 && length_now + 8 <= L && frag_pos + 8 <= frag_cap) {
            length_now += 8; frag_pos += 8;
            continue;
        }
        uint32_t f = frag_pos & 7u;
        for (; f < 8 && length_now < L && frag_pos < frag_cap; f++, frag_pos++) {
            int cls = indel_class(p, e, j, frag_pos, field16(w, f));
            if (cls == 0) length_now++;
            else if (cls == 1) n_ev++;
            else if (length_now == L - 1) length_now++;
            else { length_now += 2; n_ev++; }
        }
    }
    // adjust_chrom_spaces (src/hts_illumina.cpp:153-184): the template span equals the
    // number of positions visited; the read ends up `length_now` long.
    uint32_t S = frag_pos, len = length_now;
    const uint32_t b = G.bc_len;
    if (S <= b) {                                       // barcode covers the whole template: undefined in the reference
        atomicOr(p.status, 1u);
        S = b; len = 0;
    }
    const uint32_t space = S - b;
    const uint64_t start = (p.matepair != 0) == reverse ? frag_start : frag_start + frag_len - space;

    uint8_t* row = p.seq + (size_t)r * p.row_stride;
    uint8_t* qrow = p.qual + (size_t)r * p.row_stride;
    const uint8_t* seg = G.seq + start;
    const uint8_t* bc = p.strpool + G.bc_off;

    // ---- template -> oriented, indel-applied read as base codes in `row`
    //      (fill_read / rev_comp / barcode, src/hts_illumina.cpp:369-391; indel application
    //       from fill_read_qual, src/hts_illumina.h:213-225)
    if (len == 0) {
        // nothing to build
    } else if (n_ev == 0) {
        for (uint32_t t = 0; t < S; t++) {
            uint32_t code;
            if (t < b) code = nt_code(bc[t]);
            else {
                uint32_t i = t - b;
                code = nt_code(reverse ? seg[space - 1 - i] : seg[i]);
                if (reverse && code < 4) code ^= 2u;    // cmp_map: T<->A, C<->G
            }
            row[t] = (uint8_t)code;
        }
    } else {
        uint32_t out = 0, ln = 0;
        for (uint32_t t = 0; t < S; t++) {
            U4 w = draw_block(p.seed, j, t >> 3, PL_INDEL, e);
            int cls = indel_class(p, e, j, t, field16(w, t & 7u));
            if (cls == 2 && ln == L - 1) cls = 0;
            if (cls == 1) continue;
            uint32_t code;
            if (t < b) code = nt_code(bc[t]);
            else {
                uint32_t i = t - b;
                code = nt_code(reverse ? seg[space - 1 - i] : seg[i]);
                if (reverse && code < 4) code ^= 2u;
            }
            row[out++] = (uint8_t)code;
            ln++;
            if (cls == 2) {
                // bases[(uint64)(u * 4)], src/hts_illumina.h:216; index 4 reads the
                // string terminator, which nt_map then turns into 'N'
                row[out++] = (uint8_t)ins_base_index(slow64(p.seed, j, e, PU_INS, t));
                ln++;
            }
        }
    }

    // ---- qualities, mismatches (fill_read_qual, src/hts_illumina.h:230-256), 16 bases per step
    for (uint32_t blk = 0; blk * 16 < len; blk++) {
        uint4 cw = *reinterpret_cast<const uint4*>(row + blk * 16);
        uint32_t c[4] = {cw.x, cw.y, cw.z, cw.w};
        uint32_t sw[4] = {0, 0, 0, 0}, qw[4] = {0, 0, 0, 0};
#pragma unroll
        for (uint32_t h = 0; h < 8; h++) {
            uint32_t pos = blk * 16 + 2 * h;
            if (pos < len) {
                U4 w = draw_block(p.seed, j, pos >> 1, PL_QUAL, e);
                uint32_t code0 = (c[h >> 1] >> (16 * (h & 1))) & 0xffu;
                uint32_t r0 = do_base(p, T, e, j, pos, code0, w.w0, w.w1);
                sw[h >> 1] |= (r0 & 0xffu) << (16 * (h & 1));
                qw[h >> 1] |= (r0 >> 8) << (16 * (h & 1));
                if (pos + 1 < len) {
                    uint32_t code1 = (c[h >> 1] >> (16 * (h & 1) + 8)) & 0xffu;
                    uint32_t r1 = do_base(p, T, e, j, pos + 1, code1, w.w2, w.w3);
                    sw[h >> 1] |= (r1 & 0xffu) << (16 * (h & 1) + 8);
                    qw[h >> 1] |= (r1 >> 8) << (16 * (h & 1) + 8);
                }
            }
        }
        *reinterpret_cast<uint4*>(row + blk * 16) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
        *reinterpret_cast<uint4*>(qrow + blk * 16) = make_uint4(qw[0], qw[1], qw[2], qw[3]);
    }

    // ---- record metadata + FASTQ record length (fill_fq_lines, src/hts_illumina.cpp:285-326)
    RecMeta rm;
    rm.start = start; rm.group = g; rm.len = (uint16_t)len; rm.reverse = reverse ? 1 : 0; rm.pad = 0;
    p.rec[r] = rm;
    uint32_t nd = 1;
    for (uint64_t v = start; v >= 10; v /= 10) nd++;
    // '@' name '-' chrom '-' | digits | '-' F/R | ['/' 1/2] | '\n' | read '\n' '+' '\n' qual '\n'
    p.rec_len[r] = G.prefix_len + nd + 3u + (p.n_ends == 2 ? 2u : 0u) + 2u * len + 4u;
}

size_t gen_smem_bytes(const GenParams& p) {
    size_t words = 0;
    for (uint32_t e = 0; e < p.n_ends; e++) words += 4 * (size_t)p.L + p.end[e].entry_n + 128;
    return words * 4;
}

cudaError_t launch_gen(const GenParams& p, size_t smem_bytes, cudaStream_t s) {
    uint32_t n = p.batch_pairs * p.n_ends;
    if (n == 0) return cudaSuccess;
    uint32_t blocks = (n + 255) / 256;
    if (smem_bytes) {
        static size_t configured = 0;
        if (smem_bytes > configured) {
            cudaError_t err = cudaFuncSetAttribute(k_gen<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (err != cudaSuccess) return err;
            configured = smem_bytes;
        }
        k_gen<true><<<blocks, 256, smem_bytes, s>>>(p);
    } else {
        k_gen<false><<<blocks, 256, 0, s>>>(p);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ scan ---

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}

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
                        uint32_t scan_block, uint32_t* rec_local, uint64_t* block_tot,
                        uint64_t* block_base, uint64_t* totals_out, cudaStream_t s) {
    (void)scan_block;
    uint32_t per_file = n_records / n_ends;
    uint32_t n_blocks = (per_file + kScanBlock - 1) / kScanBlock;
    if (n_blocks == 0) return cudaSuccess;
    k_scan_local<<<n_blocks * n_ends, 1024, 0, s>>>(rec_len, n_records, n_ends, n_blocks, rec_local, block_tot);
    k_scan_tops<<<n_ends, 1024, 0, s>>>(block_tot, n_blocks, block_base, totals_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- format ---

// One warp per record; lanes copy consecutive bytes, so every store instruction
// covers one contiguous 32-byte run of the output file.
__global__ void __launch_bounds__(256)
k_fmt(const __grid_constant__ FmtParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= p.n_records) return;
    const uint32_t e = (p.n_ends == 2) ? (r & 1u) : 0u;
    const uint32_t i = (p.n_ends == 2) ? (r >> 1) : r;
    const RecMeta rm = p.rec[r];
    const GroupDev G = p.groups[rm.group];
    uint8_t* dst = p.out[e] + p.block_base[(size_t)e * p.n_scan_blocks + i / kScanBlock] + p.rec_local[r];

    // ID line: "@<genome>-<chrom>-" <start> "-" F|R ["/" 1|2] "\n"
    const uint8_t* pre = p.strpool + G.prefix_off;
    for (uint32_t t = lane; t < G.prefix_len; t += 32) dst[t] = pre[t];
    dst += G.prefix_len;
    uint32_t nd = 1;
    for (uint64_t v = rm.start; v >= 10; v /= 10) nd++;
    if (lane < nd) {
        uint64_t v = rm.start;
        for (uint32_t t = 0; t < nd - 1 - lane; t++) v /= 10;
        dst[lane] = (uint8_t)('0' + (v % 10));
    }
    dst += nd;
    if (lane == 0) {
        uint32_t o = 0;
        dst[o++] = '-';
        dst[o++] = rm.reverse ? 'R' : 'F';
        if (p.n_ends == 2) { dst[o++] = '/'; dst[o++] = (uint8_t)('1' + e); }
        dst[o++] = '\n';
    }
    dst += 3u + (p.n_ends == 2 ? 2u : 0u);
    const uint32_t len = rm.len;
    const uint8_t* srow = p.seq + (size_t)r * p.row_stride;
    const uint8_t* qrow = p.qual + (size_t)r * p.row_stride;
    for (uint32_t t = lane; t < len; t += 32) dst[t] = srow[t];
    if (lane == 0) { dst[len] = '\n'; dst[len + 1] = '+'; dst[len + 2] = '\n'; }
    dst += len + 3;
    for (uint32_t t = lane; t < len; t += 32) dst[t] = qrow[t];
    if (lane == 0) dst[len] = '\n';
}

cudaError_t launch_fmt(const FmtParams& p, cudaStream_t s) {
    if (p.n_records == 0) return cudaSuccess;
    uint32_t blocks = (p.n_records + 7) / 8;
    k_fmt<<<blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace jlp
