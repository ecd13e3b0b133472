// PacBio read generator, device side.  First version: one thread per read, the
// reference's loops restated literally (PacBioQualityError::sample's walk, src/hts_pacbio.h:296-318;
// PacBioOneGenome::append_pool, src/hts_pacbio.cpp), every draw addressed by (seed, read index, position), so a read
// does not depend on batch size, thread mapping or GPU count.
//   k_pb_plan   walks the template once to count insertions and deletions -> read_chrom_space, read_start, strand,
//               record length
//   k_pb_reads  walks it again and writes the FASTQ record
#include "jlp_pacbio.h"

#include "jlp_draws.h"

namespace jlp {

namespace {

constexpr uint32_t PL_PB = 6;

__device__ __forceinline__ uint64_t pb_draw(uint64_t seed, uint64_t j, uint32_t sub, uint32_t block, uint32_t half) {
    const U4 w = draw_block(seed, j, block, PL_PB, sub);
    return half ? hi64(w) : lo64(w);
}

// event of one template position: 0 none, 1 insertion, 2 deletion, 3 substitution
__device__ __forceinline__ uint32_t pb_event(const PbRead& R, uint32_t side, uint64_t x) {
    const uint32_t f = R.flags >> (3 * side);
    if (!((f & 4u) || x < R.thr[side][2])) return 0;
    if ((f & 1u) || x < R.thr[side][0]) return 1;
    if ((f & 2u) || x < R.thr[side][1]) return 2;
    return 3;
}

// One template position of PacBioQualityError::sample's walk (src/hts_pacbio.h:296-318): the event drawn there and
// whether it was recorded (an insertion needs room for two bases, a deletion spare template).  Returns the recorded
// event (0 none, 1 insertion, 2 deletion, 3 substitution).
struct Walk {
    uint32_t len = 0;          // the walk's current_length
    uint32_t side = 0;
    uint64_t extra_space;
    __device__ __forceinline__ uint32_t step(const PbRead& R, uint64_t x) {
        if (len == R.split_pos) side = 1;
        const uint32_t ev = pb_event(R, side, x);
        if (ev == 0) { len++; return 0; }
        if (ev == 1) {
            uint32_t rec = 0;
            if (len < R.read_length - 1) { rec = 1; len++; extra_space++; if (len == R.split_pos) side = 1; }
            len++;
            return rec;
        }
        if (ev == 2) {
            if (extra_space > 0) { extra_space--; return 2; }
            return 0;
        }
        len++;
        return 3;
    }
};
// append_pool (src/hts_pacbio.cpp) replays the recorded events position by position and counts what it emits: an
// event the walk drew but could not record is a plain base there, so its count can run ahead of the walk's.
__device__ __forceinline__ uint32_t pb_emitted(uint32_t rec) { return rec == 1 ? 2u : rec == 2 ? 0u : 1u; }

// the walk's draws, two template positions per Philox block
struct WalkDraws {
    uint64_t seed, j;
    U4 w;
    uint32_t have = 0xffffffffu;
    __device__ __forceinline__ uint64_t at(uint32_t pos) {
        if ((pos >> 1) != have) { have = pos >> 1; w = draw_block(seed, j, have, PL_PB, 1); }
        return (pos & 1u) ? hi64(w) : lo64(w);
    }
};

__global__ void __launch_bounds__(128)
k_pb_plan(PbRead* __restrict__ reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* __restrict__ groups,
          uint64_t c_rev, uint32_t* __restrict__ rec_len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PbRead R = reads[i];
    const uint64_t j = first_read + i;
    const GroupDev G = groups[R.group];
    const uint64_t chrom_len = G.len;
    const uint32_t read_length = R.read_length;
    // ---- PacBioQualityError::sample: the counts, and how many bases append_pool will emit
    uint32_t chrom_pos = 0, n_ins = 0, n_del = 0, seq_len = 0;
    Walk W;
    W.extra_space = chrom_len - read_length;
    WalkDraws D{seed, j};
    while (W.len < read_length) {
        const uint32_t rec = W.step(R, D.at(chrom_pos));
        n_ins += rec == 1;
        n_del += rec == 2;
        if (seq_len < read_length) seq_len += pb_emitted(rec);
        chrom_pos++;
    }
    // ---- read_chrom_space, read_start (one_read, src/hts_pacbio.cpp), strand (append_pool)
    const uint32_t space = read_length + n_del - n_ins;
    const U4 w = draw_block(seed, j, 1, PL_PB, 0);
    uint64_t start = 0;
    if (space < chrom_len) {
        // double u = runif_01(eng); read_start = u * (chrom_len - read_chrom_space + 1)
        const uint64_t xs = lo64(w);
        const double u = (xs == ~0ull) ? 1.0 : __ull2double_rn(xs + 1) * 5.421010862427522170037e-20;
        start = __double2ull_rz(__dmul_rn(u, __ull2double_rn(chrom_len - space + 1)));
    }
    const bool reverse = hi64(w) < c_rev;
    uint32_t nd = 1;
    for (uint64_t v = start; v >= 10; v /= 10) nd++;
    R.seg = G.seq + start;
    R.space = space;
    R.start = start;
    R.reverse = reverse ? 1u : 0u;
    R.rec_len = G.prefix_len + nd + 3u + seq_len + read_length + 4u;   // "@<genome>-<chrom>-" start "-F\n" | read '\n' '+' '\n' qual '\n'
    reads[i] = R;
    rec_len[i] = R.rec_len;
}

// cmp_map (src/str_manip.h:58-72): A<->T, C<->G, N->N, anything else -> 0
__device__ __forceinline__ uint8_t pb_complement(uint8_t c) {
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'N' ? 'N' : 0;
}

__global__ void __launch_bounds__(128)
k_pb_reads(const PbRead* __restrict__ reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* __restrict__ groups,
           const uint8_t* __restrict__ strpool, const uint32_t* __restrict__ rec_local, const uint64_t* __restrict__ block_base,
           uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PbRead R = reads[i];
    const uint64_t j = first_read + i;
    const GroupDev G = groups[R.group];
    uint8_t* o = out + block_base[i / kScanBlock] + rec_local[i];
    // ---- ID line
    {
        const uint8_t* pre = strpool + G.prefix_off;
        for (uint32_t t = 0; t < G.prefix_len; t++) *o++ = pre[t];
        uint8_t dg[20];
        uint32_t nd = 0;
        uint64_t v = R.start;
        do { dg[nd++] = (uint8_t)('0' + (uint32_t)(v % 10)); v /= 10; } while (v);
        for (uint32_t t = 0; t < nd; t++) *o++ = dg[nd - 1u - t];
        *o++ = '-'; *o++ = R.reverse ? 'R' : 'F'; *o++ = '\n';
    }
    // ---- the walk again, this time with the bases; append_pool stops when IT has emitted read_length bases
    const uint32_t read_length = R.read_length;
    uint32_t emitted = 0, pos = 0;
    Walk W;
    W.extra_space = G.len - read_length;
    WalkDraws D{seed, j};
    while (emitted < read_length) {
        const uint32_t rec = W.step(R, D.at(pos));
        // the template base at this position: fill_read + rev_comp over read_chrom_space bases
        // (append_pool can step past the template when a deletion the walk drew could not be recorded; the reference
        //  then reads what an earlier read left in its buffer -- here such a position is an 'N')
        const uint8_t base = pos >= R.space ? (uint8_t)'N' : R.reverse ? pb_complement(R.seg[R.space - 1u - pos]) : R.seg[pos];
        if (rec == 1) {
            const uint64_t xe = pb_draw(seed, j, 2, pos >> 1, pos & 1u);
            *o++ = base;
            const uint32_t r4 = ins_base_index(xe);
            *o++ = r4 == 0 ? 'T' : r4 == 1 ? 'C' : r4 == 2 ? 'A' : r4 == 3 ? 'G' : 0;
        } else if (rec == 3) {
            const uint64_t xe = pb_draw(seed, j, 2, pos >> 1, pos & 1u);
            uint64_t r3 = mul_floor_x87(xe, 3);
            if (r3 > 2) r3 = 2;
            // mm_nucleos[nt_map[base]][r3] (src/hts.h:36-46): the r3-th of T, C, A, G other than the base; NNN otherwise
            const uint32_t code = base == 'T' ? 0u : base == 'C' ? 1u : base == 'A' ? 2u : base == 'G' ? 3u : 4u;
            uint8_t sub = 'N';
            if (code < 4u) { const uint32_t k = (uint32_t)r3 + ((uint32_t)r3 >= code ? 1u : 0u); sub = k == 0 ? 'T' : k == 1 ? 'C' : k == 2 ? 'A' : 'G'; }
            *o++ = sub;
        } else if (rec == 0) {
            *o++ = base;
        }
        emitted += pb_emitted(rec);
        pos++;
    }
    *o++ = '\n'; *o++ = '+'; *o++ = '\n';
    const uint8_t ql = (uint8_t)(R.flags >> 8), qr = (uint8_t)(R.flags >> 16);
    for (uint32_t t = 0; t < read_length; t++) *o++ = t < R.split_pos ? ql : qr;
    *o++ = '\n';
}

}  // namespace

cudaError_t launch_pb_plan(PbRead* reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* groups,
                           uint64_t c_rev, uint32_t* rec_len, cudaStream_t s) {
    if (n) k_pb_plan<<<(n + 127) / 128, 128, 0, s>>>(reads, n, first_read, seed, groups, c_rev, rec_len);
    return cudaGetLastError();
}

cudaError_t launch_pb_reads(const PbRead* reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* groups,
                            const uint8_t* strpool, const uint32_t* rec_local, const uint64_t* block_base, uint8_t* out,
                            cudaStream_t s) {
    if (n) k_pb_reads<<<(n + 127) / 128, 128, 0, s>>>(reads, n, first_read, seed, groups, strpool, rec_local, block_base, out);
    return cudaGetLastError();
}

}  // namespace jlp
