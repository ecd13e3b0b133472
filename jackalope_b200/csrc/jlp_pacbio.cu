// PacBio read generator, device side: PacBioQualityError::sample's walk
// (src/hts_pacbio.h:296-318) and PacBioOneGenome::append_pool (src/hts_pacbio.cpp), every draw addressed by
// (seed, read index, position), so a read does not depend on batch size, thread mapping or GPU count.  One warp per
// read, one kernel template in two roles:
//   k_pb_warp<false>  walks the template once to count insertions and deletions -> read_chrom_space, read_start,
//                     strand, record length
//   k_pb_warp<true>   walks it again and writes the FASTQ record
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

// cmp_map (src/str_manip.h:58-72): A<->T, C<->G, N->N, anything else -> 0
__device__ __forceinline__ uint8_t pb_complement(uint8_t c) {
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'N' ? 'N' : 0;
}

// ---- one warp per read ---------------------------------------------------------------------------------------
// The template is walked 64 positions at a time, two per lane (one Philox block).  Away from the three places where
// the walk's rules couple neighbouring positions -- the switch from the left to the right error probabilities at
// split_pos, the end of the read (an insertion needs room for two bases, the loop stops), and a chromosome with no
// spare bases left for deletions -- every position's event is a function of its own draw, and a warp prefix sum of
// the emitted bases places them.  A chunk that may touch one of those places is walked by lane 0 with the serial code
// above; a read whose emitted count has ever run ahead of the walk's (an unrecordable deletion) stays there.
template <bool EMIT>
__global__ void __launch_bounds__(256)
k_pb_warp(PbRead* __restrict__ reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* __restrict__ groups,
          uint64_t c_rev, uint32_t* __restrict__ rec_len, const uint8_t* __restrict__ strpool,
          const uint32_t* __restrict__ rec_local, const uint64_t* __restrict__ block_base, uint8_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    PbRead R = reads[i];
    const uint64_t j = first_read + i;
    const GroupDev G = groups[R.group];
    const uint32_t RL = R.read_length;
    uint8_t* o = nullptr;
    if (EMIT && R.rec_len == 0) return;          // a duplicate whose template does not fit is not written (re_read)
    if (EMIT) {
        o = out + block_base[i / kScanBlock] + rec_local[i];
        const uint32_t idlen = R.rec_len - (RL + 4u) - (R.pad /* seq_len, kept by the plan */);
        if (lane == 0) {
            uint8_t* q = o;
            const uint8_t* pre = strpool + G.prefix_off;
            for (uint32_t t = 0; t < G.prefix_len; t++) *q++ = pre[t];
            uint8_t dg[20];
            uint32_t nd = 0;
            uint64_t v = R.start;
            do { dg[nd++] = (uint8_t)('0' + (uint32_t)(v % 10)); v /= 10; } while (v);
            for (uint32_t t = 0; t < nd; t++) *q++ = dg[nd - 1u - t];
            *q++ = '-'; *q++ = R.reverse ? 'R' : 'F'; *q++ = '\n';
        }
        o += idlen;
    }
    Walk W;
    W.extra_space = G.len - RL;
    uint32_t emitted = 0, pos = 0, n_ins = 0, n_del = 0;        // emitted: append_pool's count (capped bookkeeping in the plan)
    for (;;) {
        if (EMIT ? emitted >= RL : W.len >= RL) break;
        if (W.side == 0 && W.len == R.split_pos) W.side = 1;
        // A chunk is decided in parallel and kept only if no rule coupled its positions: the read must not come within
        // two bases of its end inside the chunk (an insertion needs room, the loop stops there), deletions must always be
        // recordable, and append_pool's count must not have run ahead of the walk's.  The switch to the right-hand error
        // probabilities is resolved in place: a position uses them iff the length before it has reached split_pos, which
        // the prefix sum over the left-hand events tells up to the first such position.
        bool fast = !(R.flags & kPbSerial) && emitted == W.len && W.extra_space >= 64 && W.len + 4u < RL;
        uint32_t ev0 = 0, ev1 = 0, e0 = 0, e1 = 0, incl = 0, total = 0;
        if (fast) {
            const U4 w = draw_block(seed, j, (pos >> 1) + lane, PL_PB, 1);
            const uint64_t x0 = lo64(w), x1 = hi64(w);
            ev0 = pb_event(R, W.side, x0); ev1 = pb_event(R, W.side, x1);
            e0 = pb_emitted(ev0); e1 = pb_emitted(ev1);
            incl = e0 + e1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += x; }
            total = __shfl_sync(0xffffffffu, incl, 31);
            if (W.side == 0 && W.len + total >= R.split_pos) {
                const uint32_t pre0 = W.len + incl - e0 - e1, pre1 = pre0 + e0;
                const bool r0 = pre0 >= R.split_pos, r1 = pre1 >= R.split_pos;
                const uint32_t any = __ballot_sync(0xffffffffu, r0 || r1);
                if (any) {
                    const uint32_t fl = (uint32_t)__ffs(any) - 1u;
                    const bool s0 = lane > fl || (lane == fl && r0), s1 = lane > fl || (lane == fl && (r0 || r1));
                    if (s0) ev0 = pb_event(R, 1, x0);
                    if (s1) ev1 = pb_event(R, 1, x1);
                    e0 = pb_emitted(ev0); e1 = pb_emitted(ev1);
                    incl = e0 + e1;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += x; }
                    total = __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            fast = W.len + total + 3u < RL;            // every position saw a length below read_length - 1, and the walk goes on
        }
        if (fast) {
            const uint32_t ins = __popc(__ballot_sync(0xffffffffu, ev0 == 1)) + __popc(__ballot_sync(0xffffffffu, ev1 == 1));
            const uint32_t del = __popc(__ballot_sync(0xffffffffu, ev0 == 2)) + __popc(__ballot_sync(0xffffffffu, ev1 == 2));
            if (EMIT) {
                uint8_t* q = o + emitted + (incl - e0 - e1);
                // the inserted / substituted bases of both positions come from one Philox block (positions 2k and 2k + 1
                // share block k of the edit plane)
                U4 we = {0, 0, 0, 0};
                if (((ev0 | ev1) & 1u) != 0) we = draw_block(seed, j, (pos >> 1) + lane, PL_PB, 2);     // events 1 and 3 are odd
#pragma unroll
                for (uint32_t h = 0; h < 2; h++) {
                    const uint32_t ev = h ? ev1 : ev0, p = pos + 2u * lane + h;
                    if (ev == 2) continue;
                    const uint8_t base = R.reverse ? pb_complement(R.seg[R.space - 1u - p]) : R.seg[p];
                    const uint64_t xe = h ? hi64(we) : lo64(we);
                    uint8_t first = base;
                    if (ev == 3) {
                        uint64_t r3 = mul_floor_x87(xe, 3);
                        if (r3 > 2) r3 = 2;
                        const uint32_t code = base == 'T' ? 0u : base == 'C' ? 1u : base == 'A' ? 2u : base == 'G' ? 3u : 4u;
                        first = 'N';
                        if (code < 4u) { const uint32_t k = (uint32_t)r3 + ((uint32_t)r3 >= code ? 1u : 0u); first = k == 0 ? 'T' : k == 1 ? 'C' : k == 2 ? 'A' : 'G'; }
                    }
                    *q++ = first;
                    if (ev == 1) {
                        const uint32_t r4 = ins_base_index(xe);
                        *q++ = r4 == 0 ? 'T' : r4 == 1 ? 'C' : r4 == 2 ? 'A' : r4 == 3 ? 'G' : 0;
                    }
                }
            }
            W.len += total; emitted += total;
            if (W.side == 0 && W.len > R.split_pos) W.side = 1;         // (reaching it exactly is seen at the next position)
            W.extra_space += ins; W.extra_space -= del;
            n_ins += ins; n_del += del;
            pos += 64;
        } else {
            // serial: lane 0 walks to the next multiple of 64 positions (or to the end of the read)
            if (lane == 0) {
                WalkDraws D{seed, j};
                const uint32_t stop = (pos & ~63u) + 64u;
                while (pos < stop && (EMIT ? emitted < RL : W.len < RL)) {
                    uint32_t rec = W.step(R, D.at(pos));
                    n_ins += rec == 1;
                    if (rec == 2 && EMIT && n_del >= R.del_keep) rec = 0;      // a deletion the duplicate gave up: a plain base
                    n_del += rec == 2;
                    if (EMIT) {
                        const uint8_t base = pos >= R.space ? (uint8_t)'N' : R.reverse ? pb_complement(R.seg[R.space - 1u - pos]) : R.seg[pos];
                        uint8_t* q = o + emitted;
                        if (rec == 1) {
                            const uint64_t xe = pb_draw(seed, j, 2, pos >> 1, pos & 1u);
                            *q++ = base;
                            const uint32_t r4 = ins_base_index(xe);
                            *q++ = r4 == 0 ? 'T' : r4 == 1 ? 'C' : r4 == 2 ? 'A' : r4 == 3 ? 'G' : 0;
                        } else if (rec == 3) {
                            const uint64_t xe = pb_draw(seed, j, 2, pos >> 1, pos & 1u);
                            uint64_t r3 = mul_floor_x87(xe, 3);
                            if (r3 > 2) r3 = 2;
                            const uint32_t code = base == 'T' ? 0u : base == 'C' ? 1u : base == 'A' ? 2u : base == 'G' ? 3u : 4u;
                            uint8_t sub = 'N';
                            if (code < 4u) { const uint32_t k = (uint32_t)r3 + ((uint32_t)r3 >= code ? 1u : 0u); sub = k == 0 ? 'T' : k == 1 ? 'C' : k == 2 ? 'A' : 'G'; }
                            *q++ = sub;
                        } else if (rec == 0) {
                            *q++ = base;
                        }
                        emitted += pb_emitted(rec);
                    } else if (emitted < RL) {
                        emitted += pb_emitted(rec);
                    }
                    pos++;
                }
            }
            W.len = __shfl_sync(0xffffffffu, W.len, 0);
            W.side = __shfl_sync(0xffffffffu, W.side, 0);
            W.extra_space = __shfl_sync(0xffffffffu, W.extra_space, 0);
            emitted = __shfl_sync(0xffffffffu, emitted, 0);
            pos = __shfl_sync(0xffffffffu, pos, 0);
            n_ins = __shfl_sync(0xffffffffu, n_ins, 0);
            n_del = __shfl_sync(0xffffffffu, n_del, 0);
        }
    }
    if (!EMIT) {
        if (lane == 0) {
            // ---- read_chrom_space, read_start (one_read, src/hts_pacbio.cpp), strand (append_pool)
            const uint32_t space = RL + n_del - n_ins;
            const U4 w = draw_block(seed, j, 1, PL_PB, 0);
            uint64_t start = 0;
            if (space < G.len) {
                const uint64_t xs = lo64(w);
                const double u = (xs == ~0ull) ? 1.0 : __ull2double_rn(xs + 1) * 5.421010862427522170037e-20;
                start = __double2ull_rz(__dmul_rn(u, __ull2double_rn(G.len - space + 1)));
            }
            uint32_t nd = 1;
            for (uint64_t v = start; v >= 10; v /= 10) nd++;
            R.seg = G.seq + start;
            R.space = space;
            R.start = start;
            R.reverse = hi64(w) < c_rev ? 1u : 0u;
            R.pad = emitted;                                                   // bases of the sequence line (read_length, or one more)
            R.rec_len = G.prefix_len + nd + 3u + emitted + RL + 4u;
            if (R.leader != kPbNoLeader) R.del_keep = n_del;                   // a duplicate: k_pb_dups settles start, span and record
            else R.del_keep = 0xffffffffu;
            reads[i] = R;
            rec_len[i] = R.rec_len;
        }
    } else {
        uint8_t* q = o + emitted;
        if (lane == 0) { q[0] = '\n'; q[1] = '+'; q[2] = '\n'; q[3 + RL] = '\n'; }
        const uint8_t ql = (uint8_t)(R.flags >> 8), qr = (uint8_t)(R.flags >> 16);
        for (uint32_t t = lane; t < RL; t += 32) q[3 + t] = t < R.split_pos ? ql : qr;
    }
}

// PacBioOneGenome::re_read (src/hts_pacbio.cpp): a duplicate keeps the read_start of the first read of its chain; the
// deletions its own walk recorded are given up from the back until the template fits the chromosome, and if it still
// does not fit the read is not written.  One thread per read, after k_pb_warp<false>.
__global__ void __launch_bounds__(128)
k_pb_dups(PbRead* __restrict__ reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* __restrict__ groups,
          uint32_t* __restrict__ rec_len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PbRead R = reads[i];
    if (R.leader == kPbNoLeader) return;
    const GroupDev G = groups[R.group];
    const uint64_t start = reads[R.leader].start;
    const uint32_t RL = R.read_length;
    uint64_t space = R.space;
    uint32_t keep = R.del_keep, dropped = 0;
    while (space + start > G.len && keep > 0) { keep--; space--; dropped++; }
    if (space + start > G.len) {
        R.rec_len = 0;
    } else {
        uint32_t seq_len = R.pad;
        if (dropped) {
            // append_pool now sees fewer deletions: what it emits is counted again, by the literal walk
            Walk W;
            W.extra_space = G.len - RL;
            WalkDraws D{seed, first_read + i};
            uint32_t emitted = 0, pos = 0, del_ord = 0;
            while (W.len < RL) {
                uint32_t rec = W.step(R, D.at(pos));
                if (rec == 2) { if (del_ord >= keep) rec = 0; del_ord++; }
                if (emitted < RL) emitted += pb_emitted(rec);
                pos++;
            }
            seq_len = emitted;
            R.flags |= kPbSerial;
        }
        uint32_t nd = 1;
        for (uint64_t v = start; v >= 10; v /= 10) nd++;
        R.seg = G.seq + start;
        R.start = start;
        R.space = (uint32_t)space;
        R.del_keep = keep;
        R.pad = seq_len;
        R.rec_len = G.prefix_len + nd + 3u + seq_len + RL + 4u;
    }
    reads[i] = R;
    rec_len[i] = R.rec_len;
}

}  // namespace

cudaError_t launch_pb_plan(PbRead* reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* groups,
                           uint64_t c_rev, uint32_t* rec_len, cudaStream_t s) {
    if (n) k_pb_warp<false><<<(n + 7) / 8, 256, 0, s>>>(reads, n, first_read, seed, groups, c_rev, rec_len, nullptr, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_pb_dups(PbRead* reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* groups, uint32_t* rec_len,
                           cudaStream_t s) {
    if (n) k_pb_dups<<<(n + 127) / 128, 128, 0, s>>>(reads, n, first_read, seed, groups, rec_len);
    return cudaGetLastError();
}

cudaError_t launch_pb_reads(const PbRead* reads, uint32_t n, uint64_t first_read, uint64_t seed, const GroupDev* groups,
                            const uint8_t* strpool, const uint32_t* rec_local, const uint64_t* block_base, uint8_t* out,
                            cudaStream_t s) {
    if (n) k_pb_warp<true><<<(n + 7) / 8, 256, 0, s>>>(const_cast<PbRead*>(reads), n, first_read, seed, groups, 0, nullptr, strpool, rec_local, block_base, out);
    return cudaGetLastError();
}

}  // namespace jlp
