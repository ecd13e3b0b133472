// Device-side data layout and kernel launchers (sm_100a).
#ifndef JLP_KERNELS_CUH
#define JLP_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace jlp {

// One (haplotype, chromosome) group: the unit the reference exhausts read
// quotas by (src/hts_illumina.cpp:199-201, :505-528).
struct GroupDev {
    const uint8_t* seq;    // chromosome bases in HBM (byte per base)
    uint64_t len;
    uint32_t prefix_off;   // "@<genome>-<chrom>-" in the string pool
    uint32_t prefix_len;
    uint32_t bc_off;       // barcode in the string pool
    uint32_t bc_len;
};

// Per read end, written by the quality/error kernel, read by the formatter.
struct RecMeta {
    uint64_t start;        // leftmost template coordinate (fill_fq_lines `start`)
    uint32_t group;
    uint16_t len;          // final read length
    uint8_t reverse;
    uint8_t pad;
};

struct EndDev {
    const uint32_t* meta;    // [4*L] offset << 8 | n
    const uint32_t* entry;   // coin16 | q_self << 16 | q_alias << 24
    const uint64_t* coin;    // full thresholds (slow path)
    const uint16_t* mis16;   // [256]
    const uint64_t* mis;     // [256]
    uint32_t entry_n;
    uint32_t hA;             // high-16 gate of the indel draw (0x10000 = always slow)
    uint64_t tA, tI;         // x >= tA: plain base; else x >= tI: deletion; else insertion
    uint32_t tA_all, tI_all; // threshold is 2^64 (never reached)
};

struct GenParams {
    uint64_t seed;
    uint64_t job_lo, job_hi;
    uint64_t pool_pairs;
    uint64_t batch_lo;
    uint32_t batch_pairs;
    uint32_t n_ends;
    uint32_t L;
    uint32_t matepair;
    uint32_t row_stride;
    uint32_t dup_never;      // prob_dup threshold is 0
    uint64_t c_dup;          // x < c_dup: duplicate of the previous fragment
    uint64_t c_rev;          // x < c_rev: reverse strand first
    EndDev end[2];
    const uint64_t* frag_cdf;
    uint32_t frag_n;
    uint32_t n_groups;
    uint64_t frag_min;
    const uint64_t* group_off;   // [n_groups + 1] pair-index prefix offsets
    const GroupDev* groups;
    const uint8_t* strpool;
    uint8_t* seq;                // [records][row_stride]
    uint8_t* qual;               // [records][row_stride]
    RecMeta* rec;                // [records]
    uint32_t* rec_len;           // [records] FASTQ bytes of each record
    uint32_t* status;            // device error bits (1: barcode >= template)
};

struct FmtParams {
    uint32_t n_records;
    uint32_t n_ends;
    uint32_t row_stride;
    uint32_t scan_block;         // records per scan block (per file)
    const uint8_t* seq;
    const uint8_t* qual;
    const RecMeta* rec;
    const uint32_t* rec_local;   // exclusive prefix inside the scan block
    const uint64_t* block_base;  // [n_ends][n_scan_blocks] exclusive prefix of block totals
    uint32_t n_scan_blocks;
    const GroupDev* groups;
    const uint8_t* strpool;
    uint8_t* out[2];
};

// haplotype materialisation (HapChrom::get_chrom_full)
cudaError_t launch_materialize(const uint8_t* ref, uint64_t ref_size, uint64_t n_muts,
                               const uint64_t* old_pos, const uint64_t* new_pos,
                               const int64_t* size_mod, const uint64_t* nuc_off,
                               const uint8_t* pool, uint64_t chrom_size, uint8_t* out,
                               cudaStream_t s);

// quality / error kernel; smem_bytes = 0 selects the global-table variant
cudaError_t launch_gen(const GenParams& p, size_t smem_bytes, cudaStream_t s);
size_t gen_smem_bytes(const GenParams& p);

// record-offset scan: local exclusive prefixes + per-block totals, then the
// exclusive prefix of the totals (and the grand total per file in totals_out)
cudaError_t launch_scan(const uint32_t* rec_len, uint32_t n_records, uint32_t n_ends,
                        uint32_t scan_block, uint32_t* rec_local, uint64_t* block_tot,
                        uint64_t* block_base, uint64_t* totals_out, cudaStream_t s);

// FASTQ formatter
cudaError_t launch_fmt(const FmtParams& p, cudaStream_t s);

constexpr uint32_t kScanBlock = 1024;

}  // namespace jlp
#endif
