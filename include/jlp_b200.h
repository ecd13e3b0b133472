/*
 * jlp_b200.h -- C ABI of the B200-native Illumina read generator.
 *
 * This is the drop-in boundary for jackalope's Illumina hot path.  The entry
 * points below are what the package's Rcpp glue would bind in place of the
 * bodies of
 *     illumina_ref_cpp   /root/reference/src/hts_illumina.cpp:589-649
 *     illumina_hap_cpp   /root/reference/src/hts_illumina.cpp:662-739
 * (INTEGRATION.md shows that glue).  Plain pointers and sizes only; no
 * exceptions cross the boundary; every function returns 0 on success and a
 * negative jlp_status otherwise, with a message available from
 * jlp_last_error().  The caller owns every input buffer; the library owns all
 * device and pinned memory.  There is no CPU fallback: without a CUDA device
 * jlp_ctx_create fails with JLP_ERR_NO_DEVICE.
 */
#ifndef JLP_B200_H
#define JLP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jlp_ctx jlp_ctx;

enum jlp_status {
    JLP_OK = 0,
    JLP_ERR_ARG = -1,        /* invalid argument (the reference would Rcpp::stop) */
    JLP_ERR_NO_DEVICE = -2,  /* no usable CUDA device */
    JLP_ERR_CUDA = -3,       /* CUDA runtime error */
    JLP_ERR_IO = -4,         /* cannot open / write an output file (src/io.h:288-290) */
    JLP_ERR_ABORTED = -5,    /* abort callback asked to stop (Progress::check_abort, src/hts.h:396-399) */
    JLP_ERR_UNSUPPORTED = -6 /* a feature of the reference not built yet (none on this path at present) */
};

/* ---- context: one per GPU (one process per GPU under torchrun) -------------- */

int jlp_ctx_create(int device, jlp_ctx** out);
void jlp_ctx_destroy(jlp_ctx* ctx);
/* Last error text of this context (or of context creation when ctx == NULL). */
const char* jlp_last_error(const jlp_ctx* ctx);

/* ---- genome data model (replaces reading RefGenome / HapSet through XPtr,
 *      src/hts_illumina.cpp:614,689; layouts src/ref_classes.h:36-116,
 *      src/hap_classes.h:100-104,287-296) ------------------------------------- */

/* Reference genome: chromosome c is bases[chrom_off[c] .. chrom_off[c+1]).
 * genome_name is what fill_fq_lines prints first ("REF", src/ref_classes.h:138). */
int jlp_set_genome(jlp_ctx* ctx, const char* bases, const uint64_t* chrom_off, uint64_t n_chroms,
                   const char* const* chrom_names, const char* genome_name);

/* The same upload, deferred: nothing is copied yet.  A following jlp_illumina_* call on
 * the reference genome copies, chromosome by chromosome on its own stream, the
 * chromosomes its shard of the job reads (all of them for an unsharded job, about 1/N
 * with N shards), starts generating as soon as the first batch's are resident and waits
 * for its copies before it returns.  Chromosomes no run has read stay on the host, so
 * `bases` must stay valid until jlp_genome_sync() (which uploads the rest) or until the
 * genome is replaced.  Haplotype calls upload the whole genome first. */
int jlp_set_genome_async(jlp_ctx* ctx, const char* bases, const uint64_t* chrom_off, uint64_t n_chroms,
                         const char* const* chrom_names, const char* genome_name);
int jlp_genome_sync(jlp_ctx* ctx);

/* create_genome (create_genome_cpp, src/create_sequences.cpp:162-184): random
 * chromosomes of the given lengths, nucleotides alias-sampled from pi_tcag (T, C, A, G), generated
 * straight into device memory -- the genome is resident afterwards, no upload needed.  chrom_names may be
 * NULL ("chrom<i>").  jlp_get_genome copies the bases (all chromosomes, concatenated) to the host;
 * out == NULL only asks for the size. */
int jlp_create_genome(jlp_ctx* ctx, uint64_t n_chroms, const uint64_t* lens, const double* pi_tcag, uint64_t seed,
                      const char* const* chrom_names, const char* genome_name);
int jlp_get_genome(jlp_ctx* ctx, char* out, uint64_t cap, uint64_t* len);

/* Drop all haplotypes previously added. */
int jlp_clear_haplotypes(jlp_ctx* ctx);

/* One haplotype = per chromosome the sorted AllMutations arrays
 * (old_pos, new_pos, nucleos) plus the mutated chromosome size.  nuc_off[c][i]
 * indexes nuc_pool[c]; a deletion has nuc_len 0, a substitution 1, an insertion
 * 1+k.  The chromosome is materialised on the device at once
 * (HapChrom::get_chrom_full, src/hap_classes.cpp:80-116).  Returns the
 * haplotype's index through *hap_index. */
int jlp_add_haplotype(jlp_ctx* ctx, const char* name, const uint64_t* n_muts,
                      const uint64_t* const* old_pos, const uint64_t* const* new_pos,
                      const uint64_t* const* nuc_off, const char* const* nuc_pool,
                      const uint64_t* nuc_pool_len, const uint64_t* chrom_sizes,
                      uint64_t* hap_index);

/* Copy a materialised haplotype chromosome back to the host (parity tests of
 * the materialisation kernel; ~ HapChrom::get_chrom_full). */
int jlp_get_haplotype_chrom(jlp_ctx* ctx, uint64_t hap, uint64_t chrom, char* out, uint64_t cap,
                            uint64_t* len);

/* ---- ART quality profile for one read end (0 or 1), flattened
 *      [nt in T,C,A,G][pos][k]: nq[nt*L+pos] entries per position, probs and
 *      quals concatenated in that order (the nested vectors qual_probs{1,2} /
 *      quals{1,2} of src/hts_illumina.cpp:604-609). ---------------------------- */
int jlp_set_profile(jlp_ctx* ctx, int end, uint64_t read_length, const uint32_t* nq,
                    const double* probs, const uint8_t* quals);

/* ---- run --------------------------------------------------------------------- */

typedef int (*jlp_abort_cb)(void* user);                       /* non-zero = abort */
typedef void (*jlp_progress_cb)(void* user, uint64_t reads);   /* reads written since last call */

typedef struct jlp_illumina_params {
    /* the arguments of illumina_{ref,hap}_cpp, same meaning */
    int paired;
    int matepair;
    const char* out_prefix;      /* files <prefix>[_<hap>]_R{1,2}.fq, src/hts.h:344,541 */
    int sep_files;               /* haplotype runs only */
    int compress;                /* 0 = plain FASTQ; 1..9 = zlib level, files get ".gz" appended (src/io.h:126,217) */
    const char* comp_method;     /* "gzip" | "bgzip" (src/hts.h:470); with n_threads > 1 always bgzip (src/hts.h:478-490) */
    uint64_t n_reads;
    double prob_dup;
    uint64_t n_threads;          /* host writer threads (pwrite of the pinned batch buffers), 1..64 */
    int show_progress;           /* unused by the library; see progress callback */
    uint64_t read_pool_size;     /* duplicate chains stop at pool boundaries, src/hts.h:266-267 */
    const double* haplotype_probs; /* [n_haps], haplotype runs only */
    double frag_len_shape, frag_len_scale;
    uint64_t frag_len_min, frag_len_max;
    double ins_prob1, del_prob1, ins_prob2, del_prob2;
    const char* const* barcodes; /* [1] for a reference run, [n_haps] for haplotypes; NULL = none */
    /* additions of this implementation */
    uint64_t seed;               /* run seed; the glue draws it from R's RNG (src/pcg.h:37-46) */
    uint64_t batch_pairs;        /* pairs per device batch; 0 = default */
    uint32_t shard_index;        /* this process generates jlp_shard_range(job, shard_index, shard_count) of every job */
    uint32_t shard_count;        /* 0 or 1 = everything */
    jlp_abort_cb abort_cb;
    jlp_progress_cb progress_cb;
    void* cb_user;
    int comp_engine;             /* who compresses when compress > 0: enum jlp_comp_engine */
} jlp_illumina_params;

/* Compressed output (write_reads_cpp_, src/hts.h:441-500).  The device coder writes BGZF members of one
 * dynamic-Huffman block each on the GPU, so only compressed bytes cross PCIe.  Levels 1..3: literals only (on
 * FASTQ the size zlib gives at level 1, 0.375 of the input); levels 4..6: also length/distance pairs for the
 * prefix a line shares with the line four lines earlier (the ID lines' "@<genome>-<chrom>-", the quality
 * lines' first characters): 0.34 of the input, against 0.32 for zlib at level 6, for about twice the kernel
 * time.  The host coder is zlib at `compress` on the writer threads.  AUTO: device for levels 1..6 when writing files, host for 7..9.  Memory and
 * stream sinks (jlp_illumina_to_memory / _stream) receive BGZF bytes, EOF block included, only with
 * JLP_COMP_DEVICE; otherwise they always receive plain FASTQ. */
enum jlp_comp_engine { JLP_COMP_AUTO = 0, JLP_COMP_HOST = 1, JLP_COMP_DEVICE = 2 };

typedef struct jlp_run_stats {
    uint64_t pairs;              /* read pairs (or single reads) generated by this call */
    uint64_t bytes_out[2];       /* FASTQ bytes produced per end */
    uint64_t batches;
    uint64_t kernel_launches;    /* launches of this library's kernels */
    double device_ms;            /* sum of CUDA-event time of the kernels */
    double place_ms;             /* ... of which the placement kernel (k_place) */
    double reads_ms;             /* ... of which the fused quality/error + FASTQ kernel (k_reads) */
    uint64_t d2h_bytes;
    uint64_t h2d_bytes;
    double run_ms;               /* CUDA-event time on the compute stream from the start of the call's device
                                    work to the end of its last kernel (includes waits on the D2H double buffer) */
    uint64_t z_bytes[2];         /* compressed bytes produced per end by the device coder (0 otherwise) */
    double bgzf_ms;              /* ... of device_ms, the BGZF kernels (k_bgzf, k_bgzf_scan, k_bgzf_gather) */
} jlp_run_stats;

/* Reads from the reference genome (illumina_ref_cpp). */
int jlp_illumina_ref(jlp_ctx* ctx, const jlp_illumina_params* p, jlp_run_stats* stats);
/* Reads from the haplotypes added so far (illumina_hap_cpp). */
int jlp_illumina_hap(jlp_ctx* ctx, const jlp_illumina_params* p, jlp_run_stats* stats);

/* Same generators, output captured in caller-provided HOST buffers instead of
 * files (tests; bench.py's end-to-end leg).  With sep_files the haplotypes'
 * outputs are concatenated in haplotype order.  *len receives the bytes
 * needed; JLP_ERR_ARG if a buffer is too small. */
int jlp_illumina_to_memory(jlp_ctx* ctx, int use_haplotypes, const jlp_illumina_params* p,
                           char* out1, uint64_t cap1, uint64_t* len1,
                           char* out2, uint64_t cap2, uint64_t* len2, jlp_run_stats* stats);

/* Same generators, FASTQ handed to the caller batch by batch in the library's
 * double-buffered PINNED host buffers (the buffers FileUncomp::write would be fed
 * from, src/hts.h:225-241): cb(user, job, end, data, n) is called on the calling
 * thread, R1 then R2 of a batch, batches in order; `job` counts output file sets
 * (one per haplotype with sep_files, else 0).  `data` is valid during the call
 * only.  A non-zero return stops the run with JLP_ERR_IO. */
typedef int (*jlp_chunk_cb)(void* user, uint64_t job, int end, const char* data, uint64_t n);
int jlp_illumina_stream(jlp_ctx* ctx, int use_haplotypes, const jlp_illumina_params* p,
                        jlp_chunk_cb cb, void* user, jlp_run_stats* stats);

/* Same generators with the FASTQ left in device memory and dropped (bench.py's
 * device-resident leg: everything but the D2H copy and the file write). */
int jlp_illumina_device_only(jlp_ctx* ctx, int use_haplotypes, const jlp_illumina_params* p,
                             jlp_run_stats* stats);

/* Pairs apportioned to each (haplotype, chromosome) group, haplotype-major, for
 * exactly the run `p` describes (reads_per_group is seeded from p->seed, so the
 * generators above make the same split).  A reference run has n_chroms groups. */
int jlp_illumina_group_counts(jlp_ctx* ctx, int use_haplotypes, const jlp_illumina_params* p,
                              uint64_t* counts, uint64_t cap, uint64_t* n_groups);

/* ---- PacBio reads: pacbio_ref_cpp / pacbio_hap_cpp,
 *      /root/reference/src/hts_pacbio.cpp; arguments of pacbio(), R/hts_pacbio.R -------------------------------- */

typedef struct jlp_pacbio_params {
    /* the arguments of pacbio_{ref,hap}_cpp, same meaning */
    const char* out_prefix;      /* files <prefix>[_<hap>]_R1.fq */
    int sep_files;               /* haplotype runs only */
    int compress;                /* as for Illumina; the device coder serves levels 1..6 */
    const char* comp_method;
    uint64_t n_reads;
    uint64_t n_threads;          /* accepted for compatibility; the files are written by the calling thread */
    uint64_t read_pool_size;
    const double* haplotype_probs;
    double prob_dup;             /* duplicates re-read their chain's chromosome, length and start (re_read); read_pool_size <= 65536 then */
    double scale, sigma, loc;    /* lognorm_read_length[3], [1], [2] */
    double min_read_len;
    const double* read_probs;    /* custom_read_lengths: n_custom probabilities and lengths, or NULL / 0 */
    const uint64_t* read_lens;
    uint64_t n_custom;
    uint64_t max_passes;
    double chi2_params_n[3];
    double chi2_params_s[5];
    double sqrt_params[2];
    double norm_params[2];
    double prob_thresh, prob_ins, prob_del, prob_subst;
    /* additions of this implementation */
    uint64_t seed;
    uint64_t batch_reads;        /* reads per device batch; 0 = default */
    int comp_engine;             /* enum jlp_comp_engine */
    uint32_t shard_index;        /* this process generates its share of every job: whole pools of read_pool_size reads, */
    uint32_t shard_count;        /* contiguous and near-equal over the shards (0 or 1 = everything) */
} jlp_pacbio_params;

/* Reads into files (use_haplotypes: the haplotypes added so far, else the reference genome). */
int jlp_pacbio(jlp_ctx* ctx, int use_haplotypes, const jlp_pacbio_params* p, jlp_run_stats* stats);
/* The same into a caller-provided host buffer (with sep_files the haplotypes' outputs are concatenated); out == NULL
 * generates and drops the reads on the device.  *len receives the bytes needed. */
int jlp_pacbio_to_memory(jlp_ctx* ctx, int use_haplotypes, const jlp_pacbio_params* p, char* out, uint64_t cap,
                         uint64_t* len, jlp_run_stats* stats);
/* What the run `p` describes draws for each of its reads before the per-base work: the (haplotype, chromosome) group,
 * the read length (already limited to the chromosome), split_pos, passes_left, passes_right -- the quantities the
 * reference's stateful samplers produce, which the parity tests inject into the oracle.  Arrays of n_reads entries. */
int jlp_pacbio_read_plan(jlp_ctx* ctx, int use_haplotypes, const jlp_pacbio_params* p, uint64_t* group, uint64_t* read_len,
                         uint64_t* split_pos, double* passes_left, double* passes_right);
/* The samplers alone, no device needed: n reads of a chromosome of chrom_len bases (statistical tests against
 * PacBioReadLenSampler::sample / PacBioPassSampler::sample). */
int jlp_pacbio_sample(const jlp_pacbio_params* p, uint64_t n, uint64_t chrom_len, uint64_t* read_len, uint64_t* split_pos,
                      double* passes_left, double* passes_right);

/* ---- host-side pieces exposed for CPU tests (no device needed) --------------- */

/* The apportioning jlp_illumina_group_counts reports, from sizes alone: pairs ->
 * haplotypes by hap_probs, then each haplotype's pairs -> chromosomes by size
 * (write_reads_one_filetype_ src/hts.h:334-353, add_n_reads src/hts_illumina.h:410-418
 * and :620-644).  hap_probs == NULL is a reference-genome run (n_haps ignored,
 * sizes[n_chroms]); otherwise sizes is [n_haps][n_chroms].  counts has the same shape. */
int jlp_apportion(uint64_t seed, uint64_t n_pairs, uint64_t n_haps, uint64_t n_chroms,
                  const double* hap_probs, const uint64_t* sizes, uint64_t* counts);
/* The compressed form the files sink writes for `n` bytes of FASTQ: BGZF blocks + EOF block
 * (bgzf != 0; FileBGZF, src/io.h:58-135) or concatenated gzip members (FileGZ, src/io.h:140-236).
 * *len receives the size (out == NULL only asks for it); JLP_ERR_ARG if `cap` is too small. */
int jlp_deflate(int bgzf, int level, const void* in, uint64_t n, void* out, uint64_t cap, uint64_t* len);
/* The same for the device coder: `n` host bytes are uploaded, compressed by the BGZF kernels and brought back,
 * EOF block appended (parity tests of the coder on arbitrary bytes; needs a device). */
int jlp_bgzf_device(jlp_ctx* ctx, int level, const void* in, uint64_t n, void* out, uint64_t cap, uint64_t* len);
/* Pair-index range [lo, hi) of job [job_lo, job_hi) that shard `shard_index` of
 * `shard_count` generates (jlp_illumina_params.shard_index / shard_count): contiguous
 * and near-equal, as split_int (src/util.h:245-258) splits reads over threads. */
int jlp_shard_range(uint64_t job_lo, uint64_t job_hi, uint32_t shard_index, uint32_t shard_count,
                    uint64_t* lo, uint64_t* hi);

/* reads_per_group (src/hts.h:58-103): multinomial apportioning by conditional
 * binomials.  Statistically equivalent to the reference (own engine). */
int jlp_reads_per_group(uint64_t n_reads, const double* probs, uint64_t n, uint64_t seed,
                        uint64_t* out);
/* AliasSampler::construct (src/alias_sampler.h:68-106). */
int jlp_alias_build(const double* probs, uint64_t n, double* prob_out, uint64_t* alias_out);
/* Integer thresholds equivalent to the reference's comparisons on
 * u = runif_01(x): kind 1 -> #{x : double(u) < p}; kind 2 -> #{x : !(double(u) > p)};
 * kind 3 -> #{x : u < p} (long double).  *all is set when the count is 2^64. */
int jlp_threshold(int kind, double p, uint64_t* thr, int* all);
/* The exact integer restatements used by the kernels, kinds as in
 * oracle/ref_driver.cpp jref_unif_expr (0: die, 4: N quality, 5: frag start). */
uint64_t jlp_unif_expr(int kind, uint64_t x, double p, uint64_t n);
/* Fragment-length table: frag_len = frag_min + #{i : X >= cdf[i]} reproduces
 * min(max((uint64)Gamma(shape, scale), frag_min), frag_max)
 * (src/hts_illumina.cpp:206-208).  Returns the entry count through *n. */
int jlp_frag_table(double shape, double scale, uint64_t frag_min, uint64_t frag_max,
                   uint64_t* cdf, uint64_t cap, uint64_t* n);
/* Philox4x32-10 block and a logical draw, as the kernels address them. */
void jlp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t jlp_draw_pos(uint64_t seed, uint64_t j, uint32_t end, uint32_t purpose, uint32_t pos);
uint64_t jlp_draw_pair(uint64_t seed, uint64_t j, int which);
/* logical draw of the genome generator: which = 0 die roll, 1 alias coin, of base `pos` of chromosome `chrom` */
uint64_t jlp_genome_draw(uint64_t seed, uint32_t chrom, uint64_t pos, int which);

const char* jlp_version(void);

#ifdef __cplusplus
}
#endif
#endif
