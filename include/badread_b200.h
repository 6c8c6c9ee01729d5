/*
 * badread_b200.h — C ABI of libbadread_b200.so: the drop-in boundary for Badread's per-read
 * error-injection hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (rrwick/Badread v0.4.2) is pure Python; its only native call on this path is
 * `edlib.align` (third-party).  This ABI is what a ctypes binding inside the reference would bind to replace
 *     simulate.sequence_fragment            badread/simulate.py:256-358
 *     ErrorModel.add_errors_to_kmer         badread/error_model.py:135-176   (table sampling, on device)
 *     qscore_model.get_qscores              badread/qscore_model.py:32-75
 *     QScoreModel.get_qscore                badread/qscore_model.py:273-287
 *     edlib.align(..., task='path')         call sites simulate.py:330,340; qscore_model.py:37; error_model.py:202
 * Plain pointers and sizes only; the caller owns every host buffer, the library owns device memory behind an
 * opaque bb_ctx.  All functions return 0 on success or a negative bb_status; bb_last_error() gives the text.
 * The library never calls exit() and never falls back to a CPU implementation of the hot path.
 */
#ifndef BADREAD_B200_H
#define BADREAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BB_API __attribute__((visibility("default")))
#else
#define BB_API
#endif

typedef struct bb_ctx bb_ctx;

enum bb_status {
    BB_OK = 0,
    BB_ERR_CUDA = -1,      /* CUDA runtime error (no device, launch failure, out of memory) */
    BB_ERR_ARG = -2,       /* invalid argument */
    BB_ERR_STATE = -3,     /* models / reference not uploaded yet */
    BB_ERR_CAPACITY = -4,  /* caller-provided output buffer too small (required size is reported) */
    BB_ERR_INTERNAL = -5   /* device-side invariant violated (reported with a code in bb_last_error) */
};

/* Segment kinds of a fragment descriptor. A fragment (what simulate.build_fragment returns, simulate.py:91-115)
 * is the concatenation of its segments: slices of the HBM-resident reference (either strand) and literal bytes
 * (adapters, glitch inserts, junk / random reads). */
enum bb_seg_kind {
    BB_SEG_REF_FWD = 0, /* reference bytes [src, src+len) as stored */
    BB_SEG_REF_REV = 1, /* reverse complement (misc.reverse_complement, misc.py:56-71) of reference [src, src+len) */
    BB_SEG_LITERAL = 2  /* bytes [src, src+len) of the literal pool passed with the batch */
};

typedef struct bb_segment {
    int64_t src;   /* offset into the uploaded reference or into the batch's literal pool */
    int32_t len;
    int32_t kind;  /* bb_seg_kind */
} bb_segment;

/* Per-read outputs besides the bases (header fields of simulate.py:73-75 come from these). */
typedef struct bb_read_result {
    int64_t out_off;      /* offset of this read's seq/qual in the output buffers (reads are packed without gaps,
                           * but not necessarily in batch order) */
    int32_t out_len;      /* len(seq) after trimming; 0 => the reference skips the read (simulate.py:70) */
    int32_t frag_len;     /* len(fragment) before padding ("error-free_length") */
    int32_t matches;      /* '=' columns of the final alignment */
    int32_t columns;      /* all columns; read identity = matches / columns (misc.py:228-240) */
    int32_t loop_count;   /* iterations of the k-mer loop (simulate.py:278) */
    int32_t change_count; /* applied slot changes (simulate.py:311) */
    int32_t n_alignments; /* identity re-measurements (simulate.py:325-346) */
    int32_t flags;        /* non-zero: device-side problem for this read (see bb_last_error) */
    int32_t loop_kcycles; /* diagnostics: SM kilo-cycles this read spent in the error loop ... */
    int32_t align_kcycles;/* ... and in the final alignment */
} bb_read_result;

/* ---- lifecycle -------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`. seed is `--seed` (simulate.py:34-36). Fails with BB_ERR_CUDA when
 * there is no usable GPU: there is no CPU path.
 * A context holds BADREAD_B200_SUBBATCHES (environment, default 4, 1..8) workers on the device: large batches are
 * dealt out over them and their kernel chains overlap on separate streams.  Results do not depend on it. */
BB_API int bb_create(bb_ctx **ctx, int device, uint64_t seed);
BB_API int bb_destroy(bb_ctx *ctx);
BB_API const char *bb_last_error(const bb_ctx *ctx); /* ctx may be NULL for creation errors */
BB_API const char *bb_version(void);

/* ---- one-time uploads ------------------------------------------------------------------------------- */
/* Reference contigs concatenated (upper-case ASCII, misc.load_fasta misc.py:122-153). */
BB_API int bb_upload_reference(bb_ctx *ctx, const uint8_t *bases, int64_t n_bases);

/* Error model tables (flat form of ErrorModel.alternatives / .probabilities, error_model.py:86-133):
 *  type 0 = 'random' (k = 1, no tables), 1 = 'model'.
 *  kmer_to_row[4^k]: row of each ACGT k-mer or -1; row_off[n_rows+1] entry ranges; per entry: cum (the
 *  list(accumulate(probs)) that random.choices builds), flags (bit0: ''.join(alt)==kmer, bit1: the
 *  "random change" remainder entry appended at error_model.py:151-154), k slot strings encoded as
 *  len | chars<<8 (len<=3) or len | pool_offset<<8. */
BB_API int bb_upload_error_model(bb_ctx *ctx, int k, int type, const int32_t *kmer_to_row, int64_t n_index, int32_t n_rows,
                          const int32_t *row_off, const double *cum, const uint8_t *flags, const uint32_t *slots,
                          const uint8_t *pool, int64_t pool_len);

/* Qscore model tables (flat form of QScoreModel.scores / .probabilities, qscore_model.py:178-271).
 *  keys[n_keys]: CIGAR strings over {=,X,I,D} packed 2 bits per symbol under a leading 1 bit (<= 31 symbols);
 *  row_off[n_keys+1]; scores / cum per entry. kmer_size as QScoreModel.kmer_size. */
BB_API int bb_upload_qscore_model(bb_ctx *ctx, int kmer_size, int32_t n_keys, const uint64_t *keys, const int32_t *row_off,
                           const uint8_t *scores, const double *cum);

/* ---- the hot path ----------------------------------------------------------------------------------- */
/* sequence_fragment for a batch of reads (simulate.py:256-358 for each).
 *  read_index[n]: global read ordinal, keys the per-read Philox streams (output is independent of batching and
 *  of the number of GPUs). seg_off[n+1] indexes segs. target_identity[n] as Identities.get_identity().
 *  Outputs: results[n]; seq_out / qual_out receive the trimmed reads back to back (capacity out_cap bytes each);
 *  *out_total = bytes written. If out_cap is too small: returns BB_ERR_CAPACITY with *out_total = required size
 *  (device results are kept; call bb_fetch_last_batch with larger buffers). */
BB_API int bb_sequence_batch(bb_ctx *ctx, int32_t n_reads, const uint64_t *read_index, const int32_t *seg_off,
                      const bb_segment *segs, const uint8_t *literal_pool, int64_t literal_len,
                      const double *target_identity, bb_read_result *results, uint8_t *seq_out, uint8_t *qual_out,
                      int64_t out_cap, int64_t *out_total);
BB_API int bb_fetch_last_batch(bb_ctx *ctx, bb_read_result *results, uint8_t *seq_out, uint8_t *qual_out, int64_t out_cap,
                        int64_t *out_total);

/* Page-locked host memory for the seq / qual output buffers (device-to-host copies into pageable memory go through
 * a staging buffer at a fraction of the link rate).  Optional: any host pointer is accepted by the fetch calls. */
BB_API int bb_host_alloc(void **ptr, int64_t bytes);
BB_API int bb_host_free(void *ptr);

/* Split form used by bench.py to time the device work with inputs already resident in HBM:
 * bb_batch_upload (H2D of descriptors) -> bb_batch_run (kernels only, asynchronous on the ctx stream;
 * may be called repeatedly on the same uploaded batch) -> bb_fetch_last_batch (D2H). */
BB_API int bb_batch_upload(bb_ctx *ctx, int32_t n_reads, const uint64_t *read_index, const int32_t *seg_off,
                    const bb_segment *segs, const uint8_t *literal_pool, int64_t literal_len,
                    const double *target_identity);
BB_API int bb_batch_run(bb_ctx *ctx);
BB_API int bb_synchronize(bb_ctx *ctx);
/* CUDA-event time (ms) of the last bb_batch_run on the ctx stream, total and per stage
 * (stage_ms[BB_N_STAGES], see bb_stage_name). Synchronizes. */
#define BB_N_STAGES 8
BB_API int bb_last_run_ms(bb_ctx *ctx, float *total_ms, float *stage_ms);
BB_API const char *bb_stage_name(int stage);
/* Number of kernel launches issued by this context so far. */
BB_API int64_t bb_launch_count(const bb_ctx *ctx);
/* Diagnostics: with BADREAD_B200_TRACE=1 in the environment at bb_create, every launch of a run is followed by a CUDA
 * event; this writes the last run's timeline (worker, stream, name, begin_ms, end_ms) as CSV. */
BB_API int bb_trace_dump(bb_ctx *ctx, const char *path);

/* get_qscores(seq, frag, qscore_model) on its own (qscore_model.py:32-75) for one pair; qual_out has seq_len bytes. */
BB_API int bb_get_qscores(bb_ctx *ctx, uint64_t read_index, const uint8_t *seq, int32_t seq_len, const uint8_t *frag,
                   int32_t frag_len, uint8_t *qual_out, int32_t *matches, int32_t *columns);

/* edlib.align(query, target, task='path') on the device for one pair: expanded CIGAR (one of "=XID" per column)
 * into ops_out (capacity ops_cap); *n_ops = columns, *distance = edit distance. Diagnostic / test entry point
 * for the aligner the kernels use. */
BB_API int bb_align_path(bb_ctx *ctx, const uint8_t *query, int32_t q_len, const uint8_t *target, int32_t t_len,
                  uint8_t *ops_out, int64_t ops_cap, int64_t *n_ops, int32_t *distance);

/* ---- host-side helpers (no GPU needed) -------------------------------------------------------------- */
/* error_model.align_kmers (error_model.py:179-229) for a batch of (kmer, alt) pairs: kmers is n_alts*k bytes,
 * alts are concatenated with alt_off[n_alts+1]. Writes n_alts*k encoded slots, appends long strings to pool
 * (capacity pool_cap, *pool_len updated) and flags bit0 = (''.join(slots) == kmer). */
BB_API int bb_host_align_kmers(int k, int32_t n_alts, const uint8_t *kmers, const uint8_t *alts, const int32_t *alt_off,
                        uint32_t *slots_out, uint8_t *flags_out, uint8_t *pool, int64_t pool_cap, int64_t *pool_len);
/* edlib.align(query, target, task='path') on the host for SMALL inputs (full matrix; q_len*t_len <= 2^22). */
BB_API int bb_host_align_path(const uint8_t *query, int32_t q_len, const uint8_t *target, int32_t t_len, uint8_t *ops_out,
                       int64_t ops_cap, int64_t *n_ops, int32_t *distance);


/* ---- multi-GPU: the one collective of the path ------------------------------------------------------ */
/* Reads shard over GPUs by read index (GPU g owns indices = g mod G) and never exchange data.  The only value the GPUs
 * have to agree on is the running total of emitted bases that ends the simulation (simulate.py:63): one NCCL
 * all-reduce (SUM, one int64) per batch.  NCCL is loaded at run time (dlopen of libnccl.so.2); bb_nccl_available()
 * says whether that worked.  Two ways to form the communicator:
 *   one process per GPU (torchrun / mpirun): rank 0 calls bb_comm_unique_id, the host program ships the 128 bytes to
 *     the other ranks by whatever means it has, every rank calls bb_comm_init_rank, then bb_allreduce_bases;
 *   one process, several GPUs (`badread simulate --gpus N`): bb_comm_init_all, then bb_allreduce_bases_all. */
typedef struct bb_nccl_id { char internal[128]; } bb_nccl_id;   /* ncclUniqueId */
BB_API int bb_nccl_available(void);
BB_API int bb_comm_unique_id(bb_nccl_id *id);
BB_API int bb_comm_init_rank(bb_ctx *ctx, const bb_nccl_id *id, int rank, int world);
BB_API int bb_comm_init_all(bb_ctx **ctxs, int n);
BB_API int bb_allreduce_bases(bb_ctx *ctx, int64_t local, int64_t *total);
BB_API int bb_allreduce_bases_all(bb_ctx **ctxs, int n, const int64_t *local, int64_t *total);

/* ---- fragment builder and FASTQ assembly on the host (no GPU needed) -------------------------------- */
/* The steps either side of the hot path, as native multi-threaded host code.  bb_planner_plan replaces the per-read
 * Python of build_fragment (badread/simulate.py:91-115), get_fragment / get_real_fragment / get_junk_fragment
 * (:148-253), the adapters (:361-394), add_glitches (:459-482), FragmentLengths.get_fragment_length
 * (fragment_lengths.py:47-64) and Identities.get_identity (identities.py:76-94); it emits fragment DESCRIPTORS in
 * exactly the layout bb_batch_upload / bb_sequence_batch take.  Every read draws from its own random.Random /
 * numpy RandomState keyed by (seed, read index), restated bit for bit. */
typedef struct bb_planner bb_planner;

typedef struct bb_plan_config {
    uint64_t seed;
    /* reference (misc.load_fasta, simulate.py:118-121): contigs in file order */
    int32_t n_contigs;
    const int64_t *contig_len;
    const double *contig_weight;     /* depth * length after adjust_depths (simulate.py:516-536) */
    const uint8_t *contig_flags;     /* bit 0 circular, bit 1 hairpin_left, bit 2 hairpin_right */
    const char *contig_names;        /* concatenated; contig i is [contig_name_off[i], contig_name_off[i+1]) */
    const int64_t *contig_name_off;
    /* fragment lengths (fragment_lengths.py): stdev == 0 => constant */
    double frag_mean, frag_stdev, gamma_k, gamma_t;
    /* identities (identities.py): type 0 beta (mean, max as fractions), 1 normal (mean, stdev as qscores) */
    int32_t identity_type;
    double id_mean, id_stdev, id_max, beta_a, beta_b;
    /* adapters (simulate.py:361-394): rate / amount as fractions */
    const uint8_t *start_adapter; int32_t start_adapter_len; double start_adapter_rate, start_adapter_amount;
    const uint8_t *end_adapter; int32_t end_adapter_len; double end_adapter_rate, end_adapter_amount;
    /* read types (simulate.py:168-180) and chimeras (:101-110), as fractions */
    double junk_rate, random_rate, chimera_rate, chimera_end_adapter_chance, chimera_start_adapter_chance;
    /* glitches (simulate.py:459-482) */
    double glitch_rate, glitch_size, glitch_skip;
} bb_plan_config;

/* The last plan of a planner (pointers stay valid until the next bb_planner_plan / bb_planner_destroy). */
typedef struct bb_plan_view {
    int32_t n_reads;
    const uint64_t *read_index;      /* [n] */
    const int32_t *seg_off;          /* [n+1] */
    const bb_segment *segs;
    const uint8_t *literals;
    int64_t literal_len;
    const double *target_identity;   /* [n] */
    const uint8_t *read_names;       /* [n][16]: uuid.UUID(int=random.getrandbits(128)).bytes (simulate.py:77) */
    const int64_t *info_off;         /* [n+1] into info */
    const char *info;                /* ' '.join(info) of simulate.py:97-113, before the length fields */
    const int32_t *frag_len;         /* [n] len(fragment) */
} bb_plan_view;

BB_API int bb_planner_create(bb_planner **planner, const bb_plan_config *config);
BB_API int bb_planner_destroy(bb_planner *planner);
/* Plans reads first_index, first_index + stride, ... (n_reads of them) with n_threads host threads.
 * BB_ERR_STATE: a read could not be built (the reference exits with bb_planner_error()'s message, simulate.py:164). */
BB_API int bb_planner_plan(bb_planner *planner, uint64_t first_index, uint64_t stride, int32_t n_reads, int32_t n_threads);
BB_API int bb_planner_view(const bb_planner *planner, bb_plan_view *view);
BB_API const char *bb_planner_error(const bb_planner *planner);

/* FASTQ records of simulate.py:70-86 for reads [first, n) of a finished batch in plan order, into out: empty reads
 * are skipped; stops after the read with which bases_so_far + emitted bases reaches target_bases.  *out_len = bytes
 * needed (BB_ERR_CAPACITY if out_cap is smaller), *next_read = first read not consumed. */
BB_API int bb_fastq_format(const bb_plan_view *view, const bb_read_result *results, const uint8_t *seq, const uint8_t *qual,
                    int32_t first, int64_t bases_so_far, int64_t target_bases, int32_t n_threads, uint8_t *out,
                    int64_t out_cap, int64_t *out_len, int32_t *n_emitted, int64_t *bases_emitted, int32_t *next_read);
/* The same for a batch dealt out over n_shards contexts (GPUs): read j of the batch is read j / n_shards of shard
 * j % n_shards (views[g], results[g], seq[g], qual[g]); records come out in read-index order, independent of n_shards. */
BB_API int bb_fastq_format_sharded(int32_t n_shards, const bb_plan_view *const *views, const bb_read_result *const *results,
                            const uint8_t *const *seq, const uint8_t *const *qual, int32_t first, int64_t bases_so_far,
                            int64_t target_bases, int32_t n_threads, uint8_t *out, int64_t out_cap, int64_t *out_len,
                            int32_t *n_emitted, int64_t *bases_emitted, int32_t *next_read);

/* ---- Model builders: the counting passes of `badread error_model` (badread/error_model.py:31-83) and `badread
 * qscore_model` (badread/qscore_model.py:78-161) on the GPU.  The caller has parsed the inputs and chosen the alignments
 * (badread/alignment.py:79-105) and hands them over flat: for alignment a = 0 .. n_aln-1 the aligned slice of the read
 * read[read_off[a] .. read_off[a+1]) (and its qualities), the aligned slice of the reference ref[ref_off[a] .. ref_off[a+1])
 * already on the read's strand, and the CIGAR runs ops[ops_off[a] .. ops_off[a+1]) in read orientation, each
 * (length << 2) | type with type 0 = M, 1 = I, 2 = D, starting at read offset op_read0[] / reference offset op_ref0[] within
 * the alignment.  A window's content is a 64-bit key; the library returns every distinct key with its count(s) and the first
 * window it occurred in (the reference's dicts keep insertion order and its stable sorts break ties by it), in arbitrary order:
 *   bb_count_kmer_alternatives  key = reference k-mer (2k bits from bit 63 down, A C G T = 0 1 2 3) | read k-mer length
 *                               (6 bits) | read k-mer (2 bits a base from bit 0 up); counts_out: one per key;
 *                               first_out = (alignment << 32) | reference offset of the window.  k <= 12.
 *   bb_count_cigar_qscores      key = CIGAR length (6 bits from bit 63 down) | symbols (2 bits each from bit 0 up, = X I D =
 *                               0 1 2 3, runs of 'D' cut to max_del); counts_out: 94 per key (quality 0 .. 93 of the window's
 *                               middle base); first_out = (alignment << 36) | ((window size - 1) / 2 << 32) | read offset;
 *                               overall_out[94]: the qualities of all bases.  Odd k <= 13: every odd size up to k is counted.
 * Windows that do not fit a key (and quality characters outside '!' .. '~') are not counted but listed: alignment, offset and
 * window size (negative for a bad quality character) in ovf_*; the caller evaluates those itself.  table_cap (a power of two)
 * slots are used on the device and bound the number of distinct keys; BB_ERR_CAPACITY if the table or the overflow list is
 * too small (*n_ovf then holds the required overflow capacity).  bb_model_error() describes the last failure of the calling
 * thread.  No bb_ctx is involved: the calls allocate and release what they need on `device`. */
BB_API int bb_count_kmer_alternatives(int device, int k, int32_t n_aln, const uint8_t *read, const int64_t *read_off,
                               const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops, const int32_t *op_read0,
                               const int32_t *op_ref0, const int64_t *ops_off, int64_t table_cap, uint64_t *keys_out,
                               uint64_t *first_out, uint32_t *counts_out, int64_t *n_entries, int64_t ovf_cap,
                               int32_t *ovf_aln, int32_t *ovf_pos, int32_t *ovf_k, int64_t *n_ovf);
BB_API int bb_count_cigar_qscores(int device, int k, int max_del, int32_t n_aln, const uint8_t *read, const uint8_t *qual,
                           const int64_t *read_off, const uint8_t *ref, const int64_t *ref_off, const uint32_t *ops,
                           const int32_t *op_read0, const int32_t *op_ref0, const int64_t *ops_off, int64_t table_cap,
                           uint64_t *keys_out, uint64_t *first_out, uint32_t *counts_out, int64_t *n_entries,
                           uint64_t *overall_out, int64_t ovf_cap, int32_t *ovf_aln, int32_t *ovf_pos, int32_t *ovf_k,
                           int64_t *n_ovf);
BB_API const char *bb_model_error(void);

#ifdef __cplusplus
}
#endif
#endif /* BADREAD_B200_H */
