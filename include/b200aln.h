/*
 * b200aln.h — C ABI of the B200-native `bwa aln` engine (libb200aln.so).
 *
 * Drop-in boundary for the one hot path of genome/ibwa: the gapped FM-index
 * search that turns reads into .sai suffix-array intervals.  Plain C types
 * only.  Every entry point names the reference interface it replaces
 * (file:line under the reference tree).
 *
 * Error convention (reference: utils.c:35-82): there are no return codes for
 * fatal conditions.  A CUDA failure, an unsupported option or an internal
 * overflow prints "[b200aln_*] <message> Abort!" to stderr and calls abort();
 * a partial result is never returned.  There is no CPU fallback.
 */
#ifndef B200ALN_H
#define B200ALN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == gap_opt_t (bwtaln.h:105-115): 16 x 4 bytes, also the .sai header (bwtaln.c:192). */
typedef struct {
    int32_t s_mm, s_gapo, s_gape;
    int32_t mode; /* BWA_MODE_* bits (bwtaln.h:93-101); bits 24-31 = barcode length */
    int32_t indel_end_skip, max_del_occ, max_entries;
    float fnr;
    int32_t max_diff, max_gapo, max_gape;
    int32_t max_seed_diff, seed_len;
    int32_t n_threads;
    int32_t max_top2;
    int32_t trim_qual;
} b200aln_opt_t;

/* == bwt_aln1_t (bwtaln.h:34-38): 16 bytes, the .sai record. */
typedef struct {
    uint32_t packed; /* n_mm:8 | n_gapo:8 | n_gape:8 | a:1 */
    uint32_t k, l;   /* SA interval */
    int32_t score;
} b200aln_rec_t;

/* A view of one in-memory FM-index exactly as bwt_restore_bwt() leaves it
 * (bwt.h:42-54, bwtio.c:51-70): `bwt` is the file payload after the 5-word
 * header, occ checkpoints interleaved every 128 bases. */
typedef struct {
    uint32_t primary;
    uint32_t L2[5];
    uint32_t seq_len;
    uint64_t bwt_size; /* number of 32-bit words in `bwt` */
    const uint32_t *bwt;
} b200aln_bwt_view_t;

typedef struct b200aln_ctx b200aln_ctx;

/* Library / build identification; usable without a GPU. */
const char *b200aln_version(void);

/* gap_init_opt (bwtaln.c:21-37): fill *opt with the reference defaults. */
void b200aln_opt_init(b200aln_opt_t *opt);

/* bwa_cal_maxdiff (bwtaln.c:39-51). */
int b200aln_cal_maxdiff(int len, double err, double thres);

/* Number of CUDA devices visible (0 when none / no driver). */
int b200aln_device_count(void);

/*
 * Replaces the two bwt_restore_bwt() results held by bwa_aln_core
 * (bwtaln.c:184-189): uploads both indexes to `device`, re-laid out as one
 * 32-byte occ block per 64 bases, and creates streams, pinned staging buffers
 * and search scratch.  The caller keeps ownership of the host arrays and may
 * free them after the call returns.
 */
b200aln_ctx *b200aln_open(const b200aln_bwt_view_t *bwt, const b200aln_bwt_view_t *rbwt, int device);

/* Convenience: bwt_restore_bwt(prefix.bwt / prefix.rbwt) + b200aln_open (bwtaln.c:184-189). */
b200aln_ctx *b200aln_open_prefix(const char *prefix, int device);

/* A sibling context on the same GPU that shares ctx's device index but has its own stream, pinned
 * staging and scratch: two batches can then be in flight at once (one host thread per context), so the
 * H2D / D2H copies and host work of one overlap the kernels of the other — the double buffering that
 * replaces the reference's per-batch pthread fan-out (bwtaln.c:199-218).  Close clones before `ctx`. */
b200aln_ctx *b200aln_clone(b200aln_ctx *ctx);

/* bwt_destroy x2 (bwtaln.c:238) plus device teardown. */
void b200aln_close(b200aln_ctx *ctx);

/*
 * Replaces bwa_cal_sa_reg_gap (bwtaln.c:80-140, declared bwtaln.h:148) for one
 * reference batch, on packed HOST buffers.  Read r is codes[offs[r] ..
 * offs[r]+lens[r]) in sequencing orientation, nt4 codes (0-3 = ACGT, >3 =
 * ambiguous; bntseq.c:39-56) after trimming; the reversed / reverse-
 * complemented forms the reference keeps (bwaseqio.c:189-192) are derived on
 * the device.  Batch-level semantics (max_len -> max_diff -> max_gapo clamp,
 * bwtaln.c:89-92) apply to exactly the reads of this call.
 *
 * n_aln[r] receives the record count of read r.  The records of all reads, in
 * read order, are returned in a buffer owned by the context, valid until the
 * next batch call on it; *total receives their number.  Host<->device copies
 * happen inside the call (pinned staging, asynchronous streams).
 */
const b200aln_rec_t *b200aln_batch(b200aln_ctx *ctx, int n_reads, const int32_t *lens, const int64_t *offs,
                                   const uint8_t *codes, const b200aln_opt_t *opt, int32_t *n_aln, int64_t *total);

/*
 * Same operator with inputs already resident on the device (d_lens, d_offs,
 * d_codes are device pointers on ctx's device; max_len and codes_bytes describe
 * them) and results left on the device: *d_n_aln (int32[n_reads]) and
 * *d_recs (b200aln_rec_t[*total]) point into context-owned device memory valid
 * until the next batch call.  Used to time the kernels without PCIe traffic.
 */
void b200aln_batch_device(b200aln_ctx *ctx, int n_reads, int max_len, const int32_t *d_lens, const int64_t *d_offs,
                          const uint8_t *d_codes, const b200aln_opt_t *opt, const int32_t **d_n_aln,
                          const b200aln_rec_t **d_recs, int64_t *total);

/*
 * b200aln_batch with the result formatted on the device as the bytes
 * bwa_aln_core writes for the batch (the fwrite loop of bwtaln.c:227-231): per
 * read its n_aln (int32) followed by its records.  Returns the context's
 * page-locked output buffer (valid until the next batch call on it) and its
 * size in *n_bytes; the driver only has to write() it.
 */
const void *b200aln_batch_sai(b200aln_ctx *ctx, int n_reads, const int32_t *lens, const int64_t *offs,
                              const uint8_t *codes, const b200aln_opt_t *opt, int64_t *n_bytes);

/*
 * Page-locks caller memory (cudaHostRegister) / releases it.  The batch calls
 * copy straight from page-locked input arrays; pageable ones go through the
 * context's staging buffer first.  b200aln_pin returns 0, or -1 when the range
 * cannot be locked (the memory then simply stays pageable).
 */
int b200aln_pin(void *p, size_t bytes);
void b200aln_unpin(void *p);

/*
 * Allocates, ahead of time, the per-batch device and staging buffers of n_contexts contexts that are going to be
 * opened or cloned on `device`, sized for calls of up to n_reads reads of up to max_len bases with the default
 * knobs.  A context of that device takes one such set with its first batch (waiting for one that is still being
 * allocated), so that batch does not start with 15+ GB of allocations of its own; a driver calls this on a thread
 * of its own while it is still reading the index files.  Sets that no context picks up stay allocated until
 * b200aln_prealloc_release.
 */
void b200aln_prealloc(int device, int n_contexts, int n_reads, int max_len);
/* Frees the sets of `device` (< 0: of every device) that no context has taken. */
void b200aln_prealloc_release(int device);

/* Per-call counters of the last batch on this context (instrumentation). */
typedef struct {
    double ms_h2d, ms_width, ms_search, ms_compact, ms_d2h, ms_total; /* CUDA-event times */
    uint64_t kernel_launches; /* kernels of this library launched by the call */
    uint64_t overflow_reads;  /* reads re-run with the large per-read arena */
    uint64_t pops, occ_lookups; /* stack pops / 32-byte index sectors of the fast pass; filled when the knob `count` is 1 */
} b200aln_stats_t;
void b200aln_last_stats(const b200aln_ctx *ctx, b200aln_stats_t *out);

/* CUDA-event stopwatch on the stream the engine launches on: start records an
 * event; stop records another, waits for it and returns the milliseconds
 * between them.  bench.py brackets its timed region with these. */
void b200aln_timer_start(b200aln_ctx *ctx);
double b200aln_timer_stop(b200aln_ctx *ctx);

/* Tuning knobs (optional; call before the first batch).  Unknown keys abort.
 *   batch_max_len  > 0: this context processes a SHARD of a reference batch whose longest read has this
 *                  length; the batch-level max_gapo clamp (bwtaln.c:89-92) is then taken from it, so that
 *                  shards on several GPUs reproduce the single-batch result.  0 = the call is the batch.
 *   lut_k          levels of the path-k-mer interval table built at open (default 14 = max, 0 = off).
 *   search_blocks_per_sm, width_blocks_per_sm, arena_cap, arena_cap_mid, arena_cap_big, rec_cap, rec_cap_mid,
 *   rec_cap_big, mid_lanes, big_lanes:
 *                  launch geometry and per-lane capacities; arena capacities count 64-byte records.
 *   prep_rounds    pruned pops a lane may go through per warp iteration before the warp moves on (default 1).
 *   reserve_reads  size the per-batch device buffers for at least this many reads (drivers whose launches vary in size).
 *   count          1: the fast pass runs with its pop / sector counters (b200aln_stats_t pops, occ_lookups).
 *   chunk_reads, slots, chunk_reads_device
 *                  a host-buffer call of more than 1.5 x chunk_reads reads (default 4 Mi) is cut into chunks that run
 *                  on `slots` sibling contexts (default 3; 1 = never), so that copies and kernels of different
 *                  chunks overlap; device-resident calls are one launch unless chunk_reads_device is set.
 *   q16            1 (default): 16-bit width records in the fast pass whenever max_diff < 7 and max_seed_diff < 3.
 *   order          1 (default): the fast pass takes the reads by work class, longest searches first.
 *   susp, susp_calls, susp_min
 *                  parking of a draining launch's stragglers: |susp| = lanes per warp at or below
 *                  which a warp parks what it has left once the work queue is dry; > 0 always, < 0 (default -16)
 *                  only while at least susp_calls (4) batches are in flight on this device index, 0 never.
 *   search_block   lanes per block of the fast pass: 128 (default) or 32. */
void b200aln_set_int(b200aln_ctx *ctx, const char *key, int64_t value);

/*
 * Drop-in for the reference's batch seam on its own structures:
 *   void bwa_cal_sa_reg_gap(int tid, bwt_t *const bwt[2], int n_seqs,
 *                           bwa_seq_t *seqs, const gap_opt_t *opt)   (bwtaln.h:148)
 * `seqs` is the reference's bwa_seq_t array (bwtaln.h:72-104; sizeof == 176 on
 * LP64, see b200aln_seq_layout_check).  Contract kept: fills n_aln and aln
 * (malloc'ed, freed by bwa_free_read_seq, bwaseqio.c:218), zeroes sa/type/c1/c2
 * (bwtaln.c:114), frees and nulls name/seq/rseq/qual (bwtaln.c:134-135).  The
 * context replaces the bwt[2] argument; tid is ignored (called once per batch).
 */
void b200aln_cal_sa_reg_gap(b200aln_ctx *ctx, int n_seqs, void *seqs, const b200aln_opt_t *opt);

/* sizeof / offsets this library assumes for bwa_seq_t; INTEGRATION.md shows the
 * static asserts a reference-side binding adds. */
typedef struct {
    size_t size, off_name, off_seq, off_rseq, off_qual, off_lenword, off_n_aln, off_aln, off_sa, off_c1c2;
} b200aln_seq_layout_t;
void b200aln_seq_layout(b200aln_seq_layout_t *out);

/*
 * Replaces bwa_aln_core (bwtaln.c:173-241, declared bwtaln.h:138): reads
 * `fn_fa` (FASTA/FASTQ, plain or gzip, "-" = stdin), writes the 64-byte header
 * and the per-read records to `out_fd`.  The reference's batches of 0x40000
 * reads (bwtaln.c:193) keep their meaning — the max_gapo clamp of a batch
 * comes from its longest read (bwtaln.c:89-92) — but consecutive batches that
 * agree on it go to a GPU as one launch (B200ALN_MERGE, default 8), several
 * launches are in flight per GPU (B200ALN_INFLIGHT, default 4), parsing runs
 * ahead in page-locked arrays and the .sai bytes are formatted on the device
 * (INTEGRATION.md section 1).  `device` < 0 uses every visible GPU with the
 * index replicated, launches going round the GPUs.  Returns the number of
 * reads processed.
 */
int64_t b200aln_aln_core(const char *prefix, const char *fn_fa, const b200aln_opt_t *opt, int out_fd, int device);

/*
 * Replaces bwa_open_reads / bwa_read_seq / bwa_seq_close (bwtaln.c:159-171, bwaseqio.c:21-52,89-208): BAM with
 * -b (filtered by -0/-1/-2, bwaseqio.c:89-141) or FASTA / FASTQ input,
 * plain or gzip ("-" = stdin), with the reference parser's behaviour (kseq.h:150-194): multi-line records,
 * name = first token, only graphic characters enter the sequence, reading stops at the first truncated
 * record; -q trimming (bwaseqio.c:74-87), -B barcode strip and -I quality offset applied like the reference.
 * b200aln_reader_next returns the number of reads (0 at the end) and views, valid until the next call, of
 * the packed batch that b200aln_batch takes: lens (after trimming), offs and nt4 codes (full reads).
 */
typedef struct b200aln_reader b200aln_reader;
b200aln_reader *b200aln_reader_open(const char *fn, int mode); /* mode & BWA_MODE_BAM (0x20): BAM input, -0/-1/-2 bits */
int b200aln_reader_next(b200aln_reader *r, int n_needed, int mode, int trim_qual, const int32_t **lens,
                        const int64_t **offs, const uint8_t **codes, int64_t *codes_bytes);
void b200aln_reader_close(b200aln_reader *r);

/* Creates the CUDA context of a device (the first CUDA call of a process takes a few hundred ms); the
 * CLI driver calls it on a side thread while it reads the index files.  No reference counterpart. */
void b200aln_warm_device(int device);

/* Replaces bwa_aln (bwtaln.c:243-328): same getopt string and semantics. */
int b200aln_aln_main(int argc, char *argv[]);

/*
 * Scope row N2 (next after the aln path): suffix-array row -> text position, the random-access step of
 * samse / sampe (bwa_cal_pac_pos, bwase.c:136-161) on the same device index.
 *   b200aln_sa_load   bwt_restore_sa (bwtio.c:29-49): `sa` is the array as restored (sa[0] = 0xffffffff,
 *                     sa[j] = SA(j * sa_intv)), which = 0 for <prefix>.sa (with .bwt), 1 for .rsa (with .rbwt)
 *   b200aln_bwt_sa    bwt_sa (bwt.c:69-79) for n rows at once (host buffers)
 *   b200aln_sa2seq    bwtdb_sa2seq with offset 0 (dbset.c:240-245): strand != 0 -> bwt_sa(bwt, row),
 *                     else rbwt.seq_len - (bwt_sa(rbwt, row) + len)
 */
typedef struct {
    uint32_t primary;
    uint32_t seq_len;
    int32_t sa_intv;
    uint64_t n_sa;
    const uint32_t *sa;
} b200aln_sa_view_t;
void b200aln_sa_load(b200aln_ctx *ctx, int which, const b200aln_sa_view_t *sa);
void b200aln_bwt_sa(b200aln_ctx *ctx, int which, int64_t n, const uint32_t *rows, uint32_t *pos);
void b200aln_sa2seq(b200aln_ctx *ctx, int64_t n, const uint8_t *strand, const uint32_t *rows, const int32_t *lens,
                    uint64_t *pos);

/*
 * Scope row N4: the per-read merge of several .sai streams (primary + alt indexes), i.e. alngrp_create
 * (saiset.c:45-78) for a whole batch: per read the alignments of stream 0, 1, ... in that order; with more
 * than one stream they are sorted by score exactly like the reference's ks_introsort(alignment)
 * (ksort.h:172-224 — not stable, equal scores come out in the reference's order) and cut at the first
 * score > best + s_mm.
 *   n_aln[s][r], recs[s]   stream s as b200aln_batch returns it (records packed in read order), host buffers
 *   out_off[r]             first slot of read r's group = sum of the unmerged group sizes of reads < r
 *   out_n[r]               size of the merged group;  out_recs / out_dbidx: records and the stream each came from
 * out_recs / out_dbidx must hold the sum of all counts, which is also the return value.
 */
int64_t b200aln_alngrp_merge(b200aln_ctx *ctx, int n_streams, int n_reads, const int32_t *const *n_aln,
                             const b200aln_rec_t *const *recs, int s_mm, int64_t *out_off, int32_t *out_n,
                             b200aln_rec_t *out_recs, uint32_t *out_dbidx);

/* Random 32-byte-sector gather micro-benchmark over the device index (the
 * roofline denominator of SURVEY.md §8d): n_loads independent uniformly random
 * sector reads; returns GB/s (sectors * 32 B / CUDA-event time). */
double b200aln_sector_roofline(b200aln_ctx *ctx, uint64_t n_loads, int repeats);

#ifdef __cplusplus
}
#endif
#endif /* B200ALN_H */
