/*
 * include/ambivert_align.h -- C ABI of the B200 (sm_100a) alignment backend for AmBiVErT.
 *
 * The shared object built from ambivert_b200/csrc (align_c.b200.so) is a drop-in for the
 * reference's `align_c` extension (setup.py:44-50; loaded by ambivert/align/align_ctypes.py:30-35,
 * which takes the first file in the package directory whose name contains "align_c.").
 *
 * Part 1 re-exports, symbol for symbol and struct for struct, the interface of the reference's
 * lib/align/align.h:29-125 (what align_ctypes.py:60-101 binds).  Part 2 is the batched extension
 * that the pipeline host calls once per stage instead of once per pair.
 *
 * Every entry point computes on the GPU.  There is no CPU fallback: without a usable CUDA device
 * the legacy symbols return NULL and the batched ones return ABV_ERR_NO_DEVICE.
 *
 * All sequence arguments are length-delimited, borrowed for the duration of the call and never
 * modified.  All arithmetic is 32-bit signed integer; results are bit-exact with the reference.
 */
#ifndef AMBIVERT_ALIGN_H
#define AMBIVERT_ALIGN_H

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only this interface is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* ============================ Part 1: the reference's own ABI ================================ */

/* lib/align/align.h:29 */
typedef enum { A_GAP = 0, B_GAP = 1, MATCH = 2 } FragType;

/* lib/align/align.h:31-35  (24 bytes on LP64) */
typedef struct AlignFrag {
  struct AlignFrag *next;
  FragType type;
  int sa_start, sb_start, hsp_len;
} AlignFrag;

/* lib/align/align.h:37-41  (16 bytes on LP64) */
typedef struct Alignment {
  AlignFrag *align_frag;
  int frag_count;
  int score;
} Alignment;

/* lib/align/align.h:43-46, alignment.c:22-49.  Callee-allocated (malloc) lists; the caller frees
 * each Alignment exactly once with alignment_free (ambivert.py:113, :154, :176). */
int align_frag_count(AlignFrag *f);
void align_frag_free(AlignFrag *f);
void alignment_free(Alignment *alignment);
Alignment *alignment_new(AlignFrag *align_frag, int score);

/* The eight aligners.  `*_raw` take sequences already encoded as symbol indices
 * (0 <= code < alpha_len); the others take bytes and a 256-entry byte->index map
 * (helpers.h:26-32 to_raw).  score_matrix is alpha_len*alpha_len ints, row = sa symbol.
 * gap_open/gap_extend are <= 0; a gap of length L costs gap_open + (L-1)*gap_extend.
 * Return NULL on sa_len<=0, sb_len<=0, NULL pointers, gap_open>0, gap_extend>0, allocation
 * failure (global_align.c:118-132), where the reference has undefined behaviour
 * (glocal with sb_len==1, glocal_align.c:138+172; overlap with sa_len==1, align.c:183), or when
 * no CUDA device is usable.
 *
 *   align_raw / align                 ends-free overlap      lib/align/align.c:68-243, :246-274
 *   local_align_raw / local_align     Smith-Waterman/Gotoh   lib/align/local_align.c:75-244, :247-275
 *   global_align_raw / global_align   Needleman-Wunsch/Gotoh lib/align/global_align.c:95-228, :231-260
 *   glocal_align_raw / glocal_align   sb global, sa local    lib/align/glocal_align.c:73-202, :205-234
 */
Alignment *align_raw(const unsigned char *sa, int sa_len, const unsigned char *sb, int sb_len,
                     int alpha_len, int *score_matrix, int gap_open, int gap_extend);
Alignment *align(const char *seqa, int sa_len, const char *seqb, int sb_len, int alpha_len,
                 const unsigned char *map, int *score_matrix, int gap_open, int gap_extend);
Alignment *local_align_raw(const unsigned char *sa, int sa_len, const unsigned char *sb, int sb_len,
                           int alpha_len, int *score_matrix, int gap_open, int gap_extend);
Alignment *local_align(const char *seqa, int sa_len, const char *seqb, int sb_len, int alpha_len,
                       const unsigned char *map, int *score_matrix, int gap_open, int gap_extend);
Alignment *global_align_raw(const unsigned char *sa, int sa_len, const unsigned char *sb, int sb_len,
                            int alpha_len, int *score_matrix, int gap_open, int gap_extend);
Alignment *global_align(const char *seqa, int sa_len, const char *seqb, int sb_len, int alpha_len,
                        const unsigned char *map, int *score_matrix, int gap_open, int gap_extend);
Alignment *glocal_align_raw(const unsigned char *sa, int sa_len, const unsigned char *sb, int sb_len,
                            int alpha_len, int *score_matrix, int gap_open, int gap_extend);
Alignment *glocal_align(const char *seqa, int sa_len, const char *seqb, int sb_len, int alpha_len,
                        const unsigned char *map, int *score_matrix, int gap_open, int gap_extend);

/* ============================ Part 2: batched extension ====================================== */

/* alignment modes (the four aligners above) */
#define ABV_OVERLAP 0
#define ABV_LOCAL   1
#define ABV_GLOBAL  2
#define ABV_GLOCAL  3

/* status codes (0 = success) */
#define ABV_OK              0
#define ABV_ERR_ARGS      (-1)  /* what makes the reference return NULL (global_align.c:118-122) */
#define ABV_ERR_UNDEFINED (-2)  /* a pair on which the reference has undefined behaviour */
#define ABV_ERR_CUDA      (-3)  /* CUDA runtime error; text in abv_last_error() */
#define ABV_ERR_NOMEM     (-4)
#define ABV_ERR_FRAG_CAP  (-5)  /* fragment pool too small; *frag_total holds the needed size */
#define ABV_ERR_NO_DEVICE (-6)

/* kernel selection for the many-to-many scorer (abv_set_option("path", ...)) */
#define ABV_PATH_AUTO    0      /* packed 16-bit kernel when provably exact, else the 32-bit one */
#define ABV_PATH_WF32    1      /* force the general 32-bit wavefront kernel */
#define ABV_PATH_PACKED  2      /* force the packed kernel; ABV_ERR_ARGS if it is not applicable */

/* one alignment fragment, the payload of AlignFrag without the link (align.h:31-35) */
typedef struct abv_frag { int type, sa_start, sb_start, hsp_len; } abv_frag;

typedef struct abv_job abv_job;   /* opaque: a many-to-many job whose inputs are resident in HBM */

/* device management / diagnostics */
int abv_device_count(void);               /* number of usable CUDA devices (0 if none) */
int abv_set_device(int ordinal);          /* the process's default device (one process per GPU under torchrun) */
int abv_set_thread_device(int ordinal);   /* device of the CALLING THREAD's subsequent calls; -1 = the process default.  Every
                                             device has its own context (streams, allocation cache, lock): host threads working
                                             on different devices run concurrently */
int abv_get_device(void);                 /* the calling thread's device */
const char *abv_last_error(void);         /* message of the last failing call of this thread */
const char *abv_version(void);
/* tuning / test options (defaults are what the measurements use):
 *   "path"               ABV_PATH_*            kernel selection of the many-to-many scorer
 *   "split"              0 = automatic         slices of the target pairs per query (short tail on multi-GPU shards)
 *   "fine_tail"          1                     the last work items of a packed launch are cut into finer slices (the tail of the
 *                                              persistent grid is one short item long); 0: uniform slices
 *   "tb16"               1                     0: one-to-one jobs always take the 32-bit kernel
 *   "global_v1"          0                     1: first-generation packed kernel for global mode
 *   "no_levels"          0                     1: streaming global kernel resets every pair explicitly (no level staircase)
 *   "no_border"          0                     1: the streaming kernels never take the border-lane instantiation (every query keeps 32 lanes)
 *   "no_gop_imm"         0                     1: never the instantiation with the gap addend as an immediate
 *   "concurrent_buckets" 1                     the packed launches of a job run on streams of their own, chained by tail flags; 0: one stream
 *   "split_min_items"    4                     a launch with fewer work items per resident CTA cuts every query into slices of the target pairs
 *   "tb16_extra_smem"    0                     experiment knob: unused shared memory per CTA of the packed traceback kernel (lowers occupancy)
 *   "pairs_chunk"        200000                one-to-one jobs of >= 2x this many pairs run as a pipeline of chunks; 0: never
 *   "wf32_blocks_per_sm", "mm16_blocks_per_sm", "cache_mb" (device allocation cache) */
int abv_set_option(const char *key, int value);

/*
 * Many-to-many best hit: what ambivert.py:461-485 / :487-511 compute with one ctypes call per
 * (query, target) pair and a Python arg-max.  For every query i the score of `mode` against every
 * target j (query = sa, target = sb) is computed and reduced with the reference's selection rule
 * (strict '>' in target order starting from seed_score: the first maximum wins,
 * ambivert.py:478-485): best_idx[i] = -1 and best_score[i] = seed_score when no target beats the
 * seed.
 *
 * Sequences are concatenated: sequence i is bytes [off[i], off[i+1]) of the buffer.
 * map256 != NULL: buffers hold characters, encoded on the device through map256 (to_raw);
 * map256 == NULL: buffers already hold symbol indices (< alpha_len).
 * A symbol outside the alphabet (a byte whose map entry, or a raw code, is >= alpha_len) makes every batched call
 * return ABV_ERR_ARGS: the reference reads past its score matrix there (helpers.h:26-32, global_align.c:148).  The
 * legacy single-pair symbols of part 1 keep the reference's signature (no status) and score such a symbol as the last
 * symbol of the alphabet.
 */
int abv_best_hits(int mode,
                  const char *q, const long long *q_off, int n_q,
                  const char *t, const long long *t_off, int n_t,
                  int alpha_len, const unsigned char *map256, const int *score_matrix,
                  int gap_open, int gap_extend, int seed_score,
                  int *best_score, int *best_idx);

/*
 * The same job on n_devices GPUs of one box from ONE process -- the successor of the reference's --threads
 * (ambivert.py:474-476, :1092-1095).  The queries are cut into n_devices contiguous shards balanced by
 * summed length (abv_shard_bounds), every device gets all targets, one host thread per device scores its shard, and each
 * shard's (best_score, best_idx) go from its device straight into the caller's arrays.  devices = NULL: ordinals
 * 0 .. n_devices-1; an ordinal may repeat (shards then queue on that device).  Results are identical to abv_best_hits.
 * (Under torchrun, one process per GPU, the shards are gathered by one NCCL collective instead: ambivert_b200/dist.py.)
 */
int abv_best_hits_multi(int mode,
                        const char *q, const long long *q_off, int n_q,
                        const char *t, const long long *t_off, int n_t,
                        int alpha_len, const unsigned char *map256, const int *score_matrix,
                        int gap_open, int gap_extend, int seed_score,
                        const int *devices, int n_devices, int *best_score, int *best_idx);
/* shard r of n_shards owns queries [bounds[r], bounds[r+1]); bounds has n_shards + 1 entries (host logic, no device needed) */
void abv_shard_bounds(const long long *q_off, int n_q, int n_shards, int *bounds);

/* Full n_q x n_t score matrix (row-major, scores[i*n_t + j]); same arguments. */
int abv_score_matrix(int mode,
                     const char *q, const long long *q_off, int n_q,
                     const char *t, const long long *t_off, int n_t,
                     int alpha_len, const unsigned char *map256, const int *score_matrix,
                     int gap_open, int gap_extend, int *scores);

/*
 * The same job split into stages, so that a caller can keep the inputs resident in HBM, launch
 * on its own stream and hand the per-query results to a collective (one NCCL gather of 8 bytes per
 * query) without a host round trip.
 *   prepare : validate, upload and pre-process (encode, length-bucket, pair the targets)
 *   launch  : enqueue the kernels on `stream` (a cudaStream_t; NULL = the library's stream).
 *             d_best_score/d_best_idx are DEVICE pointers to n_q ints each, or NULL to use the
 *             job's own result buffers.  Asynchronous.
 *   fetch   : wait for the last launch and copy the job's own result buffers to the host
 *   cells   : sum over all pairs of len(query)*len(target)
 */
abv_job *abv_best_hits_prepare(int mode,
                               const char *q, const long long *q_off, int n_q,
                               const char *t, const long long *t_off, int n_t,
                               int alpha_len, const unsigned char *map256, const int *score_matrix,
                               int gap_open, int gap_extend, int seed_score);
int abv_best_hits_launch(abv_job *job, void *stream, int *d_best_score, int *d_best_idx);
int abv_best_hits_fetch(abv_job *job, int *best_score, int *best_idx);
double abv_job_cells(const abv_job *job);
int abv_job_kernel_launches(const abv_job *job);   /* kernels one abv_best_hits_launch enqueues */
/* Per packed-kernel launch of the LAST abv_best_hits_launch (call after it completed): rows of
 * 4 doubles {K = packed columns per lane (K + 0.5: the bucket runs the border-lane instantiation of the streaming global
 * kernel, 31 lanes x K columns), queries, cells, milliseconds by CUDA events on the launch's stream}.  Returns the number of rows written (<= max_rows). */
int abv_job_profile(abv_job *job, double *out, int max_rows);
/* The same for the last `last_launches` calls of abv_best_hits_launch (the library keeps the event pairs of the last 32), so
 * that many launches can be timed back to back without a host synchronisation between them: rows of 5 doubles
 * {K, queries, cells, milliseconds, launch number (0 = the job's first launch)}, oldest launch first.  Call after the last
 * launch completed.  Returns the number of rows written (<= max_rows). */
int abv_job_profile_history(abv_job *job, int last_launches, double *out, int max_rows);
void abv_job_free(abv_job *job);

/*
 * One-to-one alignments with traceback: pair k aligns a[k] (sa) with b[k] (sb); what
 * ambivert.py:75-155 obtains one ctypes call at a time in merge_overlaps (:320-328) and
 * align_to_reference (:556-573).
 *   scores[k]      the alignment score (Alignment.score)
 *   frag_start[k]  index of pair k's first fragment in frags
 *   frag_count[k]  number of fragments (Alignment.frag_count); fragments are in sequence order
 *   frags          caller-allocated pool of frag_cap entries; *frag_total receives the number
 *                  of entries needed (ABV_ERR_FRAG_CAP if that exceeds frag_cap; scores and
 *                  frag_count are valid even then)
 */
int abv_align_pairs(int mode,
                    const char *a, const long long *a_off,
                    const char *b, const long long *b_off, int n,
                    int alpha_len, const unsigned char *map256, const int *score_matrix,
                    int gap_open, int gap_extend,
                    int *scores, long long *frag_start, int *frag_count,
                    abv_frag *frags, long long frag_cap, long long *frag_total);

/*
 * The string work that follows the aligner in the pipeline, done on the device as well
 * (SURVEY.md section 8f, row f1).  Both take CHARACTER input (map256 != NULL); outputs are concatenated:
 * pair k owns bytes [out_off[k], out_off[k+1]).  status[k]: 1 ok; 0 merge skipped because the aligned
 * length is below minimum_overlap (ambivert.py:324-325, "unmergable"); -1 the alignment is empty (the
 * reference raises ValueError on the NULL fragment, ambivert.py:93); -2 a mismatching base pair outside
 * ACGTN (encode_ambiguous raises KeyError, sequence_utilities.py:152-156).
 * ABV_ERR_FRAG_CAP with *out_total = needed bytes when out_cap is too small (offsets/status are valid).
 *
 * abv_gapped_pairs: the two gapped rows that smith_waterman / needleman_wunsch return
 *   (ambivert.py:95-111, :136-152): '-' in the row that has the gap, the other row's characters as given
 *   (case preserved); start_a/start_b = first fragment's starts (what smith_waterman reports).
 * abv_merge_pairs: local alignment of every forward read with its (already reverse-complemented) reverse
 *   read and the merged read of merge_overlaps (ambivert.py:327):
 *   fwd[:fwd_start] + flatten_paired_alignment(rows) + rev[rev_start + aligned rev bases:]
 *   where flatten_paired_alignment / encode_ambiguous are sequence_utilities.py:130-169.
 */
int abv_gapped_pairs(int mode,
                     const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                     int alpha_len, const unsigned char *map256, const int *score_matrix,
                     int gap_open, int gap_extend,
                     int *scores, int *status, int *start_a, int *start_b,
                     long long *out_off, char *gapped_a, char *gapped_b, long long out_cap, long long *out_total);
int abv_merge_pairs(const char *fwd, const long long *fwd_off, const char *rev, const long long *rev_off, int n,
                    int alpha_len, const unsigned char *map256, const int *score_matrix,
                    int gap_open, int gap_extend, int minimum_overlap,
                    int *scores, int *status, long long *merged_off, char *merged, long long merged_cap,
                    long long *merged_total);

/*
 * The step before the path: the clustering of read pairs.  The key of a pair is
 * md5(f_seq[trim5:trim3] + r_seq[trim5:trim3]) (AmpliconData.add_reads, ambivert.py:244-262; MD5 of RFC 1321,
 * Python slice semantics: has_trim5/has_trim3 = 0 means None, trim3 is the already negated value the reference
 * stores, ambivert.py:224); pairs with equal keys form a cluster.  Clusters are numbered in order of first
 * occurrence (the insertion order of AmpliconData.data).  Outputs: cluster_of[n]; *n_clusters; for cluster k its raw
 * 16-byte digest (hexdigest = hex of it) in digest[16k..], the number of pairs count[k] (what get_above_threshold
 * compares, ambivert.py:291-299) and the index of its first pair first[k].  digest/count/first need room for n.
 */
int abv_cluster_pairs(const char *fwd, const long long *fwd_off, const char *rev, const long long *rev_off, int n,
                      int has_trim5, int trim5, int has_trim3, int trim3,
                      int *cluster_of, int *n_clusters, unsigned char *digest, int *count, int *first);

/*
 * The step after the path: the variant scan of aligned row pairs, replacing the
 * per-site Python loop of caller() in ambivert/call_variants.py:32-128.  Pair k is the gapped sample row
 * sample[sample_off[k] .. sample_off[k+1]) and the gapped reference row ref[ref_off[k] .. ref_off[k+1]) (what
 * align_to_reference stores in AmpliconData.aligned, ambivert.py:556-573; abv_gapped_pairs produces them).
 * softmask != 0: lower-case reference is primer sequence (the reference's default); gap: the gap character ('-').
 * A record is what the reference yields, with the accumulated bases named by their place in the rows:
 *   kind       ABV_VAR_X substitution, ABV_VAR_D deletion, ABV_VAR_I insertion
 *   start      "Start": zero-based position in the gapped rows, as the reference computes it
 *   pos        "Pos":   one-based position in the ungapped reference, as the reference computes it
 *   len, run_start      "Bases" = row[run_start .. run_start+len) of the sample row (X, I) / reference row (D)
 * status[k]: 0 ok; ABV_VAR_ERR_LENGTH rows of different length (RuntimeError, call_variants.py:54-55);
 * ABV_VAR_ERR_INDEX the reference's look-ahead over a gap run leaves the row (IndexError, call_variants.py:66).
 * rec_start[k], rec_count[k]: pair k's records in recs (caller-allocated, rec_cap entries); *rec_total = entries
 * needed (ABV_ERR_FRAG_CAP if that exceeds rec_cap; status, rec_start and rec_count are valid even then).
 */
#define ABV_VAR_X 0
#define ABV_VAR_D 1
#define ABV_VAR_I 2
#define ABV_VAR_ERR_LENGTH (-1)
#define ABV_VAR_ERR_INDEX  (-2)
typedef struct abv_variant { int kind, start, pos, len, run_start; } abv_variant;
int abv_call_variants(const char *sample, const long long *sample_off, const char *ref, const long long *ref_off, int n,
                      int softmask, int gap,
                      int *status, long long *rec_start, int *rec_count,
                      abv_variant *recs, long long rec_cap, long long *rec_total);

/*
 * Page-locked host memory for buffers handed to the batched calls (cudaHostAlloc): copies from / to such buffers run at
 * full PCIe / NVLink-C2C speed and overlap the host-side planning.  Any host memory is accepted by every call; pinned
 * buffers are only faster.  abv_pinned_alloc returns NULL on failure or without a usable CUDA device.
 */
void *abv_pinned_alloc(unsigned long long bytes);
void abv_pinned_free(void *p);

/*
 * Timings of the last batched call of this thread, CUDA events on the library's stream:
 *   out[0] host->device ms, out[1] kernel ms, out[2] device->host ms, out[3] cells,
 *   out[4] kernel launches, out[5] bytes host->device, out[6] bytes device->host,
 *   out[7] ms of the dominant DP kernel alone
 * Returns the number of values written (<= n).
 */
int abv_get_timing(double *out, int n);

/*
 * Integer-pipe microbenchmark, the measured denominator of the GCUPS roofline:
 * runs a register-only dependent-chain kernel of the instruction mix `which` on every SM and
 * reports lane-operations per second.
 *   which 0: IADD3 (32-bit add)            1: VIMNMX3.S16x2        2: VIADDMNMX.S16x2
 *         3: VIADD.16x2                    4: the packed DP cell mix (2 add + 1 max3 + 2 addmax)
 *         5: the 32-bit DP cell mix        6: VIMNMX3 (32-bit)     7: 32-bit add + 16x2 max mix
 *         8: the streaming kernel's cell: 1 add + VIMNMX3.S16x2 + 2 VIADDMNMX.S16x2 (4 lane-ops per cell pair)
 */
int abv_int_peak(int which, int iters, double *lane_ops_per_s, double *sm_clock_mhz);
/* diagnostics (host logic only, no device needed): the chunk boundaries the one-to-one jobs would use for n pairs
 * ("pairs_chunk" option); writes min(count, cap) entries, returns the count (chunks + 1; 2 = one piece) */
int abv_pairs_chunk_bounds(int n, int *bounds, int cap);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* AMBIVERT_ALIGN_H */
