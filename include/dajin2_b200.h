/*
 * dajin2_b200 — C ABI of the B200-native read-level numeric core for DAJIN2.
 *
 * The reference (akikuno/DAJIN2 v0.9.3) is pure Python and has no FFI: the seams this library plugs into are
 * plain Python functions.  Each entry point below names the reference function(s) whose
 * read-scale loop it replaces (paths relative to src/DAJIN2/).  The Python shim in dajin2_b200/ binds these
 * with ctypes and keeps the reference's signatures; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - Every pointer is a DEVICE pointer unless the parameter name starts with h_.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on that
 *     stream unless stated otherwise; buffers are owned by the caller.
 *   - Return value: 0 on success, non-zero on failure; dj_last_error() returns a thread-local message.
 *   - There is no CPU fallback anywhere in this library.
 *
 * Read x locus code matrix (the encoding every later stage consumes), one byte per MIDSV token:
 *     bit 7      token is an insertion "+X|+Y|...|<anchor>"; the low bits then describe the anchor
 *     bits 6..0  anchor id a:
 *                  0..9    "=B"   a = 5*lower + b          b: A=0 C=1 G=2 T=3 N=4
 *                  10..19  "-B"   a = 10 + 5*lower + b
 *                  20..69  "*RB"  a = 20 + 25*lower + 5*r + b   (ref base r, read base b; same case)
 *                  127     token outside the supported grammar (flag bit DJ_FLAG_BAD_TOKEN is raised)
 * "lower" marks an all-lowercase token (inside an inversion).  An insertion takes the case of its anchor: midsv
 * lowercases whole tokens, so a token whose inserted bases and anchor differ in case does not occur and is not
 * looked for.
 */
#ifndef DAJIN2_B200_H
#define DAJIN2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DJ_ABI_VERSION 2

#define DJ_CODE_INVALID 0x7F
#define DJ_CODE_INS 0x80

/* per-read flag bits written by dj_encode */
#define DJ_FLAG_TOKEN_COUNT 1u /* number of tokens != n_loci */
#define DJ_FLAG_BAD_TOKEN 2u   /* a token outside the supported grammar */

/* bytes of text one trip of the encoder's main loop covers; checkpoint granularity of the token index */
#define DJ_ENC_CHUNK 1536

/* rows of dj_hist output: counts[DJ_HIST_ROWS][n_loci] (int32) */
#define DJ_HIST_MATCH 0   /* '=' with the N / lowercase skip rule   (indel_counter.py:15-31)            */
#define DJ_HIST_INS 1     /* '+'  "                                                                    */
#define DJ_HIST_DEL 2     /* '-'  "                                                                    */
#define DJ_HIST_SUB 3     /* '*'  "                                                                    */
#define DJ_HIST_SW_INS_P 4 /* startswith('+'), '+' strand, no skip rule (error_correction/strand_bias_handler.py:18-33) */
#define DJ_HIST_SW_DEL_P 5
#define DJ_HIST_SW_SUB_P 6
#define DJ_HIST_SW_INS_M 7 /* same, '-' strand */
#define DJ_HIST_SW_DEL_M 8
#define DJ_HIST_SW_SUB_M 9
#define DJ_HIST_ROWS 10

/* bits of the per-locus mask describing mutation_loci[i] (list[set[str]] in the reference) */
#define DJ_MASK_INS 1
#define DJ_MASK_DEL 2
#define DJ_MASK_SUB 4

/* One distinct 3-mer observed at a selected locus (record exported by dj_kmer_export).
 * The reference's key is the string "prev,cur,next" (clustering/kmer_generator.py:35-56). */
typedef struct {
    uint32_t sel_index;     /* index into the `sel` array passed to dj_kmer_count                    */
    uint16_t prev;          /* 0..255 code byte, 256 = '@', 257 = raw (uncompressed) insertion token */
    uint16_t cur;           /* code byte of the centre token                                         */
    uint16_t next;          /* as prev                                                               */
    uint16_t reserved;
    uint32_t count_sample;  /* reads of file 0 carrying this 3-mer at this locus                     */
    uint32_t count_control; /* reads of file 1                                                       */
    uint32_t example;       /* (file << 31) | smallest read index seen: lets the host print raw tokens */
    uint32_t slot;          /* slot in the device table (index for slot_value_id in dj_annotate)     */
} DjKmerRecord;

/* One distinct score row found by dj_dedup_rows. */
typedef struct {
    uint32_t slot;       /* slot in the device table; dj_dedup_rows wrote it to slot_of_row[]   */
    uint32_t count;      /* number of rows equal to this one                                    */
    uint32_t count_plus; /* ... of which '+' strand                                             */
    uint32_t first_row;  /* smallest row position (in the order given to dj_dedup_rows)         */
    uint32_t last_row;   /* largest row position: decides scikit-learn's ties between points (dj_bisect) */
} DjRowRecord;

int dj_abi_version(void);
const char *dj_last_error(void);
/* h_out[0]=SM count, [1]=cc major, [2]=cc minor.  Fails when no CUDA device is usable. */
int dj_device_info(int32_t *h_out);

/* Row pitch (bytes) of the code matrix for n_loci tokens per read (multiple of 32). */
int32_t dj_code_pitch(int32_t n_loci);
/* Number of uint32 checkpoints per read for rows of at most max_row_len bytes. */
int32_t dj_ckpt_pitch(int32_t max_row_len);

/*
 * MIDSV text -> code matrix (+ per-read scalars).  Replaces the `json.loads` + `split(",")` every reference
 * pass starts with (utils/fileio.py:49-52) and computes, in the same pass,
 *   match_score[r]  = calc_match(MIDSV)                      (classification/classifier.py:11-24)
 *   n_tokens[r], flags[r]                                     (validation; the reference assumes n_loci tokens)
 *   ckpt[r][j]      = number of tokens that start before byte j*DJ_ENC_CHUNK - (row_off[r] & 15) of the row
 * Read r's text is text[row_off[r] .. row_off[r]+row_len[r]); 16 readable bytes must follow the last row.
 */
int dj_encode(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
              int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, uint8_t *codes, int32_t *match_score,
              int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt, void *stream);

/* The round-1 generic encoder (every 16-byte chunk through the comma slots, no checkpoints), kept as an on-device
 * cross-check of dj_encode: same codes, match_score, n_tokens and flags for any input. */
int dj_encode_v1(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                 int32_t n_loci, int32_t code_pitch, uint8_t *codes, int32_t *match_score, int32_t *n_tokens,
                 uint32_t *flags, void *stream);

/* dj_encode with the allele sequence as an accelerator: ref_seq[n_loci] (device, ASCII bases).  Stretches of plain
 * matches are verified against the reference rendered as "=X,=X,..." instead of being translated token by token;
 * everything else -- and every read whose text disagrees with the sequence -- takes the path of dj_encode.  Output
 * identical to dj_encode for any input; ref_seq == NULL (or a sequence too long for shared memory) calls dj_encode.
 * Same reference interface as dj_encode; the sequence is the `sequence` argument the reference passes next to the
 * file (summarize_indels(path, sequence), indel_counter.py:58; score_allele, classifier.py:27). */
int dj_encode_ref(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                  int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, const uint8_t *ref_seq, uint8_t *codes,
                  int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt, void *stream);
/* dynamic shared memory dj_encode_ref needs per CTA for this reference length */
int64_t dj_encode_ref_smem_bytes(int32_t n_loci, int32_t code_pitch);

/*
 * Per-locus histograms of the code matrix over `rows` (NULL = all n_rows reads in order).
 * counts[DJ_HIST_ROWS][n_loci] is zeroed and filled.  Replaces count_indels (indel_counter.py:15-31) and the
 * second full pass of count_strandless_of_indels (error_correction/strand_bias_handler.py:18-33).
 */
int dj_hist(const uint8_t *codes, int32_t code_pitch, const uint8_t *strand, const int32_t *rows,
            int64_t n_rows, int32_t n_loci, int32_t *counts, void *stream);

/*
 * 10-wide window cosine distances of is_dissimilar_loci (mutation_processing/anomaly_detector.py:28-58):
 * dist[i] = cosine_distance(sample[i:i+10], control[i:i+10]), dist_tail[i] = the same on [i+1:i+10].
 */
int dj_window_cosine(const double *sample, const double *control, int32_t n_loci, double *dist,
                     double *dist_tail, void *stream);

/* Bytes of device memory a 3-mer table with `capacity` slots (power of two) needs. */
int64_t dj_kmer_table_bytes(uint32_t capacity);
int dj_kmer_table_clear(void *table, uint32_t capacity, void *stream);

/*
 * Count the distinct 3-mers at the selected loci (generate_mutation_kmers + call_count,
 * clustering/kmer_generator.py:35-56, clustering/score_handler.py:15-16).  "@,@,@" is not stored: its count is
 * n_rows minus the stored counts of that locus.  `which` = 0 (sample file) or 1 (control file).
 * sel[] holds loci in [1, n_loci-2] whose mask is non-zero, ascending.
 */
int dj_kmer_count(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                  const uint8_t *text, const int64_t *row_off, const int32_t *row_len, const uint32_t *ckpt,
                  int32_t ckpt_pitch, const uint8_t *locimask, int32_t n_loci, const int32_t *sel,
                  int32_t n_sel, int32_t which, void *table, uint32_t capacity, void *stream);

/* Synchronous: copies the occupied slots to the HOST array h_out (at most max_records). */
int dj_kmer_export(const void *table, uint32_t capacity, DjKmerRecord *h_out, uint32_t max_records,
                   uint32_t *h_n_records, void *stream);

/*
 * Score rows (annotate_score, clustering/score_handler.py:136-148) as value ids:
 * ids[t][c] for row t (read rows[t]) and column c (locus col_locus[c]).  col_sel[c] is the index of that locus
 * in the `sel` array used for counting, or -1 when the locus can only carry "@,@,@".  A 3-mer found in the
 * table yields slot_value_id[slot]; "@,@,@" yields col_at_id[c]; control rows (is_control) get 0 for every
 * stored 3-mer, exactly as score_handler.py:143-144 does.
 */
int dj_annotate(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                const uint8_t *text, const int64_t *row_off, const int32_t *row_len, const uint32_t *ckpt,
                int32_t ckpt_pitch, const uint8_t *locimask, int32_t n_loci, const int32_t *col_locus,
                const int32_t *col_sel, const uint16_t *col_at_id, int32_t n_cols, int32_t is_control,
                const void *table, uint32_t capacity, const uint16_t *slot_value_id, uint16_t *ids,
                void *stream);

/* Raw text of token loci[k] of read reads[k] -> out[k][0..out_len[k]) (truncated to max_len). */
int dj_fetch_tokens(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, const uint32_t *ckpt,
                    int32_t ckpt_pitch, const int32_t *reads, const int32_t *loci, int32_t n, uint8_t *out,
                    int32_t *out_len, int32_t max_len, void *stream);

/* Row de-duplication of the id matrix: slot_of_row[t] = table slot of row t's content. */
int64_t dj_row_table_bytes(uint32_t capacity);
int dj_row_table_clear(void *table, uint32_t capacity, void *stream);
int dj_dedup_rows(const uint16_t *ids, int64_t n_rows, int32_t n_cols, const uint8_t *strand,
                  const int32_t *rows, void *table, uint32_t capacity, int32_t *slot_of_row, void *stream);
int dj_dedup_export(const void *table, uint32_t capacity, DjRowRecord *h_out, uint32_t max_records,
                    uint32_t *h_n_records, void *stream);

/*
 * One bisection of sklearn's BisectingKMeans on weighted points (sklearn/cluster/_bisect_k_means.py:298-363,
 * Lloyd: sklearn/cluster/_kmeans.py:630-758, sparse kernels sklearn/cluster/_k_means_lloyd.pyx:221-420,
 * _k_means_common.pyx:55-312).  Points whose leaf[] equals `target` are split in two starting from the points
 * seed0 / seed1.  sub_label[p] in {0,1} is written for those points (untouched elsewhere).
 * last_row[p] = position, in the row order of the matrix the reference would have built, of the LAST row equal to
 * point p: scikit-learn's empty-cluster relocation takes the farthest row and, among rows at the same distance, the
 * last one (np.argpartition), so ties between points are decided by it.
 * stats[0..1] = inertia of the two halves, stats[2] = iterations, stats[3..4] = weights of the halves.
 * The call that call site clustering/clustering.py:35,68 makes through scikit-learn.
 */
int dj_bisect(const double *X, const double *weight, int32_t n_points, int32_t n_cols, const int32_t *leaf,
              const int32_t *last_row, int32_t target, int32_t seed0, int32_t seed1, double tol, int32_t max_iter,
              int8_t *sub_label, double *centers, double *stats, void *stream);

/*
 * Brute-force read x centroid assignment for n_rows score rows given as value ids (no de-duplication):
 * label[t] = argmin_j |x_t - c_j|^2, plus per-centroid weighted sums.  Diagnostic / evidence path for the
 * dense contraction; the production path clusters the de-duplicated points with dj_bisect.
 */
int dj_assign_rows(const uint16_t *ids, int64_t n_rows, int32_t n_cols, const double *value_of_id,
                   const double *centers, int32_t n_centers, int32_t *label, double *sums, double *counts,
                   void *stream);

/*
 * Weighted silhouette score (sklearn/metrics/cluster/_unsupervised.py:59-300) of points with multiplicities:
 * every point stands for weight[p] identical rows.  h-less: score[0] receives the mean over all rows.
 */
int dj_silhouette(const double *X, const double *weight, const int32_t *label, int32_t n_points,
                  int32_t n_cols, int32_t n_clusters, double *score, void *stream);

/*
 * One iteration of optimize_labels (clustering/clustering.py:29-55) without a host round trip in between: locates the
 * two seeds scikit-learn draws for the bisection (seed_given[i] >= 0: a point; else the seed_rank[i]-th row, in file
 * order, of the sample rows whose point lies in leaf `target`: slot_of_row / point_of_slot as dj_dedup_rows made them),
 * runs the 2-means of that leaf (as dj_bisect), moves its members to child_a / child_b in `leaf` (device, in place),
 * labels every point with order_of_leaf_host[leaf] (left-to-right order of the leaves after the split), counts the
 * clusters that hold more than 1 % of the control weight (points >= n_sample_points; control_coverage = their total)
 * and, if want_silhouette, computes the weighted silhouette score (as dj_silhouette).  One device -> host copy.
 * sub_host[p] (host, n_points bytes) = 0 / 1 for the members of the split leaf, 0xFF elsewhere.
 */
typedef struct {
    double inertia[2];   /* of the two halves: the next leaf to split is the one with the largest */
    double weight[2];
    double silhouette;   /* undefined unless want_silhouette */
    int32_t iterations;
    int32_t control_clusters;
    int32_t sample_labels_distinct; /* 0: every sample point carries the same label (the reference then scores -1) */
    int32_t seed_point[2];
} DjClusterStepOut;
int dj_cluster_step(const double *X, const double *weight, int32_t n_points, int32_t n_sample_points, int32_t n_cols,
                    int32_t *leaf, const int32_t *last_row, const int32_t *slot_of_row, int64_t n_rows,
                    const int32_t *point_of_slot, int32_t target, int32_t child_a, int32_t child_b,
                    const int32_t *seed_given, const int64_t *seed_rank, const int32_t *order_of_leaf_host,
                    int32_t n_leaf_ids, int32_t n_clusters, double control_coverage, double tol, int32_t max_iter,
                    int32_t want_silhouette, int8_t *sub_host, DjClusterStepOut *out, void *stream);

/* j-th member selection for sklearn's random init: members of leaf `target` in row order. */
int dj_member_counts(const int32_t *slot_of_row, int64_t n_rows, const int32_t *leaf_of_slot, int32_t target,
                     int32_t *block_counts, void *stream); /* one count per 1024 rows */
int dj_member_locate(const int32_t *slot_of_row, int64_t n_rows, const int32_t *leaf_of_slot, int32_t target,
                     int32_t block, int32_t rank_in_block, int32_t *out_row, void *stream);

/* labels[t] = label_of_slot[slot_of_row[t]] */
int dj_gather_labels(const int32_t *slot_of_row, int64_t n_rows, const int32_t *label_of_slot,
                     int32_t *labels, void *stream);

/*
 * Per-read N features of the sequence-error read filter: n_count[t] = number of tokens of read rows[t] for which
 * is_n_tag holds (utils/midsv_handler.py:8-14), n_maxrun[t] = the longest run of such tokens.  Replaces the pass
 * convert_nm_tag + extract_n_features make over every token of the control and sample files
 * (preprocess/error_correction/sequence_error_handler.py:26-41, :48-53, :104-107) and the N count of the consensus
 * stage's control filter (consensus/consensus_mutation_analyzer.py:85-88).  rows == NULL: all n_rows reads in order.
 */
int dj_read_n_features(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                       int32_t n_loci, int32_t *n_count, int32_t *n_maxrun, void *stream);

/* Words of a per-locus bit string (bit b of word k = locus 32 * k + b) as dj_sv_scan expects them. */
int32_t dj_loci_words(int32_t n_loci);

/*
 * Structural-variant features of every read (get_index_and_sv_size, preprocess/structural_variants/
 * sv_detector.py:53-97), all three SV types in one scan of the code matrix:
 *   type 0  deletion   a maximal run of >= min_size tokens that start with '-' on loci set in del_loci
 *                      ('-' in mutation_loci[i], :67): start = first locus of the run, size = its length
 *   type 1  inversion  a maximal run of >= min_size lowercase tokens (:70-73)
 *   type 2  insertion  a token that starts with '+' on a locus set in ins_loci with >= min_size inserted bases
 *                      (tags[i].count("|"), :76-78): start = the locus, size = the number of inserted bases
 * records[k] = {type, t, start, size} (int32 x 4, t = position in `rows`), in no particular order; *n_records
 * (device) receives the number of records found, which may exceed `capacity` (only `capacity` are stored).
 * del_loci / ins_loci: dj_loci_words(n_loci) words each.
 */
int dj_sv_scan(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows, int32_t n_loci,
               const uint8_t *text, const int64_t *row_off, const int32_t *row_len, const uint32_t *ckpt,
               int32_t ckpt_pitch, const uint32_t *del_loci, const uint32_t *ins_loci, int32_t min_size,
               int32_t *records, uint32_t capacity, uint32_t *n_records, void *stream);

/*
 * SAM alignments (cs tag, long form) -> MIDSV text on the device: the per-read string
 * work of midsv.transform(path_sam, qscore=False, keep={"FLAG"}) as preprocess/alignment/midsv_caller.py:84-85 calls
 * it (third-party midsv >= 0.13.1, absent from the reference tree: PARITY UNPINNED; rules as restated in
 * oracle/midsv_transform.py and checked with the reference's own inverse, utils/midsv_handler.py:158-161).
 *   cs          the cs tags (without "cs:Z:") back to back; 16-byte aligned, >= 16 readable bytes after the last one;
 *               NUL bytes inside a tag are skipped (slack behind a tag the caller has shortened in place)
 *   aln_off     n_alignments + 1 offsets into cs
 *   aln_pos     0-based reference position of each alignment (POS - 1)
 *   aln_lower   1: the alignment is on the other strand than the first alignment of its read -> lowercase tokens
 *   read_first  n_reads + 1: alignments read_first[r] .. read_first[r+1]-1 belong to read r, ordered by position
 * "=ACG" -> "=A,=C,=G"; "*ag" -> "*AG"; "-ac" -> "-A,-C"; "+ac" joins the next token as "+A|+C|<next>";
 * "~gt123ag" -> 123 x "=N"; gaps between the alignments of a read and the flanks up to ref_len tokens -> "=N".
 *   ref_seq     NULL, or the ref_len bases of the reference sequence: the tokens filled between the pieces of a read
 *               (gaps between alignments, introns) are then written as "-<base>" instead of "=N", i.e. with
 *               replace_internal_n_to_d (midsv_caller.py:108-126) already applied; the flanks stay "=N"
 * dj_cs_midsv_measure: row_len[r] = bytes of read r's MIDSV string, 0 for a dropped read; row_status[r] = 0 or
 *   1 (two alignments share a reference base) | 2 (not a long-form cs tag) | 4 (longer than ref_len tokens / 2 GB)
 *   | 8 (informational, the read is kept: an alignment holds "=N" tokens of its own); row_n[r] (row_n may be NULL) =
 *   the number of "=N" tokens the filler writes, i.e. all is_n_tag tokens of the row unless bit 8 is set -- what
 *   filter_samples_by_n_proportion (midsv_caller.py:145-155) counts.
 * dj_cs_midsv_write: writes the string of every read with row_len[r] > 0 at text + row_off[r], followed by '\n'
 *   (row_len[r] + 1 bytes); the result is the text / row_off / row_len triple dj_encode takes.  ref_seq as given to
 *   dj_cs_midsv_measure.
 */
int dj_cs_midsv_measure(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                        const int64_t *read_first, int64_t n_reads, int32_t ref_len, const uint8_t *ref_seq,
                        int32_t *row_len, int32_t *row_status, int32_t *row_n, void *stream);
int dj_cs_midsv_write(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                      const int64_t *read_first, int64_t n_reads, int32_t ref_len, const uint8_t *ref_seq,
                      const int64_t *row_off, const int32_t *row_len, uint8_t *text, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DAJIN2_B200_H */
