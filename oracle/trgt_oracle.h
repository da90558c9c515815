/*
 * trgt_oracle.h -- CPU restatement of the TRGT hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for the B200 engine.  It restates, in plain C, the
 * algorithms of the reference (PacificBiosciences/trgt v3.0.0) for the path
 * named in BASELINE.json: wavefront alignment (src/wfaligner.rs over WFA2-lib)
 * and the motif HMM (src/hmm/).  Every function cites the reference file:line
 * it follows.  It is NOT part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call it.
 *
 * Pinning: checked against every golden vector the reference's own unit tests
 * hold for this path (tests/test_oracle_golden.py).
 * Parity UNPINNED (no reference test fixes the result, upstream WFA2-lib source
 * absent from /root/reference): BiWFA (MemoryUltraLow) CIGAR tie-breaking and
 * the default wfadaptive heuristic; this oracle implements exact
 * (Heuristic::None) unidirectional WFA for every aligner configuration.
 */
#ifndef TRGT_ORACLE_H
#define TRGT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ HMM -- */

typedef struct tro_hmm tro_hmm;

/* build_hmm: src/hmm/builder.rs:4-78.  motifs = concatenated bytes, CSR offsets[n+1].
 * Motif bytes must be in ACGTN (the caller applies replace_invalid_bases).
 * Returns NULL on an invalid base (the reference panics, builder.rs:182). */
tro_hmm *tro_hmm_build(const uint8_t *motifs, const uint32_t *offsets, uint32_t n_motifs);
void tro_hmm_free(tro_hmm *h);
int tro_hmm_num_states(const tro_hmm *h);
/* raw table access (for tests): ln emission of `state` for symbol 0..4 (#ATCG) */
double tro_hmm_em(const tro_hmm *h, int state, int sym);
int tro_hmm_num_in(const tro_hmm *h, int state);
int tro_hmm_in_state(const tro_hmm *h, int state, int i);
double tro_hmm_in_lp(const tro_hmm *h, int state, int i);

/* Hmm::label: src/hmm/hmm_model.rs:144-156 (Viterbi :54-114, traceback :125-142).
 * query = ACGT bytes (no sentinels).  Writes the state path into out (capacity cap)
 * and returns its length, 0 for an empty query, -1 on invalid base, -2 if cap is
 * too small. */
int64_t tro_hmm_label(const tro_hmm *h, const uint8_t *query, uint32_t len,
                      uint32_t *out, uint64_t cap);

/* remove_imperfect_motifs: src/hmm/operations.rs:6-80.  Returns new length. */
int64_t tro_remove_imperfect_motifs(const tro_hmm *h, const uint32_t *states, uint64_t n_states,
                                    const uint8_t *query, uint32_t qlen, uint32_t max_motif_len,
                                    uint32_t *out, uint64_t cap);

typedef struct {
  uint32_t motif_index, start, end;
} tro_span;

/* Hmm::label_motifs: src/hmm/hmm_model.rs:158-200.  Returns number of spans. */
int64_t tro_label_motifs(const tro_hmm *h, const uint32_t *states, uint64_t n_states,
                         tro_span *out, uint64_t cap);

/* calc_purity: src/hmm/purity.rs:6-41 via get_events src/hmm/events.rs:17-86. */
double tro_calc_purity(const tro_hmm *h, const uint32_t *states, uint64_t n_states,
                       const uint8_t *query, uint32_t qlen);

/* get_base_match: src/hmm/events.rs:88-117 */
uint8_t tro_get_base_match(const tro_hmm *h, int state);

/* replace_invalid_bases: src/hmm/utils.rs:29-42; allowed = "ATCG" or "ATCGN". In place. */
void tro_replace_invalid_bases(uint8_t *seq, uint32_t len, const char *allowed);

/* label_with_hmm for ONE allele: src/trgt/workflows/tr.rs:464-489.
 * `allele` is the raw allele (replace_invalid_bases is applied inside).
 * motif_counts[n_motifs] out; spans_out = collapsed labels (skip spans dropped);
 * returns number of collapsed spans (0 <=> labels None), or <0 on error. */
int64_t tro_annotate_allele(const tro_hmm *h, const uint8_t *allele, uint32_t len,
                            uint32_t *motif_counts, tro_span *spans_out, uint64_t span_cap,
                            double *purity_out);

/* ------------------------------------------------------------------ WFA -- */

enum { TRO_INDEL = 0, TRO_EDIT = 1, TRO_LINEAR = 2, TRO_AFFINE = 3, TRO_AFFINE2P = 4 };

/* wfa2 status codes as surfaced by src/wfaligner.rs:132-159 */
enum {
  TRO_STATUS_OK = 0,
  TRO_STATUS_MAX_STEPS = -100,
  TRO_STATUS_OOM = -200,
  TRO_STATUS_UNATTAINABLE = -300
};

typedef struct {
  int metric;            /* TRO_* */
  int x;                 /* mismatch */
  int o1, e1;            /* gap open/extend (linear: e1 = indel penalty) */
  int o2, e2;            /* second affine piece */
  int ends_free;         /* 0 = end-to-end */
  int pattern_begin_free, pattern_end_free, text_begin_free, text_end_free;
  int score_only;        /* AlignmentScope::Score */
  int max_steps;         /* <=0: unlimited */
  /* optional (0 = off): WFA2-lib's adaptive heuristic wfadaptive(min_len, max_dist), applied at every score
   * (steps_between_cutoffs = 1).  Restated from the published rule, parity unpinned: a MEASURING device only. */
  int wfadaptive_min_len, wfadaptive_max_dist;
} tro_wfa_params;

typedef struct {
  int status;
  int score;             /* sign convention of wfaligner.rs:530 (edit/indel +cost, others -cost) */
  uint8_t *ops;          /* 'M','X','I','D' in forward order; malloc'ed; NULL if score_only */
  int64_t n_ops;
  int end_k, end_offset; /* diagonal / text offset at which the wavefront terminated */
} tro_wfa_result;

/* wavefront_align as driven by WFAligner::align_end_to_end / align_ends_free
 * (src/wfaligner.rs:489-528).  The DP itself lives in WFA2-lib (wfa2-sys 0.1.0,
 * git rev 4342b3b0..., not vendored); rules restated in SURVEY.md section 8c. */
int tro_wfa_align(const tro_wfa_params *p, const uint8_t *pattern, int plen,
                  const uint8_t *text, int tlen, tro_wfa_result *res);
void tro_wfa_result_free(tro_wfa_result *res);

/* accessors: wfaligner.rs:988 count_matches, :864 get_alignment_span, :932 get_sam_cigar,
 * :1002 cigar_score, :595 cigar_score_clipped */
int tro_count_matches(const uint8_t *ops, int64_t n);
void tro_alignment_span(const uint8_t *ops, int64_t n, int ends_free, int plen, int tlen,
                        int *xs, int *xe, int *ys, int *ye);
int64_t tro_sam_cigar(const uint8_t *ops, int64_t n, int show_mismatches, uint32_t *out,
                      uint64_t cap);
int tro_cigar_score(const tro_wfa_params *p, const uint8_t *ops, int64_t n);
int tro_cigar_score_clipped(const tro_wfa_params *p, const uint8_t *ops, int64_t n, int flank_len);

/* ----------------------------------------------------- callers (a1-a6) -- */

typedef struct {
  int32_t found;         /* Option::is_some */
  uint32_t start, end;
} tro_opt_span;

/* find_spans for ONE read: src/trgt/genotype/span_locater.rs:7-30.
 * via (optional): 0 none, 1 exact window hit, 2 WFA fallback accepted, 3 WFA fallback rejected.
 * matches (optional): count_matches of the fallback alignment (or piece_len for exact). */
tro_opt_span tro_find_span(const uint8_t *piece, int piece_len, const uint8_t *seq, int seq_len,
                           int x, int o, int e, double threshold, int *via, int *matches);

/* find_tr_spans combine rule: span_locater.rs:53-67 */
tro_opt_span tro_combine_spans(tro_opt_span lf, tro_opt_span rf);

/* utils::align for ONE (backbone, seq): src/utils/align.rs:14-28 -> run-length SAM cigar words
 * (len<<4|op with '='=7 'X'=8 'I'=1 'D'=2).  Gap-affine (2,5,1), exact WFA. */
int64_t tro_align_consensus(const uint8_t *backbone, int blen, const uint8_t *seq, int slen,
                            uint32_t *out, uint64_t cap, int *score);

/* get_dist: src/trgt/genotype/genotype_cluster.rs:236-248 */
double tro_get_dist(const uint8_t *a, int alen, const uint8_t *b, int blen);

/* ------------------------------------------------- next row: consensus -- */

/* repair_consensus: src/trgt/genotype/consensus.rs:5-72 (get_ins_consensus :94-111).  words / word_off:
 * run-length SAM CIGAR of every seq against the backbone (utils::align output).  Returns the consensus
 * length, -1 on an unexpected base or op (the reference panics), -2 if cap is too small.
 * PARITY UNPINNED: the reference's tests for it are commented out (consensus.rs:168-215). */
int64_t tro_repair_consensus(const uint8_t *backbone, uint32_t blen, const uint8_t *seqs, const uint64_t *seq_off,
                             uint32_t n_seqs, const uint32_t *words, const uint64_t *word_off, uint8_t *out,
                             uint64_t cap);

/* ------------------------------------------ next row: read clipping -- */

/* Result of clip_cigar + the query range clip_to_region copies.  The clipped CIGAR is
 * [first_word, ops[first_op+1 .. first_op+n_ops-2], last_word] (first_word == last_word if n_ops == 1). */
typedef struct {
  int64_t ref_start;       /* clipped_ref_start */
  uint64_t query_start;    /* clipped_query_start */
  uint64_t query_end;      /* query_start + sum of the clipped ops' query lengths */
  uint32_t first_op;       /* index in ops of the first clipped op */
  uint32_t n_ops;          /* number of clipped ops */
  uint32_t first_word, last_word; /* BAM-encoded (len<<4)|op, after splitting */
} tro_clip;

/* clip_cigar: src/trgt/reads/clip_region.rs:105-186 (+ :19-38).  1 = overlap (out filled), 0 = None,
 * -1 = the reference panics (split of an op without reference length). */
int tro_clip_cigar(const uint32_t *ops, uint32_t n_ops, int64_t ref_start, int64_t region_start,
                   int64_t region_end, tro_clip *out);

/* rec.seq().as_bytes() (read.rs:104) for bases [start, start+len) of a BAM 4-bit sequence */
void tro_decode_seq4(const uint8_t *packed, uint64_t start, uint32_t len, uint8_t *out);

/* ------------------------------------------ next row: BAMlet clipping -- */

/* Result of HiFiRead::clip_bases (clip_bases.rs:9-56): bases / quals [base_start, base_end), methylation entries
 * [meth_start, meth_end), and -- if the read has a CIGAR -- its clipped form
 * [first_word, ops[first_op+1 .. first_op+n_ops-2], last_word] at reference position ref_pos. */
typedef struct {
  int64_t ref_pos;
  uint64_t base_start, base_end;
  uint32_t meth_start, meth_end;
  uint32_t first_op, n_ops;
  uint32_t first_word, last_word;
  int32_t has_cigar;
  int32_t pad;
} tro_bclip;

/* which methylation entries a clip to bases [start, end) keeps: clip_region.rs:40-58, clip_bases.rs:23-44 */
void tro_meth_range(const uint8_t *bases, uint64_t len, uint64_t start, uint64_t end, uint32_t *m0, uint32_t *m1);
/* clip_bases.rs:9-119.  1 = Some, 0 = None, -1 = the reference panics */
int tro_clip_bases(const uint32_t *ops, uint32_t n_ops, int64_t ref_pos, const uint8_t *bases, uint64_t len,
                   uint64_t left_len, uint64_t right_len, tro_bclip *out);
/* the clip of BamWriter::write, write_bam.rs:80-92.  1 = written, 0 = skipped (short flanks / None), -1 = panic */
int tro_bamlet_clip(const uint32_t *ops, uint32_t n_ops, int64_t ref_pos, const uint8_t *bases, uint64_t len,
                    uint64_t span_start, uint64_t span_end, uint64_t flank_len, tro_bclip *out);

/* ------------------------------------------ next row: VCF sample fields -- */

/* One field (0 AL, 1 MC, 2 MS, 3 AP) of one locus: src/trgt/writers/write_vcf.rs:267-343.  Returns the
 * field's length; the bytes are written when they fit cap. */
size_t tro_vcf_field(int field, uint32_t n_alleles, const uint64_t *allele_len, const uint64_t *mc_off,
                     const uint32_t *mc, const uint64_t *span_off, const tro_span *spans, const double *purity,
                     char *out, size_t cap);

/* ------------------------------------------- batched CPU baseline path -- */

/* Batched, multi-threaded drivers (batch_oracle.c) over the per-item functions above.  They take the
 * packed CSR inputs of include/trgt_engine.h and fill the same packed outputs; used ONLY by the
 * parity tests and by bench.py's cpu_baseline / --impl reference legs.  One task per locus (or
 * group / allele chunk) over n_threads workers, as the reference's rayon pool runs them
 * (src/commands/genotype.rs:178-187). */
typedef struct {
  int32_t via, matches, score;
  uint32_t start, end;
} tro_flank_hit;

/* find_tr_spans for a chunk of loci: span_locater.rs:32-68.  hits_out may be NULL. */
int tro_flank_batch(const uint8_t *left, const uint64_t *left_off, const uint8_t *right, const uint64_t *right_off,
                    const uint8_t *reads, const uint64_t *read_off, const uint32_t *locus_read_off, uint32_t n_loci,
                    int x, int o, int e, double min_flank_id_frac, tro_opt_span *spans_out, tro_flank_hit *hits_out,
                    int n_threads);

/* utils::align for many (backbone, seqs) groups: src/utils/align.rs:14-28.
 * offsets_out[n_seqs+1]; *words_out is malloc'ed (tro_free). */
int tro_align_batch(const uint8_t *bb, const uint64_t *bb_off, const uint8_t *seqs, const uint64_t *seq_off,
                    const uint32_t *group_off, uint32_t n_groups, uint64_t *offsets_out, uint32_t **words_out,
                    int32_t *scores_out, int n_threads);

/* label_with_hmm for many loci: tr.rs:454-492.  mc_off[n_alleles+1] = prefix sum of the motif count of
 * each allele's locus; span_off_out[n_alleles+1]; *spans_out is malloc'ed (tro_free). */
int tro_hmm_batch(const uint8_t *motifs, const uint64_t *motif_off, const uint32_t *locus_motif_off, uint32_t n_loci,
                  const uint8_t *alleles, const uint64_t *allele_off, const uint32_t *allele_locus, uint32_t n_alleles,
                  const uint64_t *mc_off, uint32_t *mc_out, uint64_t *span_off_out, tro_span **spans_out,
                  double *purity_out, int32_t *status_out, int n_threads);

void tro_free(void *p);

/* ---------------------------------------------- next row: cluster-genotyper glue (cluster_oracle.c) -- */

typedef struct {
  uint32_t cluster1, cluster2; /* cluster1 < cluster2; observations are 0..n-1, step i creates cluster n+i */
  double dissimilarity;
  uint32_t size;
} tro_linkage_step;

/* kodama::linkage(dists, n, Method::Ward) as called at genotype_cluster.rs:161; dists is modified in place */
int tro_ward_linkage(double *dists, uint32_t n, tro_linkage_step *steps_out);
/* cluster(): genotype_cluster.rs:154-227 -> number of groups, group_out[n] */
int tro_cluster(uint32_t n, double *dists, uint32_t *group_out);
/* central_read(): genotype_cluster.rs:12-39 */
uint32_t tro_central_read(uint32_t num_seqs, const uint32_t *group, uint32_t group_size, const double *dists);
/* genotype() :57-72 up to the make_consensus calls: group1 / group2 / others and their central reads */
int tro_cluster_locus(uint32_t n, double *dists, uint8_t *sel_out, uint32_t *central_out);

#ifdef __cplusplus
}
#endif
#endif
