/*
 * trgt_engine.h -- C ABI of the B200 tandem-repeat DP engine (libtrgt_b200.so).
 *
 * This is the drop-in boundary for the hot path of `trgt genotype` (PacificBiosciences/trgt
 * v3.0.0): the host (Rust in the reference) keeps its Locus/Genotyper surface and per-locus
 * worker logic and replaces four leaf calls by batched, stateless calls into this library.
 * All pointers are plain host pointers unless a function says "device"; no torch / C++ types
 * cross the boundary; nothing throws or aborts across it.  The reference reaches its aligner
 * through the one-alignment-at-a-time WFA2-lib FFI re-exported by src/wfa2.rs:2 (call sites
 * src/wfaligner.rs:371,458,467,479-485,492-498,519-525,944-949,999) and calls its HMM as plain
 * Rust (src/hmm/); each entry point below names the reference interface it replaces.
 *
 * Conventions
 *  - sequences are ASCII bytes in one contiguous buffer with CSR offsets (uint64_t[n+1]);
 *  - a call blocks until its batch is complete and results are in host memory;
 *  - return value: 0 = OK, <0 = engine failure (trgt_engine_last_error() has the text);
 *    per-item failures are reported in per-item status arrays with the WFA2 status codes
 *    the reference surfaces (src/wfaligner.rs:132-159);
 *  - entry points may be called from any host thread; calls on one engine serialise.
 *  - there is NO CPU fallback: without a CUDA device trgt_engine_create fails.
 */
#ifndef TRGT_ENGINE_H
#define TRGT_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRGT_OK 0
#define TRGT_ERR_CUDA (-1)
#define TRGT_ERR_ARG (-2)
#define TRGT_ERR_NO_DEVICE (-4)
#define TRGT_ERR_INTERNAL (-5)

/* per-item status, numerically the WF_STATUS_* codes of WFA2-lib */
#define TRGT_ITEM_OK 0
#define TRGT_ITEM_MAX_STEPS (-100)
#define TRGT_ITEM_OOM (-200)
#define TRGT_ITEM_UNATTAINABLE (-300)
#define TRGT_ITEM_INVALID_BASE (-400) /* reference panics: hmm_model.rs:250, builder.rs:182 */
#define TRGT_ITEM_INVALID_OP (-500)   /* reference panics: clip_region.rs:150,178 "Unexpected operation" */

typedef struct trgt_engine trgt_engine_t;
/* inputs of one phase kept resident in HBM (resident-batch interface below) */
typedef struct trgt_flank_batch trgt_flank_batch_t;
typedef struct trgt_align_batch trgt_align_batch_t;
typedef struct trgt_hmm_batch trgt_hmm_batch_t;

/* CSR set of byte sequences */
typedef struct {
  const uint8_t *data;
  const uint64_t *offsets; /* [n+1] */
  uint64_t n;
} trgt_seqs_t;

/* --aln-scoring (src/cli.rs:271-280): gap-affine penalties of the flank aligner */
typedef struct {
  int32_t mismatch, gap_open, gap_extend;
} trgt_scoring_t;

/* Option<(usize,usize)> */
typedef struct {
  int32_t found;
  uint32_t start, end;
} trgt_span_t;

/* how one (read, flank) pair was resolved; optional diagnostic output of trgt_flank_spans */
#define TRGT_VIA_NONE 0
#define TRGT_VIA_EXACT 1        /* span_locater.rs:10-12 */
#define TRGT_VIA_WFA 2          /* span_locater.rs:14-25, accepted */
#define TRGT_VIA_WFA_REJECTED 3 /* count_matches below threshold */
typedef struct {
  int32_t via;
  int32_t matches; /* count_matches (piece length for an exact hit) */
  int32_t score;   /* WFA score (<= 0), 0 for an exact hit */
  uint32_t start, end;
} trgt_flank_hit_t;

/* ---- lifetime ---------------------------------------------------------- */

/* One engine per GPU.  Replaces the thread-local aligner trio of
 * src/commands/genotype.rs:94-103 (THREAD_WFA_FLANK / _CONSENSUS / _ED). */
int32_t trgt_engine_create(int32_t device, trgt_engine_t **out);
void trgt_engine_destroy(trgt_engine_t *eng);
const char *trgt_engine_last_error(const trgt_engine_t *eng);
/* text of the last failure of trgt_engine_create (no engine to ask) */
const char *trgt_last_create_error(void);
/* cudaStream_t the engine launches on (for external CUDA-event timing) */
void *trgt_engine_stream(trgt_engine_t *eng);
int32_t trgt_engine_sm_count(const trgt_engine_t *eng);
/* block until everything queued on the engine stream has finished */
int32_t trgt_engine_sync(trgt_engine_t *eng);
/* cap (bytes) on the device workspace one wave of work may use (HMM back-pointers, WFA trace
 * history); larger batches are processed in several waves.  Default 24 GiB. */
void trgt_engine_set_workspace_budget(trgt_engine_t *eng, size_t bytes);
/* Phase A settles a (read, flank) pair that misses the exact search on chip when its alignment
 * cost is <= max_cost (seed filter + banded wavefront, results identical to the full-width
 * alignment); costlier pairs take the full-width path.  0 sends every miss down the full-width
 * path.  Default 16. */
void trgt_engine_set_flank_band_budget(trgt_engine_t *eng, int32_t max_cost);

/* Phase C runs the alleles of single-motif loci (one motif of 1..7 bases: every locus of a genome-wide catalog)
 * through kernels specialised per motif length (one lane per allele, score column in registers, one packed
 * back-pointer word per column); all other loci take the generic kernels.  Results are identical either way;
 * 0 sends every allele down the generic path (diagnostic).  Default on.  Applies to batches uploaded afterwards. */
void trgt_engine_set_hmm_lane_path(trgt_engine_t *eng, int32_t on);

/* pinned host memory for the caller's packing buffers (DMA-direct H2D/D2H) */
void *trgt_host_alloc(size_t bytes);
void trgt_host_free(void *p);

/* ---- producer of phase A's input: read clipping and BAM bases (next row, rank 2) ---- */

/* clip_cigar plus the query range HiFiRead::clip_to_region copies
 * (src/trgt/reads/clip_region.rs:19-38,105-186).  The clipped CIGAR is
 * [first_word, ops[first_op+1 .. first_op+n_ops-2], last_word]; the clipped bases are
 * bases[query_start .. query_end) of the record. */
typedef struct {
  int64_t ref_start;      /* clipped_ref_start */
  uint64_t query_start;   /* clipped_query_start */
  uint64_t query_end;
  uint32_t first_op;      /* index in the read's ops of the first clipped op */
  uint32_t n_ops;         /* number of clipped ops */
  uint32_t first_word, last_word; /* BAM encoding (len<<4)|op after splitting; equal when n_ops == 1 */
  int32_t status;         /* 1 = overlaps the region, 0 = None (read dropped by clip_reads),
                             TRGT_ITEM_INVALID_OP = the reference panics */
} trgt_clip_t;

/* clip_reads for a chunk of loci: src/trgt/workflows/tr.rs:186-196 -> HiFiRead::clip_to_region.
 *   cigar_ops: BAM-encoded CIGAR words ((len<<4)|op, op in MIDNSHP=X) of all reads, as bam_get_cigar
 *   stores them; cigar_offsets[n_reads+1] delimits each read; ref_starts[n_reads] = rec.reference_start()
 *   regions[2*n_loci] = (start, end) the reads of each locus are clipped to (locus region +- radius)
 *   clips_out[n_reads] */
int32_t trgt_clip_reads(trgt_engine_t *eng, const uint32_t *cigar_ops, const uint64_t *cigar_offsets,
                        const int64_t *ref_starts, uint64_t n_reads, const int64_t *regions,
                        const uint32_t *locus_read_offsets, uint32_t n_loci, trgt_clip_t *clips_out);

/* Read bases as a BAM record stores them (htslib bam_get_seq / rust-htslib Seq::encoded: codes of
 * "=ACMGRSVTWYHKDBN", two bases per byte, first base in the high nibble).  Read i is bases
 * [starts[i], starts[i] + lengths[i]) of `data`, starts counted in bases (nibbles); starts must be
 * non-decreasing.  The host copies the bytes covering bases[query_start..query_end) of each record
 * (trgt_clip_t) into `data` -- no decoding, no shifting -- and sets starts[i] = 2*byte_offset + (query_start & 1). */
typedef struct {
  const uint8_t *data;
  uint64_t data_bytes;
  const uint64_t *starts;   /* [n] */
  const uint32_t *lengths;  /* [n] */
  uint64_t n;
} trgt_seq4_t;

/* ---- phase A: flank location (a1-a4) ------------------------------------ */

/* find_tr_spans for a chunk of loci: src/trgt/genotype/span_locater.rs:32-68 (find_spans :7-30,
 * WFA fallback = WFAligner::align_ends_free src/wfaligner.rs:503-528 with the aligner of
 * src/commands/genotype.rs:66-80: gap-affine, Heuristic::None, MemoryHigh; then count_matches
 * :988 and get_alignment_span :864).
 *   left_pieces/right_pieces: n_loci flank pieces (lf[len-P..], rf[..P]; span_locater.rs:38-39)
 *   reads: all clipped reads of the chunk; locus_read_offsets[n_loci+1] delimits each locus
 *   spans_out[n_reads]; hits_out[2*n_reads] optional (may be NULL): [2*r] left, [2*r+1] right */
int32_t trgt_flank_spans(trgt_engine_t *eng, const trgt_seqs_t *left_pieces,
                         const trgt_seqs_t *right_pieces, const trgt_seqs_t *reads,
                         const uint32_t *locus_read_offsets, uint32_t n_loci,
                         trgt_scoring_t scoring, double min_flank_id_frac,
                         trgt_span_t *spans_out, trgt_flank_hit_t *hits_out);

/* trgt_flank_spans on reads handed over as BAM 4-bit bases: replaces `rec.seq().as_bytes()`
 * (src/trgt/reads/read.rs:104) + the copy of clip_to_region (clip_region.rs:29-31) + find_tr_spans.
 * Half the bytes cross PCIe; the engine decodes to the ASCII reads of phase A in HBM (k_unpack_seq4).
 * Spans are in bases of the clipped read, as in trgt_flank_spans. */
int32_t trgt_flank_spans_seq4(trgt_engine_t *eng, const trgt_seqs_t *left_pieces,
                              const trgt_seqs_t *right_pieces, const trgt_seq4_t *reads,
                              const uint32_t *locus_read_offsets, uint32_t n_loci,
                              trgt_scoring_t scoring, double min_flank_id_frac,
                              trgt_span_t *spans_out, trgt_flank_hit_t *hits_out);

/* decode only (for callers that want the clipped ASCII reads back, e.g. to cut allele sequences):
 * ascii_out receives the reads back to back, ascii_offsets_out[n+1] their CSR offsets */
int32_t trgt_seq4_decode(trgt_engine_t *eng, const trgt_seq4_t *reads, uint8_t *ascii_out,
                         uint64_t *ascii_offsets_out);

/* ---- consumer of phase A's output: the reads of the BAMlet (next row, rank 4) ---- */

#define TRGT_BAMLET_SKIPPED 0   /* no span, or "unexpectedly short flanks" (write_bam.rs:80-83, :88-91): no record */
#define TRGT_BAMLET_WRITE 1

/* HiFiRead::clip_bases (src/trgt/reads/clip_bases.rs:9-119) as BamWriter::write asks for it
 * (src/trgt/writers/write_bam.rs:72-92: flank_len bases either side of the repeat span): the record's bases and
 * quals are [base_start, base_end) of the clipped read, its methylation profile entries [meth_start, meth_end)
 * (one entry per "CG" of the read, clip_bases.rs:23-44), its CIGAR
 * [first_word, ops[first_op+1 .. first_op+n_ops-2], last_word] at ref_pos (n_ops == 0 for a read without CIGAR:
 * the writer marks it unmapped at the locus start, write_bam.rs:108-112).  The record itself (rust-htslib
 * Record::set / push_aux, :95-140) stays with the host's BAM writer. */
typedef struct {
  int64_t ref_pos;
  uint32_t base_start, base_end;
  uint32_t meth_start, meth_end;
  uint32_t first_op, n_ops;
  uint32_t first_word, last_word;
  int32_t status;   /* TRGT_BAMLET_WRITE, TRGT_BAMLET_SKIPPED, TRGT_ITEM_INVALID_OP where the reference panics / exits
                       (clip_bases.rs:61,80,106) */
  uint32_t pad;
} trgt_bamlet_clip_t;

/* clip_bases for every read of a resident phase-A batch that has run (the clipped reads and their spans are in
 * HBM already: only the CIGARs go up and 48 bytes per read come back).
 *   cigar_ops / cigar_offsets[n_reads+1]: the reads' CIGARs after clip_to_region, BAM words (len<<4)|op; a read
 *   without CIGAR has an empty range; ref_starts[n_reads] = cigar.ref_pos; clips_out[n_reads];
 *   batch == NULL: the batch of the last one-shot trgt_flank_spans / trgt_flank_spans_seq4 call on this engine */
int32_t trgt_bamlet_clip(trgt_engine_t *eng, trgt_flank_batch_t *batch, const uint32_t *cigar_ops,
                         const uint64_t *cigar_offsets, const int64_t *ref_starts, uint32_t flank_len,
                         trgt_bamlet_clip_t *clips_out);

/* ---- phase B: consensus alignments (a5) and edit distances (a6) ---------- */

typedef struct {
  uint64_t n;              /* number of alignments */
  const uint64_t *offsets; /* [n+1] into words */
  const uint32_t *words;   /* run-length SAM CIGAR: (len<<4)|op, '='=7 'X'=8 'I'=1 'D'=2,
                              exactly what WFAligner::decode_sam_cigar expects (wfaligner.rs:961) */
  const int32_t *scores;   /* WFA score, -cost */
  const int32_t *status;   /* per-item TRGT_ITEM_* */
} trgt_cigars_t;

/* utils::align for many (backbone, seqs) groups: src/utils/align.rs:14-28
 * (WFAligner::align_end_to_end src/wfaligner.rs:489, gap-affine (2,5,1) fixed, then
 * get_sam_cigar(true) :932).  group_seq_offsets[n_groups+1] delimits each backbone's seqs.
 * Exact WFA; the reference's BiWFA + default heuristic are not pinned by any reference test
 * (see DESIGN.md "parity unpinned").  `out` points into engine-owned memory that stays valid
 * until the next call on this engine (as WFA2 owns its CIGAR until the next wavefront_align). */
int32_t trgt_align_e2e(trgt_engine_t *eng, const trgt_seqs_t *backbones, const trgt_seqs_t *seqs,
                       const uint32_t *group_seq_offsets, uint32_t n_groups, trgt_cigars_t *out);

/* one sequence per group */
typedef struct {
  uint64_t n;
  const uint64_t *offsets; /* [n+1] into data */
  const uint8_t *data;
  const int32_t *status;   /* per-group TRGT_ITEM_* */
} trgt_seqs_out_t;

/* utils::align followed by repair_consensus for many (backbone, seqs) groups, as every genotyper
 * chains them (genotype_size.rs:35-36, genotype_cluster.rs:52-53, genotype_flank.rs:19-20):
 * src/trgt/genotype/consensus.rs:5-72 (+ get_ins_consensus :94-111).  The CIGARs stay on the device;
 * out receives one repaired consensus per group (engine-owned pinned memory, valid until the next
 * align / consensus call on this engine).  A member base outside ACGT makes the reference panic
 * (consensus.rs:81); here the group gets TRGT_ITEM_INVALID_BASE. */
int32_t trgt_consensus(trgt_engine_t *eng, const trgt_seqs_t *backbones, const trgt_seqs_t *seqs,
                       const uint32_t *group_seq_offsets, uint32_t n_groups, trgt_seqs_out_t *out);

/* get_dist_matrix for many loci: src/trgt/genotype/genotype_cluster.rs:236-286.
 * locus_seq_offsets[n_loci+1] delimits each locus' TR sequences; dists_out receives, locus after
 * locus, the condensed upper triangles (index i*n - i(i+1)/2 + (j-i-1)), sqrt applied. */
int32_t trgt_edit_dist(trgt_engine_t *eng, const trgt_seqs_t *seqs,
                       const uint32_t *locus_seq_offsets, uint32_t n_loci, double *dists_out);

/* ---- cluster-genotyper glue between get_dist_matrix and make_consensus (next row, rank 3) ---- */

/* For many loci: get_dist_matrix (genotype_cluster.rs:250-286), then cluster() (:154-227: Ward linkage of
 * kodama::linkage on the distance matrix, the dendrogram cut that leaves both sides at least
 * max(2, round(n / 100)) sequences, alternate split when there is none), the choice of the two largest groups
 * (:64-69: stable sort by size, the last two) and central_read (:12-39) of each -- what genotype() needs to call
 * make_consensus (:41-55) twice.
 *   group_out[n_seqs]: 0 = member of group1 (the largest group), 1 = group2, 2 = neither (an outlier that
 *     genotype() assigns to the closer consensus afterwards, :125-142)
 *   central_out[2 * n_loci]: index within the locus of the backbone of group1 / group2, 0xFFFFFFFF when the locus
 *     has no such group (fewer than two sequences)
 *   n_groups_out[n_loci] (may be NULL): number of groups cluster() returned
 * kodama (0.3.0) is not part of the reference tree: see DESIGN.md for what pins this row. */
int32_t trgt_cluster(trgt_engine_t *eng, const trgt_seqs_t *seqs, const uint32_t *locus_seq_offsets, uint32_t n_loci,
                     int32_t *group_out, uint32_t *central_out, uint32_t *n_groups_out);

/* The same on the repeat sequences of a flank batch that has been run (batch == NULL: the last one-shot call),
 * read where they lie in HBM: reads[locus_offsets[l] .. locus_offsets[l+1]) are the read indices (within the
 * batch) of locus l's spanning reads in genotyping order (tr.rs:138-165: margin filter, stable sort by repeat
 * length, down-sampling -- host logic on the spans).  Indices in group_out / central_out are positions in that
 * per-locus list.  Nothing but the index lists crosses PCIe. */
int32_t trgt_cluster_trs(trgt_engine_t *eng, trgt_flank_batch_t *batch, const uint32_t *reads,
                         const uint32_t *locus_offsets, uint32_t n_loci, int32_t *group_out, uint32_t *central_out,
                         uint32_t *n_groups_out);

/* trgt_consensus with backbones and members named by read index of a flank batch (make_consensus :41-55 for many
 * groups): the repeat sequences are gathered on the device, aligned and voted on; only the repaired consensuses
 * come back. */
int32_t trgt_consensus_trs(trgt_engine_t *eng, trgt_flank_batch_t *batch, const uint32_t *backbone_reads,
                           const uint32_t *member_reads, const uint32_t *group_offsets, uint32_t n_groups,
                           trgt_seqs_out_t *out);

/* trgt_align_e2e with backbones and members named by read index of a flank batch (the repeat sequences are gathered
 * on the device from the batch's reads and spans): utils::align (src/utils/align.rs:14-28) for hosts that took their
 * repeat sequences from trgt_flank_trs -- 4 bytes per member go up instead of its bases and a 64-bit offset.
 * batch == NULL: the batch of the last one-shot trgt_flank_spans* call.  `out` as for trgt_align_e2e. */
int32_t trgt_align_trs(trgt_engine_t *eng, trgt_flank_batch_t *batch, const uint32_t *backbone_reads,
                       const uint32_t *member_reads, const uint32_t *group_offsets, uint32_t n_groups,
                       trgt_cigars_t *out);

/* ---- phase C: motif HMM (a7-a13) ---------------------------------------- */

typedef struct {
  uint32_t motif_index, start, end;
} trgt_motif_span_t;

typedef struct {
  uint64_t n;                        /* number of alleles */
  const uint64_t *motif_count_offsets; /* [n+1] */
  const uint32_t *motif_counts;        /* MC: copies per motif of the allele's locus */
  const uint64_t *span_offsets;        /* [n+1] */
  const trgt_motif_span_t *spans;      /* MS: collapsed labels, skip spans dropped; empty <=> None */
  const double *purity;                /* AP (NaN for an empty allele) */
  const int32_t *status;               /* TRGT_ITEM_* */
  /* optional raw Viterbi state paths (Hmm::label), filled only when want_paths != 0 */
  const uint64_t *path_offsets;        /* [n+1] or NULL */
  const uint32_t *paths;
} trgt_annotations_t;

/* label_with_hmm for many loci: src/trgt/workflows/tr.rs:454-492 = build_hmm
 * (src/hmm/builder.rs:4) once per locus, then per allele Hmm::label (hmm_model.rs:144),
 * calc_purity (purity.rs:6), remove_imperfect_motifs(.., 6) (operations.rs:6), label_motifs
 * (hmm_model.rs:158), skip-span filter, count_motifs / collapse_labels (utils.rs:3,11).
 * replace_invalid_bases (utils.rs:29) is applied inside to motifs (ATCGN) and alleles (ATCG).
 *   motifs: all motifs of all loci; locus_motif_offsets[n_loci+1] delimits each locus
 *   alleles: sequences to annotate; allele_locus[n_alleles] = locus of each
 * Also serves filter_impure_trs (tr.rs:400-452): pass reads' TR sequences as alleles. */
int32_t trgt_hmm_label(trgt_engine_t *eng, const trgt_seqs_t *motifs,
                       const uint32_t *locus_motif_offsets, uint32_t n_loci,
                       const trgt_seqs_t *alleles, const uint32_t *allele_locus,
                       int32_t want_paths, trgt_annotations_t *out);

/* ---- resident-batch interface (same phases, inputs kept in HBM) ---------- */

/* upload once, run many times: lets a caller (and bench.py) separate PCIe from kernel time */
int32_t trgt_flank_upload(trgt_engine_t *eng, const trgt_seqs_t *left_pieces,
                          const trgt_seqs_t *right_pieces, const trgt_seqs_t *reads,
                          const uint32_t *locus_read_offsets, uint32_t n_loci,
                          trgt_scoring_t scoring, double min_flank_id_frac,
                          trgt_flank_batch_t **out);
/* the same resident batch from BAM 4-bit reads (decoded on the device at upload) */
int32_t trgt_flank_upload_seq4(trgt_engine_t *eng, const trgt_seqs_t *left_pieces,
                               const trgt_seqs_t *right_pieces, const trgt_seq4_t *reads,
                               const uint32_t *locus_read_offsets, uint32_t n_loci,
                               trgt_scoring_t scoring, double min_flank_id_frac,
                               trgt_flank_batch_t **out);
/* the *_run calls enqueue on the engine stream; they synchronise internally where a later launch
 * is sized by an earlier one (work-list length, workspace), but results are only guaranteed in
 * place after the matching *_download (or trgt_engine_sync) */
int32_t trgt_flank_run(trgt_engine_t *eng, trgt_flank_batch_t *batch);
int32_t trgt_flank_download(trgt_engine_t *eng, trgt_flank_batch_t *batch,
                            trgt_span_t *spans_out, trgt_flank_hit_t *hits_out);   /* syncs */
void trgt_flank_free(trgt_engine_t *eng, trgt_flank_batch_t *batch);
/* device pointers of a resident flank batch (reads, CSR offsets, spans, hits) for device-side
 * consumers; n_wfa = (read, flank) pairs the last run sent to the WFA fallback */
int32_t trgt_flank_device_views(trgt_flank_batch_t *batch, const void **d_reads, const void **d_read_off,
                                const void **d_spans, const void **d_hits, uint32_t *n_reads,
                                uint32_t *n_wfa);
/* how the last run settled the pairs that missed the exact search: out[0] = handed to the second cost
 * tier, out[1] = handed to the wide-band kernel, out[2] = needed the full-width kernels */
int32_t trgt_flank_fallback_counts(trgt_flank_batch_t *batch, uint32_t out[3]);

/* The repeat sequences the genotypers work on, `trs` of src/trgt/workflows/tr.rs:58-62
 * (read.bases[span.0..span.1] per spanning read), cut on the device from the reads of a flank batch
 * (batch == NULL: the last trgt_flank_spans / trgt_flank_spans_seq4 call on this engine).  One sequence
 * per read, empty for reads without a span; out->status is NULL.  out points into engine-owned pinned
 * memory that stays valid until the next trgt_flank_trs on the same batch. */
int32_t trgt_flank_trs(trgt_engine_t *eng, trgt_flank_batch_t *batch, trgt_seqs_out_t *out);

int32_t trgt_align_upload(trgt_engine_t *eng, const trgt_seqs_t *backbones,
                          const trgt_seqs_t *seqs, const uint32_t *group_seq_offsets,
                          uint32_t n_groups, trgt_align_batch_t **out);
int32_t trgt_align_run(trgt_engine_t *eng, trgt_align_batch_t *batch);
int32_t trgt_align_download(trgt_engine_t *eng, trgt_align_batch_t *batch, trgt_cigars_t *out);
void trgt_align_free(trgt_engine_t *eng, trgt_align_batch_t *batch);

int32_t trgt_hmm_upload(trgt_engine_t *eng, const trgt_seqs_t *motifs,
                        const uint32_t *locus_motif_offsets, uint32_t n_loci,
                        const trgt_seqs_t *alleles, const uint32_t *allele_locus,
                        int32_t want_paths, trgt_hmm_batch_t **out);
int32_t trgt_hmm_run(trgt_engine_t *eng, trgt_hmm_batch_t *batch);
int32_t trgt_hmm_download(trgt_engine_t *eng, trgt_hmm_batch_t *batch, trgt_annotations_t *out);
void trgt_hmm_free(trgt_engine_t *eng, trgt_hmm_batch_t *batch);

/* ---- VCF sample fields behind phase C (next row, rank 4) --------------------- */

/* The sample fields the engine's results determine, as the reference's writer encodes them
 * (src/trgt/writers/write_vcf.rs: encode_al :267-277, encode_mc :286-299, encode_ms :307-323,
 * encode_ap :332-343), for every locus of an HMM batch that has been run (batch == NULL: the last
 * trgt_hmm_label call).  The alleles of a locus must be consecutive in the batch, in genotype order.
 * out->n = 4 * n_loci strings, locus after locus in the order AL, MC, MS, AP ("33,33", "11,11",
 * "0(0-33),0(0-33)", "1.000000,1.000000" for the tutorial locus, docs/tutorial.md:44); a locus without
 * alleles gets four empty strings; out->status is NULL.  `{:.6}` is rounded half to even on the exact
 * binary value, as Rust prints it.  out points into engine-owned pinned memory valid until the next
 * trgt_vcf_fields on the same batch.  GT, ALLR, SD and AM come from the host genotyper. */
int32_t trgt_vcf_fields(trgt_engine_t *eng, trgt_hmm_batch_t *batch, trgt_seqs_out_t *out);

/* ---- instrumentation ------------------------------------------------------ */

/* When enabled, every kernel launch is bracketed by CUDA events on the engine stream. */
void trgt_engine_set_profiling(trgt_engine_t *eng, int32_t on);
void trgt_engine_reset_stats(trgt_engine_t *eng);
/* number of distinct kernels seen; i-th: name, launches, total device ms (syncs the stream) */
int32_t trgt_engine_kernel_count(trgt_engine_t *eng);
int32_t trgt_engine_kernel_stat(trgt_engine_t *eng, int32_t i, const char **name,
                                uint64_t *launches, double *total_ms);
/* launches since the last reset (counted whether or not profiling is on) */
uint64_t trgt_engine_launches(const trgt_engine_t *eng);

#ifdef __cplusplus
}
#endif
#endif
