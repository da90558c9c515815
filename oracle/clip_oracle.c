/*
 * clip_oracle.c -- CPU restatement of read clipping and BAM base decoding (TEST INFRASTRUCTURE ONLY;
 * see trgt_oracle.h).  Follows src/trgt/reads/clip_region.rs:19-186 (clip_to_region, clip_cigar,
 * get_reference_end), the op lengths of src/trgt/reads/cigar.rs:8-31 and the base decoding of
 * src/trgt/reads/read.rs:104 (`rec.seq().as_bytes()`, i.e. htslib's 4-bit alphabet "=ACMGRSVTWYHKDBN",
 * first base in the high nibble).
 *
 * Pinned on the six unit tests of clip_region.rs:187-296 (tests/test_oracle_golden.py).
 */
#include <string.h>

#include "trgt_oracle.h"

/* BAM op codes: M I D N S H P = X  ->  0..8 (rust_htslib::bam::record::Cigar order) */
static int64_t op_ref_len(uint32_t w) { /* cigar.rs:9-19 */
  const uint32_t op = w & 15u;
  return (op == 0 || op == 3 || op == 2 || op == 7 || op == 8) ? (int64_t)(w >> 4) : 0;
}

static int64_t op_query_len(uint32_t w) { /* cigar.rs:21-31 */
  const uint32_t op = w & 15u;
  return (op == 0 || op == 7 || op == 8 || op == 1 || op == 4) ? (int64_t)(w >> 4) : 0;
}

static int op_splittable(uint32_t w) { /* clip_region.rs:143-150, 171-178: anything else panics */
  const uint32_t op = w & 15u;
  return op == 0 || op == 3 || op == 2 || op == 7 || op == 8;
}

/* clip_cigar: clip_region.rs:105-186, plus the query range clip_to_region :19-38 copies.
 * ops = BAM-encoded CIGAR ((len<<4)|op).  Returns 1 and fills *out when the alignment overlaps the
 * region, 0 when it does not (None), -1 where the reference panics ("Unexpected operation"). */
int tro_clip_cigar(const uint32_t *ops, uint32_t n_ops, int64_t ref_start, int64_t region_start,
                   int64_t region_end, tro_clip *out) {
  memset(out, 0, sizeof *out);
  int64_t read_end = ref_start; /* get_reference_end :84-90 */
  for (uint32_t i = 0; i < n_ops; i++) read_end += op_ref_len(ops[i]);
  if (read_end <= region_start || region_end <= ref_start) return 0; /* :109-111 */

  int64_t ref_pos = ref_start, query_pos = 0;
  uint32_t cur = 0;
  /* skip operations outside of the target region :122-126 */
  while (cur < n_ops && ref_pos + op_ref_len(ops[cur]) <= region_start) {
    ref_pos += op_ref_len(ops[cur]);
    query_pos += op_query_len(ops[cur]);
    cur++;
  }
  int64_t clipped_ref_start = ref_pos, clipped_query_start = query_pos;
  int64_t clipped_query_len = 0;
  uint32_t n_clipped = 0;
  out->first_op = cur;
#define PUSH(word)                                  \
  do {                                              \
    if (n_clipped == 0) out->first_word = (word);   \
    out->last_word = (word);                        \
    n_clipped++;                                    \
    clipped_query_len += op_query_len(word);        \
  } while (0)
  /* split operation overlapping the left flank :132-161 */
  if (ref_pos < region_start) {
    if (cur >= n_ops) return -1; /* current_op.unwrap() on None */
    const int64_t ref_outside_len = region_start - ref_pos;
    const int64_t op_len = op_ref_len(ops[cur]);
    const int64_t clipped_len = ref_pos + op_len <= region_end ? op_len - ref_outside_len : region_end - region_start;
    if (!op_splittable(ops[cur])) return -1;
    const uint32_t first = ((uint32_t)clipped_len << 4) | (ops[cur] & 15u);
    PUSH(first);
    clipped_ref_start += ref_outside_len;
    if (op_query_len(first) != 0) clipped_query_start += ref_outside_len;
    ref_pos += op_ref_len(ops[cur]);
    query_pos += op_query_len(ops[cur]);
    cur++;
  }
  /* copy operations contained within the region :164-169 */
  while (cur < n_ops && ref_pos + op_ref_len(ops[cur]) <= region_end) {
    PUSH(ops[cur]);
    ref_pos += op_ref_len(ops[cur]);
    query_pos += op_query_len(ops[cur]);
    cur++;
  }
  /* split operation overlapping the right flank :172-186 */
  if (cur < n_ops && ref_pos < region_end) {
    if (!op_splittable(ops[cur])) return -1;
    const uint32_t last = ((uint32_t)(region_end - ref_pos) << 4) | (ops[cur] & 15u);
    PUSH(last);
  }
#undef PUSH
  out->ref_start = clipped_ref_start;
  out->query_start = (uint64_t)clipped_query_start;
  out->query_end = (uint64_t)(clipped_query_start + clipped_query_len);
  out->n_ops = n_clipped;
  return 1;
}

/* rec.seq().as_bytes() restricted to bases [start, start+len): read.rs:104 then clip_region.rs:29-31 */
void tro_decode_seq4(const uint8_t *packed, uint64_t start, uint32_t len, uint8_t *out) {
  static const char alphabet[] = "=ACMGRSVTWYHKDBN";
  for (uint32_t i = 0; i < len; i++) {
    const uint64_t n = start + i;
    const uint8_t byte = packed[n >> 1];
    out[i] = (uint8_t)alphabet[(n & 1u) ? (byte & 15u) : (byte >> 4)];
  }
}

/* ---- BAMlet output: HiFiRead::clip_bases (src/trgt/reads/clip_bases.rs:9-119) as BamWriter::write calls it
 * (src/trgt/writers/write_bam.rs:72-92), and the CpG rule both clips share.
 * Pinned on the six unit tests of clip_bases.rs:121-229 and on the methylation vectors of the six
 * clip_region.rs tests (tests/test_oracle_golden.py). */

/* The i-th "CG" dinucleotide of a read (by the index of its C, 0 <= index < len - 1) owns meth[i]; a clip to
 * bases [start, end) keeps the entries whose C lies inside (clip_region.rs:40-58, clip_bases.rs:23-44).
 * -> the kept entries are meth[*m0 .. *m1). */
void tro_meth_range(const uint8_t *bases, uint64_t len, uint64_t start, uint64_t end, uint32_t *m0, uint32_t *m1) {
  uint32_t before = 0, inside = 0;
  for (uint64_t index = 0; index + 1 < len; index++) {
    if (bases[index] == 'C' && bases[index + 1] == 'G') {
      if (start <= index && index < end) inside++;
      else if (index < start) before++;
    }
  }
  *m0 = before;
  *m1 = before + inside;
}

static int op_has_query_kind(uint32_t w) { /* clip_bases.rs:73-81, 98-106: Match Diff Ins Equal SoftClip */
  const uint32_t op = w & 15u;
  return op == 0 || op == 8 || op == 1 || op == 7 || op == 4;
}

/* clip_bases.  n_ops == 0 <=> the read has no CIGAR (cigar: None).  1 = Some (out filled), 0 = None (:10-12),
 * -1 = the reference panics / exits (assert :61, unexpected operation :80,106). */
int tro_clip_bases(const uint32_t *ops, uint32_t n_ops, int64_t ref_pos_in, const uint8_t *bases, uint64_t len,
                   uint64_t left_len, uint64_t right_len, tro_bclip *out) {
  memset(out, 0, sizeof *out);
  if (left_len + right_len >= len) return 0; /* :10-12 */
  out->base_start = left_len;                /* :14-15: bases and quals */
  out->base_end = len - right_len;
  tro_meth_range(bases, len, left_len, len - right_len, &out->meth_start, &out->meth_end); /* :23-44 */
  out->has_cigar = n_ops != 0;
  if (n_ops == 0) return 1;
  /* clip_cigar :59-119 */
  uint64_t align_query_len = 0;
  for (uint32_t i = 0; i < n_ops; i++) align_query_len += (uint64_t)op_query_len(ops[i]);
  if (align_query_len < left_len + right_len) return -1; /* assert :61 */
  uint64_t keep_len = align_query_len - left_len - right_len;
  uint64_t left = left_len;
  uint32_t cur = 0;
  uint32_t cur_word = ops[0];
  int64_t ref_pos = ref_pos_in;
  while (left != 0) { /* :68-92 */
    if (cur >= n_ops) return -1; /* current_op.unwrap() on None */
    const uint64_t query_len = (uint64_t)op_query_len(cur_word);
    if (query_len > left) {
      if (!op_has_query_kind(cur_word)) return -1;
      cur_word = ((uint32_t)(query_len - left) << 4) | (cur_word & 15u);
      if (op_ref_len(cur_word) != 0) ref_pos += (int64_t)left;
      left = 0;
    } else {
      left -= query_len;
      ref_pos += op_ref_len(cur_word);
      cur++;
      cur_word = cur < n_ops ? ops[cur] : 0;
    }
  }
  uint32_t n_clipped = 0;
  out->first_op = cur;
  while (cur < n_ops && keep_len != 0) { /* :94-113 */
    const uint64_t query_len = (uint64_t)op_query_len(cur_word);
    uint32_t word;
    if (query_len > keep_len) {
      if (!op_has_query_kind(cur_word)) return -1;
      word = ((uint32_t)keep_len << 4) | (cur_word & 15u);
      keep_len = 0;
    } else {
      keep_len -= query_len;
      word = cur_word;
      cur++;
      cur_word = cur < n_ops ? ops[cur] : 0;
    }
    if (n_clipped == 0) out->first_word = word;
    out->last_word = word;
    n_clipped++;
  }
  out->n_ops = n_clipped;
  out->ref_pos = ref_pos;
  return 1;
}

/* The clip BamWriter::write asks for (write_bam.rs:80-92): flank_len bases either side of the repeat span.
 * 1 = record written (out filled), 0 = skipped with "unexpectedly short flanks" (:80-83, :88-91), -1 = panic. */
int tro_bamlet_clip(const uint32_t *ops, uint32_t n_ops, int64_t ref_pos, const uint8_t *bases, uint64_t len,
                    uint64_t span_start, uint64_t span_end, uint64_t flank_len, tro_bclip *out) {
  memset(out, 0, sizeof *out);
  if (span_start < flank_len || len < span_end + flank_len) return 0;
  return tro_clip_bases(ops, n_ops, ref_pos, bases, len, span_start - flank_len, len - span_end - flank_len, out);
}
