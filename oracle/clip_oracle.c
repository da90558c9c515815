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
