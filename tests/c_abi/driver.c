/* driver.c -- TEST INFRASTRUCTURE: a plain C99 caller of include/trgt_engine.h, linked against
 * trgt_b200/libtrgt_b200.so the way a foreign-language host (cgo, Rust FFI, JNI) would bind it.
 *   driver nogpu : trgt_engine_create must fail loudly (no CPU fallback), prints the error text
 *   driver gpu   : one locus, three reads, through trgt_flank_spans -> "span r found start end" lines, then every other
 *                  one-shot entry of the path on small inputs whose answers the reference's docs and tests give */
#include <stdio.h>
#include <string.h>

#include "../../include/trgt_engine.h"

static void fill(char *dst, int n, unsigned seed) {
  static const char acgt[] = "ACGT";
  for (int i = 0; i < n; i++) {
    seed = seed * 1664525u + 1013904223u;
    dst[i] = acgt[(seed >> 24) & 3];
  }
}

int main(int argc, char **argv) {
  trgt_engine_t *eng = NULL;
  const int32_t rc = trgt_engine_create(0, &eng);
  if (argc > 1 && strcmp(argv[1], "nogpu") == 0) {
    printf("create rc=%d error=%s\n", (int)rc, trgt_last_create_error());
    if (rc == TRGT_OK) trgt_engine_destroy(eng);
    return rc == TRGT_OK ? 1 : 0;
  }
  if (rc != TRGT_OK) {
    printf("create failed rc=%d: %s\n", (int)rc, trgt_last_create_error());
    return 2;
  }
  enum { P = 250, TR = 30, CTX = 100 };
  static char left[P], right[P], reads[3 * (2 * CTX + 2 * P + TR) + 16];
  fill(left, P, 1u);
  fill(right, P, 2u);
  uint64_t read_off[4] = {0, 0, 0, 0};
  char *w = reads;
  for (int r = 0; r < 3; r++) {
    fill(w, CTX, 10u + r); w += CTX;
    memcpy(w, left, P); w += P;
    if (r == 1) w[-100] = w[-100] == 'A' ? 'C' : 'A';       /* one mismatch: WFA fallback */
    for (int i = 0; i < TR + 3 * r; i++) *w++ = "CAG"[i % 3];
    if (r < 2) { memcpy(w, right, P); w += P; }              /* read 2 lacks the right flank */
    fill(w, CTX, 20u + r); w += CTX;
    read_off[r + 1] = (uint64_t)(w - reads);
  }
  const uint64_t piece_off[2] = {0, P};
  const uint32_t locus_read_off[2] = {0, 3};
  trgt_seqs_t lp = {(const uint8_t *)left, piece_off, 1}, rp = {(const uint8_t *)right, piece_off, 1};
  trgt_seqs_t rd = {(const uint8_t *)reads, read_off, 3};
  trgt_scoring_t sc = {2, 5, 1};
  trgt_span_t spans[3];
  trgt_flank_hit_t hits[6];
  const int32_t frc = trgt_flank_spans(eng, &lp, &rp, &rd, locus_read_off, 1, sc, 0.7, spans, hits);
  if (frc != TRGT_OK) {
    printf("trgt_flank_spans rc=%d: %s\n", (int)frc, trgt_engine_last_error(eng));
    return 3;
  }
  for (int r = 0; r < 3; r++)
    printf("span %d %d %u %u via %d %d\n", r, (int)spans[r].found, spans[r].start, spans[r].end, (int)hits[2 * r].via,
           (int)hits[2 * r + 1].via);
  printf("launches %llu\n", (unsigned long long)trgt_engine_launches(eng));

  /* the repeat sequences of the spanning reads, cut on the device from the batch of the call above */
  trgt_seqs_out_t trs;
  if (trgt_flank_trs(eng, NULL, &trs) != TRGT_OK) return 4;
  for (uint64_t r = 0; r < trs.n; r++)
    printf("tr %llu %.*s\n", (unsigned long long)r, (int)(trs.offsets[r + 1] - trs.offsets[r]),
           (const char *)trs.data + trs.offsets[r]);

  /* clip_bases for the BAMlet on the same batch: CIGAR = the whole read as one '=' run, 50 bases of flank kept */
  {
    uint32_t ops[3];
    uint64_t op_off[4] = {0, 1, 2, 3};
    int64_t ref_starts[3] = {1000, 2000, 3000};
    for (int r = 0; r < 3; r++) ops[r] = ((uint32_t)(read_off[r + 1] - read_off[r]) << 4) | 7u;
    trgt_bamlet_clip_t bc[3];
    if (trgt_bamlet_clip(eng, NULL, ops, op_off, ref_starts, 50, bc) != TRGT_OK) return 5;
    for (int r = 0; r < 3; r++)
      printf("bamlet %d status %d bases %u %u ref %lld ops %u first %u=\n", r, (int)bc[r].status, bc[r].base_start,
             bc[r].base_end, (long long)bc[r].ref_pos, bc[r].n_ops, bc[r].first_word >> 4);
  }

  /* phase B: utils::align of two members against their backbone, then repair_consensus, then get_dist_matrix */
  static const char bb[] = "CAGCAGCAGCAGCAGCAGCAGCAGCAGCAG";
  static const char members[] = "CAGCAGCAGCAGCAGCAGCAGCAGCAGCAG" "CAGCAGCAGCAGCAGCAGCAGCAGCAG" "CAGCAGCAGCATCAGCAGCAGCAGCAGCAG";
  const uint64_t bb_off[2] = {0, 30}, mem_off[4] = {0, 30, 57, 87};
  const uint32_t group_off[2] = {0, 3};
  trgt_seqs_t bbs = {(const uint8_t *)bb, bb_off, 1}, mem = {(const uint8_t *)members, mem_off, 3};
  trgt_cigars_t cig;
  if (trgt_align_e2e(eng, &bbs, &mem, group_off, 1, &cig) != TRGT_OK) return 6;
  for (uint64_t i = 0; i < cig.n; i++) {
    printf("cigar %llu score %d:", (unsigned long long)i, (int)cig.scores[i]);
    for (uint64_t w = cig.offsets[i]; w < cig.offsets[i + 1]; w++)
      printf(" %u%c", cig.words[w] >> 4, "MIDNSHP=X"[cig.words[w] & 15u]);
    printf("\n");
  }
  trgt_seqs_out_t cons;
  if (trgt_consensus(eng, &bbs, &mem, group_off, 1, &cons) != TRGT_OK) return 7;
  printf("consensus %.*s\n", (int)(cons.offsets[1] - cons.offsets[0]), (const char *)cons.data);
  double dist[3];
  if (trgt_edit_dist(eng, &mem, group_off, 1, dist) != TRGT_OK) return 8;
  printf("dist %.6f %.6f %.6f\n", dist[0], dist[1], dist[2]);

  /* phase C: label_with_hmm on the two alleles of the tutorial's locus, then the VCF sample fields */
  static const char motif[] = "CAG";
  static const char alleles[] = "CAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAG" "CAGCAGCAGCAGCAGCAGCAGCAGCAGCAGCAG";
  const uint64_t motif_off[2] = {0, 3}, allele_off[3] = {0, 33, 66};
  const uint32_t locus_motif_off[2] = {0, 1}, allele_locus[2] = {0, 0};
  trgt_seqs_t mo = {(const uint8_t *)motif, motif_off, 1}, al = {(const uint8_t *)alleles, allele_off, 2};
  trgt_annotations_t ann;
  if (trgt_hmm_label(eng, &mo, locus_motif_off, 1, &al, allele_locus, 0, &ann) != TRGT_OK) return 9;
  for (uint64_t a = 0; a < ann.n; a++)
    printf("allele %llu MC %u MS %u(%u-%u) AP %.6f\n", (unsigned long long)a, ann.motif_counts[ann.motif_count_offsets[a]],
           ann.spans[ann.span_offsets[a]].motif_index, ann.spans[ann.span_offsets[a]].start,
           ann.spans[ann.span_offsets[a]].end, ann.purity[a]);
  trgt_seqs_out_t vcf;
  if (trgt_vcf_fields(eng, NULL, &vcf) != TRGT_OK) return 10;
  printf("vcf");
  for (uint64_t f = 0; f < vcf.n; f++)
    printf(" %.*s", (int)(vcf.offsets[f + 1] - vcf.offsets[f]), (const char *)vcf.data + vcf.offsets[f]);
  printf("\n");

  /* producer: clip_reads on the reference's own example (clip_region.rs:257-269) */
  {
    const uint32_t cops[6] = {(3u << 4) | 7u, (2u << 4) | 2u, (2u << 4) | 7u, (1u << 4) | 8u, (2u << 4) | 7u, (5u << 4) | 1u};
    const uint64_t coff[2] = {0, 6};
    const int64_t cref[1] = {10}, region[2] = {12, 17};
    const uint32_t lro[2] = {0, 1};
    trgt_clip_t clip;
    if (trgt_clip_reads(eng, cops, coff, cref, 1, region, lro, 1, &clip) != TRGT_OK) return 11;
    printf("clip status %d ref %lld query %llu %llu ops %u\n", (int)clip.status, (long long)clip.ref_start,
           (unsigned long long)clip.query_start, (unsigned long long)clip.query_end, clip.n_ops);
  }
  trgt_engine_destroy(eng);
  return 0;
}
