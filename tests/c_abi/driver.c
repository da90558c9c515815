/* driver.c -- TEST INFRASTRUCTURE: a plain C99 caller of include/trgt_engine.h, linked against
 * trgt_b200/libtrgt_b200.so the way a foreign-language host (cgo, Rust FFI, JNI) would bind it.
 *   driver nogpu : trgt_engine_create must fail loudly (no CPU fallback), prints the error text
 *   driver gpu   : one locus, three reads, through trgt_flank_spans -> "span r found start end" lines */
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
  trgt_engine_destroy(eng);
  return 0;
}
