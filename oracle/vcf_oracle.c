/*
 * vcf_oracle.c -- CPU restatement of the VCF sample fields the engine's results determine (TEST
 * INFRASTRUCTURE ONLY; see trgt_oracle.h).  Follows src/trgt/writers/write_vcf.rs: encode_al :267-277,
 * encode_mc :286-299, encode_ms :307-323, encode_ap :332-343.
 *
 * Pinned on the tutorial's record (docs/tutorial.md:44: AL 33,33  MC 11,11  MS 0(0-33),0(0-33)
 * AP 1.000000,1.000000).  `{:.6}` is glibc's "%.6f" here: both it and Rust's formatter print the correctly
 * rounded decimal of the exact binary value (ties to even); no reference test fixes a tie, so that detail is
 * parity-unpinned.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "trgt_oracle.h"

static size_t put(char *out, size_t cap, size_t n, const char *s) {
  const size_t l = strlen(s);
  if (out && n + l <= cap) memcpy(out + n, s, l);
  return n + l;
}

/* one field (0 AL, 1 MC, 2 MS, 3 AP) of one locus with n_alleles alleles; returns its length (the bytes are
 * written if they fit cap) */
size_t tro_vcf_field(int field, uint32_t n_alleles, const uint64_t *allele_len, const uint64_t *mc_off,
                     const uint32_t *mc, const uint64_t *span_off, const tro_span *spans, const double *purity,
                     char *out, size_t cap) {
  size_t n = 0;
  char buf[64];
  for (uint32_t a = 0; a < n_alleles; a++) {
    if (a) n = put(out, cap, n, ",");
    if (field == 0) { /* :267-277 */
      snprintf(buf, sizeof buf, "%llu", (unsigned long long)allele_len[a]);
      n = put(out, cap, n, buf);
    } else if (field == 1) { /* :286-299 */
      for (uint64_t i = mc_off[a]; i < mc_off[a + 1]; i++) {
        if (i > mc_off[a]) n = put(out, cap, n, "_");
        snprintf(buf, sizeof buf, "%u", mc[i]);
        n = put(out, cap, n, buf);
      }
    } else if (field == 2) { /* :307-323 */
      if (span_off[a + 1] == span_off[a]) { n = put(out, cap, n, "."); continue; }
      for (uint64_t i = span_off[a]; i < span_off[a + 1]; i++) {
        if (i > span_off[a]) n = put(out, cap, n, "_");
        snprintf(buf, sizeof buf, "%u(%u-%u)", spans[i].motif_index, spans[i].start, spans[i].end);
        n = put(out, cap, n, buf);
      }
    } else { /* :332-343 */
      if (isnan(purity[a])) { n = put(out, cap, n, "."); continue; }
      snprintf(buf, sizeof buf, "%.6f", purity[a]);
      n = put(out, cap, n, buf);
    }
  }
  return n;
}
