// vcf_core.h -- the VCF sample fields the engine owns, encoded behind phase C.
//
// Replaces, for a whole chunk of loci at once, the per-allele string building of the reference's writer:
//   encode_al  src/trgt/writers/write_vcf.rs:267-277   allele lengths          "33,33"
//   encode_mc  :286-299                                  motif counts            "11_2,9_2"
//   encode_ms  :307-323                                  motif spans             "0(0-33)_1(33-39),."
//   encode_ap  :332-343                                  allele purity, {:.6}    "1.000000,."
// (GT, ALLR, SD and AM come from the host genotyper and stay there.)  With the fields encoded on the device
// the annotations never have to be unpacked on the host: phase C returns four byte strings per locus.
//
// One lane per locus, run twice (count -> exclusive scan -> write) through the same routine and a writer
// that either counts or stores.  `{:.6}` is exact: the double is scaled by 10^6 in 128-bit integer
// arithmetic and rounded half to even on its exact binary value, which is what Rust's (and glibc's)
// correctly rounded float formatting prints.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/trgt_engine.h"
#include "coop.h"

namespace trgt {

struct VcfWriter {
  uint8_t *out;  // nullptr: count only
  uint64_t n;
  TRGT_HD void put(char c) {
    if (out) out[n] = (uint8_t)c;
    n++;
  }
  TRGT_HD void put_u64(unsigned long long v) {
    char buf[20];
    int k = 0;
    do { buf[k++] = (char)('0' + (int)(v % 10ull)); v /= 10ull; } while (v);
    while (k) put(buf[--k]);
  }
};

// format!("{:.6}", v) for a finite v with |v| * 10^6 < 2^63; returns false otherwise (nothing written)
TRGT_HD bool vcf_put_fixed6(VcfWriter &w, double v) {
  unsigned long long bits;
  memcpy(&bits, &v, 8);
  const bool neg = (bits >> 63) != 0;
  const int be = (int)((bits >> 52) & 0x7FFull);
  unsigned long long m = bits & 0xFFFFFFFFFFFFFull;
  if (be == 0x7FF) return false;  // inf / nan
  int e;                          // v = m * 2^e
  if (be == 0) e = -1074; else { m |= 1ull << 52; e = be - 1075; }
  unsigned long long q;           // round_half_even(|v| * 10^6)
  if (m == 0) {
    q = 0;
  } else if (e >= 0) {
    return false;  // |v| >= 2^52: not a value any of these fields can take
  } else {
    const int k = -e;
    if (k >= 75) {
      q = 0;  // m * 10^6 < 2^73: the value is below 1/4
    } else {
      const unsigned __int128 p = (unsigned __int128)m * 1000000ull;
      const unsigned __int128 one = 1;
      unsigned __int128 qq = p >> k;
      const unsigned __int128 rem = p & ((one << k) - 1), half = one << (k - 1);
      if (rem > half || (rem == half && (qq & 1))) qq += 1;
      if (qq >> 63) return false;
      q = (unsigned long long)qq;
    }
  }
  if (neg) w.put('-');
  w.put_u64(q / 1000000ull);
  w.put('.');
  unsigned long long f = q % 1000000ull, d = 100000ull;
  for (int i = 0; i < 6; i++) { w.put((char)('0' + (int)(f / d))); f %= d; d /= 10ull; }
  return true;
}

struct VcfLocus {
  uint32_t a0, a1;                       // the locus' alleles, in genotype order
  const uint64_t *allele_off;            // [.. n_alleles+1]
  const unsigned long long *mc_off;      // [.. n_alleles+1]
  const uint32_t *mc;
  const unsigned long long *span_off;    // [.. n_alleles+1]
  const trgt_motif_span_t *spans;
  const double *purity;
  const int32_t *status;
};

// field 0 AL, 1 MC, 2 MS, 3 AP of one locus
TRGT_HD void vcf_encode_field(const VcfLocus &L, int field, VcfWriter &w) {
  for (uint32_t a = L.a0; a < L.a1; a++) {
    if (a > L.a0) w.put(',');
    const bool bad = L.status[a] != 0;  // the reference panics on such an allele (hmm_model.rs:250)
    if (field == 0) {
      w.put_u64(L.allele_off[a + 1] - L.allele_off[a]);
    } else if (field == 1) {
      if (bad) { w.put('.'); continue; }
      for (unsigned long long i = L.mc_off[a]; i < L.mc_off[a + 1]; i++) {
        if (i > L.mc_off[a]) w.put('_');
        w.put_u64(L.mc[i]);
      }
    } else if (field == 2) {
      if (bad || L.span_off[a + 1] == L.span_off[a]) { w.put('.'); continue; }  // labels: None
      for (unsigned long long i = L.span_off[a]; i < L.span_off[a + 1]; i++) {
        if (i > L.span_off[a]) w.put('_');
        w.put_u64(L.spans[i].motif_index); w.put('(');
        w.put_u64(L.spans[i].start); w.put('-');
        w.put_u64(L.spans[i].end); w.put(')');
      }
    } else {
      const double v = L.purity[a];
      if (bad || v != v || !vcf_put_fixed6(w, v)) w.put('.');
    }
  }
}

}  // namespace trgt
