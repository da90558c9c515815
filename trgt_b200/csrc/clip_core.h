// clip_core.h -- read clipping and BAM base decoding (the producer of phase A's input).
//
// Replaces, for a whole chunk of loci at once,
//   * clip_cigar / the query range of HiFiRead::clip_to_region   src/trgt/reads/clip_region.rs:19-38,105-186
//     (caller clip_reads, src/trgt/workflows/tr.rs:186-196: region = locus region +- 2*search_flank_len);
//   * `rec.seq().as_bytes()`                                      src/trgt/reads/read.rs:104
//     (htslib's 4-bit alphabet "=ACMGRSVTWYHKDBN", first base in the high nibble).
// The reference decodes every whole read to ASCII on the host and copies the clipped part out.  Here
// the host hands the engine the clipped part of the BAM record as it is stored (two bases per byte):
// half the bytes cross PCIe, and k_unpack_seq4 writes the ASCII reads of phase A straight into HBM.
// The clip walk touches only the CIGAR (a few dozen words per read): one thread per read.
#pragma once
#include <stdint.h>

#include "../../include/trgt_engine.h"
#include "coop.h"

namespace trgt {

// BAM op codes: M I D N S H P = X -> 0..8
TRGT_HD long long clip_ref_len(uint32_t w) {  // cigar.rs:9-19
  const uint32_t op = w & 15u;
  return ((0x18Du >> op) & 1u) ? (long long)(w >> 4) : 0;  // M D N = X
}
TRGT_HD long long clip_query_len(uint32_t w) {  // cigar.rs:21-31
  const uint32_t op = w & 15u;
  return ((0x193u >> op) & 1u) ? (long long)(w >> 4) : 0;  // M I S = X
}

// clip_region.rs:105-186 on one read.  status: 1 overlap, 0 None, TRGT_ITEM_INVALID_OP where the
// reference panics (an op without reference length would have to be split).
TRGT_HD trgt_clip_t clip_cigar_one(const uint32_t *ops, uint32_t n_ops, long long ref_start, long long region_start,
                                   long long region_end) {
  trgt_clip_t c;
  c.ref_start = 0; c.query_start = 0; c.query_end = 0; c.first_op = 0; c.n_ops = 0; c.first_word = 0; c.last_word = 0;
  c.status = 0;
  long long read_end = ref_start;  // get_reference_end :84-90
  for (uint32_t i = 0; i < n_ops; i++) read_end += clip_ref_len(ops[i]);
  if (read_end <= region_start || region_end <= ref_start) return c;
  long long ref_pos = ref_start, query_pos = 0;
  uint32_t cur = 0;
  while (cur < n_ops) {  // operations left of the region :122-126
    const uint32_t w = ops[cur];
    if (ref_pos + clip_ref_len(w) > region_start) break;
    ref_pos += clip_ref_len(w);
    query_pos += clip_query_len(w);
    cur++;
  }
  long long c_ref = ref_pos, c_q = query_pos, q_len = 0;
  uint32_t n = 0, first = 0, last = 0;
  c.first_op = cur;
  if (ref_pos < region_start) {  // operation split by the region start :132-161
    if (cur >= n_ops) { c.status = TRGT_ITEM_INVALID_OP; return c; }
    const uint32_t w = ops[cur];
    const long long outside = region_start - ref_pos, len = clip_ref_len(w);
    const long long keep = ref_pos + len <= region_end ? len - outside : region_end - region_start;
    first = last = ((uint32_t)keep << 4) | (w & 15u);
    n = 1;
    c_ref += outside;
    if (clip_query_len(first) != 0) c_q += outside;
    q_len += clip_query_len(first);
    ref_pos += len;
    query_pos += clip_query_len(w);
    cur++;
  }
  while (cur < n_ops) {  // operations inside the region :164-169
    const uint32_t w = ops[cur];
    if (ref_pos + clip_ref_len(w) > region_end) break;
    if (n == 0) first = w;
    last = w;
    n++;
    q_len += clip_query_len(w);
    ref_pos += clip_ref_len(w);
    query_pos += clip_query_len(w);
    cur++;
  }
  if (cur < n_ops && ref_pos < region_end) {  // operation split by the region end :172-186
    const uint32_t w = ops[cur];
    if (clip_ref_len(w) == 0) { c.status = TRGT_ITEM_INVALID_OP; return c; }
    const uint32_t cut = ((uint32_t)(region_end - ref_pos) << 4) | (w & 15u);
    if (n == 0) first = cut;
    last = cut;
    n++;
    q_len += clip_query_len(cut);
  }
  c.ref_start = c_ref;
  c.query_start = (uint64_t)c_q;
  c.query_end = (uint64_t)(c_q + q_len);
  c.n_ops = n; c.first_word = first; c.last_word = last;
  c.status = 1;
  return c;
}

// ---------------------------------------------------------------- BAM 4-bit bases -> ASCII ----

// "=ACMGRSVTWYHKDBN": letters of codes 0..7 and 8..15, code c in byte (c & 7)
#define TRGT_SEQ4_LO 0x565352474D43413Dull
#define TRGT_SEQ4_HI 0x4E42444B48595754ull

TRGT_HD uint8_t seq4_letter(uint32_t code) {
  return (uint8_t)(((code & 8u) ? TRGT_SEQ4_HI : TRGT_SEQ4_LO) >> (8u * (code & 7u)));
}

// base i of a BAM sequence (first base in the high nibble)
TRGT_HD uint32_t seq4_code_at(const uint8_t *data, long long nib) {
  const uint8_t b = data[nib >> 1];
  return (nib & 1) ? (b & 15u) : (uint32_t)(b >> 4);
}

struct alignas(16) Seq4Word {
  uint32_t x, y, z, w;
};

// swap the two nibbles of every byte: afterwards base i of the word sits in bits [4i, 4i+4)
TRGT_HD uint64_t seq4_swap(uint64_t v) {
  return ((v & 0x0F0F0F0F0F0F0F0Full) << 4) | ((v >> 4) & 0x0F0F0F0F0F0F0F0Full);
}

// four codes (bits [4i,4i+4) of the low 16 bits) -> four ASCII bytes, base i in byte i
TRGT_HD uint32_t seq4_decode4(uint32_t g) {
#if defined(__CUDA_ARCH__)
  const uint32_t sel = g & 0x7777u;
  const uint32_t lo = __byte_perm((uint32_t)TRGT_SEQ4_LO, (uint32_t)(TRGT_SEQ4_LO >> 32), sel);
  const uint32_t hi = __byte_perm((uint32_t)TRGT_SEQ4_HI, (uint32_t)(TRGT_SEQ4_HI >> 32), sel);
  // bit 3 of code i -> bit 0 of byte i -> the whole byte (no carries: 1 * 0xFF fits a byte)
  const uint32_t x = ((g & 0x8u) >> 3) | ((g & 0x80u) << 1) | ((g & 0x800u) << 5) | ((g & 0x8000u) << 9);
  const uint32_t m = x * 0xFFu;
  return (lo & ~m) | (hi & m);
#else
  return (uint32_t)seq4_letter(g & 15u) | ((uint32_t)seq4_letter((g >> 4) & 15u) << 8) |
         ((uint32_t)seq4_letter((g >> 8) & 15u) << 16) | ((uint32_t)seq4_letter((g >> 12) & 15u) << 24);
#endif
}

// Sixteen ASCII bases starting at nibble index `nib` (may be negative or run past the read: the
// packed buffer carries 16 bytes of padding on both sides).  out[0..3] = bases 0-3, 4-7, 8-11, 12-15.
TRGT_HD void seq4_decode16(const uint8_t *data, long long nib, uint32_t out[4]) {
  const long long byte = nib >> 1;  // floor, also for negative nib
  const uint8_t *a = data + byte;
  const unsigned mis = (unsigned)((uintptr_t)a & 7u);
  const uint64_t *w = (const uint64_t *)(a - mis);
  const uint64_t w0 = seq4_swap(w[0]), w1 = seq4_swap(w[1]);
  const unsigned sh = mis * 8u + (unsigned)(nib & 1) * 4u;  // < 64
  const uint64_t v = sh ? ((w0 >> sh) | (w1 << (64u - sh))) : w0;
  out[0] = seq4_decode4((uint32_t)v & 0xFFFFu);
  out[1] = seq4_decode4((uint32_t)(v >> 16) & 0xFFFFu);
  out[2] = seq4_decode4((uint32_t)(v >> 32) & 0xFFFFu);
  out[3] = seq4_decode4((uint32_t)(v >> 48) & 0xFFFFu);
}

// One read: bases [start, start+len) of `data` -> out[o .. o+len).  The aligned 16-byte words that lie
// wholly inside the read are written one per lane and step (no divergence); the < 32 bytes left at the
// two ends are then written one byte per lane.
template <class G>
TRGT_HD void seq4_unpack_read(const G &g, const uint8_t *data, uint64_t start, uint32_t len, uint8_t *out_base,
                              uint64_t o) {
  if (len == 0) return;
  const uint32_t head = (uint32_t)((16u - (uint32_t)(o & 15u)) & 15u);  // bytes before the first aligned word
  const uint32_t nfull = len > head ? (len - head) >> 4 : 0;           // aligned words wholly inside
  uint8_t *dst = out_base + o + head;                                   // 16-byte aligned
  const long long nib0 = (long long)start + head;
  for (uint32_t c = (uint32_t)g.lane(); c < nfull; c += (uint32_t)g.size()) {
    uint32_t q[4];
    seq4_decode16(data, nib0 + 16ll * c, q);
    Seq4Word v;
    v.x = q[0]; v.y = q[1]; v.z = q[2]; v.w = q[3];
    *(Seq4Word *)(dst + 16u * c) = v;
  }
  // ends: bases [0, head) and [head + 16 * nfull, len)
  const uint32_t h = head < len ? head : len;
  const uint32_t tail0 = head + 16u * nfull;
  const uint32_t ntail = len > tail0 ? len - tail0 : 0;
  for (uint32_t k = (uint32_t)g.lane(); k < h + ntail; k += (uint32_t)g.size()) {
    const uint32_t i = k < h ? k : tail0 + (k - h);
    out_base[o + i] = seq4_letter(seq4_code_at(data, (long long)start + i));
  }
}

// ---------------------------------------------------------------- BAMlet output: clip_bases ---
//
// HiFiRead::clip_bases (src/trgt/reads/clip_bases.rs:9-119) for the clip BamWriter::write asks for
// (src/trgt/writers/write_bam.rs:72-92).  The clipped reads and their spans are resident after phase A, so only
// the CIGAR walk (one lane) and the count of "CG" dinucleotides left of / inside the kept bases (the whole
// group: entry i of the methylation profile belongs to the read's i-th CG, clip_bases.rs:23-44) are computed.

TRGT_HD bool bamlet_query_kind(uint32_t w) {  // clip_bases.rs:73-81, 98-106: Match Diff Ins Equal SoftClip
  return ((0x193u >> (w & 15u)) & 1u) != 0;
}

// clip_cigar of clip_bases.rs:59-119 on one read; fills ref_pos / first_op / n_ops / first_word / last_word,
// returns false where the reference panics or exits
TRGT_HD bool bamlet_clip_cigar(const uint32_t *ops, uint32_t n_ops, long long ref_pos, unsigned long long left,
                               unsigned long long right, trgt_bamlet_clip_t *c) {
  unsigned long long qlen = 0;
  for (uint32_t i = 0; i < n_ops; i++) qlen += (unsigned long long)clip_query_len(ops[i]);
  if (qlen < left + right) return false;  // assert :61
  unsigned long long keep = qlen - left - right;
  uint32_t cur = 0, w = ops[0];
  while (left != 0) {  // :68-92
    if (cur >= n_ops) return false;
    const unsigned long long q = (unsigned long long)clip_query_len(w);
    if (q > left) {  // split the operation
      if (!bamlet_query_kind(w)) return false;
      w = ((uint32_t)(q - left) << 4) | (w & 15u);
      if (clip_ref_len(w) != 0) ref_pos += (long long)left;
      left = 0;
    } else {
      left -= q;
      ref_pos += clip_ref_len(w);
      cur++;
      w = cur < n_ops ? ops[cur] : 0u;
    }
  }
  uint32_t n = 0;
  c->first_op = cur;
  while (cur < n_ops && keep != 0) {  // :94-113
    const unsigned long long q = (unsigned long long)clip_query_len(w);
    uint32_t word;
    if (q > keep) {
      if (!bamlet_query_kind(w)) return false;
      word = ((uint32_t)keep << 4) | (w & 15u);
      keep = 0;
    } else {
      keep -= q;
      word = w;
      cur++;
      w = cur < n_ops ? ops[cur] : 0u;
    }
    if (n == 0) c->first_word = word;
    c->last_word = word;
    n++;
  }
  c->n_ops = n;
  c->ref_pos = ref_pos;
  return true;
}

// One read by one group of lanes.  found / span: the read's repeat span from phase A.
template <class G>
TRGT_HD trgt_bamlet_clip_t bamlet_clip_read(const G &g, const uint8_t *bases, uint32_t len, bool found, uint32_t span_start,
                                            uint32_t span_end, uint32_t flank_len, const uint32_t *ops, uint32_t n_ops,
                                            long long ref_pos) {
  trgt_bamlet_clip_t c;
  c.ref_pos = 0; c.base_start = 0; c.base_end = 0; c.meth_start = 0; c.meth_end = 0; c.first_op = 0; c.n_ops = 0;
  c.first_word = 0; c.last_word = 0; c.status = TRGT_BAMLET_SKIPPED; c.pad = 0;
  if (!found || span_start < flank_len || (unsigned long long)len < (unsigned long long)span_end + flank_len) return c;  // write_bam.rs:80-83
  const uint32_t left = span_start - flank_len, right = len - span_end - flank_len;
  if ((unsigned long long)left + right >= len) return c;  // clip_bases.rs:10-12 -> None -> :88-91
  c.base_start = left;
  c.base_end = len - right;
  int before = 0, inside = 0;  // "CG" whose C lies left of / inside [base_start, base_end)
  for (uint32_t i = (uint32_t)g.lane(); i + 1 < len; i += (uint32_t)g.size()) {
    const int cg = bases[i] == 'C' && bases[i + 1] == 'G';
    before += cg && i < c.base_start;
    inside += cg && i >= c.base_start && i < c.base_end;
  }
  { int tot; g.excl_scan_i(before, &tot); before = tot; }
  { int tot; g.excl_scan_i(inside, &tot); inside = tot; }
  c.meth_start = (uint32_t)before;
  c.meth_end = (uint32_t)(before + inside);
  c.status = TRGT_BAMLET_WRITE;
  if (n_ops != 0 && !bamlet_clip_cigar(ops, n_ops, ref_pos, left, right, &c)) c.status = TRGT_ITEM_INVALID_OP;
  return c;
}

}  // namespace trgt
