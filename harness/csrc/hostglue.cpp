// hostglue.cpp -- host-side stand-in for the reference's genotyper between the GPU phases (part of
// libtrgt_harness.so, used identically by the GPU arm and the CPU reference arm of bench.py).
//
// In the reference the scalar genotype logic (src/trgt/genotype/*) sits between span location and
// consensus alignment / HMM annotation and is out of scope here.  For synthetic reads whose
// haplotype of origin is known, this glue does what that logic would hand on: per locus and
// haplotype the repeat sequences read[span.start..span.end] of the spanning reads
// (tr.rs:139-165), the first of them as the backbone every member is aligned to
// (genotype_cluster.rs:41-56 -> utils::align), and the backbones as the allele sequences the
// HMM annotates (tr.rs:77).  Pure byte shuffling, multi-threaded over loci.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

extern "C" {

typedef struct {
  int32_t found;
  uint32_t start, end;
} glue_span;

typedef struct {
  uint32_t n_groups, n_seqs;
  uint8_t *bb; uint64_t *bb_off;      // [n_groups+1]
  uint8_t *seqs; uint64_t *seq_off;   // [n_seqs+1]
  uint32_t *group_seq_off;            // [n_groups+1]
  uint32_t *group_locus;              // [n_groups]
  uint32_t *seq_read;                 // [n_seqs]
} glue_out;

// Grow-only buffers reused from pass to pass; with the engine's trgt_host_alloc / trgt_host_free as
// allocator they are pinned, so phases B and C upload from them at full PCIe rate.
typedef void *(*glue_alloc_fn)(size_t);
typedef void (*glue_free_fn)(void *);
typedef struct {
  glue_alloc_fn alloc;
  glue_free_fn release;
  void *buf[7];
  size_t cap[7];
} glue_ctx;

}  // extern "C"

namespace {

template <class F>
void par_loci(uint32_t n, uint32_t threads, F f) {
  uint32_t nt = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
  if (nt > 64) nt = 64;
  if (nt > n) nt = n ? n : 1;
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < nt; t++)
    th.emplace_back([&, t]() { f((uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt)); });
  for (auto &x : th) x.join();
}

}  // namespace

namespace {

void *ctx_reserve(glue_ctx *c, int slot, size_t bytes) {
  if (!c) return malloc(bytes);
  if (c->buf[slot] && c->cap[slot] >= bytes) return c->buf[slot];
  if (c->buf[slot]) c->release(c->buf[slot]);
  const size_t ncap = bytes + bytes / 4 + 4096;
  c->buf[slot] = c->alloc(ncap);
  c->cap[slot] = c->buf[slot] ? ncap : 0;
  return c->buf[slot];
}

}  // namespace

extern "C" {

glue_ctx *glue_ctx_create(glue_alloc_fn a, glue_free_fn f) {
  glue_ctx *c = (glue_ctx *)calloc(1, sizeof(glue_ctx));
  c->alloc = a ? a : (glue_alloc_fn)malloc;
  c->release = f ? f : (glue_free_fn)free;
  return c;
}

void glue_ctx_destroy(glue_ctx *c) {
  if (!c) return;
  for (int i = 0; i < 7; i++) if (c->buf[i]) c->release(c->buf[i]);
  free(c);
}

// The reads as consecutive BAM records would hold them (trgt_seq4_t of include/trgt_engine.h): read r
// starts in the high nibble of a byte when r is even and in the low nibble when r is odd, as clips at even
// and odd query positions do.  Two calls: out == NULL returns the packed size in bytes.
uint64_t seq4_pack(const uint8_t *reads, const uint64_t *read_off, uint64_t n_reads, uint8_t *out, uint64_t *starts,
                   uint32_t *lengths, uint32_t threads) {
  std::vector<uint64_t> boff((size_t)n_reads + 1, 0);
  for (uint64_t r = 0; r < n_reads; r++) boff[r + 1] = boff[r] + ((r & 1u) + (read_off[r + 1] - read_off[r]) + 1) / 2;
  if (!out) return boff[n_reads];
  uint8_t code[256];
  memset(code, 15, sizeof code);
  const char *alphabet = "=ACMGRSVTWYHKDBN";
  for (int i = 0; i < 16; i++) code[(uint8_t)alphabet[i]] = (uint8_t)i;
  par_loci((uint32_t)n_reads, threads, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t r = lo; r < hi; r++) {
      const uint8_t *src = reads + read_off[r];
      const uint64_t len = read_off[r + 1] - read_off[r];
      const uint64_t par = r & 1u;
      uint8_t *dst = out + boff[r];
      memset(dst, 0, (size_t)(boff[r + 1] - boff[r]));
      for (uint64_t i = 0; i < len; i++) {
        const uint64_t nib = par + i;
        dst[nib >> 1] |= (nib & 1u) ? code[src[i]] : (uint8_t)(code[src[i]] << 4);
      }
      starts[r] = 2 * boff[r] + par;
      lengths[r] = (uint32_t)len;
    }
  });
  return boff[n_reads];
}

namespace {
// bases [pos, pos+len) of read r: from the ASCII reads, or decoded from the BAM 4-bit bases
// (rec.seq().as_bytes(), read.rs:104, restricted to the repeat -- the host never decodes the flanks)
struct ReadSource {
  const uint8_t *reads; const uint64_t *read_off;
  const uint8_t *seq4; const uint64_t *starts;
  bool trs;  // reads / read_off hold the repeat sequences themselves (trgt_flank_trs), not whole reads
  void copy(uint8_t *dst, uint32_t r, uint64_t pos, uint64_t len) const {
    if (trs) { memcpy(dst, reads + read_off[r], len); return; }
    if (!seq4) { memcpy(dst, reads + read_off[r] + pos, len); return; }
    static const char alphabet[] = "=ACMGRSVTWYHKDBN";
    uint64_t nib = starts[r] + pos;
    for (uint64_t i = 0; i < len; i++, nib++) {
      const uint8_t b = seq4[nib >> 1];
      dst[i] = (uint8_t)alphabet[(nib & 1u) ? (b & 15u) : (b >> 4)];
    }
  }
};
}  // namespace

// ctx == NULL: plain malloc, release with glue_free; otherwise the outputs live in ctx's buffers and
// stay valid until the next glue_build on that ctx.  seq4 != NULL: the reads are BAM 4-bit bases
// (seq4, starts) and `reads` / `read_off` are ignored.  trs_mode != 0: `reads` / `read_off` are the repeat
// sequences of trgt_flank_trs (one per read, empty without a span).
int glue_build(const uint8_t *reads, const uint64_t *read_off, const uint32_t *locus_read_off, uint32_t n_loci,
               const glue_span *spans, const uint8_t *read_hap, uint32_t threads, glue_ctx *ctx, glue_out *out,
               const uint8_t *seq4, const uint64_t *seq4_starts, int trs_mode) {
  const ReadSource rs{reads, read_off, seq4, seq4_starts, trs_mode != 0};
  // pass 1: per-locus counts
  std::vector<uint32_t> g_cnt((size_t)n_loci + 1, 0), s_cnt((size_t)n_loci + 1, 0);
  std::vector<uint64_t> sb_cnt((size_t)n_loci + 1, 0), bb_cnt((size_t)n_loci + 1, 0);
  par_loci(n_loci, threads, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t l = lo; l < hi; l++) {
      bool have[2] = {false, false};
      uint32_t ns = 0, ng = 0;
      uint64_t sb = 0, bb = 0;
      for (uint32_t r = locus_read_off[l]; r < locus_read_off[l + 1]; r++) {
        if (!spans[r].found) continue;
        const int h = read_hap[r] ? 1 : 0;
        const uint64_t len = spans[r].end - spans[r].start;
        if (!have[h]) { have[h] = true; ng++; bb += len; }
        ns++;
        sb += len;
      }
      g_cnt[l] = ng; s_cnt[l] = ns; sb_cnt[l] = sb; bb_cnt[l] = bb;
    }
  });
  std::vector<uint32_t> g0((size_t)n_loci + 1, 0), s0((size_t)n_loci + 1, 0);
  std::vector<uint64_t> sb0((size_t)n_loci + 1, 0), bb0((size_t)n_loci + 1, 0);
  for (uint32_t l = 0; l < n_loci; l++) {
    g0[l + 1] = g0[l] + g_cnt[l];
    s0[l + 1] = s0[l] + s_cnt[l];
    sb0[l + 1] = sb0[l] + sb_cnt[l];
    bb0[l + 1] = bb0[l] + bb_cnt[l];
  }
  const uint32_t ng = g0[n_loci], ns = s0[n_loci];
  out->n_groups = ng;
  out->n_seqs = ns;
  out->bb = (uint8_t *)ctx_reserve(ctx, 0, (size_t)bb0[n_loci] + 16);
  out->bb_off = (uint64_t *)ctx_reserve(ctx, 1, sizeof(uint64_t) * ((size_t)ng + 1));
  out->seqs = (uint8_t *)ctx_reserve(ctx, 2, (size_t)sb0[n_loci] + 16);
  out->seq_off = (uint64_t *)ctx_reserve(ctx, 3, sizeof(uint64_t) * ((size_t)ns + 1));
  out->group_seq_off = (uint32_t *)ctx_reserve(ctx, 4, sizeof(uint32_t) * ((size_t)ng + 1));
  out->group_locus = (uint32_t *)ctx_reserve(ctx, 5, sizeof(uint32_t) * ((size_t)ng + 1));
  out->seq_read = (uint32_t *)ctx_reserve(ctx, 6, sizeof(uint32_t) * ((size_t)ns + 1));
  if (!out->bb || !out->bb_off || !out->seqs || !out->seq_off || !out->group_seq_off || !out->group_locus || !out->seq_read)
    return -1;
  out->bb_off[ng] = bb0[n_loci];
  out->seq_off[ns] = sb0[n_loci];
  out->group_seq_off[ng] = ns;
  // pass 2: fill, haplotype 0's group first
  par_loci(n_loci, threads, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t l = lo; l < hi; l++) {
      uint32_t g = g0[l], s = s0[l];
      uint64_t sb = sb0[l], bb = bb0[l];
      for (int h = 0; h < 2; h++) {
        bool first = true;
        for (uint32_t r = locus_read_off[l]; r < locus_read_off[l + 1]; r++) {
          if (!spans[r].found || (read_hap[r] ? 1 : 0) != h) continue;
          const uint64_t len = spans[r].end - spans[r].start;
          if (first) {
            first = false;
            out->bb_off[g] = bb;
            out->group_seq_off[g] = s;
            out->group_locus[g] = l;
            rs.copy(out->bb + bb, r, spans[r].start, len);
            bb += len;
            g++;
          }
          out->seq_off[s] = sb;
          out->seq_read[s] = r;
          if (out->group_seq_off[g - 1] == s) memcpy(out->seqs + sb, out->bb + bb - len, len);  // the backbone is the first member
          else rs.copy(out->seqs + sb, r, spans[r].start, len);
          sb += len;
          s++;
        }
      }
    }
  });
  return 0;
}

void glue_free(glue_out *o) {
  free(o->bb); free(o->bb_off); free(o->seqs); free(o->seq_off);
  free(o->group_seq_off); free(o->group_locus); free(o->seq_read);
  memset(o, 0, sizeof(*o));
}

}  // extern "C"
