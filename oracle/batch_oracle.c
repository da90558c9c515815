/*
 * batch_oracle.c -- batched, multi-threaded drivers over the CPU oracle (TEST INFRASTRUCTURE ONLY;
 * see trgt_oracle.h).  They take the same packed inputs as the engine's C ABI and fill the same
 * packed outputs, so (a) full-size parity checks are plain array comparisons and (b) bench.py can
 * time the reference-equivalent CPU path on the box's host cores the way the reference runs it:
 * one task per locus over a thread pool (src/commands/genotype.rs:178-187).
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "trgt_oracle.h"

typedef void (*tro_task_fn)(void *ctx, uint64_t lo, uint64_t hi);

typedef struct {
  tro_task_fn fn;
  void *ctx;
  uint64_t n, chunk;
  atomic_ullong next;
} tro_pool;

static void *pool_worker(void *arg) {
  tro_pool *p = (tro_pool *)arg;
  for (;;) {
    const uint64_t lo = atomic_fetch_add(&p->next, p->chunk);
    if (lo >= p->n) break;
    const uint64_t hi = lo + p->chunk < p->n ? lo + p->chunk : p->n;
    p->fn(p->ctx, lo, hi);
  }
  return NULL;
}

static void parallel_for(uint64_t n, uint64_t chunk, int n_threads, tro_task_fn fn, void *ctx) {
  tro_pool p;
  p.fn = fn; p.ctx = ctx; p.n = n; p.chunk = chunk ? chunk : 1;
  atomic_init(&p.next, 0);
  if (n_threads <= 1) {
    pool_worker(&p);
    return;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  int started = 0;
  for (int t = 0; t < n_threads; t++)
    if (pthread_create(&th[t], NULL, pool_worker, &p) == 0) started++; else break;
  if (started == 0) pool_worker(&p);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(th);
}

/* ---------------------------------------------------------------- phase A -- */

typedef struct {
  const uint8_t *left; const uint64_t *left_off;
  const uint8_t *right; const uint64_t *right_off;
  const uint8_t *reads; const uint64_t *read_off;
  const uint32_t *locus_read_off;
  int x, o, e;
  double frac;
  tro_opt_span *spans;
  tro_flank_hit *hits;
} flank_ctx;

static void flank_task(void *vctx, uint64_t lo, uint64_t hi) {
  flank_ctx *c = (flank_ctx *)vctx;
  for (uint64_t l = lo; l < hi; l++) {
    const uint8_t *lp = c->left + c->left_off[l], *rp = c->right + c->right_off[l];
    const int lpl = (int)(c->left_off[l + 1] - c->left_off[l]), rpl = (int)(c->right_off[l + 1] - c->right_off[l]);
    for (uint32_t r = c->locus_read_off[l]; r < c->locus_read_off[l + 1]; r++) {
      const uint8_t *seq = c->reads + c->read_off[r];
      const int sl = (int)(c->read_off[r + 1] - c->read_off[r]);
      int via[2], nm[2];
      const tro_opt_span a = tro_find_span(lp, lpl, seq, sl, c->x, c->o, c->e, (double)lpl * c->frac, &via[0], &nm[0]);
      const tro_opt_span b = tro_find_span(rp, rpl, seq, sl, c->x, c->o, c->e, (double)rpl * c->frac, &via[1], &nm[1]);
      c->spans[r] = tro_combine_spans(a, b);
      if (c->hits) {
        const tro_opt_span s[2] = {a, b};
        for (int k = 0; k < 2; k++) {
          tro_flank_hit *h = &c->hits[2 * (uint64_t)r + k];
          h->via = via[k]; h->matches = nm[k]; h->score = 0;
          h->start = s[k].found ? s[k].start : 0;
          h->end = s[k].found ? s[k].end : 0;
        }
      }
    }
  }
}

int tro_flank_batch(const uint8_t *left, const uint64_t *left_off, const uint8_t *right, const uint64_t *right_off,
                    const uint8_t *reads, const uint64_t *read_off, const uint32_t *locus_read_off, uint32_t n_loci,
                    int x, int o, int e, double min_flank_id_frac, tro_opt_span *spans_out, tro_flank_hit *hits_out,
                    int n_threads) {
  flank_ctx c = {left, left_off, right, right_off, reads, read_off, locus_read_off, x, o, e, min_flank_id_frac,
                 spans_out, hits_out};
  parallel_for(n_loci, 4, n_threads, flank_task, &c);
  return 0;
}

/* ---------------------------------------------------------------- phase B -- */

typedef struct {
  const uint8_t *bb; const uint64_t *bb_off;
  const uint8_t *seqs; const uint64_t *seq_off;
  const uint32_t *group_off;
  uint32_t **words;   /* per sequence, malloc'ed */
  uint32_t *n_words;
  int32_t *scores;
} align_ctx;

static void align_task(void *vctx, uint64_t lo, uint64_t hi) {
  align_ctx *c = (align_ctx *)vctx;
  for (uint64_t g = lo; g < hi; g++) {
    const uint8_t *bb = c->bb + c->bb_off[g];
    const int bl = (int)(c->bb_off[g + 1] - c->bb_off[g]);
    for (uint32_t s = c->group_off[g]; s < c->group_off[g + 1]; s++) {
      const uint8_t *seq = c->seqs + c->seq_off[s];
      const int sl = (int)(c->seq_off[s + 1] - c->seq_off[s]);
      const uint64_t cap = (uint64_t)bl + (uint64_t)sl + 2;
      uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * cap);
      int score = 0;
      const int64_t n = tro_align_consensus(bb, bl, seq, sl, buf, cap, &score);
      c->words[s] = buf;
      c->n_words[s] = n > 0 ? (uint32_t)n : 0;
      c->scores[s] = score;
    }
  }
}

/* offsets_out[n_seqs+1]; *words_out is malloc'ed (free with tro_free) */
int tro_align_batch(const uint8_t *bb, const uint64_t *bb_off, const uint8_t *seqs, const uint64_t *seq_off,
                    const uint32_t *group_off, uint32_t n_groups, uint64_t *offsets_out, uint32_t **words_out,
                    int32_t *scores_out, int n_threads) {
  const uint32_t n_seqs = n_groups ? group_off[n_groups] : 0;
  align_ctx c;
  c.bb = bb; c.bb_off = bb_off; c.seqs = seqs; c.seq_off = seq_off; c.group_off = group_off;
  c.words = (uint32_t **)calloc(n_seqs ? n_seqs : 1, sizeof(uint32_t *));
  c.n_words = (uint32_t *)calloc(n_seqs ? n_seqs : 1, sizeof(uint32_t));
  c.scores = scores_out;
  parallel_for(n_groups, 2, n_threads, align_task, &c);
  uint64_t tot = 0;
  for (uint32_t s = 0; s < n_seqs; s++) {
    offsets_out[s] = tot;
    tot += c.n_words[s];
  }
  offsets_out[n_seqs] = tot;
  uint32_t *w = (uint32_t *)malloc(sizeof(uint32_t) * (tot ? tot : 1));
  for (uint32_t s = 0; s < n_seqs; s++) {
    if (c.n_words[s]) memcpy(w + offsets_out[s], c.words[s], sizeof(uint32_t) * c.n_words[s]);
    free(c.words[s]);
  }
  free(c.words);
  free(c.n_words);
  *words_out = w;
  return 0;
}

/* ---------------------------------------------------------------- phase C -- */

typedef struct {
  const uint8_t *motifs; const uint64_t *motif_off; const uint32_t *locus_motif_off;
  const uint8_t *alleles; const uint64_t *allele_off; const uint32_t *allele_locus;
  const uint64_t *mc_off;
  uint32_t *mc;
  tro_span **spans;
  uint32_t *n_spans;
  double *purity;
  int32_t *status;
} hmm_ctx;

static void hmm_task(void *vctx, uint64_t lo, uint64_t hi) {
  hmm_ctx *c = (hmm_ctx *)vctx;
  tro_hmm *h = NULL;
  uint32_t h_locus = 0xFFFFFFFFu;
  for (uint64_t a = lo; a < hi; a++) {
    const uint32_t l = c->allele_locus[a];
    if (!h || l != h_locus) {  /* build_hmm once per locus, tr.rs:461 */
      tro_hmm_free(h);
      const uint32_t m0 = c->locus_motif_off[l], nm = c->locus_motif_off[l + 1] - m0;
      /* replace_invalid_bases(motif, ATCGN), tr.rs:455-460 */
      const uint64_t b0 = c->motif_off[m0], b1 = c->motif_off[m0 + nm];
      uint8_t *mb = (uint8_t *)malloc((size_t)(b1 - b0) + 1);
      uint32_t *mo = (uint32_t *)malloc(sizeof(uint32_t) * (nm + 1));
      memcpy(mb, c->motifs + b0, (size_t)(b1 - b0));
      for (uint32_t m = 0; m <= nm; m++) mo[m] = (uint32_t)(c->motif_off[m0 + m] - b0);
      for (uint32_t m = 0; m < nm; m++) tro_replace_invalid_bases(mb + mo[m], mo[m + 1] - mo[m], "ATCGN");
      h = tro_hmm_build(mb, mo, nm);
      free(mb);
      free(mo);
      h_locus = l;
    }
    const uint8_t *al = c->alleles + c->allele_off[a];
    const uint32_t L = (uint32_t)(c->allele_off[a + 1] - c->allele_off[a]);
    c->n_spans[a] = 0;
    c->spans[a] = NULL;
    if (!h) {
      c->status[a] = -400;
      continue;
    }
    tro_span *sp = (tro_span *)malloc(sizeof(tro_span) * ((size_t)L + 1));
    double pur = 0.0;
    const int64_t n = tro_annotate_allele(h, al, L, c->mc + c->mc_off[a], sp, (uint64_t)L + 1, &pur);
    c->purity[a] = pur;
    c->status[a] = n < 0 ? (int32_t)n : 0;
    c->n_spans[a] = n > 0 ? (uint32_t)n : 0;
    c->spans[a] = sp;
  }
  tro_hmm_free(h);
}

/* mc_off[n_alleles+1] given by the caller (motif count of each allele's locus, prefix-summed);
 * span_off_out[n_alleles+1]; *spans_out malloc'ed (free with tro_free) */
int tro_hmm_batch(const uint8_t *motifs, const uint64_t *motif_off, const uint32_t *locus_motif_off, uint32_t n_loci,
                  const uint8_t *alleles, const uint64_t *allele_off, const uint32_t *allele_locus, uint32_t n_alleles,
                  const uint64_t *mc_off, uint32_t *mc_out, uint64_t *span_off_out, tro_span **spans_out,
                  double *purity_out, int32_t *status_out, int n_threads) {
  (void)n_loci;
  hmm_ctx c;
  c.motifs = motifs; c.motif_off = motif_off; c.locus_motif_off = locus_motif_off;
  c.alleles = alleles; c.allele_off = allele_off; c.allele_locus = allele_locus;
  c.mc_off = mc_off; c.mc = mc_out;
  c.spans = (tro_span **)calloc(n_alleles ? n_alleles : 1, sizeof(tro_span *));
  c.n_spans = (uint32_t *)calloc(n_alleles ? n_alleles : 1, sizeof(uint32_t));
  c.purity = purity_out; c.status = status_out;
  parallel_for(n_alleles, 16, n_threads, hmm_task, &c);
  uint64_t tot = 0;
  for (uint32_t a = 0; a < n_alleles; a++) {
    span_off_out[a] = tot;
    tot += c.n_spans[a];
  }
  span_off_out[n_alleles] = tot;
  tro_span *sp = (tro_span *)malloc(sizeof(tro_span) * (tot ? tot : 1));
  for (uint32_t a = 0; a < n_alleles; a++) {
    if (c.n_spans[a]) memcpy(sp + span_off_out[a], c.spans[a], sizeof(tro_span) * c.n_spans[a]);
    free(c.spans[a]);
  }
  free(c.spans);
  free(c.n_spans);
  *spans_out = sp;
  return 0;
}

void tro_free(void *p) { free(p); }
