/*
 * wfa_oracle.c -- CPU restatement of the wavefront alignment the reference drives through
 * src/wfaligner.rs (TEST INFRASTRUCTURE ONLY; see trgt_oracle.h).
 *
 * The DP lives in WFA2-lib (wfa2-sys 0.1.0 @ git rev 4342b3b06278656fa51c0a33b4eb0b67d53bfa8c,
 * Cargo.toml:36; sources NOT under /root/reference).  This file restates the published WFA
 * recurrences (Marco-Sola et al. 2021) with the tie-breaking and ends-free conventions pinned
 * by the reference's own golden tests (src/wfaligner.rs:1137-1828); rules in SURVEY.md 8c:
 *   - offsets are text positions h on diagonal k = h - v;
 *   - I[s][k] = max(M[s-o-e][k-1], I[s-e][k-1]) + 1;  D[s][k] = max(M[s-o-e][k+1], D[s-e][k+1]);
 *     M[s][k] = max(M[s-x][k]+1, I[s][k], D[s][k]), NULL when h>T or v>P or negative; then extend;
 *   - termination scans diagonals low k -> high k;
 *   - backtrace picks max over (offset<<4 | tag), tags M=9 > D2e=8 > D2o=7 > D1e=6 > D1o=5 >
 *     I2e=4 > I2o=3 > I1e=2 > I1o=1.
 * Full wavefront history is kept (MemoryHigh equivalent); exact, no heuristic.
 */
#include "trgt_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OFF_NULL (INT_MIN / 2)

enum { C_M = 0, C_I1 = 1, C_D1 = 2, C_I2 = 3, C_D2 = 4, N_COMP = 5 };
enum { T_I1O = 1, T_I1E = 2, T_I2O = 3, T_I2E = 4, T_D1O = 5, T_D1E = 6, T_D2O = 7, T_D2E = 8, T_M = 9 };

typedef struct {
  int lo, hi;        /* inclusive; lo > hi <=> null */
  int *off[N_COMP];  /* off[c][k - lo] */
  int elo, ehi;      /* what a heuristic cut-off left of [lo, hi] (cells outside are nulled); = lo, hi without one */
} wf_t;

typedef struct {
  wf_t *wf;
  int n, cap;
} wf_hist;

static inline int wf_get(const wf_hist *H, int s, int c, int k) {
  if (s < 0 || s >= H->n) return OFF_NULL;
  const wf_t *w = &H->wf[s];
  if (k < w->lo || k > w->hi || !w->off[c]) return OFF_NULL;
  return w->off[c][k - w->lo];
}
static inline int wf_present(const wf_hist *H, int s, int c) {
  return s >= 0 && s < H->n && H->wf[s].elo <= H->wf[s].ehi && H->wf[s].off[c] != NULL;
}
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* Wavefront offsets come from a per-thread slab that is reused from one alignment to the next (as WFA2-lib's
 * mm_allocator does), not from one malloc per wavefront component. */
typedef struct slab_chunk {
  struct slab_chunk *next;
  size_t cap, used;
} slab_chunk;
static __thread slab_chunk *tl_slab = NULL;    /* chunks of this thread, the first one is the current one */
static __thread slab_chunk *tl_slab_free = NULL;

static int *slab_ints(size_t n) {
  const size_t bytes = (n * sizeof(int) + 15) & ~(size_t)15;
  if (!tl_slab || tl_slab->used + bytes > tl_slab->cap) {
    slab_chunk *c = NULL, **pp = &tl_slab_free;
    while (*pp && (*pp)->cap < bytes) pp = &(*pp)->next;
    if (*pp) {
      c = *pp;
      *pp = c->next;
    } else {
      const size_t cap = bytes > ((size_t)4 << 20) ? bytes : ((size_t)4 << 20);
      c = (slab_chunk *)malloc(sizeof(slab_chunk) + cap);
      c->cap = cap;
    }
    c->used = 0;
    c->next = tl_slab;
    tl_slab = c;
  }
  int *p = (int *)((char *)(tl_slab + 1) + tl_slab->used);
  tl_slab->used += bytes;
  return p;
}
static void slab_reset(void) { /* everything back to the free list; the memory stays with the thread */
  while (tl_slab) {
    slab_chunk *c = tl_slab;
    tl_slab = c->next;
    c->next = tl_slab_free;
    tl_slab_free = c;
  }
}

static wf_t *hist_push(wf_hist *H) {
  if (H->n == H->cap) {
    H->cap = H->cap ? H->cap * 2 : 64;
    H->wf = (wf_t *)realloc(H->wf, sizeof(wf_t) * (size_t)H->cap);
  }
  wf_t *w = &H->wf[H->n++];
  memset(w, 0, sizeof(*w));
  w->lo = 1;
  w->hi = -1;
  w->elo = 1;
  w->ehi = -1;
  return w;
}
static void hist_free(wf_hist *H) {
  slab_reset();
  free(H->wf);
}

static void extend(wf_t *w, const uint8_t *p, int P, const uint8_t *t, int T) {
  for (int k = w->lo; k <= w->hi; k++) {
    int h = w->off[C_M][k - w->lo];
    if (h < 0) continue;
    int v = h - k;
    while (v + 8 <= P && h + 8 <= T) { /* eight bases per step, as WFA2-lib's extend kernel compares blocks */
      uint64_t a, b;
      memcpy(&a, p + v, 8);
      memcpy(&b, t + h, 8);
      const uint64_t d = a ^ b;
      if (d) {
        const int m = __builtin_ctzll(d) >> 3;
        v += m;
        h += m;
        goto done;
      }
      v += 8;
      h += 8;
    }
    while (v < P && h < T && p[v] == t[h]) {
      v++;
      h++;
    }
  done:
    w->off[C_M][k - w->lo] = h;
  }
}

/* WFA2-lib's adaptive wavefront reduction (Marco-Sola et al. 2021, section 2.4; `wfadaptive(min_wavefront_length,
 * max_distance_threshold, steps_between_cutoffs)`, the default heuristic of wavefront_aligner_attr_default with
 * (10, 50, 1) -- WFA2-lib is not vendored in the reference tree, so this is a restatement of the published rule,
 * PARITY UNPINNED, used only to MEASURE how far the reference's default-heuristic aligners can stray from exact
 * WFA).  After the wavefront of a score has been extended and found not to terminate:
 * if it spans at least min_len diagonals, every diagonal's remaining distance max(P - v, T - h) is compared with
 * the smallest one and diagonals more than max_dist behind it are cut off from both ends, never across the
 * target diagonal T - P; the I / D components are clamped to the M range. */
static void wfadaptive_cutoff(wf_t *w, int P, int T, int min_len, int max_dist) {
  if (w->elo > w->ehi) return;
  if (w->ehi - w->elo + 1 < min_len) return;
  int min_d = INT_MAX;
  for (int k = w->elo; k <= w->ehi; k++) {
    const int h = w->off[C_M][k - w->lo];
    if (h < 0) continue;
    const int v = h - k, d = imax(P - v, T - h);
    if (d < min_d) min_d = d;
  }
  if (min_d == INT_MAX) return;
  const int ak = T - P;
  int lo = w->elo, hi = w->ehi;
  for (int k = w->elo; k < ak && k <= w->ehi; k++) {
    const int h = w->off[C_M][k - w->lo];
    const int d = h < 0 ? INT_MAX : imax(P - (h - k), T - h);
    if (d != INT_MAX && d - min_d <= max_dist) break;
    lo++;
  }
  for (int k = w->ehi; k > ak && k >= lo; k--) {
    const int h = w->off[C_M][k - w->lo];
    const int d = h < 0 ? INT_MAX : imax(P - (h - k), T - h);
    if (d != INT_MAX && d - min_d <= max_dist) break;
    hi--;
  }
  for (int k = w->lo; k <= w->hi; k++)
    if (k < lo || k > hi)
      for (int c = 0; c < N_COMP; c++)
        if (w->off[c]) w->off[c][k - w->lo] = OFF_NULL;
  w->elo = lo;
  w->ehi = hi;
}

/* returns 1 and sets (k, offset) for the first (lowest k) diagonal that satisfies the end condition */
static int terminated(const wf_t *w, const tro_wfa_params *prm, int P, int T, int *ek, int *eo) {
  if (w->lo > w->hi) return 0;
  if (!prm->ends_free) {
    const int k = T - P;
    if (k < w->lo || k > w->hi) return 0;
    const int h = w->off[C_M][k - w->lo];
    if (h >= T) {
      *ek = k;
      *eo = h;
      return 1;
    }
    return 0;
  }
  for (int k = w->lo; k <= w->hi; k++) {
    const int h = w->off[C_M][k - w->lo];
    if (h < 0) continue;
    const int v = h - k;
    if (v < 0 || v > P || h > T) continue;
    if ((h >= T && P - v <= prm->pattern_end_free) || (v >= P && T - h <= prm->text_end_free)) {
      *ek = k;
      *eo = h;
      return 1;
    }
  }
  return 0;
}

static inline int64_t pg(int off, int tag) {
  return off < 0 ? (int64_t)OFF_NULL : (((int64_t)off << 4) | tag);
}
static inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

typedef struct {
  uint8_t *buf;
  int64_t n, cap;
} opvec;
static void op_push(opvec *v, uint8_t op, int64_t count) {
  if (count <= 0) return;
  if (v->n + count > v->cap) {
    while (v->n + count > v->cap) v->cap = v->cap ? v->cap * 2 : 256;
    v->buf = (uint8_t *)realloc(v->buf, (size_t)v->cap);
  }
  memset(v->buf + v->n, op, (size_t)count);
  v->n += count;
}

int tro_wfa_align(const tro_wfa_params *prm, const uint8_t *p, int P, const uint8_t *t, int T,
                  tro_wfa_result *res) {
  memset(res, 0, sizeof(*res));
  const int metric = prm->metric;
  const int x = (metric == TRO_EDIT) ? 1 : prm->x;
  /* per-metric source distances */
  int o1e1, e1, o2e2 = 0, e2 = 0;
  if (metric == TRO_INDEL || metric == TRO_EDIT) {
    o1e1 = 1;
    e1 = 0;
  } else if (metric == TRO_LINEAR) {
    o1e1 = prm->e1;
    e1 = 0;
  } else {
    o1e1 = prm->o1 + prm->e1;
    e1 = prm->e1;
    o2e2 = prm->o2 + prm->e2;
    e2 = prm->e2;
  }
  const int affine = (metric == TRO_AFFINE || metric == TRO_AFFINE2P);
  const int two = (metric == TRO_AFFINE2P);
  const int use_x = (metric != TRO_INDEL);

  wf_hist H = {0};
  /* initial wavefront: src/wfaligner.rs:464-487 semantics */
  {
    wf_t *w = hist_push(&H);
    if (prm->ends_free) {
      w->lo = -prm->pattern_begin_free;
      w->hi = prm->text_begin_free;
    } else {
      w->lo = 0;
      w->hi = 0;
    }
    w->elo = w->lo;
    w->ehi = w->hi;
    w->off[C_M] = slab_ints((size_t)(w->hi - w->lo + 1));
    for (int k = w->lo; k <= w->hi; k++) w->off[C_M][k - w->lo] = k >= 0 ? k : 0;
    extend(w, p, P, t, T);
  }
  int s = 0, ek = 0, eo = 0;
  int status = TRO_STATUS_OK;
  for (;;) {
    if (terminated(&H.wf[s], prm, P, T, &ek, &eo)) break;
    if (prm->wfadaptive_min_len > 0 && !prm->ends_free)
      wfadaptive_cutoff(&H.wf[s], P, T, prm->wfadaptive_min_len, prm->wfadaptive_max_dist);
    s++;
    if (prm->max_steps > 0 && s > prm->max_steps) {
      status = TRO_STATUS_MAX_STEPS;
      break;
    }
    wf_t *w = hist_push(&H);
    /* gather sources */
    int lo = INT_MAX, hi = INT_MIN;
#define SRC(ss, c)                                 \
  if (wf_present(&H, (ss), (c))) {                 \
    lo = imin(lo, H.wf[(ss)].elo);                 \
    hi = imax(hi, H.wf[(ss)].ehi);                 \
  }
    if (use_x) SRC(s - x, C_M);
    SRC(s - o1e1, C_M);
    if (affine) {
      SRC(s - e1, C_I1);
      SRC(s - e1, C_D1);
    }
    if (two) {
      SRC(s - o2e2, C_M);
      SRC(s - e2, C_I2);
      SRC(s - e2, C_D2);
    }
#undef SRC
    if (lo > hi) continue; /* null wavefront */
    lo -= 1;
    hi += 1;
    /* diagonals outside [-P, T] can never hold a valid M offset and cannot feed one */
    lo = imax(lo, -P);
    hi = imin(hi, T);
    if (lo > hi) continue;
    w->lo = lo;
    w->hi = hi;
    w->elo = lo;
    w->ehi = hi;
    const size_t width = (size_t)(hi - lo + 1);
    w->off[C_M] = slab_ints(width);
    if (affine) {
      w->off[C_I1] = slab_ints(width);
      w->off[C_D1] = slab_ints(width);
    }
    if (two) {
      w->off[C_I2] = slab_ints(width);
      w->off[C_D2] = slab_ints(width);
    }
    for (int k = lo; k <= hi; k++) {
      int ins, del;
      if (affine) {
        const int i1 = imax(wf_get(&H, s - o1e1, C_M, k - 1), wf_get(&H, s - e1, C_I1, k - 1)) + 1;
        const int d1 = imax(wf_get(&H, s - o1e1, C_M, k + 1), wf_get(&H, s - e1, C_D1, k + 1));
        w->off[C_I1][k - lo] = i1;
        w->off[C_D1][k - lo] = d1;
        ins = i1;
        del = d1;
        if (two) {
          const int i2 = imax(wf_get(&H, s - o2e2, C_M, k - 1), wf_get(&H, s - e2, C_I2, k - 1)) + 1;
          const int d2 = imax(wf_get(&H, s - o2e2, C_M, k + 1), wf_get(&H, s - e2, C_D2, k + 1));
          w->off[C_I2][k - lo] = i2;
          w->off[C_D2][k - lo] = d2;
          ins = imax(ins, i2);
          del = imax(del, d2);
        }
      } else {
        ins = wf_get(&H, s - o1e1, C_M, k - 1) + 1;
        del = wf_get(&H, s - o1e1, C_M, k + 1);
      }
      const int mm = use_x ? wf_get(&H, s - x, C_M, k) + 1 : OFF_NULL;
      int mx = imax(mm, imax(ins, del));
      const int h = mx, v = mx - k;
      if (mx < 0 || h > T || v > P || v < 0) mx = OFF_NULL;
      w->off[C_M][k - lo] = mx;
    }
    extend(w, p, P, t, T);
  }

  res->status = status;
  if (status != TRO_STATUS_OK) {
    res->score = INT_MIN; /* wfaligner.rs:1448: failed alignments report i32::MIN */
    hist_free(&H);
    return status;
  }
  res->score = (metric == TRO_INDEL || metric == TRO_EDIT) ? s : -s;
  res->end_k = ek;
  res->end_offset = eo;
  if (prm->score_only) {
    hist_free(&H);
    return status;
  }

  /* ---- backtrace (reverse order, then flipped) ---- */
  opvec rev = {0};
  int k = ek, off = eo;
  int v = off - k, h = off;
  const int tail_d = P - v, tail_i = T - h; /* ends-free: unaligned pattern tail as D, text tail as I */
  int mt = C_M, sc = s;
  while (v > 0 && h > 0 && sc > 0) {
    const int ms = sc - x, go1 = sc - o1e1, ge1 = sc - e1, go2 = sc - o2e2, ge2 = sc - e2;
    int64_t mx;
    if (affine) {
      const int64_t c_m = pg(wf_get(&H, ms, C_M, k) + 1, T_M);
      const int64_t c_i1o = pg(wf_get(&H, go1, C_M, k - 1) + 1, T_I1O);
      const int64_t c_i1e = pg(wf_get(&H, ge1, C_I1, k - 1) + 1, T_I1E);
      const int64_t c_d1o = pg(wf_get(&H, go1, C_M, k + 1), T_D1O);
      const int64_t c_d1e = pg(wf_get(&H, ge1, C_D1, k + 1), T_D1E);
      int64_t c_i2o = OFF_NULL, c_i2e = OFF_NULL, c_d2o = OFF_NULL, c_d2e = OFF_NULL;
      if (two) {
        c_i2o = pg(wf_get(&H, go2, C_M, k - 1) + 1, T_I2O);
        c_i2e = pg(wf_get(&H, ge2, C_I2, k - 1) + 1, T_I2E);
        c_d2o = pg(wf_get(&H, go2, C_M, k + 1), T_D2O);
        c_d2e = pg(wf_get(&H, ge2, C_D2, k + 1), T_D2E);
      }
      switch (mt) {
        case C_M:
          mx = max64(c_m, max64(max64(max64(c_i1o, c_i1e), max64(c_i2o, c_i2e)),
                                max64(max64(c_d1o, c_d1e), max64(c_d2o, c_d2e))));
          break;
        case C_I1: mx = max64(c_i1o, c_i1e); break;
        case C_I2: mx = max64(c_i2o, c_i2e); break;
        case C_D1: mx = max64(c_d1o, c_d1e); break;
        default: mx = max64(c_d2o, c_d2e); break;
      }
    } else {
      const int64_t c_m = use_x ? pg(wf_get(&H, ms, C_M, k) + 1, T_M) : (int64_t)OFF_NULL;
      const int64_t c_io = pg(wf_get(&H, go1, C_M, k - 1) + 1, T_I1O);
      const int64_t c_do = pg(wf_get(&H, go1, C_M, k + 1), T_D1O);
      mx = max64(c_m, max64(c_io, c_do));
    }
    if (mx < 0) break; /* no predecessor: cannot happen for a completed alignment */
    if (mt == C_M) {
      const int mo = (int)(mx >> 4);
      op_push(&rev, 'M', off - mo);
      off = mo;
      v = off - k;
      h = off;
      if (v <= 0 || h <= 0) break;
    }
    switch ((int)(mx & 15)) {
      case T_M: sc = ms; mt = C_M; op_push(&rev, 'X', 1); off -= 1; break;
      case T_I1O: sc = go1; mt = C_M; op_push(&rev, 'I', 1); k -= 1; off -= 1; break;
      case T_I1E: sc = ge1; mt = C_I1; op_push(&rev, 'I', 1); k -= 1; off -= 1; break;
      case T_I2O: sc = go2; mt = C_M; op_push(&rev, 'I', 1); k -= 1; off -= 1; break;
      case T_I2E: sc = ge2; mt = C_I2; op_push(&rev, 'I', 1); k -= 1; off -= 1; break;
      case T_D1O: sc = go1; mt = C_M; op_push(&rev, 'D', 1); k += 1; break;
      case T_D1E: sc = ge1; mt = C_D1; op_push(&rev, 'D', 1); k += 1; break;
      case T_D2O: sc = go2; mt = C_M; op_push(&rev, 'D', 1); k += 1; break;
      case T_D2E: sc = ge2; mt = C_D2; op_push(&rev, 'D', 1); k += 1; break;
      default: break;
    }
    v = off - k;
    h = off;
  }
  if (mt == C_M && v > 0 && h > 0) {
    const int n = imin(v, h);
    op_push(&rev, 'M', n);
    v -= n;
    h -= n;
  }
  op_push(&rev, 'D', v);
  op_push(&rev, 'I', h);

  const int64_t n_ops = rev.n + (tail_d > 0 ? tail_d : 0) + (tail_i > 0 ? tail_i : 0);
  uint8_t *ops = (uint8_t *)malloc((size_t)(n_ops ? n_ops : 1));
  for (int64_t i = 0; i < rev.n; i++) ops[i] = rev.buf[rev.n - 1 - i];
  int64_t w = rev.n;
  for (int i = 0; i < tail_d; i++) ops[w++] = 'D';
  for (int i = 0; i < tail_i; i++) ops[w++] = 'I';
  free(rev.buf);
  res->ops = ops;
  res->n_ops = n_ops;
  hist_free(&H);
  return status;
}

void tro_wfa_result_free(tro_wfa_result *res) {
  free(res->ops);
  res->ops = NULL;
  res->n_ops = 0;
}

/* cigar_count_matches (wfaligner.rs:988-1000) */
int tro_count_matches(const uint8_t *ops, int64_t n) {
  int c = 0;
  for (int64_t i = 0; i < n; i++) c += (ops[i] == 'M');
  return c;
}

/* get_alignment_span: wfaligner.rs:864-908 */
void tro_alignment_span(const uint8_t *ops, int64_t n, int ends_free, int plen, int tlen, int *xs,
                        int *xe, int *ys, int *ye) {
  if (!ends_free) {
    *xs = 0; *xe = plen; *ys = 0; *ye = tlen;
    return;
  }
  int pi = 0, ti = 0, started = 0;
  *xs = *xe = *ys = *ye = 0;
  for (int64_t i = 0; i < n; i++) {
    switch (ops[i]) {
      case 'I': ti++; break;
      case 'D': pi++; break;
      default:
        if (!started) {
          *xs = pi;
          *ys = ti;
          started = 1;
        }
        pi++;
        ti++;
        *xe = pi;
        *ye = ti;
    }
  }
}

/* get_sam_cigar (wfaligner.rs:932-959 -> WFA2 cigar_get_CIGAR): run-length, (len<<4)|op */
int64_t tro_sam_cigar(const uint8_t *ops, int64_t n, int show_mismatches, uint32_t *out,
                      uint64_t cap) {
  uint64_t n_out = 0;
  int64_t i = 0;
  while (i < n) {
    uint8_t op = ops[i];
    if (!show_mismatches && op == 'X') op = 'M';
    int64_t j = i + 1;
    while (j < n) {
      uint8_t o2 = ops[j];
      if (!show_mismatches && o2 == 'X') o2 = 'M';
      if (o2 != op) break;
      j++;
    }
    uint32_t code;
    switch (op) {
      case 'M': code = show_mismatches ? 7u : 0u; break;
      case 'X': code = 8u; break;
      case 'I': code = 1u; break;
      default: code = 2u; break;
    }
    if (n_out >= cap) return -2;
    out[n_out++] = ((uint32_t)(j - i) << 4) | code;
    i = j;
  }
  return (int64_t)n_out;
}

static int run_score(const tro_wfa_params *p, uint8_t op, int len) {
  /* wfaligner.rs:534-593; match penalty is 0 for every configuration the path builds */
  switch (p->metric) {
    case TRO_INDEL:
    case TRO_EDIT: return op == 'M' ? 0 : len;
    case TRO_LINEAR: return op == 'M' ? 0 : (op == 'X' ? len * p->x : len * p->e1);
    case TRO_AFFINE: return op == 'M' ? 0 : (op == 'X' ? len * p->x : p->o1 + p->e1 * len);
    default: {
      if (op == 'M') return 0;
      if (op == 'X') return len * p->x;
      const int s1 = p->o1 + p->e1 * len, s2 = p->o2 + p->e2 * len;
      return s1 < s2 ? s1 : s2;
    }
  }
}

static int score_range(const tro_wfa_params *p, const uint8_t *ops, int64_t b, int64_t e) {
  int score = 0;
  int64_t i = b;
  while (i < e) {
    int64_t j = i + 1;
    while (j < e && ops[j] == ops[i]) j++;
    score += run_score(p, ops[i], (int)(j - i));
    i = j;
  }
  return (p->metric == TRO_INDEL || p->metric == TRO_EDIT) ? score : -score;
}

int tro_cigar_score(const tro_wfa_params *p, const uint8_t *ops, int64_t n) {
  return score_range(p, ops, 0, n);
}

/* cigar_score_clipped: wfaligner.rs:595-705 */
int tro_cigar_score_clipped(const tro_wfa_params *p, const uint8_t *ops, int64_t n, int flank_len) {
  const int64_t b = flank_len;
  int64_t e = n - flank_len;
  if (e < b) e = b;
  if (b >= e) return 0;
  return score_range(p, ops, b, e);
}

/* ----------------------------------------------------------- callers -- */

tro_opt_span tro_find_span(const uint8_t *piece, int piece_len, const uint8_t *seq, int seq_len,
                           int x, int o, int e, double threshold, int *via, int *matches) {
  tro_opt_span r = {0, 0, 0};
  /* span_locater.rs:10-12: first exact window */
  if (piece_len > 0) {
    for (int s = 0; s + piece_len <= seq_len; s++) {
      if (seq[s] == piece[0] && memcmp(seq + s, piece, (size_t)piece_len) == 0) {
        r.found = 1;
        r.start = (uint32_t)s;
        r.end = (uint32_t)(s + piece_len);
        if (via) *via = 1;
        if (matches) *matches = piece_len;
        return r;
      }
    }
  }
  /* span_locater.rs:14-25: WFA ends-free fallback, flank aligner of commands/genotype.rs:66-80 */
  tro_wfa_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.metric = TRO_AFFINE;
  prm.x = x;
  prm.o1 = o;
  prm.e1 = e;
  prm.ends_free = 1;
  prm.text_begin_free = seq_len;
  prm.text_end_free = seq_len;
  tro_wfa_result res;
  tro_wfa_align(&prm, piece, piece_len, seq, seq_len, &res);
  const int nm = tro_count_matches(res.ops, res.n_ops);
  if (matches) *matches = nm;
  if ((double)nm >= threshold) {
    int xs, xe, ys, ye;
    tro_alignment_span(res.ops, res.n_ops, 1, piece_len, seq_len, &xs, &xe, &ys, &ye);
    r.found = 1;
    r.start = (uint32_t)ys;
    r.end = (uint32_t)ye;
    if (via) *via = 2;
  } else {
    if (via) *via = 3;
  }
  tro_wfa_result_free(&res);
  return r;
}

tro_opt_span tro_combine_spans(tro_opt_span lf, tro_opt_span rf) {
  tro_opt_span r = {0, 0, 0};
  if (lf.found && rf.found && lf.end <= rf.start) {
    r.found = 1;
    r.start = lf.end;
    r.end = rf.start;
  }
  return r;
}

int64_t tro_align_consensus(const uint8_t *backbone, int blen, const uint8_t *seq, int slen,
                            uint32_t *out, uint64_t cap, int *score) {
  tro_wfa_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.metric = TRO_AFFINE;
  prm.x = 2;
  prm.o1 = 5;
  prm.e1 = 1;
  tro_wfa_result res;
  tro_wfa_align(&prm, backbone, blen, seq, slen, &res);
  if (score) *score = res.score;
  const int64_t n = tro_sam_cigar(res.ops, res.n_ops, 1, out, cap);
  tro_wfa_result_free(&res);
  return n;
}

double tro_get_dist(const uint8_t *a, int alen, const uint8_t *b, int blen) {
  const size_t MAX_OPS = 10000; /* genotype_cluster.rs:237 */
  if ((size_t)alen * (size_t)blen > MAX_OPS) {
    return sqrt((double)(alen > blen ? alen - blen : blen - alen));
  }
  tro_wfa_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.metric = TRO_EDIT;
  prm.score_only = 1;
  tro_wfa_result res;
  tro_wfa_align(&prm, a, alen, b, blen, &res);
  return sqrt((double)res.score);
}
