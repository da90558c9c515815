/*
 * consensus_oracle.c -- CPU restatement of repair_consensus (TEST INFRASTRUCTURE ONLY; see
 * trgt_oracle.h).  Follows src/trgt/genotype/consensus.rs:5-111 of the reference.
 *
 * Parity UNPINNED: the reference's own tests for this function are commented out
 * (consensus.rs:168-215) and use an older API; their three examples are replayed in
 * tests/test_oracle_golden.py as informative checks only.
 */
#include <stdlib.h>
#include <string.h>

#include "trgt_oracle.h"

typedef struct {
  uint32_t seq, x, len;
} ins_rec;

typedef struct {
  ins_rec *v;
  uint32_t n, cap;
} ins_list;

static void ins_push(ins_list *l, ins_rec r) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 4;
    l->v = (ins_rec *)realloc(l->v, sizeof(ins_rec) * l->cap);
  }
  l->v[l->n++] = r;
}

/* lexicographic byte order of two inserted strings (Rust String Ord) */
static int ins_cmp(const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb) {
  const uint32_t m = la < lb ? la : lb;
  const int c = m ? memcmp(a, b, m) : 0;
  if (c) return c;
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

/* repair_consensus: consensus.rs:5-72.  cigars: run-length SAM words ((len<<4)|op) of each seq against the
 * backbone (utils::align).  Returns the consensus length, -1 on an unexpected base / op, -2 if cap is too small. */
int64_t tro_repair_consensus(const uint8_t *backbone, uint32_t blen, const uint8_t *seqs, const uint64_t *seq_off,
                             uint32_t n_seqs, const uint32_t *words, const uint64_t *word_off, uint8_t *out,
                             uint64_t cap) {
  (void)backbone;
  int32_t *counts = (int32_t *)calloc((size_t)blen * 5 + 1, sizeof(int32_t));
  ins_list *ins = (ins_list *)calloc((size_t)blen + 1, sizeof(ins_list));
  int bad = 0;
  for (uint32_t s = 0; s < n_seqs && !bad; s++) {
    const uint8_t *seq = seqs + seq_off[s];
    const uint64_t slen = seq_off[s + 1] - seq_off[s];
    uint64_t x = 0, y = 0;
    for (uint64_t w = word_off[s]; w < word_off[s + 1] && !bad; w++) {
      const uint32_t len = words[w] >> 4, op = words[w] & 15u;
      switch (op) {
        case 7: case 0: case 8: /* '=' 'M' 'X': consensus.rs:15-27 */
          if (x + len > slen || y + len > blen) { bad = 1; break; }
          for (uint32_t i = 0; i < len; i++) {
            int bi;
            switch (seq[x + i]) { /* summarize_matches :74-86 */
              case 'A': bi = 0; break;
              case 'T': bi = 1; break;
              case 'C': bi = 2; break;
              case 'G': bi = 3; break;
              default: bi = -1;
            }
            if (bi < 0) { bad = 1; break; }
            counts[(y + i) * 5 + bi]++;
          }
          x += len; y += len;
          break;
        case 2: /* 'D' :28-31 */
          if (y + len > blen) { bad = 1; break; }
          for (uint32_t i = 0; i < len; i++) counts[(y + i) * 5 + 4]++;
          y += len;
          break;
        case 1: { /* 'I' :32-36 */
          if (x + len > slen || y > blen) { bad = 1; break; }
          ins_rec r = {s, (uint32_t)x, len};
          ins_push(&ins[y], r);
          x += len;
          break;
        }
        default: bad = 1;
      }
    }
  }
  int64_t n_out = 0;
  int small = 0;
  for (uint32_t p = 0; p < blen && !bad; p++) {
    /* consensus.rs:57-59 + get_ins_consensus :94-111 */
    const ins_list *l = &ins[p];
    if (l->n > n_seqs / 2) {
      const uint32_t without = n_seqs - l->n;
      int best = -1;
      uint32_t best_cnt = 0;
      for (uint32_t a = 0; a < l->n; a++) {
        const uint8_t *sa = seqs + seq_off[l->v[a].seq] + l->v[a].x;
        uint32_t cnt = 0;
        for (uint32_t b = 0; b < l->n; b++) {
          const uint8_t *sb = seqs + seq_off[l->v[b].seq] + l->v[b].x;
          if (ins_cmp(sa, l->v[a].len, sb, l->v[b].len) == 0) cnt++;
        }
        /* sort, group, stable sort by count descending: highest count, ties -> smallest string */
        int better = 0;
        if (best < 0 || cnt > best_cnt) better = 1;
        else if (cnt == best_cnt) {
          const uint8_t *sc = seqs + seq_off[l->v[best].seq] + l->v[best].x;
          if (ins_cmp(sa, l->v[a].len, sc, l->v[best].len) < 0) better = 1;
        }
        if (better) { best = (int)a; best_cnt = cnt; }
      }
      if (best >= 0 && best_cnt > without) {
        const uint8_t *sa = seqs + seq_off[l->v[best].seq] + l->v[best].x;
        for (uint32_t i = 0; i < l->v[best].len; i++) {
          if ((uint64_t)n_out < cap) out[n_out] = sa[i]; else small = 1;
          n_out++;
        }
      }
    }
    /* max_by_key returns the LAST maximum: ties prefer '-' > G > C > T > A (consensus.rs:44-53) */
    int bi = 0;
    for (int i = 1; i < 5; i++)
      if (counts[(size_t)p * 5 + i] >= counts[(size_t)p * 5 + bi]) bi = i;
    if (bi != 4) {
      if ((uint64_t)n_out < cap) out[n_out] = (uint8_t)"ATCG"[bi]; else small = 1;
      n_out++;
    }
  }
  for (uint32_t p = 0; p <= blen; p++) free(ins[p].v);
  free(ins);
  free(counts);
  if (bad) return -1;
  return small ? -2 : n_out;
}
