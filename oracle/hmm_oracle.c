/*
 * hmm_oracle.c -- CPU restatement of the reference motif HMM (TEST INFRASTRUCTURE ONLY;
 * see trgt_oracle.h).  Follows the files under /root/reference/src/hmm/ literally: full
 * score / back-pointer matrices, Kahn-layer state ordering, forward post-processing
 * of the state path.  The CUDA engine restructures all of this; this file is what
 * it is checked against.
 */
#include "trgt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct tro_hmm {
  int S;
  int n_motifs; /* real motifs; the skip block is hmm-motif index n_motifs */
  double *ems;  /* [S][5] ln emissions, symbols # A T C G */
  int *in_off;  /* [S+1] CSR */
  int *in_states;
  double *in_lps;
  int n_in, cap_in;
  int *hm_start, *hm_end; /* [n_motifs+1] start/end state of each motif block (HmmMotif) */
  uint8_t **motif;        /* copies of the motif strings */
  int *motif_len;
  /* scratch for set_trans while building: per-state lists */
  int **tmp_states;
  double **tmp_probs;
  int *tmp_n;
};

/* Hmm::set_ems: hmm_model.rs:49-52 */
static void set_ems(tro_hmm *h, int st, const double p[5]) {
  for (int i = 0; i < 5; i++) h->ems[st * 5 + i] = log(p[i]);
}

/* Hmm::set_trans: hmm_model.rs:44-47 */
static void set_trans(tro_hmm *h, int st, int n, const int *states, const double *probs) {
  free(h->tmp_states[st]);
  free(h->tmp_probs[st]);
  h->tmp_states[st] = (int *)malloc(sizeof(int) * (n ? n : 1));
  h->tmp_probs[st] = (double *)malloc(sizeof(double) * (n ? n : 1));
  for (int i = 0; i < n; i++) {
    h->tmp_states[st][i] = states[i];
    h->tmp_probs[st][i] = log(probs[i]);
  }
  h->tmp_n[st] = n;
}

/* get_match_emissions: builder.rs:175-184 */
static int match_emissions(uint8_t base, double out[5]) {
  const double lo = 0.03, hi = 0.90;
  out[0] = 0.0;
  for (int i = 1; i < 5; i++) out[i] = lo;
  switch (base) {
    case 'A': out[1] = hi; return 0;
    case 'T': out[2] = hi; return 0;
    case 'C': out[3] = hi; return 0;
    case 'G': out[4] = hi; return 0;
    case 'N':
      for (int i = 1; i < 5; i++) out[i] = 0.25;
      return 0;
    default: return -1;
  }
}

static const double EMS_SILENT[5] = {0, 0, 0, 0, 0};
static const double EMS_TERMINAL[5] = {1.0, 0, 0, 0, 0};
static const double EMS_UNIFORM[5] = {0, 0.25, 0.25, 0.25, 0.25};

/* define_motif_block: builder.rs:80-173 */
static int define_motif_block(tro_hmm *h, int ms, const uint8_t *motif, int n) {
  const int first_match = ms + 1;
  const int first_ins = first_match + n;
  const int first_del = first_ins + n; /* n-1 deletion states */
  const int me = ms + 3 * n;

  const double match_prob = 0.90;
  const double ins_to_ins = 0.25;
  const double match_to_indel = (1.00 - match_prob) / 2.00;
  const double del_to_match = 0.50;
  const double mismatch_seed_prob = 2.00 * (1.00 - match_prob) / (double)(n * (n - 1));

  for (int i = 0; i < n; i++) {
    const int st = first_match + i;
    double em[5];
    if (match_emissions(motif[i], em)) return -1;
    set_ems(h, st, em);
    if (i == 0) {
      int s[1] = {ms};
      double p[1] = {match_prob};
      set_trans(h, st, 1, s, p);
    } else {
      const double mismatch_prob = mismatch_seed_prob * (double)(n - i);
      if (i == 1) {
        int s[3] = {st - 1, ms, first_ins + i - 1};
        double p[3] = {match_prob, mismatch_prob, 1.0 - ins_to_ins};
        set_trans(h, st, 3, s, p);
      } else {
        int s[4] = {st - 1, ms, first_ins + i - 1, first_del + i - 2};
        double p[4] = {match_prob, mismatch_prob, 1.0 - ins_to_ins, del_to_match};
        set_trans(h, st, 4, s, p);
      }
    }
  }
  for (int i = 0; i < n; i++) {
    const int st = first_ins + i;
    set_ems(h, st, EMS_UNIFORM);
    int s[2] = {st, first_match + i};
    double p[2] = {ins_to_ins, match_to_indel};
    set_trans(h, st, 2, s, p);
  }
  for (int i = 0; i + 1 < n; i++) {
    const int st = first_del + i;
    set_ems(h, st, EMS_SILENT);
    if (i == 0) {
      int s[1] = {first_match + i};
      double p[1] = {match_to_indel};
      set_trans(h, st, 1, s, p);
    } else {
      int s[2] = {first_match + i, st - 1};
      double p[2] = {match_to_indel, 1.0 - del_to_match};
      set_trans(h, st, 2, s, p);
    }
  }
  set_ems(h, me, EMS_SILENT);
  if (n > 1) {
    int s[3] = {first_match + n - 1, first_ins + n - 1, first_del + n - 2};
    double p[3] = {match_prob, 1.0 - ins_to_ins, 1.0};
    set_trans(h, me, 3, s, p);
  } else {
    int s[2] = {first_match + n - 1, first_ins + n - 1};
    double p[2] = {match_prob, 1.0 - ins_to_ins};
    set_trans(h, me, 2, s, p);
  }
  return 0;
}

tro_hmm *tro_hmm_build(const uint8_t *motifs, const uint32_t *offsets, uint32_t n_motifs) {
  int S = 7;
  for (uint32_t m = 0; m < n_motifs; m++) {
    int n = (int)(offsets[m + 1] - offsets[m]);
    if (n <= 0) return NULL;
    S += 3 * n + 1;
  }
  tro_hmm *h = (tro_hmm *)calloc(1, sizeof(tro_hmm));
  h->S = S;
  h->n_motifs = (int)n_motifs;
  h->ems = (double *)malloc(sizeof(double) * 5 * S);
  for (int i = 0; i < 5 * S; i++) h->ems[i] = -INFINITY; /* Hmm::new: hmm_model.rs:29 */
  h->tmp_states = (int **)calloc(S, sizeof(int *));
  h->tmp_probs = (double **)calloc(S, sizeof(double *));
  h->tmp_n = (int *)calloc(S, sizeof(int));
  h->hm_start = (int *)malloc(sizeof(int) * (n_motifs + 1));
  h->hm_end = (int *)malloc(sizeof(int) * (n_motifs + 1));
  h->motif = (uint8_t **)calloc(n_motifs + 1, sizeof(uint8_t *));
  h->motif_len = (int *)calloc(n_motifs + 1, sizeof(int));

  const int start = 0, end = S - 1, rs = 1, re = S - 2;
  set_ems(h, start, EMS_TERMINAL);
  set_ems(h, end, EMS_TERMINAL);
  {
    int s[1] = {re};
    double p[1] = {0.10};
    set_trans(h, end, 1, s, p);
  }
  set_ems(h, rs, EMS_SILENT);
  {
    int s[2] = {start, re};
    double p[2] = {1.00, 1.00};
    set_trans(h, rs, 2, s, p);
  }
  const double rs_to_ms = 1.00, me_to_re = 0.50;
  int *mes = (int *)malloc(sizeof(int) * (n_motifs + 1));
  double *me_probs = (double *)malloc(sizeof(double) * (n_motifs + 1));
  int ms = rs + 1;
  int bad = 0;
  for (uint32_t m = 0; m < n_motifs; m++) {
    const int n = (int)(offsets[m + 1] - offsets[m]);
    const uint8_t *motif = motifs + offsets[m];
    const int me = ms + 3 * n;
    set_ems(h, ms, EMS_SILENT);
    int s[2] = {rs, me};
    double p[2] = {rs_to_ms, 1.0 - me_to_re};
    set_trans(h, ms, 2, s, p);
    if (define_motif_block(h, ms, motif, n)) bad = 1;
    mes[m] = me;
    h->hm_start[m] = ms;
    h->hm_end[m] = me;
    h->motif[m] = (uint8_t *)malloc(n);
    memcpy(h->motif[m], motif, n);
    h->motif_len[m] = n;
    ms += 3 * n + 1;
  }
  /* skip block: builder.rs:41-53 */
  {
    const int skip_state = ms + 1, me = ms + 2;
    set_ems(h, ms, EMS_SILENT);
    int s[2] = {rs, me};
    double p[2] = {rs_to_ms, 1.0 - me_to_re};
    set_trans(h, ms, 2, s, p);
    const double skip_to_skip = 0.5;
    set_ems(h, skip_state, EMS_UNIFORM);
    int s2[2] = {ms, skip_state};
    double p2[2] = {1.0, skip_to_skip};
    set_trans(h, skip_state, 2, s2, p2);
    set_ems(h, me, EMS_SILENT);
    int s3[1] = {skip_state};
    double p3[1] = {1.0 - skip_to_skip};
    set_trans(h, me, 1, s3, p3);
    mes[n_motifs] = me;
    h->hm_start[n_motifs] = ms;
    h->hm_end[n_motifs] = me;
  }
  set_ems(h, re, EMS_SILENT);
  for (uint32_t m = 0; m <= n_motifs; m++) me_probs[m] = me_to_re;
  set_trans(h, re, (int)n_motifs + 1, mes, me_probs);
  free(mes);
  free(me_probs);

  /* flatten to CSR */
  h->in_off = (int *)malloc(sizeof(int) * (S + 1));
  int tot = 0;
  for (int s = 0; s < S; s++) {
    h->in_off[s] = tot;
    tot += h->tmp_n[s];
  }
  h->in_off[S] = tot;
  h->in_states = (int *)malloc(sizeof(int) * (tot ? tot : 1));
  h->in_lps = (double *)malloc(sizeof(double) * (tot ? tot : 1));
  for (int s = 0; s < S; s++) {
    for (int i = 0; i < h->tmp_n[s]; i++) {
      h->in_states[h->in_off[s] + i] = h->tmp_states[s][i];
      h->in_lps[h->in_off[s] + i] = h->tmp_probs[s][i];
    }
    free(h->tmp_states[s]);
    free(h->tmp_probs[s]);
  }
  free(h->tmp_states);
  free(h->tmp_probs);
  free(h->tmp_n);
  h->tmp_states = NULL;
  h->tmp_probs = NULL;
  h->tmp_n = NULL;
  if (bad) {
    tro_hmm_free(h);
    return NULL;
  }
  return h;
}

void tro_hmm_free(tro_hmm *h) {
  if (!h) return;
  free(h->ems);
  free(h->in_off);
  free(h->in_states);
  free(h->in_lps);
  free(h->hm_start);
  free(h->hm_end);
  if (h->motif)
    for (int m = 0; m <= h->n_motifs; m++) free(h->motif[m]);
  free(h->motif);
  free(h->motif_len);
  free(h);
}

int tro_hmm_num_states(const tro_hmm *h) { return h->S; }
double tro_hmm_em(const tro_hmm *h, int state, int sym) { return h->ems[state * 5 + sym]; }
int tro_hmm_num_in(const tro_hmm *h, int state) { return h->in_off[state + 1] - h->in_off[state]; }
int tro_hmm_in_state(const tro_hmm *h, int state, int i) { return h->in_states[h->in_off[state] + i]; }
double tro_hmm_in_lp(const tro_hmm *h, int state, int i) { return h->in_lps[h->in_off[state] + i]; }

static int is_silent(const tro_hmm *h, int st) {
  for (int i = 0; i < 5; i++)
    if (!isinf(h->ems[st * 5 + i])) return 0;
  return 1;
}

/* Hmm::emits_base: hmm_model.rs:202-204 */
static int emits_base(const tro_hmm *h, int st) {
  for (int i = 1; i < 5; i++)
    if (isfinite(h->ems[st * 5 + i])) return 1;
  return 0;
}

static int emits_any(const tro_hmm *h, int st) {
  for (int i = 0; i < 5; i++)
    if (isfinite(h->ems[st * 5 + i])) return 1;
  return 0;
}

/* Hmm::order_states: hmm_model.rs:206-240 */
static int *order_states(const tro_hmm *h) {
  const int S = h->S;
  int *order = (int *)malloc(sizeof(int) * S);
  int *silent = (int *)malloc(sizeof(int) * S);
  int *unused = (int *)malloc(sizeof(int) * S);
  char *in_set = (char *)calloc(S, 1);
  int n_order = 0, n_silent = 0;
  for (int s = 0; s < S; s++) {
    if (is_silent(h, s)) {
      silent[n_silent++] = s;
      in_set[s] = 1;
    } else {
      order[n_order++] = s;
    }
  }
  while (n_silent > 0) {
    int n_unused = 0;
    for (int i = 0; i < n_silent; i++) {
      const int st = silent[i];
      int has_incoming_silent = 0;
      for (int e = h->in_off[st]; e < h->in_off[st + 1]; e++)
        if (in_set[h->in_states[e]]) has_incoming_silent = 1;
      if (!has_incoming_silent)
        order[n_order++] = st;
      else
        unused[n_unused++] = st;
    }
    if (n_unused >= n_silent) break; /* reference asserts; cannot happen for built models */
    memset(in_set, 0, S);
    for (int i = 0; i < n_unused; i++) {
      silent[i] = unused[i];
      in_set[unused[i]] = 1;
    }
    n_silent = n_unused;
  }
  free(silent);
  free(unused);
  free(in_set);
  return order;
}

static int encode_base(uint8_t b) {
  switch (b) {
    case '#': return 0;
    case 'A': return 1;
    case 'T': return 2;
    case 'C': return 3;
    case 'G': return 4;
    default: return -1;
  }
}

int64_t tro_hmm_label(const tro_hmm *h, const uint8_t *query, uint32_t len, uint32_t *out,
                      uint64_t cap) {
  if (len == 0) return 0; /* hmm_model.rs:145-147 */
  const int S = h->S;
  const size_t n = (size_t)len + 2;
  uint8_t *q = (uint8_t *)malloc(n);
  q[0] = 0;
  q[n - 1] = 0;
  for (uint32_t i = 0; i < len; i++) {
    int c = encode_base(query[i]);
    if (c < 0) {
      free(q);
      return -1;
    }
    q[i + 1] = (uint8_t)c;
  }
  /* generate_mats: hmm_model.rs:99-114 */
  int *order = order_states(h);
  double *scores = (double *)malloc(sizeof(double) * S * n);
  int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * S * n);
  for (size_t i = 0; i < (size_t)S * n; i++) {
    scores[i] = -INFINITY;
    prev[i] = -1;
  }
  char *silent = (char *)malloc(S);
  for (int s = 0; s < S; s++) silent[s] = (char)is_silent(h, s);
  for (size_t index = 0; index < n; index++) {
    for (int oi = 0; oi < S; oi++) {
      /* calc_viterbi_score: hmm_model.rs:54-97 */
      const int st = order[oi];
      const int symbol = q[index];
      const double em_term = silent[st] ? 0.0 : h->ems[st * 5 + symbol];
      const int lookback = silent[st] ? 0 : 1;
      const int nin = h->in_off[st + 1] - h->in_off[st];
      if (index == 0 && nin != 0 && lookback == 1) continue;
      double max_score = -INFINITY;
      int best = -1;
      for (int e = 0; e < nin; e++) {
        const int ps = h->in_states[h->in_off[st] + e];
        const double prev_score = scores[(size_t)ps * n + (index - lookback)];
        const double trans_lp = h->in_lps[h->in_off[st] + e];
        const double sc = prev_score + trans_lp + em_term;
        if (sc > max_score) {
          best = ps;
          max_score = sc;
        }
      }
      if (index == 0 && nin == 0 && isfinite(em_term)) {
        max_score = em_term;
        best = st;
      }
      if (best >= 0) {
        scores[(size_t)st * n + index] = max_score;
        prev[(size_t)st * n + index] = best;
      }
    }
  }
  /* traceback: hmm_model.rs:125-142 */
  int64_t cnt = 0;
  int rc_small = 0;
  {
    int state = S - 1;
    size_t index = n - 1;
    /* first pass counts, second fills reversed */
    uint32_t *tmp = NULL;
    size_t tcap = 0;
    while (state != 0) {
      if ((size_t)cnt == tcap) {
        tcap = tcap ? tcap * 2 : 1024;
        tmp = (uint32_t *)realloc(tmp, sizeof(uint32_t) * tcap);
      }
      tmp[cnt++] = (uint32_t)state;
      const int p = prev[(size_t)state * n + index];
      if (p < 0) { /* reference would panic on unwrap; cannot happen for built models */
        free(tmp);
        free(q); free(order); free(scores); free(prev); free(silent);
        return -3;
      }
      if (emits_any(h, state)) index -= 1;
      state = p;
    }
    if ((uint64_t)cnt + 1 > cap) {
      rc_small = 1;
    } else {
      out[0] = 0;
      for (int64_t i = 0; i < cnt; i++) out[1 + i] = tmp[cnt - 1 - i];
    }
    cnt += 1;
    free(tmp);
  }
  free(q);
  free(order);
  free(scores);
  free(prev);
  free(silent);
  return rc_small ? -2 : cnt;
}

static int find_motif_by_start(const tro_hmm *h, int st) {
  for (int m = 0; m <= h->n_motifs; m++)
    if (h->hm_start[m] == st) return m;
  return -1;
}
static int is_motif_end(const tro_hmm *h, int st) {
  for (int m = 0; m <= h->n_motifs; m++)
    if (h->hm_end[m] == st) return 1;
  return 0;
}

int64_t tro_remove_imperfect_motifs(const tro_hmm *h, const uint32_t *states, uint64_t n_states,
                                    const uint8_t *query, uint32_t qlen, uint32_t max_motif_len,
                                    uint32_t *out, uint64_t cap) {
  (void)qlen;
  if (n_states == 0) return 0;
  if (n_states <= 4) return -3; /* reference asserts len > 4 */
  uint64_t n_out = 0;
#define PUSH(v)                      \
  do {                               \
    if (n_out >= cap) return -2;     \
    out[n_out++] = (uint32_t)(v);    \
  } while (0)
  PUSH(states[0]);
  PUSH(states[1]);
  const uint32_t run_end_state = (uint32_t)(h->S - 2);
  uint64_t si = 2;
  uint32_t base_index = 0;
  while (si != n_states) {
    const uint64_t copy_begin = si;
    const uint32_t seq_begin = base_index;
    uint32_t consumed = 0;
    if (find_motif_by_start(h, (int)states[si]) < 0) return -3; /* reference asserts */
    while (!is_motif_end(h, (int)states[si])) {
      if (emits_base(h, (int)states[si])) {
        base_index++;
        consumed++;
      }
      si++;
    }
    const uint64_t copy_end = si; /* index of the motif-end state */
    si++;
    const int m = find_motif_by_start(h, (int)states[copy_begin]);
    const uint32_t motif_len = (uint32_t)((h->hm_end[m] - h->hm_start[m]) / 3);
    int keep = 1;
    const int is_skip = (m == h->n_motifs);
    if (!is_skip && motif_len <= max_motif_len) {
      const int n = h->motif_len[m];
      if (consumed < (uint32_t)n) {
        keep = 0;
      } else {
        for (int i = 0; i < n; i++) {
          const uint8_t expected = h->motif[m][i];
          const uint8_t observed = query[seq_begin + i];
          if (expected != 'N' && observed != expected) keep = 0;
        }
      }
    }
    if (keep) {
      for (uint64_t i = copy_begin; i <= copy_end; i++) PUSH(states[i]);
    } else {
      uint32_t bases = 0;
      for (uint64_t i = copy_begin; i <= copy_end; i++) bases += (uint32_t)emits_base(h, (int)states[i]);
      const int skip_ms = h->hm_start[h->n_motifs];
      PUSH(skip_ms);
      for (uint32_t i = 0; i < bases; i++) PUSH(skip_ms + 1);
      PUSH(h->hm_end[h->n_motifs]);
    }
    if (states[si] == run_end_state) {
      PUSH(states[si]);
      PUSH(states[si + 1]);
      si += 2;
    }
  }
#undef PUSH
  return (int64_t)n_out;
}

int64_t tro_label_motifs(const tro_hmm *h, const uint32_t *states, uint64_t n_states, tro_span *out,
                         uint64_t cap) {
  uint64_t n_out = 0;
  uint64_t si = 0;
  uint32_t last_end = 0;
  while (si < n_states) {
    const int m = find_motif_by_start(h, (int)states[si]);
    if (m >= 0) {
      uint32_t span = 0;
      const uint32_t end_state = (uint32_t)h->hm_end[m];
      while (states[si] != end_state) {
        span += (uint32_t)emits_base(h, (int)states[si]);
        si++;
      }
      while (si < n_states && states[si] == end_state) {
        span += (uint32_t)emits_base(h, (int)states[si]);
        si++;
      }
      if (n_out >= cap) return -2;
      out[n_out].motif_index = (uint32_t)m;
      out[n_out].start = last_end;
      out[n_out].end = last_end + span;
      last_end += span;
      n_out++;
    } else {
      si++;
    }
  }
  return (int64_t)n_out;
}

uint8_t tro_get_base_match(const tro_hmm *h, int state) {
  const double *ems = h->ems + state * 5;
  if (!emits_base(h, state)) return ' ';
  double mx = ems[0];
  for (int i = 1; i < 5; i++)
    if (ems[i] > mx) mx = ems[i];
  int n_top = 0, top = -1;
  for (int i = 0; i < 5; i++)
    if (ems[i] == mx) {
      n_top++;
      if (top < 0) top = i;
    }
  if (n_top == 1) return (uint8_t)"#ATCG"[top];
  if (n_top == 4) return 'N';
  return ' ';
}

/* get_events + calc_purity: events.rs:17-86, purity.rs:6-41 */
double tro_calc_purity(const tro_hmm *h, const uint32_t *states, uint64_t n_states,
                       const uint8_t *query, uint32_t qlen) {
  if (qlen == 0) return NAN;
  const int S = h->S;
  int *state_to_motif = (int *)malloc(sizeof(int) * S);
  for (int s = 0; s < S; s++) state_to_motif[s] = -1;
  for (int m = 0; m <= h->n_motifs; m++)
    for (int s = h->hm_start[m]; s <= h->hm_end[m]; s++) state_to_motif[s] = m;
  uint64_t n_match = 0, n_mismatch = 0, n_ins = 0, n_del = 0, n_skip = 0;
  uint32_t base_index = 0;
  for (uint64_t si = 0; si < n_states; si++) {
    const int st = (int)states[si];
    const int m = state_to_motif[st];
    if (m == -1) continue; /* Trans */
    if (st == h->hm_start[m]) {
      const int next_state = (int)states[si + 1];
      n_del += (uint64_t)(next_state - st - 1);
      continue;
    }
    if (st == h->hm_end[m]) continue;
    if (m == h->n_motifs) {
      n_skip++;
      base_index++;
      continue;
    }
    const int offset = st - h->hm_start[m] - 1;
    const int motif_len = h->motif_len[m];
    switch (offset / motif_len) {
      case 0: {
        const uint8_t base = query[base_index];
        const uint8_t expected = tro_get_base_match(h, st);
        if (base == expected || expected == 'N')
          n_match++;
        else
          n_mismatch++;
        base_index++;
        break;
      }
      case 1:
        n_ins++;
        base_index++;
        break;
      case 2: n_del++; break;
      default: break;
    }
  }
  free(state_to_motif);
  const double edit_dist = (double)(n_del + n_ins + n_mismatch + n_skip);
  const uint64_t ref_len = n_match + n_mismatch + n_del + n_skip;
  const double max_dist = (double)(ref_len > qlen ? ref_len : qlen);
  return (max_dist - edit_dist) / max_dist;
}

void tro_replace_invalid_bases(uint8_t *seq, uint32_t len, const char *allowed) {
  const size_t na = strlen(allowed);
  for (uint32_t i = 0; i < len; i++) {
    if (!memchr(allowed, seq[i], na)) seq[i] = (uint8_t)allowed[i % na];
  }
}

int64_t tro_annotate_allele(const tro_hmm *h, const uint8_t *allele, uint32_t len,
                            uint32_t *motif_counts, tro_span *spans_out, uint64_t span_cap,
                            double *purity_out) {
  for (int m = 0; m < h->n_motifs; m++) motif_counts[m] = 0;
  uint8_t *seq = (uint8_t *)malloc(len ? len : 1);
  memcpy(seq, allele, len);
  tro_replace_invalid_bases(seq, len, "ATCG");
  /* generous bound on path length: every column visits <= S states */
  const uint64_t cap = ((uint64_t)len + 2) * (uint64_t)(h->S) + 8;
  uint32_t *states = (uint32_t *)malloc(sizeof(uint32_t) * cap);
  uint32_t *states2 = (uint32_t *)malloc(sizeof(uint32_t) * cap);
  int64_t rc = 0;
  int64_t n = tro_hmm_label(h, seq, len, states, cap);
  if (n < 0) {
    rc = n;
    goto done;
  }
  *purity_out = tro_calc_purity(h, states, (uint64_t)n, seq, len);
  int64_t n2 = tro_remove_imperfect_motifs(h, states, (uint64_t)n, seq, len, 6, states2, cap);
  if (n2 < 0) {
    rc = n2;
    goto done;
  }
  {
    tro_span *spans = (tro_span *)malloc(sizeof(tro_span) * ((size_t)len + 1));
    int64_t ns = tro_label_motifs(h, states2, (uint64_t)n2, spans, (uint64_t)len + 1);
    if (ns < 0) {
      free(spans);
      rc = ns;
      goto done;
    }
    /* tr.rs:471-476: drop skip spans, count, collapse (utils.rs:11-27) */
    uint64_t n_out = 0;
    int have = 0;
    for (int64_t i = 0; i < ns; i++) {
      if (spans[i].motif_index >= (uint32_t)h->n_motifs) continue;
      motif_counts[spans[i].motif_index]++;
      if (have && spans_out[n_out - 1].motif_index == spans[i].motif_index &&
          spans_out[n_out - 1].end == spans[i].start) {
        spans_out[n_out - 1].end = spans[i].end;
      } else {
        if (n_out >= span_cap) {
          free(spans);
          rc = -2;
          goto done;
        }
        spans_out[n_out++] = spans[i];
        have = 1;
      }
    }
    free(spans);
    rc = (int64_t)n_out;
  }
done:
  free(seq);
  free(states);
  free(states2);
  return rc;
}
