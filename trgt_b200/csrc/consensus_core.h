// consensus_core.h -- repair_consensus fused behind the consensus alignments.
//
// Replaces src/trgt/genotype/consensus.rs:5-72 (repair_consensus) and :94-111 (get_ins_consensus) of
// the reference, which every genotyper runs right after utils::align (genotype_size.rs:35-36,
// genotype_cluster.rs:52-53, genotype_flank.rs:19-20).  With the vote on the device the CIGARs never
// leave HBM: phase B returns one repaired consensus per group.
//
// One group of lanes per (backbone, members) group:
//   1. members in parallel: walk the member's run-length CIGAR, count A/T/C/G/- per backbone column
//      (atomic adds into a 6-int row per column: 5 counts + number of insertions anchored there) and
//      append every insertion run to the group's record list;
//   2. columns in parallel: the column's base is the LAST maximum of [A,T,C,G,-] (Rust's max_by_key:
//      ties prefer '-' > G > C > T > A); a column where more than half of the members carry an insertion
//      takes the most frequent inserted string (ties: lexicographically smallest, as sort + stable
//      sort-by-count gives) if it outnumbers the members without insertion; insertions after the last
//      column are ignored, as in the reference;
//   3. exclusive scan of the per-column output lengths, then the bytes are written in place.
// Run twice per group: a counting pass (total length -> CSR offsets) and a writing pass.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "coop.h"

namespace trgt {

struct ConsRec {
  uint32_t y, seq, x, len;  // anchored before backbone column y; bytes seq[x .. x+len)
};

struct ConsGroup {
  int B;                              // backbone length
  uint32_t s0, n;                     // members are sequences s0 .. s0+n-1
  const uint8_t *seqs;                // all member bytes
  const uint64_t *seq_off;            // [.. n_seqs+1]
  const uint32_t *words;              // run-length SAM CIGARs (utils::align output)
  const unsigned long long *word_off; // [.. n_seqs+1]
};

TRGT_HD int cons_atomic_add(int *p, int v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#endif
}

TRGT_HD int cons_cmp(const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb) {
  const uint32_t m = la < lb ? la : lb;
  for (uint32_t i = 0; i < m; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

// decision for one column: base index (0..4) and the chosen insertion record (-1: none)
TRGT_HD void cons_decide(const ConsGroup &gr, const int *row, int p, const ConsRec *recs, int n_rec, int *base,
                         int *ins) {
  int bi = 0;
  for (int i = 1; i < 5; i++)
    if (row[i] >= row[bi]) bi = i;
  *base = bi;
  *ins = -1;
  const int k = row[5];
  if (k > (int)(gr.n / 2)) {  // consensus.rs:57
    const int without = (int)gr.n - k;
    // the column's own records first (one pass over the group's list, in list order), then the quadratic count among
    // those few: a group of long alleles has hundreds of records and a handful per column
    enum { MINE = 64 };
    int mine[MINE];
    int m = 0;
    bool all = true;
    for (int a = 0; a < n_rec; a++)
      if ((int)recs[a].y == p) {
        if (m < MINE) mine[m++] = a; else all = false;
      }
    int best = -1, best_cnt = 0;
    const int na = all ? m : n_rec;
    for (int ia = 0; ia < na; ia++) {
      const int a = all ? mine[ia] : ia;
      if ((int)recs[a].y != p) continue;
      const uint8_t *sa = gr.seqs + gr.seq_off[recs[a].seq] + recs[a].x;
      int cnt = 0;
      for (int ib = 0; ib < na; ib++) {
        const int b = all ? mine[ib] : ib;
        if ((int)recs[b].y == p &&
            cons_cmp(sa, recs[a].len, gr.seqs + gr.seq_off[recs[b].seq] + recs[b].x, recs[b].len) == 0)
          cnt++;
      }
      bool better = best < 0 || cnt > best_cnt;
      if (!better && cnt == best_cnt)
        better = cons_cmp(sa, recs[a].len, gr.seqs + gr.seq_off[recs[best].seq] + recs[best].x, recs[best].len) < 0;
      if (better) { best = a; best_cnt = cnt; }
    }
    if (best >= 0 && best_cnt > without) *ins = best;  // get_ins_consensus :106-110
  }
}

// counts: 6*B ints, recs: rec_cap records, shared: 2 ints -- all group-owned scratch.
// out == nullptr: counting pass.  Returns the consensus length, or -1 on an unexpected base / CIGAR op
// (the reference panics), -2 if the record list overflows.
template <class G>
TRGT_HD long long consensus_vote(const G &g, const ConsGroup &gr, int *counts, ConsRec *recs, uint32_t rec_cap,
                                 int *shared, uint8_t *out) {
  const int B = gr.B;
  for (int i = g.lane(); i < 6 * B; i += g.size()) counts[i] = 0;
  if (g.lane() == 0) { shared[0] = 0; shared[1] = 0; }
  g.sync();
  // 1. members in parallel
  for (uint32_t m = (uint32_t)g.lane(); m < gr.n; m += (uint32_t)g.size()) {
    const uint32_t s = gr.s0 + m;
    const uint8_t *seq = gr.seqs + gr.seq_off[s];
    const unsigned long long slen = gr.seq_off[s + 1] - gr.seq_off[s];
    unsigned long long x = 0, y = 0;
    bool bad = false;
    for (unsigned long long w = gr.word_off[s]; w < gr.word_off[s + 1] && !bad; w++) {
      const uint32_t len = gr.words[w] >> 4, op = gr.words[w] & 15u;
      if (op == 7u || op == 0u || op == 8u) {  // '=' 'M' 'X'
        if (x + len > slen || y + len > (unsigned long long)B) { bad = true; break; }
        for (uint32_t i = 0; i < len; i++) {
          const uint8_t c = seq[x + i];
          const int bi = c == 'A' ? 0 : (c == 'T' ? 1 : (c == 'C' ? 2 : (c == 'G' ? 3 : -1)));
          if (bi < 0) { bad = true; break; }
          cons_atomic_add(&counts[(y + i) * 6 + bi], 1);
        }
        x += len; y += len;
      } else if (op == 2u) {  // 'D'
        if (y + len > (unsigned long long)B) { bad = true; break; }
        for (uint32_t i = 0; i < len; i++) cons_atomic_add(&counts[(y + i) * 6 + 4], 1);
        y += len;
      } else if (op == 1u) {  // 'I'
        if (x + len > slen || y > (unsigned long long)B) { bad = true; break; }
        if (y < (unsigned long long)B) {  // insertions after the last column never reach the output
          cons_atomic_add(&counts[y * 6 + 5], 1);
          const int slot = cons_atomic_add(&shared[0], 1);
          if ((uint32_t)slot < rec_cap) {
            ConsRec r;
            r.y = (uint32_t)y; r.seq = s; r.x = (uint32_t)x; r.len = len;
            recs[slot] = r;
          }
        }
        x += len;
      } else {
        bad = true;
      }
    }
    if (bad) cons_atomic_add(&shared[1], 1);
  }
  g.sync();
  const int n_rec_all = shared[0];
  const int n_bad = shared[1];
  g.sync();
  if (n_bad) return -1;
  if ((uint32_t)n_rec_all > rec_cap) return -2;
  // 2 + 3. columns in parallel, chunk by chunk, with a running offset
  long long total = 0;
  for (int pb = 0; pb < B; pb += g.size()) {
    const int p = pb + g.lane();
    int base = 4, ins = -1, len = 0;
    if (p < B) {
      cons_decide(gr, counts + (size_t)p * 6, p, recs, n_rec_all, &base, &ins);
      len = (base != 4 ? 1 : 0) + (ins >= 0 ? (int)recs[ins].len : 0);
    }
    int chunk_total = 0;
    const int before = g.excl_scan_i(len, &chunk_total);
    if (out != nullptr && p < B) {
      uint8_t *o = out + total + before;
      if (ins >= 0) {
        const uint8_t *src = gr.seqs + gr.seq_off[recs[ins].seq] + recs[ins].x;
        for (uint32_t i = 0; i < recs[ins].len; i++) *o++ = src[i];
      }
      if (base != 4) *o = (uint8_t)(base == 0 ? 'A' : (base == 1 ? 'T' : (base == 2 ? 'C' : 'G')));
    }
    total += chunk_total;
  }
  g.sync();
  return total;
}

}  // namespace trgt
