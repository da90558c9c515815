// cluster_core.h -- the cluster genotyper's glue between get_dist_matrix and make_consensus, one cooperating
// group (a warp) per locus.
//
// Replaces (reference, PacificBiosciences/trgt v3.0.0):
//   cluster()       src/trgt/genotype/genotype_cluster.rs:154-227  (kodama::linkage(.., Method::Ward) :161)
//   central_read()  src/trgt/genotype/genotype_cluster.rs:12-39
//   the choice of group1 / group2 in genotype(), :64-69
//
// kodama 0.3.0 is not vendored in the reference tree; its Ward linkage is the nearest-neighbour chain of
// fastcluster on squared dissimilarities with an in-place Lance-Williams update, a stable sort of the merges and
// a union-find relabelling (see oracle/cluster_oracle.c for what is pinned on scipy and what is not).  The
// reference hands `&mut dists` to linkage and reads the same array in central_read afterwards, so the matrix the
// chain leaves behind is part of the semantics; this core updates the condensed matrix in place the same way.
//
// Parallel shape: the chain itself is sequential (n - 1 merges), but each of its steps is a scan over the
// locus' <= max_depth sequences -- nearest-neighbour search (first minimum in index order, as the sequential
// strict '<' scan finds it) and Lance-Williams update -- which the lanes stride.  Every product and sum is
// rounded separately (no fused multiply-add), so the matrix is bit-identical to the oracle's.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "coop.h"

namespace trgt {

#if defined(__CUDA_ARCH__)
#define TRGT_FMUL(a, b) __dmul_rn((a), (b))
#define TRGT_FADD(a, b) __dadd_rn((a), (b))
#define TRGT_FSUB(a, b) __dsub_rn((a), (b))
#else
#define TRGT_FMUL(a, b) ((a) * (b))
#define TRGT_FADD(a, b) ((a) + (b))
#define TRGT_FSUB(a, b) ((a) - (b))
#endif

TRGT_HD size_t cl_cidx(size_t n, size_t i, size_t j) { return i * n - i * (i + 1) / 2 + (j - i - 1); }
TRGT_HD double *cl_dref(double *d, size_t n, size_t i, size_t j) { return i < j ? &d[cl_cidx(n, i, j)] : &d[cl_cidx(n, j, i)]; }

// per-locus scratch, carved from one allocation of cl_ws_bytes(n_max) bytes (8-byte aligned)
struct ClusterWs {
  double *raw_d;      // [n] merge heights (squared), chain order
  double *step_d;     // [n] merge heights, sorted and relabelled
  double *cand_v;     // [64] nearest-neighbour candidates of the lanes / central_read sums reuse raw_d
  uint32_t *raw_a, *raw_b;       // [n]
  uint32_t *step_c1, *step_c2, *step_size;  // [n]
  uint32_t *chain;    // [n + 2]
  uint32_t *size;     // [n]
  uint32_t *parent, *csize;      // [2n]
  int32_t *membership;           // [2n]
  uint32_t *group;    // [n]
  uint32_t *members;  // [n]
  uint32_t *cand_i;   // [64]
  uint8_t *active;    // [n]
};

TRGT_HD size_t cl_ws_bytes(size_t n) {
  const size_t n8 = (n + 8) & ~(size_t)7;
  return sizeof(double) * (2 * n8 + 64) + sizeof(uint32_t) * (9 * n8 + 2 * (2 * n8) + 8 + 64) + sizeof(int32_t) * (2 * n8) + n8 + 64;
}

TRGT_HD ClusterWs cl_carve(void *base, size_t n) {
  const size_t n8 = (n + 8) & ~(size_t)7;
  ClusterWs w;
  double *d = (double *)base;
  w.raw_d = d; d += n8;
  w.step_d = d; d += n8;
  w.cand_v = d; d += 64;
  uint32_t *u = (uint32_t *)d;
  w.raw_a = u; u += n8;
  w.raw_b = u; u += n8;
  w.step_c1 = u; u += n8;
  w.step_c2 = u; u += n8;
  w.step_size = u; u += n8;
  w.chain = u; u += n8 + 8;
  w.size = u; u += n8;
  w.parent = u; u += 2 * n8;
  w.csize = u; u += 2 * n8;
  w.group = u; u += n8;
  w.members = u; u += n8;
  w.cand_i = u; u += 64;
  w.membership = (int32_t *)u; u += 2 * n8;
  w.active = (uint8_t *)u;
  return w;
}

// first minimum of d(x, b) over the active x in [lo, hi), x != b, in index order; *idx = n if there is none
template <class G>
TRGT_HD double cl_nearest(const G &g, double *dists, uint32_t n, const ClusterWs &w, uint32_t b, uint32_t lo, uint32_t hi,
                          uint32_t *idx) {
  double bv = INFINITY;
  uint32_t bi = n;
  for (uint32_t x = lo + (uint32_t)g.lane(); x < hi; x += (uint32_t)g.size()) {
    if (!w.active[x] || x == b) continue;
    const double v = *cl_dref(dists, n, x, b);
    if (bi == n || v < bv) { bv = v; bi = x; }
  }
  if (g.lane() < 64) { w.cand_v[g.lane()] = bv; w.cand_i[g.lane()] = bi; }
  g.sync();
  const int nl = g.size() < 64 ? g.size() : 64;
  double rv = INFINITY;
  uint32_t ri = n;
  for (int l = 0; l < nl; l++) {  // every lane combines the same candidates: smallest value, then smallest index
    const uint32_t ci = w.cand_i[l];
    if (ci == n) continue;
    const double cv = w.cand_v[l];
    if (ri == n || cv < rv || (cv == rv && ci < ri)) { rv = cv; ri = ci; }
  }
  g.sync();
  *idx = ri;
  return rv;
}

// kodama::linkage(dists, n, Method::Ward): dists is modified in place; the n - 1 merges end up in
// w.step_c1 / step_c2 / step_d / step_size (sorted by height, step i creates cluster n + i)
template <class G>
TRGT_HD void cl_ward_linkage(const G &g, double *dists, uint32_t n, const ClusterWs &w) {
  const size_t np = (size_t)n * (n - 1) / 2;
  for (size_t k = (size_t)g.lane(); k < np; k += (size_t)g.size()) dists[k] = TRGT_FMUL(dists[k], dists[k]);
  for (uint32_t i = (uint32_t)g.lane(); i < n; i += (uint32_t)g.size()) { w.active[i] = 1; w.size[i] = 1; }
  g.sync();
  uint32_t tip = 0;
  for (uint32_t step = 0; step + 1 < n; step++) {
    uint32_t a, b;
    double min;
    if (tip <= 3) {
      a = 0;
      while (!w.active[a]) a++;
      if (g.lane() == 0) w.chain[0] = a;
      tip = 1;
      min = cl_nearest(g, dists, n, w, a, a + 1, n, &b);
    } else {
      tip -= 3;
      a = w.chain[tip - 1];
      b = w.chain[tip];
      min = *cl_dref(dists, n, a, b);
    }
    for (;;) {
      if (g.lane() == 0) w.chain[tip] = b;
      uint32_t x;
      const double v = cl_nearest(g, dists, n, w, b, 0, n, &x);  // (includes the g.sync()s that publish chain[tip])
      if (x != n && v < min) { min = v; a = x; }
      b = a;
      a = w.chain[tip++];
      if (b == w.chain[tip - 2]) break;
    }
    const uint32_t i1 = a < b ? a : b, i2 = a < b ? b : a;
    const double sa = (double)w.size[i1], sb = (double)w.size[i2];
    g.sync();  // every lane has read size[] and the chain
    if (g.lane() == 0) {
      w.raw_a[step] = a; w.raw_b[step] = b; w.raw_d[step] = min;
      w.size[i2] += w.size[i1];
      w.active[i1] = 0;
    }
    g.sync();
    for (uint32_t x = (uint32_t)g.lane(); x < n; x += (uint32_t)g.size()) {
      if (!w.active[x] || x == i2) continue;
      const double sx = (double)w.size[x];
      const double da = *cl_dref(dists, n, x, i1);
      double *db = cl_dref(dists, n, x, i2);
      const double numerator =
          TRGT_FSUB(TRGT_FADD(TRGT_FMUL(TRGT_FADD(sx, sa), da), TRGT_FMUL(TRGT_FADD(sx, sb), *db)), TRGT_FMUL(sx, min));
      const double denom = TRGT_FADD(TRGT_FADD(sa, sb), sx);
      *db = numerator / denom;
    }
    g.sync();
  }
  // stable sort by height (rank = number of merges that come before), then union-find relabelling
  for (uint32_t s = (uint32_t)g.lane(); s + 1 < n; s += (uint32_t)g.size()) {
    uint32_t rank = 0;
    const double ds = w.raw_d[s];
    for (uint32_t t = 0; t + 1 < n; t++) rank += (w.raw_d[t] < ds || (w.raw_d[t] == ds && t < s)) ? 1u : 0u;
    w.step_c1[rank] = w.raw_a[s];
    w.step_c2[rank] = w.raw_b[s];
    w.step_d[rank] = ds;
  }
  g.sync();
  if (g.lane() == 0) {
    for (uint32_t i = 0; i < 2 * n - 1; i++) { w.parent[i] = i; w.csize[i] = i < n ? 1u : 0u; }
    for (uint32_t s = 0; s + 1 < n; s++) {
      uint32_t ra = w.step_c1[s], rb = w.step_c2[s];
      while (w.parent[ra] != ra) ra = w.parent[ra];
      while (w.parent[rb] != rb) rb = w.parent[rb];
      const uint32_t lab = n + s;
      w.parent[ra] = lab; w.parent[rb] = lab;
      w.csize[lab] = w.csize[ra] + w.csize[rb];
      w.step_c1[s] = ra < rb ? ra : rb;
      w.step_c2[s] = ra < rb ? rb : ra;
      w.step_d[s] = sqrt(w.step_d[s]);
      w.step_size[s] = w.csize[lab];
    }
  }
  g.sync();
}

// cluster(): group id of every sequence in w.group, returns the number of groups
template <class G>
TRGT_HD int cl_cluster(const G &g, double *dists, uint32_t n, const ClusterWs &w) {
  if (n == 1) { if (g.lane() == 0) w.group[0] = 0; g.sync(); return 1; }
  if (n == 2) { if (g.lane() == 0) { w.group[0] = 0; w.group[1] = 1; } g.sync(); return 2; }
  cl_ward_linkage(g, dists, n, w);
  int num_groups = 0;
  if (g.lane() == 0) {
    double cutoff = 0.0;
    uint32_t min_cluster_size = (uint32_t)floor(0.01 * (double)n + 0.5);  // f64::round
    if (min_cluster_size < 2) min_cluster_size = 2;
    for (uint32_t s = n - 1; s-- > 0;) {
      const uint32_t c1 = w.step_c1[s], c2 = w.step_c2[s];
      const uint32_t s1 = c1 < n ? 1u : w.step_size[c1 - n], s2 = c2 < n ? 1u : w.step_size[c2 - n];
      if ((s1 < s2 ? s1 : s2) >= min_cluster_size) { cutoff = w.step_d[s] - 0.0001; break; }
    }
    if (cutoff == 0.0) {  // homozygous: split reads across alleles equally
      for (uint32_t i = 0; i < n; i++) w.group[i] = i & 1u;
      num_groups = 2;
    } else {
      for (uint32_t i = 0; i < 2 * n - 1; i++) w.membership[i] = -1;
      for (uint32_t s = n - 1; s-- > 0;) {
        const uint32_t cl = s + n;
        if (w.step_d[s] <= cutoff) {
          if (w.membership[cl] < 0) w.membership[cl] = num_groups++;
          w.membership[w.step_c1[s]] = w.membership[cl];
          w.membership[w.step_c2[s]] = w.membership[cl];
        }
      }
      for (uint32_t i = 0; i < n; i++) w.group[i] = w.membership[i] >= 0 ? (uint32_t)w.membership[i] : (uint32_t)num_groups++;
    }
  }
  num_groups = g.bcast0(num_groups);
  g.sync();
  return num_groups;
}

// central_read() of the members (ascending indices) of one group, on the matrix as the linkage left it
template <class G>
TRGT_HD uint32_t cl_central_read(const G &g, const double *dists, uint32_t n, const ClusterWs &w, uint32_t gsz) {
  if (gsz <= 2) return w.members[0];
  for (uint32_t k = (uint32_t)g.lane(); k < gsz; k += (uint32_t)g.size()) {
    double sum = 0.0;  // the reference adds the partners of k in ascending order
    for (uint32_t p = 0; p < gsz; p++) {
      if (p == k) continue;
      const size_t i1 = p < k ? w.members[p] : w.members[k], i2 = p < k ? w.members[k] : w.members[p];
      sum = TRGT_FADD(sum, dists[(size_t)n * i1 - i1 * (i1 + 3) / 2 + i2 - 1]);
    }
    w.raw_d[k] = sum;
  }
  g.sync();
  uint32_t best = 0;
  for (uint32_t k = 1; k < gsz; k++)
    if (w.raw_d[k] < w.raw_d[best]) best = k;  // min_by: the first minimum
  const uint32_t r = w.members[best];
  g.sync();
  return r;
}

// genotype() :57-72 up to the make_consensus calls: sel[i] = 0 (group1: the largest group, the later one among
// equals), 1 (group2), 2 (neither); central[2] = their central reads (0xFFFFFFFF: no such group).  Returns the
// number of groups.
template <class G>
TRGT_HD int cl_cluster_locus(const G &g, double *dists, uint32_t n, const ClusterWs &w, int32_t *sel, uint32_t *central) {
  if (n == 0) {
    if (g.lane() == 0) { central[0] = 0xFFFFFFFFu; central[1] = 0xFFFFFFFFu; }
    return 0;
  }
  const int ng = cl_cluster(g, dists, n, w);
  int g1 = -1, g2 = -1;
  if (g.lane() == 0) {
    // group sizes in csize (free after the linkage); stable ascending sort + two pops = last maximum, then the next
    for (int q = 0; q < ng; q++) w.csize[q] = 0;
    for (uint32_t i = 0; i < n; i++) w.csize[w.group[i]]++;
    for (int q = 0; q < ng; q++)
      if (g1 < 0 || w.csize[q] >= w.csize[g1]) g1 = q;
    for (int q = 0; q < ng; q++)
      if (q != g1 && (g2 < 0 || w.csize[q] >= w.csize[g2])) g2 = q;
  }
  g1 = g.bcast0(g1);
  g2 = g.bcast0(g2);
  uint32_t cen[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
  for (int which = 0; which < 2; which++) {
    const int q = which == 0 ? g1 : g2;
    if (q < 0) continue;
    int m = 0;
    if (g.lane() == 0) {
      for (uint32_t i = 0; i < n; i++)
        if ((int)w.group[i] == q) w.members[m++] = i;
    }
    m = g.bcast0(m);
    g.sync();
    cen[which] = cl_central_read(g, dists, n, w, (uint32_t)m);
  }
  for (uint32_t i = (uint32_t)g.lane(); i < n; i += (uint32_t)g.size())
    sel[i] = (int)w.group[i] == g1 ? 0 : ((int)w.group[i] == g2 ? 1 : 2);
  if (g.lane() == 0) { central[0] = cen[0]; central[1] = cen[1]; }
  g.sync();
  return ng;
}

}  // namespace trgt
