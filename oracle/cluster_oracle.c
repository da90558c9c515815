/*
 * cluster_oracle.c -- TEST INFRASTRUCTURE ONLY (see trgt_oracle.h).
 *
 * The cluster genotyper's glue between get_dist_matrix and make_consensus:
 *   cluster()       src/trgt/genotype/genotype_cluster.rs:154-227
 *   central_read()  src/trgt/genotype/genotype_cluster.rs:12-39
 *   the choice of the two largest groups, :64-69
 *
 * cluster() calls kodama::linkage(dists, n, Method::Ward) (kodama 0.3.0, Cargo.lock:839-842), which is
 * NOT vendored in /root/reference.  kodama documents itself as a port of fastcluster (Muellner 2013); for
 * Ward it runs the nearest-neighbour chain on the SQUARED dissimilarities, updates the condensed matrix IN
 * PLACE with the Lance-Williams formula, sorts the merges by dissimilarity (stable), relabels them with a
 * union-find so that step i creates cluster n + i, and takes square roots of the merge heights.  That
 * published algorithm is restated here (tro_ward_linkage).  PARITY: pinned on
 * scipy.cluster.hierarchy.linkage(method="ward") -- the same fastcluster algorithm -- for merge structure,
 * cluster sizes and heights (tests/test_cluster.py); UNPINNED against kodama itself for the floating-point
 * association order of the Lance-Williams update (this file uses kodama's
 * ((x+a) d_a + (x+b) d_b - x d_ab) / (a+b+x)) and for the order of equal-height merges.
 *
 * The reference passes `&mut dists` to linkage and afterwards reads the SAME array in central_read
 * (genotype_cluster.rs:60,74,82-83,161): central_read therefore sees the squared, partially updated matrix
 * the chain leaves behind.  tro_ward_linkage reproduces that side effect; tro_cluster_locus chains the three
 * steps the way genotype() does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "trgt_oracle.h"

/* condensed index of (i, j), i < j */
static size_t cidx(size_t n, size_t i, size_t j) { return i * n - i * (i + 1) / 2 + (j - i - 1); }
static double *dref(double *d, size_t n, size_t i, size_t j) { return i < j ? &d[cidx(n, i, j)] : &d[cidx(n, j, i)]; }

typedef struct {
  uint32_t a, b;
  double d;
  uint32_t order; /* position in which the chain produced the merge (stable sort key) */
} raw_step;

static int cmp_step(const void *x, const void *y) {
  const raw_step *a = (const raw_step *)x, *b = (const raw_step *)y;
  if (a->d < b->d) return -1;
  if (a->d > b->d) return 1;
  return a->order < b->order ? -1 : (a->order > b->order ? 1 : 0);
}

static uint32_t uf_find(uint32_t *parent, uint32_t x) {
  while (parent[x] != x) {
    parent[x] = parent[parent[x]];
    x = parent[x];
  }
  return x;
}

/* kodama::linkage(dists, n, Method::Ward): dists (condensed, n(n-1)/2) is MODIFIED IN PLACE.
 * steps_out[n-1]: cluster1 < cluster2 (labels: observations 0..n-1, step i creates n+i), dissimilarity, size. */
int tro_ward_linkage(double *dists, uint32_t n, tro_linkage_step *steps_out) {
  if (n < 2) return 0;
  const size_t np = (size_t)n * (n - 1) / 2;
  for (size_t k = 0; k < np; k++) dists[k] = dists[k] * dists[k]; /* Ward works on squared dissimilarities */
  uint32_t *chain = (uint32_t *)malloc(sizeof(uint32_t) * (n + 2));
  uint8_t *active = (uint8_t *)malloc(n);
  uint32_t *size = (uint32_t *)malloc(sizeof(uint32_t) * n);
  raw_step *raw = (raw_step *)malloc(sizeof(raw_step) * (n - 1));
  if (!chain || !active || !size || !raw) { free(chain); free(active); free(size); free(raw); return -1; }
  for (uint32_t i = 0; i < n; i++) { active[i] = 1; size[i] = 1; }
  uint32_t tip = 0;
  for (uint32_t step = 0; step + 1 < n; step++) {
    uint32_t a, b;
    double min;
    if (tip <= 3) { /* new chain from the first active observation; its nearest neighbour by a strict '<' scan */
      a = 0;
      while (!active[a]) a++;
      chain[0] = a;
      tip = 1;
      b = a + 1;
      while (!active[b]) b++;
      min = *dref(dists, n, a, b);
      for (uint32_t i = b + 1; i < n; i++)
        if (active[i] && *dref(dists, n, a, i) < min) { min = *dref(dists, n, a, i); b = i; }
    } else {
      tip -= 3;
      a = chain[tip - 1];
      b = chain[tip];
      min = *dref(dists, n, a, b);
    }
    do { /* grow the chain until two observations are each other's nearest neighbours (ties keep the predecessor) */
      chain[tip] = b;
      for (uint32_t i = 0; i < b; i++)
        if (active[i] && *dref(dists, n, i, b) < min) { min = *dref(dists, n, i, b); a = i; }
      for (uint32_t i = b + 1; i < n; i++)
        if (active[i] && *dref(dists, n, b, i) < min) { min = *dref(dists, n, b, i); a = i; }
      b = a;
      a = chain[tip++];
    } while (b != chain[tip - 2]);
    raw[step].a = a; raw[step].b = b; raw[step].d = min; raw[step].order = step;
    uint32_t i1 = a < b ? a : b, i2 = a < b ? b : a; /* the smaller index is retired, the larger one is the merged cluster */
    const double sa = (double)size[i1], sb = (double)size[i2];
    size[i2] += size[i1];
    active[i1] = 0;
    for (uint32_t x = 0; x < n; x++) {
      if (!active[x] || x == i2) continue;
      const double sx = (double)size[x];
      const double da = *dref(dists, n, x, i1);
      double *db = dref(dists, n, x, i2);
      const double numerator = ((sx + sa) * da) + ((sx + sb) * *db) - (sx * min);
      const double denom = sa + sb + sx;
      *db = numerator / denom;
    }
  }
  qsort(raw, n - 1, sizeof(raw_step), cmp_step);
  /* relabel: step i creates cluster n + i; cluster1 < cluster2 */
  uint32_t *parent = (uint32_t *)malloc(sizeof(uint32_t) * (2 * (size_t)n - 1));
  uint32_t *csize = (uint32_t *)malloc(sizeof(uint32_t) * (2 * (size_t)n - 1));
  if (!parent || !csize) { free(parent); free(csize); free(chain); free(active); free(size); free(raw); return -1; }
  for (uint32_t i = 0; i < 2 * n - 1; i++) { parent[i] = i; csize[i] = i < n ? 1 : 0; }
  for (uint32_t s = 0; s + 1 < n; s++) {
    uint32_t ra = uf_find(parent, raw[s].a), rb = uf_find(parent, raw[s].b);
    const uint32_t lab = n + s;
    parent[ra] = lab; parent[rb] = lab;
    csize[lab] = csize[ra] + csize[rb];
    steps_out[s].cluster1 = ra < rb ? ra : rb;
    steps_out[s].cluster2 = ra < rb ? rb : ra;
    steps_out[s].dissimilarity = sqrt(raw[s].d);
    steps_out[s].size = csize[lab];
  }
  free(parent); free(csize); free(chain); free(active); free(size); free(raw);
  return 0;
}

/* cluster(): genotype_cluster.rs:154-227.  group_out[n] = group id of every sequence (ids in the reference's
 * order of creation); returns the number of groups.  dists is modified in place (see above). */
int tro_cluster(uint32_t n, double *dists, uint32_t *group_out) {
  if (n < 2) { if (n == 1) group_out[0] = 0; return (int)n; }
  if (n == 2) { group_out[0] = 0; group_out[1] = 1; return 2; } /* vec![vec![0], vec![1]] */
  tro_linkage_step *steps = (tro_linkage_step *)malloc(sizeof(tro_linkage_step) * (n - 1));
  if (!steps || tro_ward_linkage(dists, n, steps) != 0) { free(steps); return -1; }
  double cutoff = 0.0;
  /* (MIN_SMALLER_FRAC * n).round(): f64::round rounds half away from zero */
  uint32_t min_cluster_size = (uint32_t)floor(0.01 * (double)n + 0.5);
  if (min_cluster_size < 2) min_cluster_size = 2;
  for (uint32_t s = n - 1; s-- > 0;) {
    const uint32_t c1 = steps[s].cluster1, c2 = steps[s].cluster2;
    const uint32_t s1 = c1 < n ? 1 : steps[c1 - n].size, s2 = c2 < n ? 1 : steps[c2 - n].size;
    if ((s1 < s2 ? s1 : s2) >= min_cluster_size) { cutoff = steps[s].dissimilarity - 0.0001; break; }
  }
  if (cutoff == 0.0) { /* homozygous: split reads across alleles equally */
    for (uint32_t i = 0; i < n; i++) group_out[i] = i & 1u;
    free(steps);
    return 2;
  }
  int num_groups = 0;
  const uint32_t num_nodes = 2 * n - 1;
  int *membership = (int *)malloc(sizeof(int) * num_nodes);
  if (!membership) { free(steps); return -1; }
  for (uint32_t i = 0; i < num_nodes; i++) membership[i] = -1;
  for (uint32_t s = n - 1; s-- > 0;) {
    const uint32_t cl = s + n;
    if (steps[s].dissimilarity <= cutoff) {
      if (membership[cl] < 0) membership[cl] = num_groups++;
      membership[steps[s].cluster1] = membership[cl];
      membership[steps[s].cluster2] = membership[cl];
    }
  }
  for (uint32_t i = 0; i < n; i++) group_out[i] = membership[i] >= 0 ? (uint32_t)membership[i] : (uint32_t)num_groups++;
  free(membership);
  free(steps);
  return num_groups;
}

/* central_read(): genotype_cluster.rs:12-39; group = indices into the locus' sequences */
uint32_t tro_central_read(uint32_t num_seqs, const uint32_t *group, uint32_t group_size, const double *dists) {
  if (group_size <= 2) return group[0];
  double *sums = (double *)calloc(group_size, sizeof(double));
  for (uint32_t i = 0; i + 1 < group_size; i++)
    for (uint32_t j = i + 1; j < group_size; j++) {
      const size_t i1 = group[i], i2 = group[j];
      const size_t mat = (size_t)num_seqs * i1 - i1 * (i1 + 3) / 2 + i2 - 1;
      sums[i] += dists[mat];
      sums[j] += dists[mat];
    }
  uint32_t best = 0; /* min_by: the first minimum */
  for (uint32_t i = 1; i < group_size; i++)
    if (sums[i] < sums[best]) best = i;
  free(sums);
  return group[best];
}

/* genotype() up to the two make_consensus calls (:57-72) for one locus of n >= 1 repeat sequences:
 * dists = get_dist_matrix (condensed; modified), sel_out[n] = 0 for the members of group1 (the largest group,
 * the later one among equals: sort_by_key is stable and pop() takes from the back), 1 for group2, 2 for the
 * others; central_out[2] = central_read of group1 / group2 (0xFFFFFFFF when there is no second group). */
int tro_cluster_locus(uint32_t n, double *dists, uint8_t *sel_out, uint32_t *central_out) {
  central_out[0] = central_out[1] = 0xFFFFFFFFu;
  if (n == 0) return 0;
  if (n == 1) { sel_out[0] = 0; central_out[0] = 0; return 1; }
  uint32_t *group = (uint32_t *)malloc(sizeof(uint32_t) * n);
  const int ng = tro_cluster(n, dists, group);
  if (ng < 0) { free(group); return -1; }
  uint32_t *gsize = (uint32_t *)calloc((size_t)ng, sizeof(uint32_t));
  for (uint32_t i = 0; i < n; i++) gsize[group[i]]++;
  /* stable ascending sort by size, then pop twice: the largest (last among equals), then the next */
  int g1 = -1, g2 = -1;
  for (int g = 0; g < ng; g++)
    if (g1 < 0 || gsize[g] >= gsize[g1]) g1 = g;
  for (int g = 0; g < ng; g++)
    if (g != g1 && (g2 < 0 || gsize[g] >= gsize[g2])) g2 = g;
  uint32_t *members = (uint32_t *)malloc(sizeof(uint32_t) * n);
  for (int which = 0; which < 2; which++) {
    const int g = which == 0 ? g1 : g2;
    if (g < 0) continue;
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; i++)
      if ((int)group[i] == g) members[m++] = i;
    central_out[which] = tro_central_read(n, members, m, dists);
  }
  for (uint32_t i = 0; i < n; i++) sel_out[i] = (int)group[i] == g1 ? 0 : ((int)group[i] == g2 ? 1 : 2);
  free(members); free(gsize); free(group);
  return ng;
}
