// wfa_core.h -- gap-affine wavefront alignment restructured for one cooperating group per pair.
//
// Replaces the alignment the reference performs through WFA2-lib (wfa2-sys 0.1.0, un-vendored):
//   WFAligner::align_ends_free / align_end_to_end    src/wfaligner.rs:489-528
//   count_matches :988, get_alignment_span :864, get_sam_cigar :932 (run-length SAM words)
// as called by find_spans (src/trgt/genotype/span_locater.rs:7-30), utils::align
// (src/utils/align.rs:14-28) and get_dist (src/trgt/genotype/genotype_cluster.rs:236-248).
//
// Semantics (recurrences, NULL handling, lowest-diagonal termination, back-trace priority
// M > D-ext > D-open > I-ext > I-open) are those pinned by the reference's golden tests, see
// SURVEY.md section 8(c).  The organisation is this repo's own:
//
//   pass 1  wfa_score_ring   forward wavefronts with only the last max(x,o+e)+1 M and e+1 I/D
//                            wavefronts alive (a ring, on chip when it fits): finds the optimal
//                            score s*, the terminating diagonal k* and offset.  No history.
//   pass 2  wfa_trace        recomputes ONLY the dependency cone of (s*, k*):
//                            |k - k*| <= (s* - s) / e at score s.  Every predecessor of a cone cell is
//                            a cone cell, so the recomputed offsets equal the full computation's,
//                            and the back-trace never leaves the cone.  For the wide-and-shallow
//                            flank problem (T+1 start diagonals, s* ~ 6) this replaces the
//                            reference's 12 B x (T+1) x (s*+1) history by a few hundred bytes.
//
// Lanes of the group stride over diagonals; all communication is through the wavefront arrays
// followed by g.sync(), so the same code runs serially in the CPU unit tests.
#pragma once
#include <limits.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "coop.h"

namespace trgt {

#define TRGT_WFA_NULL (INT_MIN / 2)

// per-item status (numerically WFA2's WF_STATUS_*, src/wfaligner.rs:132-159)
#define TRGT_WFA_OK 0
#define TRGT_WFA_MAX_STEPS (-100)
#define TRGT_WFA_OOM (-200)

struct WfaProb {
  const uint8_t *p; int P;  // pattern (consumed by 'D')
  const uint8_t *t; int T;  // text (consumed by 'I')
  int x, oe, e;             // mismatch, gap_open + gap_extend, gap_extend
  int pbf, pef, tbf, tef;   // ends-free allowances (all 0 = end-to-end)
  int blo, bhi;             // diagonal band the forward pass may use ([-P, T] = exact, unrestricted)
};

TRGT_HD void wfa_unband(WfaProb &pr) { pr.blo = -pr.P; pr.bhi = pr.T; }

// 8 bytes starting at any byte address, assembled from the three aligned 32-bit words that cover
// them (two funnel shifts).  May touch up to 11 bytes past `a`: every sequence buffer of the engine
// is padded by 16 bytes.
TRGT_HD uint64_t wfa_ld64u(const uint8_t *a) {
#if defined(__CUDA_ARCH__)
  const uintptr_t addr = (uintptr_t)a;
  const uint32_t *w = (const uint32_t *)(addr & ~(uintptr_t)3);
  const unsigned sh = (unsigned)(addr & 3u) * 8u;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
#else
  uint64_t v;
  memcpy(&v, a, 8);
  return v;
#endif
}

TRGT_HD int wfa_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

// number of leading bytes (< n) on which a and b agree, 8 bytes per step
TRGT_HD int wfa_match_len(const uint8_t *a, const uint8_t *b, int n) {
  int i = 0;
  while (i < n) {
    const uint64_t x = wfa_ld64u(a + i) ^ wfa_ld64u(b + i);
    if (x) {
      i += wfa_ctz64(x) >> 3;
      return i < n ? i : n;
    }
    i += 8;
  }
  return n;
}

// the same, with the lanes of a group each comparing one 8-byte chunk per step (all lanes must
// call with the same arguments)
template <class G>
TRGT_HD int wfa_coop_match_len(const G &g, const uint8_t *a, const uint8_t *b, int n) {
  for (int base = 0; base < n; base += 8 * g.size()) {
    const int i = base + 8 * g.lane();
    int m = INT_MAX;
    if (i < n) {
      const uint64_t x = wfa_ld64u(a + i) ^ wfa_ld64u(b + i);
      if (x) m = i + (wfa_ctz64(x) >> 3);
    }
    m = g.min_i(m);
    if (m != INT_MAX) return m < n ? m : n;
  }
  return n;
}

struct WfaEnd {
  int status;
  int s;    // optimal cost (score = -s)
  int k;    // terminating diagonal
  int off;  // terminating offset (text position)
};

// One stored wavefront.  Arrays are indexed by k - base; [lo,hi] is what is stored, [tlo,thi]
// the range the full computation would hold (they differ only in cone mode).
struct WfaView {
  int *m, *i, *d;
  int base, lo, hi, tlo, thi;
};

TRGT_HD int wfa_imax(int a, int b) { return a > b ? a : b; }
TRGT_HD int wfa_imin(int a, int b) { return a < b ? a : b; }

TRGT_HD int wfa_at(const int *a, const WfaView &v, int k) {
  return (a != nullptr && k >= v.lo && k <= v.hi) ? a[k - v.base] : TRGT_WFA_NULL;
}

TRGT_HD WfaView wfa_null_view() {
  WfaView v;
  v.m = v.i = v.d = nullptr;
  v.base = 0; v.lo = 1; v.hi = 0; v.tlo = 1; v.thi = 0;
  return v;
}

// Match extension along diagonal k from text offset h, first 8 bytes only (done by the lane that
// owns the diagonal).  *more is set when all 8 matched and bases remain: the group finishes such
// diagonals together in wfa_extend_finish.
TRGT_HD int wfa_extend8(const WfaProb &pr, int k, int h, int *more) {
  const int v = h - k;
  const int n = wfa_imin(pr.P - v, pr.T - h);
  *more = 0;
  if (n <= 0 || pr.p[v] != pr.t[h]) return h;  // most diagonals of an ends-free wavefront stop here
  const uint64_t x = wfa_ld64u(pr.p + v) ^ wfa_ld64u(pr.t + h);
  const int m = x ? (wfa_ctz64(x) >> 3) : 8;
  if (m >= n) return h + n;
  if (m < 8) return h + m;
  *more = 1;
  return h + 8;
}

// Long extensions: one diagonal at a time, 8 bytes per lane per step.  `more`, k, h are this lane's
// pending diagonal (more = 0: none); returns this lane's final offset.
template <class G>
TRGT_HD int wfa_extend_finish(const G &g, const WfaProb &pr, int more, int k, int h) {
  for (;;) {
    const int leader = g.min_i(more ? g.lane() : INT_MAX);
    if (leader == INT_MAX) return h;
    const int lk = g.bcast(k, leader), lh = g.bcast(h, leader);
    const int lv = lh - lk;
    const int n = wfa_imin(pr.P - lv, pr.T - lh);
    const int m = wfa_coop_match_len(g, pr.p + lv, pr.t + lh, n);
    if (g.lane() == leader) { h = lh + m; more = 0; }
  }
}

// true diagonal range of wavefront s from its sources' true ranges; returns false if null
TRGT_HD bool wfa_next_range(const WfaProb &pr, const WfaView &vx, const WfaView &vo, const WfaView &ve,
                            int *lo_out, int *hi_out) {
  int lo = INT_MAX, hi = INT_MIN;
  if (vx.tlo <= vx.thi) { lo = wfa_imin(lo, vx.tlo); hi = wfa_imax(hi, vx.thi); }
  if (vo.tlo <= vo.thi) { lo = wfa_imin(lo, vo.tlo); hi = wfa_imax(hi, vo.thi); }
  if (ve.tlo <= ve.thi && ve.i != nullptr) { lo = wfa_imin(lo, ve.tlo); hi = wfa_imax(hi, ve.thi); }
  if (lo > hi) return false;
  lo -= 1; hi += 1;
  lo = wfa_imax(lo, pr.blo);  // blo >= -P, bhi <= T: cells outside [-P, T] can never be valid
  hi = wfa_imin(hi, pr.bhi);
  if (lo > hi) return false;
  *lo_out = lo; *hi_out = hi;
  return true;
}

// cells dst.lo..dst.hi of one wavefront from its three source wavefronts, then match extension
template <class G>
TRGT_HD void wfa_compute(const G &g, const WfaProb &pr, const WfaView &vx, const WfaView &vo,
                         const WfaView &ve, const WfaView &dst) {
  for (int kb = dst.lo; kb <= dst.hi; kb += g.size()) {  // uniform trip count: the finish step is collective
    const int k = kb + g.lane();
    int mx = TRGT_WFA_NULL, more = 0;
    if (k <= dst.hi) {
      const int i1 = wfa_imax(wfa_at(vo.m, vo, k - 1), wfa_at(ve.i, ve, k - 1)) + 1;
      const int d1 = wfa_imax(wfa_at(vo.m, vo, k + 1), wfa_at(ve.d, ve, k + 1));
      const int mm = wfa_at(vx.m, vx, k) + 1;
      mx = wfa_imax(mm, wfa_imax(i1, d1));
      const int h = mx, v = mx - k;
      if (mx < 0 || h > pr.T || v > pr.P || v < 0) mx = TRGT_WFA_NULL;
      else mx = wfa_extend8(pr, k, mx, &more);
      dst.i[k - dst.base] = i1;
      dst.d[k - dst.base] = d1;
    }
    mx = wfa_extend_finish(g, pr, more, k, mx);
    if (k <= dst.hi) dst.m[k - dst.base] = mx;
  }
}

// score-0 wavefront: M[k] = max(k, 0) on [-pbf, tbf], extended
template <class G>
TRGT_HD void wfa_init(const G &g, const WfaProb &pr, const WfaView &dst) {
  for (int kb = dst.lo; kb <= dst.hi; kb += g.size()) {
    const int k = kb + g.lane();
    int h = TRGT_WFA_NULL, more = 0;
    if (k <= dst.hi) h = wfa_extend8(pr, k, k >= 0 ? k : 0, &more);
    h = wfa_extend_finish(g, pr, more, k, h);
    if (k <= dst.hi) dst.m[k - dst.base] = h;
  }
}

// lowest diagonal of `w` that satisfies the end condition, INT_MAX if none
template <class G>
TRGT_HD int wfa_terminated(const G &g, const WfaProb &pr, const WfaView &w) {
  int best = INT_MAX;
  for (int k = w.lo + g.lane(); k <= w.hi; k += g.size()) {
    const int h = w.m[k - w.base];
    if (h < 0) continue;
    const int v = h - k;
    if (v < 0 || v > pr.P || h > pr.T) continue;
    if ((h >= pr.T && pr.P - v <= pr.pef) || (v >= pr.P && pr.T - h <= pr.tef)) {
      best = k;
      break;
    }
  }
  return g.min_i(best);
}

// ---------------------------------------------------------------- pass 1: ring ----------

TRGT_HD int wfa_ring_depth_m(const WfaProb &pr) { return wfa_imax(pr.x, pr.oe) + 1; }
TRGT_HD int wfa_ring_depth_g(const WfaProb &pr) { return pr.e + 1; }
// stride of one ring slot: every diagonal a valid cell can live on
TRGT_HD int wfa_ring_stride(const WfaProb &pr) { return pr.bhi - pr.blo + 1; }
TRGT_HD size_t wfa_ring_ints(const WfaProb &pr) {
  const size_t dm = (size_t)wfa_ring_depth_m(pr), dg = (size_t)wfa_ring_depth_g(pr);
  return 2 * dm + (dm + 2 * dg) * (size_t)wfa_ring_stride(pr);
}
// generous bound on the optimal cost: align min(P,T) columns as mismatches plus one gap
TRGT_HD int wfa_score_cap(const WfaProb &pr) {
  const long long c = (long long)pr.x * wfa_imin(pr.P, pr.T) + 2LL * pr.oe + (long long)pr.e * (pr.P + pr.T) + 8;
  return c > (long long)(INT_MAX / 4) ? INT_MAX / 4 : (int)c;
}

TRGT_HD WfaView wfa_ring_view(const WfaProb &pr, int *ring, int s) {
  if (s < 0) return wfa_null_view();
  const int dm = wfa_ring_depth_m(pr), dg = wfa_ring_depth_g(pr), W = wfa_ring_stride(pr);
  int *meta = ring;
  int *mbase = ring + 2 * dm;
  int *ibase = mbase + (size_t)dm * W;
  int *dbase = ibase + (size_t)dg * W;
  WfaView v;
  v.base = pr.blo;
  v.tlo = v.lo = meta[2 * (s % dm)];
  v.thi = v.hi = meta[2 * (s % dm) + 1];
  v.m = mbase + (size_t)(s % dm) * W;
  if (s >= 1) {
    v.i = ibase + (size_t)(s % dg) * W;
    v.d = dbase + (size_t)(s % dg) * W;
  } else {
    v.i = v.d = nullptr;
  }
  return v;
}

// Forward pass without history.  `ring` holds wfa_ring_ints(pr) ints (shared or global memory).
template <class G>
TRGT_HD WfaEnd wfa_score_ring(const G &g, const WfaProb &pr, int *ring, int s_cap) {
  WfaEnd out;
  out.status = TRGT_WFA_OK; out.s = 0; out.k = 0; out.off = 0;
  const int dm = wfa_ring_depth_m(pr);
  int *meta = ring;
  if (g.lane() == 0) {
    meta[0] = wfa_imax(-pr.pbf, pr.blo);
    meta[1] = wfa_imin(pr.tbf, pr.bhi);
  }
  g.sync();
  {
    WfaView w0 = wfa_ring_view(pr, ring, 0);
    wfa_init(g, pr, w0);
  }
  g.sync();
  int s = 0;
  for (;;) {
    WfaView cur = wfa_ring_view(pr, ring, s);
    if (cur.lo <= cur.hi) {
      const int kmin = wfa_terminated(g, pr, cur);
      if (kmin != INT_MAX) {
        out.s = s; out.k = kmin; out.off = cur.m[kmin - cur.base];
        return out;
      }
    }
    s++;
    if (s > s_cap) { out.status = TRGT_WFA_MAX_STEPS; out.s = s; return out; }
    const WfaView vx = wfa_ring_view(pr, ring, s - pr.x);
    const WfaView vo = wfa_ring_view(pr, ring, s - pr.oe);
    const WfaView ve = wfa_ring_view(pr, ring, s - pr.e);
    int lo, hi;
    const bool live = wfa_next_range(pr, vx, vo, ve, &lo, &hi);
    g.sync();  // everyone has read the slot that is about to be recycled
    if (g.lane() == 0) {
      meta[2 * (s % dm)] = live ? lo : 1;
      meta[2 * (s % dm) + 1] = live ? hi : 0;
    }
    if (live) {
      WfaView dst = wfa_ring_view(pr, ring, s);
      dst.lo = dst.tlo = lo; dst.hi = dst.thi = hi;
      wfa_compute(g, pr, vx, vo, ve, dst);
    }
    g.sync();
  }
}

// ---------------------------------------------------------------- pass 2: cone + back-trace ----

#define TRGT_WFA_META 6  // ints per score in the history header: tlo, thi, lo, hi, data offset, unused

TRGT_HD int wfa_cone_radius(const WfaProb &pr, int s_end, int s) { return (s_end - s) / pr.e; }

// ints of workspace wfa_trace needs (upper bound: assumes every cone row is as wide as allowed)
TRGT_HD size_t wfa_trace_ints(const WfaProb &pr, int s_end) {
  size_t n = (size_t)TRGT_WFA_META * ((size_t)s_end + 1);
  const long long wmax = (long long)pr.P + pr.T + 1;
  for (int s = 0; s <= s_end; s++) {
    long long w = 2LL * wfa_cone_radius(pr, s_end, s) + 1;
    if (w > wmax) w = wmax;
    n += (size_t)w * (s == 0 ? 1 : 3);
  }
  return n;
}

TRGT_HD WfaView wfa_hist_view(int *ws, int s) {
  if (s < 0) return wfa_null_view();
  const int *meta = ws + (size_t)TRGT_WFA_META * s;
  WfaView v;
  v.tlo = meta[0]; v.thi = meta[1]; v.lo = meta[2]; v.hi = meta[3];
  v.base = v.lo;
  const int w = v.hi >= v.lo ? v.hi - v.lo + 1 : 0;
  int *data = ws + (size_t)(unsigned)meta[4] + (((size_t)(unsigned)meta[5]) << 32);
  v.m = data;
  if (s >= 1 && w > 0) { v.i = data + w; v.d = data + 2 * (size_t)w; }
  else if (s >= 1) { v.i = data; v.d = data; }  // present but empty: keeps "has I/D" for range tracking
  else { v.i = v.d = nullptr; }
  return v;
}

// Recompute the dependency cone of (s_end, k_end) with full history in ws.
// Returns 0, or TRGT_WFA_OOM if cap_ints is too small.
template <class G>
TRGT_HD int wfa_trace_forward(const G &g, const WfaProb &pr, int s_end, int k_end, int *ws, size_t cap_ints) {
  size_t top = (size_t)TRGT_WFA_META * ((size_t)s_end + 1);
  if (top > cap_ints) return TRGT_WFA_OOM;
  for (int s = 0; s <= s_end; s++) {
    int tlo, thi;
    bool live;
    WfaView vx = wfa_null_view(), vo = wfa_null_view(), ve = wfa_null_view();
    if (s == 0) {
      tlo = wfa_imax(-pr.pbf, pr.blo); thi = wfa_imin(pr.tbf, pr.bhi); live = tlo <= thi;
    } else {
      vx = wfa_hist_view(ws, s - pr.x);
      vo = wfa_hist_view(ws, s - pr.oe);
      ve = wfa_hist_view(ws, s - pr.e);
      live = wfa_next_range(pr, vx, vo, ve, &tlo, &thi);
      if (!live) { tlo = 1; thi = 0; }
    }
    const int R = wfa_cone_radius(pr, s_end, s);
    int lo = live ? wfa_imax(tlo, k_end - R) : 1;
    int hi = live ? wfa_imin(thi, k_end + R) : 0;
    if (lo > hi) { lo = 1; hi = 0; }
    const size_t w = hi >= lo ? (size_t)(hi - lo + 1) : 0;
    const size_t need = w * (s == 0 ? 1 : 3);
    if (top + need > cap_ints) return TRGT_WFA_OOM;
    g.sync();
    if (g.lane() == 0) {
      int *meta = ws + (size_t)TRGT_WFA_META * s;
      meta[0] = tlo; meta[1] = thi; meta[2] = lo; meta[3] = hi;
      meta[4] = (int)(unsigned)(top & 0xffffffffu);
      meta[5] = (int)(unsigned)(top >> 32);
    }
    g.sync();
    if (w > 0) {
      WfaView dst = wfa_hist_view(ws, s);
      if (s == 0) wfa_init(g, pr, dst);
      else wfa_compute(g, pr, vx, vo, ve, dst);
    }
    top += need;
  }
  g.sync();
  return 0;
}

// Forward pass inside the band [pr.blo, pr.bhi] that keeps every wavefront (same layout as
// wfa_trace_forward, so wfa_backtrace can run on it) and stops at the first score whose wavefront
// satisfies the end condition.  Cells a band drops only ever lower the offsets of cells that are not
// on a cost-optimal alignment inside the band, so when every optimal alignment is known to lie
// inside the band (flank_seed_band) both the end cell and the back-trace equal the unbanded run's.
// Narrow-band specialisation: the band is at most one lane per diagonal.  Each lane owns one
// diagonal for the whole pass, rows have a fixed stride (no per-score bookkeeping to reload) and
// presence of a score is one bit of a mask.  Every band cell is evaluated at every live score; cells
// the general routine would leave outside a wavefront come out as (drifted) NULLs, which every
// consumer treats as NULL.  Writes the same history layout as the general routine.
// Which scores can have an M wavefront at all under a scoring (bit s, s <= s_cap < 64), and which a gap (I / D)
// wavefront: a property of the scoring alone.  A wavefront whose predecessors are all null is all null and reads
// exactly like an absent one, so the banded passes only compute the scores named here.
TRGT_HD unsigned long long wfa_live_scores(int x, int oe, int e, int s_cap, unsigned long long *live_gap) {
  unsigned long long live_m = 1ull, live_g = 0ull;
  for (int s = 1; s <= s_cap && s < 64; s++) {
    const int sx = s - x, so = s - oe, se = s - e;
    const bool g = (so >= 0 && ((live_m >> so) & 1ull)) || (se >= 1 && ((live_g >> se) & 1ull));
    if (g) live_g |= 1ull << s;
    if (g || (sx >= 0 && ((live_m >> sx) & 1ull))) live_m |= 1ull << s;
  }
  if (live_gap) *live_gap = live_g;
  return live_m;
}

template <class G>
TRGT_HD WfaEnd wfa_forward_band_hist_narrow(const G &g, const WfaProb &pr, int s_cap, int *ws, size_t cap_ints) {
  WfaEnd out;
  out.status = TRGT_WFA_OK; out.s = 0; out.k = 0; out.off = 0;
  const int W = pr.bhi - pr.blo + 1;
  const size_t hdr = (size_t)TRGT_WFA_META * ((size_t)s_cap + 1);
  const int idx = g.lane();
  const int k = pr.blo + idx;
  const bool mine = idx < W;
  unsigned long long present = 0;  // bit s: wavefront s is live
  for (int s = 0; s <= s_cap; s++) {
    const int sx = s - pr.x, so = s - pr.oe, se = s - pr.e;
    const bool px = sx >= 0 && ((present >> sx) & 1ull), po = so >= 0 && ((present >> so) & 1ull);
    const bool pe = se >= 1 && ((present >> se) & 1ull);  // score 0 has no I/D components
    const bool live = s == 0 || px || po || pe;
    const size_t base = hdr + (size_t)3 * W * s;
    if (live && base + (size_t)3 * W > cap_ints) { out.status = TRGT_WFA_OOM; return out; }
    if (g.lane() == 0) {
      int *meta = ws + (size_t)TRGT_WFA_META * s;
      meta[0] = meta[2] = live ? pr.blo : 1;
      meta[1] = meta[3] = live ? pr.bhi : 0;
      meta[4] = (int)(unsigned)(base & 0xffffffffu);
      meta[5] = (int)(unsigned)(base >> 32);
    }
    if (!live) continue;
    int mx = TRGT_WFA_NULL, more = 0;
    if (mine) {
      if (s == 0) {
        if (k >= -pr.pbf && k <= pr.tbf) mx = wfa_extend8(pr, k, k >= 0 ? k : 0, &more);
      } else {
        const int *mo = ws + hdr + (size_t)3 * W * (so > 0 ? so : 0);
        const int *me = ws + hdr + (size_t)3 * W * (se > 0 ? se : 0);
        const int *mxs = ws + hdr + (size_t)3 * W * (sx > 0 ? sx : 0);
        const int o_l = (po && idx > 0) ? mo[idx - 1] : TRGT_WFA_NULL;
        const int o_r = (po && idx + 1 < W) ? mo[idx + 1] : TRGT_WFA_NULL;
        const int i_l = (pe && idx > 0) ? me[W + idx - 1] : TRGT_WFA_NULL;
        const int d_r = (pe && idx + 1 < W) ? me[2 * W + idx + 1] : TRGT_WFA_NULL;
        const int i1 = wfa_imax(o_l, i_l) + 1;
        const int d1 = wfa_imax(o_r, d_r);
        const int mm = (px ? mxs[idx] : TRGT_WFA_NULL) + 1;
        mx = wfa_imax(mm, wfa_imax(i1, d1));
        const int h = mx, v = mx - k;
        if (mx < 0 || h > pr.T || v > pr.P || v < 0) mx = TRGT_WFA_NULL;
        else mx = wfa_extend8(pr, k, mx, &more);
        ws[base + W + idx] = i1;
        ws[base + 2 * W + idx] = d1;
      }
    }
    mx = wfa_extend_finish(g, pr, more, k, mx);
    if (mine) ws[base + idx] = mx;
    present |= 1ull << s;
    // end condition, lowest diagonal first
    int endk = INT_MAX;
    if (mine && mx >= 0) {
      const int h = mx, v = mx - k;
      if (v >= 0 && v <= pr.P && h <= pr.T &&
          ((h >= pr.T && pr.P - v <= pr.pef) || (v >= pr.P && pr.T - h <= pr.tef))) endk = k;
    }
    endk = g.min_i(endk);
    g.sync();  // rows of this score are visible to every lane from here on
    if (endk != INT_MAX) {
      out.s = s; out.k = endk; out.off = ws[base + (endk - pr.blo)];
      return out;
    }
  }
  out.status = TRGT_WFA_MAX_STEPS;
  return out;
}

template <class G>
TRGT_HD WfaEnd wfa_forward_band_hist(const G &g, const WfaProb &pr, int s_cap, int *ws, size_t cap_ints) {
  WfaEnd out;
  out.status = TRGT_WFA_OK; out.s = 0; out.k = 0; out.off = 0;
  size_t top = (size_t)TRGT_WFA_META * ((size_t)s_cap + 1);
  if (top > cap_ints) { out.status = TRGT_WFA_OOM; return out; }
  if (pr.bhi - pr.blo + 1 <= g.size() && g.size() > 1 && s_cap < 64)
    return wfa_forward_band_hist_narrow(g, pr, s_cap, ws, cap_ints);
  for (int s = 0; s <= s_cap; s++) {
    int lo, hi;
    bool live;
    WfaView vx = wfa_null_view(), vo = wfa_null_view(), ve = wfa_null_view();
    if (s == 0) {
      lo = wfa_imax(-pr.pbf, pr.blo); hi = wfa_imin(pr.tbf, pr.bhi); live = lo <= hi;
    } else {
      vx = wfa_hist_view(ws, s - pr.x);
      vo = wfa_hist_view(ws, s - pr.oe);
      ve = wfa_hist_view(ws, s - pr.e);
      live = wfa_next_range(pr, vx, vo, ve, &lo, &hi);
    }
    if (!live) { lo = 1; hi = 0; }
    const size_t w = live ? (size_t)(hi - lo + 1) : 0;
    const size_t need = w * (s == 0 ? 1 : 3);
    if (top + need > cap_ints) { out.status = TRGT_WFA_OOM; return out; }
    g.sync();
    if (g.lane() == 0) {
      int *meta = ws + (size_t)TRGT_WFA_META * s;
      meta[0] = lo; meta[1] = hi; meta[2] = lo; meta[3] = hi;
      meta[4] = (int)(unsigned)(top & 0xffffffffu);
      meta[5] = (int)(unsigned)(top >> 32);
    }
    g.sync();
    if (live) {
      const WfaView dst = wfa_hist_view(ws, s);
      if (s == 0) wfa_init(g, pr, dst);
      else wfa_compute(g, pr, vx, vo, ve, dst);
      g.sync();
      const int kmin = wfa_terminated(g, pr, dst);
      if (kmin != INT_MAX) {
        out.s = s; out.k = kmin; out.off = dst.m[kmin - dst.base];
        return out;
      }
    }
    top += need;
  }
  out.status = TRGT_WFA_MAX_STEPS;
  return out;
}

// The wavefront history as wfa_trace_forward / wfa_forward_band_hist* leave it in `ws`, read cell by cell
// (TRGT_WFA_NULL for a cell that is not stored).
struct WfaWsRow {
  WfaView v;
  TRGT_HD int m(int k) const { return wfa_at(v.m, v, k); }
  TRGT_HD int i(int k) const { return wfa_at(v.i, v, k); }
  TRGT_HD int d(int k) const { return wfa_at(v.d, v, k); }
};
struct WfaWsHist {
  int *ws;
  typedef WfaWsRow Row;
  TRGT_HD Row row(int s) const { Row r; r.v = wfa_hist_view(ws, s); return r; }  // the score's header, read once
};

// Back-trace from (s_end, k_end) through a wavefront history (any type with a row(s) whose m / i / d read a cell,
// TRGT_WFA_NULL for one that is not stored; see WfaWsHist); ops are handed to the sink from the LAST operation to the first.  One lane.
template <class Hist, class Sink>
TRGT_HD void wfa_backtrace_h(const WfaProb &pr, int s_end, int k_end, int off_end, const Hist &hist, Sink &sink) {
  enum { T_I1O = 1, T_I1E = 2, T_D1O = 5, T_D1E = 6, T_M = 9 };
  enum { C_M = 0, C_I = 1, C_D = 2 };
  int k = k_end, off = off_end;
  int v = off - k, h = off;
  sink.op('I', pr.T - h);  // free text tail
  sink.op('D', pr.P - v);  // unaligned pattern tail
  int mt = C_M, sc = s_end;
#define TRGT_PG(o, tag) ((o) < 0 ? (long long)TRGT_WFA_NULL : ((((long long)(o)) << 4) | (tag)))
  while (v > 0 && h > 0 && sc > 0) {
    const int sx = sc - pr.x, so = sc - pr.oe, se = sc - pr.e;
    const typename Hist::Row rx = hist.row(sx), ro = hist.row(so), re = hist.row(se);
    const long long c_m = TRGT_PG(rx.m(k) + 1, T_M);
    const long long c_io = TRGT_PG(ro.m(k - 1) + 1, T_I1O);
    const long long c_ie = TRGT_PG(re.i(k - 1) + 1, T_I1E);
    const long long c_do = TRGT_PG(ro.m(k + 1), T_D1O);
    const long long c_de = TRGT_PG(re.d(k + 1), T_D1E);
    long long mx;
    if (mt == C_M) {
      mx = c_m;
      if (c_io > mx) mx = c_io;
      if (c_ie > mx) mx = c_ie;
      if (c_do > mx) mx = c_do;
      if (c_de > mx) mx = c_de;
    } else if (mt == C_I) {
      mx = c_io > c_ie ? c_io : c_ie;
    } else {
      mx = c_do > c_de ? c_do : c_de;
    }
    if (mx < 0) break;
    if (mt == C_M) {
      const int mo = (int)(mx >> 4);
      sink.op('M', off - mo);
      off = mo;
      v = off - k; h = off;
      if (v <= 0 || h <= 0) break;
    }
    switch ((int)(mx & 15)) {
      case T_M:   sc = sx; mt = C_M; sink.op('X', 1); off -= 1; break;
      case T_I1O: sc = so; mt = C_M; sink.op('I', 1); k -= 1; off -= 1; break;
      case T_I1E: sc = se; mt = C_I; sink.op('I', 1); k -= 1; off -= 1; break;
      case T_D1O: sc = so; mt = C_M; sink.op('D', 1); k += 1; break;
      case T_D1E: sc = se; mt = C_D; sink.op('D', 1); k += 1; break;
      default: break;
    }
    v = off - k; h = off;
  }
#undef TRGT_PG
  if (mt == C_M && v > 0 && h > 0) {
    const int n = wfa_imin(v, h);
    sink.op('M', n);
    v -= n; h -= n;
  }
  sink.op('D', v);
  sink.op('I', h);
}

template <class Sink>
TRGT_HD void wfa_backtrace(const WfaProb &pr, int s_end, int k_end, int off_end, int *ws, Sink &sink) {
  const WfaWsHist hist{ws};
  wfa_backtrace_h(pr, s_end, k_end, off_end, hist, sink);
}

// End-to-end alignment of a short pair in one pass: with cost cap S every cell the full computation
// can reach lies on |k| <= R = (S - o) / e (a path's drift is its total gap length), so the band
// [-R, R] loses nothing at all -- not even non-optimal cells -- and its history can be back-traced
// directly.  Needs 2R+1 <= group size.  Returns status MAX_STEPS if the cost exceeds S.
template <class G>
TRGT_HD WfaEnd wfa_e2e_narrow(const G &g, const WfaProb &pr, int S, int *ws, size_t cap_ints) {
  WfaEnd out;
  out.status = TRGT_WFA_OOM; out.s = 0; out.k = 0; out.off = 0;
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  WfaProb bp = pr;
  bp.blo = wfa_imax(-pr.P, -R);
  bp.bhi = wfa_imin(pr.T, R);
  if (bp.bhi - bp.blo + 1 > g.size() || S >= 64 || (size_t)TRGT_WFA_META * ((size_t)S + 1) > cap_ints) return out;
  return wfa_forward_band_hist_narrow(g, bp, S, ws, cap_ints);
}

// count_matches (wfaligner.rs:988) and the text span of get_alignment_span (:864-908)
struct WfaFlankSink {
  int T;
  int matches, trailing_i, i_run, seen, ye;
  TRGT_HD explicit WfaFlankSink(int T_) : T(T_), matches(0), trailing_i(0), i_run(0), seen(0), ye(0) {}
  TRGT_HD void op(char c, int n) {
    if (n <= 0) return;
    if (c == 'M' || c == 'X') {
      if (c == 'M') matches += n;
      if (!seen) { seen = 1; ye = T - trailing_i; }
      i_run = 0;
    } else if (c == 'I') {
      if (!seen) trailing_i += n; else i_run += n;
    }
  }
  TRGT_HD int ystart() const { return seen ? i_run : 0; }
  TRGT_HD int yend() const { return seen ? ye : 0; }
};

// run-length SAM words (len << 4 | op) with '=' 7, 'X' 8, 'I' 1, 'D' 2 (get_sam_cigar(true),
// wfaligner.rs:932-959); collected last-to-first, then flipped by finish()
struct WfaCigarSink {
  uint32_t *out;
  uint32_t cap, n;
  uint32_t cur_code, cur_len;
  int overflow;
  TRGT_HD WfaCigarSink(uint32_t *o, uint32_t c) : out(o), cap(c), n(0), cur_code(0xF), cur_len(0), overflow(0) {}
  TRGT_HD void flush() {
    if (cur_len == 0) return;
    if (n < cap) out[n] = (cur_len << 4) | cur_code; else overflow = 1;
    n++;
    cur_len = 0;
  }
  TRGT_HD void op(char c, int cnt) {
    if (cnt <= 0) return;
    const uint32_t code = c == 'M' ? 7u : (c == 'X' ? 8u : (c == 'I' ? 1u : 2u));
    if (code != cur_code) { flush(); cur_code = code; }
    cur_len += (uint32_t)cnt;
  }
  TRGT_HD uint32_t finish() {
    flush();
    if (!overflow)
      for (uint32_t a = 0, b = n; a + 1 < b; a++) { b--; const uint32_t t = out[a]; out[a] = out[b]; out[b] = t; }
    return n;
  }
};

// ---------------------------------------------------------------- exact flank scan ----------

// first start s with t[s..s+P) == piece, or -1     span_locater.rs:10-12
// Each lane tests 4 consecutive starts per step against the piece's first 8 bytes (two 8-byte
// loads, three shifted windows); a key hit is verified 8 bytes at a time.
template <class G>
TRGT_HD int flank_scan(const G &g, const uint8_t *piece, int P, const uint8_t *t, int T) {
  const int n_starts = T - P + 1;
  if (P < 8) {
    for (int base = 0; base < n_starts; base += g.size()) {
      const int s = base + g.lane();
      int hit = INT_MAX;
      if (s < n_starts) {
        int j = 0;
        while (j < P && t[s + j] == piece[j]) j++;
        if (j == P) hit = s;
      }
      hit = g.min_i(hit);
      if (hit != INT_MAX) return hit;
    }
    return -1;
  }
  const uint64_t key = wfa_ld64u(piece);
  for (int base = 0; base < n_starts; base += 4 * g.size()) {
    const int s0 = base + 4 * g.lane();
    unsigned cand = 0;  // bit a: start s0 + a passes the 8-byte key test
    if (s0 < n_starts) {
      const uint64_t w0 = wfa_ld64u(t + s0), w1 = wfa_ld64u(t + s0 + 8);
      for (int a = 0; a < 4; a++) {
        const uint64_t win = a ? ((w0 >> (8 * a)) | (w1 << (64 - 8 * a))) : w0;
        if (win == key && s0 + a < n_starts) cand |= 1u << a;
      }
    }
    // candidates in increasing order, each verified by the whole group (8 bytes per lane per step)
    for (;;) {
      int first = INT_MAX;
      for (int a = 3; a >= 0; a--) if (cand & (1u << a)) first = s0 + a;
      const int c = g.min_i(first);
      if (c == INT_MAX) break;
      if (wfa_coop_match_len(g, piece + 8, t + c + 8, P - 8) == P - 8) return c;
      if (first == c) cand &= ~(1u << (c - s0));
    }
  }
  return -1;
}

// ---------------------------------------------------------------- seed filter ----------------

// Number of disjoint pattern blocks the seed filter cuts the piece into for cost cap S: one more than the
// blocks an alignment of cost <= S can damage, so that one block survives verbatim.  A mismatch damages one
// block for x, a one-base gap one block for o+e; a longer gap can straddle block boundaries: g deleted
// bases touch at most 1 + ceil((g-1)/blen) blocks, and blen >= TRGT_SEED_BLEN_MIN is enforced by every
// caller, so j+1 blocks cost at least o + (TRGT_SEED_BLEN_MIN*(j-1) + 2) e.  The bound is the best ratio
// blocks/cost over the events that fit the cap at all.  (For the presets 2,5,1 and 1,0,1 this equals
// S / min(x, o+e) + 1; scorings with a cheap gap open, e.g. 3,3,1, need more blocks.)
#define TRGT_SEED_BLEN_MIN 12
TRGT_HD int flank_seed_blocks(const WfaProb &pr, int S) {
  const int o = pr.oe - pr.e;
  int d = S / wfa_imin(pr.x, pr.oe);
  for (int j = 1; j < 16; j++) {
    const int cost = o + (TRGT_SEED_BLEN_MIN * (j - 1) + 2) * pr.e;
    if (cost > S) break;
    d = wfa_imax(d, S * (j + 1) / cost);
  }
  return d + 1;
}

// Diagonal band that provably contains every alignment of the whole pattern with cost <= S.
// Such an alignment has at most S / min(x, o+e) mismatches and gaps, so of nb = that + 1
// disjoint pattern blocks one is copied without any edit and shows up as an exact occurrence in the
// text; the total gap length is at most (S - o) / e, which bounds how far the path strays from that
// occurrence's diagonal.  Band = [min - R, max + R] over all exact block occurrences.
// keys: scratch for 32 uint64 owned by the group.  Returns false if no band can be given (no
// occurrence, blocks too short to be selective).
template <class G>
TRGT_HD bool flank_seed_band(const G &g, const WfaProb &pr, int S, uint64_t *keys, int *klo, int *khi) {
  const int nb = flank_seed_blocks(pr, S);
  if (nb > 32) return false;
  const int blen = pr.P / nb;
  if (blen < TRGT_SEED_BLEN_MIN || pr.T < blen) return false;
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  for (int b = g.lane(); b < nb; b += g.size()) keys[b] = wfa_ld64u(pr.p + b * blen);
  g.sync();
  int kmin = INT_MAX, kmax = INT_MIN;
  const int n_pos = pr.T - blen + 1;
  for (int jb = 0; jb < n_pos; jb += g.size()) {
    const int j = jb + g.lane();
    unsigned cand = 0;  // bit b: block b passes the 8-byte key test at text position j
    if (j < n_pos) {
      const uint64_t tw = wfa_ld64u(pr.t + j);
      for (int b = 0; b < nb; b++) if (tw == keys[b]) cand |= 1u << b;
    }
    // key hits are rare: verify them one at a time with the whole group
    for (;;) {
      const int leader = g.min_i(cand ? g.lane() : INT_MAX);
      if (leader == INT_MAX) break;
      const unsigned lc = (unsigned)g.bcast((int)cand, leader);
      int b = 0;
      while (!(lc & (1u << b))) b++;
      const int lj = jb + leader;
      if (wfa_coop_match_len(g, pr.p + b * blen + 8, pr.t + lj + 8, blen - 8) == blen - 8) {
        const int k = lj - b * blen;
        kmin = wfa_imin(kmin, k);
        kmax = wfa_imax(kmax, k);
      }
      if (g.lane() == leader) cand &= ~(1u << b);
    }
  }
  g.sync();
  if (kmin == INT_MAX) return false;
  *klo = wfa_imax(-pr.P, kmin - R);
  *khi = wfa_imin(pr.T, kmax + R);
  // Offsets on band diagonals must stay clear of the text end: there the full computation drops
  // cells whose insertion candidate runs past T, which a banded run could not reproduce.
  if (*khi + pr.P >= pr.T) return false;
  return true;
}

// ---------------------------------------------------------------- 8-mer index of a flank piece ---

// Table of every 8-mer of one piece (key = its 8 bytes, value = its offset), built once per locus and
// shared by all of the locus' reads.  With it, both the exact search and the seed filter only look at a few
// *probe* positions of the read instead of every position: an occurrence of a length-n pattern piece covers
// exactly one of the text positions (i+1)(n-7)-1, and the 8-mer found there must be one of the piece's own.
//
// Layout (TRGT_KIDX_SLOTS 16-bit words = 1344 bytes per piece): 256 buckets by 8 bits of the mixed key,
// bound[b] .. bound[b+1] delimiting bucket b's entries (bound[0] = 0), then the entries themselves,
// entry = 7-bit fingerprint << 9 | offset in the piece.  A fingerprint clash only costs a failed
// verification: every candidate is compared byte for byte before it counts.  Built by counting sort --
// count per bucket, exclusive scan, scatter -- three converged passes without a compare-and-swap loop.
#define TRGT_KIDX_BUCKETS 256u
#define TRGT_KIDX_ENT0 264u    // first entry: after bound[257], rounded to 16 bytes
#define TRGT_KIDX_SLOTS 672u   // 264 + room for 400 entries, rounded to 16 bytes
#define TRGT_KIDX_MAX_P 400    // offsets fit 9 bits
#define TRGT_CAND_CAP 64

struct KmerIndex {
  uint16_t *slot;  // [TRGT_KIDX_SLOTS]
};

TRGT_HD uint64_t kidx_mix(uint64_t k) { return k * 0x9E3779B97F4A7C15ull; }
TRGT_HD uint32_t kidx_home(uint64_t mixed) { return (uint32_t)(mixed >> 56); }
TRGT_HD uint32_t kidx_fp(uint64_t mixed) { return (uint32_t)(mixed >> 40) & 0x7Fu; }

// every entry `v` of the bucket of `mixed`
#define TRGT_KIDX_FOR(idx, mixed, v)                                                                            \
  for (uint32_t ki_ = (idx).slot[kidx_home(mixed)], ke_ = (idx).slot[kidx_home(mixed) + 1u];                   \
       ki_ < ke_ && (((v) = (idx).slot[TRGT_KIDX_ENT0 + ki_]), true); ki_++)

// slot[i] += 1, returning the old value (the 16-bit word shares its 32-bit word with a neighbour; counts stay far
// below 65536, so the add never carries across)
TRGT_HD uint32_t kidx_fetch_inc(uint16_t *slot, uint32_t i) {
#if defined(__CUDA_ARCH__)
  const unsigned sh = (i & 1u) * 16u;
  const unsigned int old = atomicAdd((unsigned int *)slot + (i >> 1), 1u << sh);
  return (old >> sh) & 0xFFFFu;
#else
  return __atomic_fetch_add(&slot[i], (uint16_t)1, __ATOMIC_ACQ_REL);
#endif
}

template <class G>
TRGT_HD void kidx_build(const G &g, const KmerIndex &idx, const uint8_t *piece, int P) {
  uint16_t *bound = idx.slot;  // during the build bound[1 + b] = count, then first free position of bucket b
  for (uint32_t i = (uint32_t)g.lane(); i < TRGT_KIDX_ENT0 / 4; i += (uint32_t)g.size())  // tables are 8-byte aligned
    ((uint64_t *)idx.slot)[i] = 0ull;
  g.sync();
  for (int i = g.lane(); i + 8 <= P; i += g.size()) kidx_fetch_inc(bound, 1u + kidx_home(kidx_mix(wfa_ld64u(piece + i))));
  g.sync();
  {  // exclusive scan of the 256 counts: a run of consecutive buckets per lane, then across the lanes
    const uint32_t per = (TRGT_KIDX_BUCKETS + (uint32_t)g.size() - 1u) / (uint32_t)g.size();
    const uint32_t b0 = (uint32_t)g.lane() * per;
    int sum = 0;
    for (uint32_t b = b0; b < b0 + per && b < TRGT_KIDX_BUCKETS; b++) sum += bound[1u + b];
    int total;
    int run = g.excl_scan_i(sum, &total);
    for (uint32_t b = b0; b < b0 + per && b < TRGT_KIDX_BUCKETS; b++) {
      const int c = bound[1u + b];
      bound[1u + b] = (uint16_t)run;
      run += c;
    }
  }
  g.sync();
  for (int i = g.lane(); i + 8 <= P; i += g.size()) {  // scatter; bound[1 + b] ends up at the end of bucket b
    const uint64_t mixed = kidx_mix(wfa_ld64u(piece + i));
    const uint32_t pos = kidx_fetch_inc(bound, 1u + kidx_home(mixed));
    idx.slot[TRGT_KIDX_ENT0 + pos] = (uint16_t)((kidx_fp(mixed) << 9) | (uint32_t)i);
  }
  g.sync();
}

// first start s with t[s..s+P) == piece via the index; -1 none, -2 too many candidates (use flank_scan)
// cand: TRGT_CAND_CAP + 1 ints of group scratch
template <class G>
TRGT_HD int flank_scan_indexed(const G &g, const KmerIndex &idx, const uint8_t *piece, int P, const uint8_t *t, int T,
                               int *cand) {
  const int n_starts = T - P + 1;
  if (n_starts <= 0) return -1;
  const int step = P - 7;
  const int n_probes = (T - 7) / step;  // probes j_i = (i+1)*step - 1 with j_i + 8 <= T
  for (int pb = 0; pb < n_probes; pb += g.size()) {
    const int i = pb + g.lane();
    const int j = (i + 1) * step - 1;
    uint32_t fp = 0;
    uint64_t mixed = 0;
    int cnt = 0, c0 = 0;  // candidates of this lane's probe; c0 = the first one
    if (i < n_probes) {
      mixed = kidx_mix(wfa_ld64u(t + j));
      fp = kidx_fp(mixed);
      uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
        const int s = j - (int)(v & 511u);
        if ((v >> 9) == fp && s >= 0 && s < n_starts) { if (cnt == 0) c0 = s; cnt++; }
      }
    }
    for (;;) {  // probes in increasing order: their candidate ranges are disjoint and increasing
      const int leader = g.min_i(cnt > 0 ? g.lane() : INT_MAX);
      if (leader == INT_MAX) break;
      const int ln = g.bcast(cnt, leader);
      int best = INT_MAX;
      if (ln == 1) {  // the usual case: one candidate, straight from the leader's register
        const int s = g.bcast(c0, leader);
        if (wfa_coop_match_len(g, piece, t + s, P) == P) best = s;
      } else {
        if (g.lane() == leader) {
          int n = 0;
          uint32_t v;
          TRGT_KIDX_FOR(idx, mixed, v) {
            const int s = j - (int)(v & 511u);
            if ((v >> 9) == fp && s >= 0 && s < n_starts) { if (n < TRGT_CAND_CAP) cand[n] = s; n++; }
          }
          cand[TRGT_CAND_CAP] = n;
        }
        g.sync();
        const int n = cand[TRGT_CAND_CAP];
        if (n <= TRGT_CAND_CAP)
          for (int c = 0; c < n; c++) {
            const int s = cand[c];
            if (s < best && wfa_coop_match_len(g, piece, t + s, P) == P) best = s;
          }
        g.sync();
        if (n > TRGT_CAND_CAP) return -2;
      }
      if (g.lane() == leader) cnt = 0;
      if (best != INT_MAX) return best;
    }
  }
  return -1;
}

// ---------------------------------------------------------------- exact search, one lane per pair ---

// The exact search again, but a lane per (read, flank) pair instead of a group per pair: the probes of
// flank_scan_indexed walked by one lane, and a candidate start verified by that lane alone against
// the text where it lies (16 aligned bytes per load).  So that no byte ever has to be realigned, the
// locus keeps FXT_COPIES copies of its piece on chip, copy k shifted by k bytes: byte 8 + k + i of
// copy k is piece[i].  A candidate at text address A compares aligned text words with aligned words
// of copy A & 7; the bytes outside [A, A + P) of the first and last word are masked.
#define FXT_COPIES 8
#define FXT_PMAX 256    // pieces this path takes
#define FXT_STRIDE 288  // bytes per copy: 8 of front slack + 7 of shift + FXT_PMAX + 15 of tail, rounded to 16

struct alignas(16) FxtU128 {
  uint64_t lo, hi;
};

// valid-byte masks of a little-endian 8-byte word: bytes >= lo / bytes < hi
TRGT_HD uint64_t fxt_mask_from(int lo) { return lo <= 0 ? ~0ull : (lo >= 8 ? 0ull : (~0ull << (8 * lo))); }
TRGT_HD uint64_t fxt_mask_to(int hi) { return hi >= 8 ? ~0ull : (hi <= 0 ? 0ull : ((1ull << (8 * hi)) - 1ull)); }

// the FXT_COPIES shifted copies of one piece (16 <= P <= FXT_PMAX), lanes stride over the words
template <class G>
TRGT_HD void fxt_build_copies(const G &g, const uint8_t *piece, int P, uint8_t *copies) {
  const int W = FXT_STRIDE / 8;
  // copy 0 (the piece at byte 8, zeros around it) from the piece, then copy k = copy 0 moved up k bytes: two
  // words of copy 0 give a word of each of the other copies
  uint64_t *c0 = (uint64_t *)copies;
  for (int wd = g.lane(); wd < W; wd += g.size()) {
    const int i0 = 8 * wd - 8;
    uint64_t v = 0;
    if (i0 >= 0 && i0 < P) {
      const int hi = i0 + 8 > P ? i0 + 8 - P : 0;
      v = wfa_ld64u(piece + (hi ? P - 8 : i0));
      if (hi) v >>= 8 * hi;
    }
    c0[wd] = v;
  }
  g.sync();
  for (int wd = g.lane(); wd < W; wd += g.size()) {
    const uint64_t a = c0[wd], b = wd ? c0[wd - 1] : 0ull;
#pragma unroll
    for (int k = 1; k < FXT_COPIES; k++)
      *(uint64_t *)(copies + k * FXT_STRIDE + 8 * wd) = (a << (8 * k)) | (b >> (64 - 8 * k));
  }
  g.sync();
}

// t[s .. s+P) == piece ?   a = t + s; reads whole 16-byte words around [a, a+P)
TRGT_HD bool fxt_verify(const uint8_t *copies, int P, const uint8_t *a) {
  const uintptr_t A = (uintptr_t)a;
  const int a16 = (int)(A & 15u), k = a16 & 7, h = a16 >> 3;
  const FxtU128 *tx = (const FxtU128 *)(A - (uintptr_t)a16);
  const uint8_t *pc = copies + k * FXT_STRIDE + 8 - 8 * h;  // the copy's bytes that face tx[0]
  const int C = (a16 + P + 15) >> 4;                        // 16-byte words touched
  {  // first word: bytes below a16 do not belong to the candidate (nor, if C == 1, those from a16 + P on)
    const FxtU128 tv = tx[0];
    const uint64_t d0 = (tv.lo ^ *(const uint64_t *)pc) & fxt_mask_from(a16) & fxt_mask_to(a16 + P);
    const uint64_t d1 = (tv.hi ^ *(const uint64_t *)(pc + 8)) & fxt_mask_from(a16 - 8) & fxt_mask_to(a16 + P - 8);
    if (d0 | d1) return false;
  }
  if (C == 1) return true;
  int c = 1;
  for (; c + 4 <= C - 1; c += 4) {  // four words per step: the loads go out together
    const FxtU128 t0 = tx[c], t1 = tx[c + 1], t2 = tx[c + 2], t3 = tx[c + 3];
    const uint64_t *q = (const uint64_t *)(pc + 16 * c);
    const uint64_t d = (t0.lo ^ q[0]) | (t0.hi ^ q[1]) | (t1.lo ^ q[2]) | (t1.hi ^ q[3]) | (t2.lo ^ q[4]) |
                       (t2.hi ^ q[5]) | (t3.lo ^ q[6]) | (t3.hi ^ q[7]);
    if (d) return false;
  }
  for (; c < C - 1; c++) {
    const FxtU128 tv = tx[c];
    const uint64_t *q = (const uint64_t *)(pc + 16 * c);
    if ((tv.lo ^ q[0]) | (tv.hi ^ q[1])) return false;
  }
  {  // last word
    const int end = a16 + P - 16 * (C - 1);  // valid bytes of it, 1..16
    const FxtU128 tv = tx[C - 1];
    const uint64_t *q = (const uint64_t *)(pc + 16 * (C - 1));
    const uint64_t d0 = (tv.lo ^ q[0]) & fxt_mask_to(end);
    const uint64_t d1 = (tv.hi ^ q[1]) & fxt_mask_to(end - 8);
    if (d0 | d1) return false;
  }
  return true;
}

// first start s with t[s..s+P) == piece, or -1 (span_locater.rs:10-12), by ONE lane.
// Probes in increasing order: their candidate ranges are disjoint and increasing, so the smallest
// verified candidate of the first probe that has one is the first occurrence.
// Lanes that run a per-lane routine side by side drift apart in its data-dependent loops and are not
// brought back together until the routine returns.  `lanes` names the lanes of the warp that entered the
// routine together; at a few points they all pass they wait for each other, so that the expensive stretches
// (a verification's sixteen-byte compares) run with the whole warp.  0 (and the CPU test build): no-op.
#if defined(__CUDA_ARCH__)
#define TRGT_CONVERGE(lanes) do { if (lanes) __syncwarp(lanes); } while (0)
#define TRGT_ANY(lanes, p) ((lanes) ? __any_sync((lanes), (p)) : (p))
#else
#define TRGT_CONVERGE(lanes) do { (void)(lanes); } while (0)
#define TRGT_ANY(lanes, p) ((void)(lanes), (p))
#endif

// ask L2 for the sectors a verification of t[s .. s+P) will read (no registers held; the verification's
// dependent rounds of loads then wait for L2 instead of DRAM)
#ifndef TRGT_FXT_PREFETCH
#define TRGT_FXT_PREFETCH 128  // bytes between prefetches (measured: 128 = 64 = 32 > none), 0 = off
#endif
TRGT_HD void fxt_prefetch(const uint8_t *a, int P) {
#if defined(__CUDA_ARCH__)
  if (TRGT_FXT_PREFETCH > 0) {
    const uintptr_t b = (uintptr_t)a & ~(uintptr_t)31, e = (uintptr_t)a + (uintptr_t)P;
    for (uintptr_t q = b; q < e; q += (TRGT_FXT_PREFETCH > 0 ? TRGT_FXT_PREFETCH : 32))
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q));
  }
#else
  (void)a; (void)P;
#endif
}

TRGT_HD int flank_exact_thread(const KmerIndex &idx, const uint8_t *copies, int P, const uint8_t *t, int T,
                               unsigned lanes = 0) {
  const int n_starts = T - P + 1;
  const int step = P - 7;
  const int n_probes = n_starts > 0 ? (T - 7) / step : 0;
  // Two loops, so that the lanes of a warp (each on its own pair) stay together: the probes' candidates
  // are collected first (in increasing order of probe, hence of candidate range), then verified.
  int c0 = -1, c1 = -1, c2 = -1, c3 = -1;  // candidate starts
  int n = 0;
  bool many = false;
  for (int ib = 0; ib < n_probes && !many; ib += 4) {  // four probes per step: their loads go out together
    uint64_t key[4];
#pragma unroll
    for (int u = 0; u < 4; u++) key[u] = ib + u < n_probes ? wfa_ld64u(t + (ib + u + 1) * step - 1) : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (ib + u >= n_probes) break;
      const int j = (ib + u + 1) * step - 1;
      const uint64_t mixed = kidx_mix(key[u]);
      const uint32_t fp = kidx_fp(mixed);
      uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
        if ((v >> 9) != fp) continue;
        const int s = j - (int)(v & 511u);
        if (s < 0 || s >= n_starts) continue;
        if (n == 0) c0 = s; else if (n == 1) c1 = s; else if (n == 2) c2 = s; else if (n == 3) c3 = s; else many = true;
        if (n == 0) fxt_prefetch(t + s, P);  // the first candidate is the answer for ~4 pairs of 5
        n++;
      }
    }
  }
  // candidates of probe i lie in [i * step, (i + 1) * step): the first occurrence is the smallest verified
  // candidate within the first probe range that has one.  Four rounds for everybody, one candidate per round.
  int best = INT_MAX, best_range = INT_MAX;
#pragma unroll 1
  for (int c = 0; c < 4; c++) {
    TRGT_CONVERGE(lanes);
    if (many || c >= n) continue;
    const int s = c == 0 ? c0 : c == 1 ? c1 : c == 2 ? c2 : c3;
    const int range = s / step;  // index of the probe whose range holds s
    if (range > best_range || s >= best) continue;
    if (fxt_verify(copies, P, t + s)) { best = s; best_range = range; }
  }
  TRGT_CONVERGE(lanes);
  if (!many) return best == INT_MAX ? -1 : best;
  // many candidates (repetitive piece): probe by probe
  for (int i = 0; i < n_probes; i++) {
    const int j = (i + 1) * step - 1;
    const uint64_t mixed = kidx_mix(wfa_ld64u(t + j));
    const uint32_t fp = kidx_fp(mixed);
    int pbest = INT_MAX;
    uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
      if ((v >> 9) != fp) continue;
      const int s = j - (int)(v & 511u);
      if (s < 0 || s >= n_starts || s >= pbest) continue;
      if (fxt_verify(copies, P, t + s)) pbest = s;
    }
    if (pbest != INT_MAX) return pbest;
  }
  return -1;
}

// flank_seed_band through the index: 1 band found, 0 no band can be given, -2 too many candidates
template <class G>
TRGT_HD int flank_seed_band_indexed(const G &g, const KmerIndex &idx, const WfaProb &pr, int S, int *cand, int *klo,
                                    int *khi) {
  const int nb = flank_seed_blocks(pr, S);
  if (nb > 32) return 0;
  const int blen = pr.P / nb;
  if (blen < TRGT_SEED_BLEN_MIN || pr.T < blen) return 0;
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  const int step = blen - 7;
  const int n_probes = (pr.T - 7) / step;
  int kmin = INT_MAX, kmax = INT_MIN;
  for (int pb = 0; pb < n_probes; pb += g.size()) {
    const int i = pb + g.lane();
    const int j = (i + 1) * step - 1;
    uint32_t fp = 0;
    uint64_t mixed = 0;
    int cnt = 0, c0 = 0;  // candidate = block start in the text << 5 | block
    if (i < n_probes) {
      mixed = kidx_mix(wfa_ld64u(pr.t + j));
      fp = kidx_fp(mixed);
      uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
        if ((v >> 9) != fp) continue;
        const int d = (int)(v & 511u), b = d / blen, q = j - (d - b * blen);
        if (b < nb && d - b * blen + 8 <= blen && q >= 0 && q + blen <= pr.T) { if (cnt == 0) c0 = (q << 5) | b; cnt++; }
      }
    }
    for (;;) {
      const int leader = g.min_i(cnt > 0 ? g.lane() : INT_MAX);
      if (leader == INT_MAX) break;
      const int ln = g.bcast(cnt, leader);
      if (ln == 1) {
        const int c = g.bcast(c0, leader);
        const int q = c >> 5, b = c & 31;
        if (wfa_coop_match_len(g, pr.p + b * blen, pr.t + q, blen) == blen) {
          kmin = wfa_imin(kmin, q - b * blen);
          kmax = wfa_imax(kmax, q - b * blen);
        }
      } else {
        if (cand == nullptr) return -2;  // no list scratch (one-lane callers): leave it to the cooperative path
        if (g.lane() == leader) {
          int n = 0;
          uint32_t v;
          TRGT_KIDX_FOR(idx, mixed, v) {
            if ((v >> 9) != fp) continue;
            const int d = (int)(v & 511u), b = d / blen, q = j - (d - b * blen);
            if (b < nb && d - b * blen + 8 <= blen && q >= 0 && q + blen <= pr.T) {
              if (n < TRGT_CAND_CAP) cand[n] = (q << 5) | b;
              n++;
            }
          }
          cand[TRGT_CAND_CAP] = n;
        }
        g.sync();
        const int n = cand[TRGT_CAND_CAP];
        if (n <= TRGT_CAND_CAP)
          for (int c = 0; c < n; c++) {
            const int q = cand[c] >> 5, b = cand[c] & 31;
            if (wfa_coop_match_len(g, pr.p + b * blen, pr.t + q, blen) == blen) {
              kmin = wfa_imin(kmin, q - b * blen);
              kmax = wfa_imax(kmax, q - b * blen);
            }
          }
        g.sync();
        if (n > TRGT_CAND_CAP) return -2;
      }
      if (g.lane() == leader) cnt = 0;
    }
  }
  if (kmin == INT_MAX) return 0;
  *klo = wfa_imax(-pr.P, kmin - R);
  *khi = wfa_imin(pr.T, kmax + R);
  if (*khi + pr.P >= pr.T) return 0;  // see flank_seed_band
  return 1;
}

struct FlankHit {
  int via;      // TRGT_VIA_* of include/trgt_engine.h: 2 accepted, 3 rejected
  int matches;  // count_matches
  int score;    // -cost
  int start, end;
};

// WFA fallback of find_spans (span_locater.rs:14-25) for one (piece, read) pair without ever
// building the T+1 wide wavefront: seed filter -> banded forward pass with history -> back-trace.
// Two cost tiers: first the cost of a single mismatch or one-base gap (most HiFi misses), then the
// caller's budget S.  `pr` is the unbanded flank problem (pbf = pef = 0, tbf = tef = T).
// ws: ws_ints ints of group scratch (on chip), keys: 32 uint64.  Returns 0 and fills *hit (lane 0's
// copy is authoritative), or 1 if this pair needs the full-width path (no seed, cost > S, scratch
// too small).
// idx (optional): 8-mer index of pr.p with cand = TRGT_CAND_CAP + 1 ints of scratch; without it, or
// when it overflows, the seed filter scans every text position.
template <class G>
TRGT_HD int flank_locate_banded(const G &g, const WfaProb &pr, int S, double min_flank_id_frac, uint64_t *keys,
                                int *ws, size_t ws_ints, FlankHit *hit, const KmerIndex *idx = nullptr,
                                int *cand = nullptr) {
  const int tier1 = wfa_imin(S, wfa_imax(pr.x, pr.oe));
  for (int tier = 0; tier < 2; tier++) {
    const int cap = tier == 0 ? tier1 : S;
    if (tier == 1 && S <= tier1) break;
    int klo, khi;
    int have = -2;
    if (idx) have = flank_seed_band_indexed(g, *idx, pr, cap, cand, &klo, &khi);
    if (have == -2) have = flank_seed_band(g, pr, cap, keys, &klo, &khi) ? 1 : 0;
    if (!have) continue;
    WfaProb bp = pr;
    bp.blo = klo; bp.bhi = khi;
    WfaEnd end = wfa_forward_band_hist(g, bp, cap, ws, ws_ints);
    g.sync();
    if (end.status == TRGT_WFA_OOM && wfa_ring_ints(bp) <= ws_ints) {
      // the in-band history does not fit: score-only ring pass in the band, then the (smaller) cone
      end = wfa_score_ring(g, bp, ws, cap);
      g.sync();
      if (end.status == TRGT_WFA_OK) {
        if (wfa_trace_ints(pr, end.s) > ws_ints || wfa_trace_forward(g, pr, end.s, end.k, ws, ws_ints) != 0)
          end.status = TRGT_WFA_OOM;
      }
    }
    if (end.status != TRGT_WFA_OK) continue;
    if (g.lane() == 0) {
      WfaFlankSink sink(pr.T);
      wfa_backtrace(pr, end.s, end.k, end.off, ws, sink);
      hit->matches = sink.matches;
      hit->score = -end.s;
      if ((double)sink.matches >= (double)pr.P * min_flank_id_frac) {  // span_locater.rs:19-25, :46
        hit->via = 2; hit->start = sink.ystart(); hit->end = sink.yend();
      } else {
        hit->via = 3; hit->start = 0; hit->end = 0;
      }
    }
    g.sync();
    return 0;
  }
  return 1;
}

// Lean variant for k_flank_band: index-only seed filter and the narrow-band forward pass only; any
// pair that would need the linear seed scan, a band wider than one lane per diagonal, or more
// scratch is handed to the full-width path instead (returns 1).  Keeps the kernel's code small
// enough to stay resident in the instruction cache.
template <class G>
TRGT_HD int flank_locate_banded_lean(const G &g, const WfaProb &pr, int S, double min_flank_id_frac, int *ws,
                                     size_t ws_ints, FlankHit *hit, const KmerIndex &idx, int *cand,
                                     int first_tier = 0, int last_tier = 1) {
  const int tier1 = wfa_imin(S, wfa_imax(pr.x, pr.oe));
#pragma unroll 1
  for (int tier = first_tier; tier <= last_tier; tier++) {
    const int cap = tier == 0 ? tier1 : S;
    if (tier == 1 && S <= tier1) break;
    int klo, khi;
    if (flank_seed_band_indexed(g, idx, pr, cap, cand, &klo, &khi) != 1) continue;
    WfaProb bp = pr;
    bp.blo = klo; bp.bhi = khi;
    if (khi - klo + 1 > g.size() || cap >= 64 ||
        (size_t)TRGT_WFA_META * ((size_t)cap + 1) > ws_ints)
      continue;
    const WfaEnd end = wfa_forward_band_hist_narrow(g, bp, cap, ws, ws_ints);
    g.sync();
    if (end.status != TRGT_WFA_OK) continue;
    if (g.lane() == 0) {
      WfaFlankSink sink(pr.T);
      wfa_backtrace(pr, end.s, end.k, end.off, ws, sink);
      hit->matches = sink.matches;
      hit->score = -end.s;
      if ((double)sink.matches >= (double)pr.P * min_flank_id_frac) {  // span_locater.rs:19-25, :46
        hit->via = 2; hit->start = sink.ystart(); hit->end = sink.yend();
      } else {
        hit->via = 3; hit->start = 0; hit->end = 0;
      }
    }
    g.sync();
    return 0;
  }
  return 1;
}

// ---------------------------------------------------------------- first cost tier, one lane per pair ---

// The first cost tier of the fallback (cost <= max(x, o+e): one mismatch or one 1-bp gap, ~89 % of HiFi
// misses) by ONE lane: such a pair lives on 3-4 diagonals, which is no work for a group, and one lane per
// pair needs no collectives at all.  Same three steps as flank_locate_banded_lean -- index seed filter,
// narrow-band forward pass with history, back-trace -- with the history (<= FT1_WS_INTS ints, same layout
// as wfa_forward_band_hist_narrow so that wfa_backtrace runs on it unchanged) in the lane's own scratch.
#define FT1_WMAX 4      // widest band taken
#define FT1_SMAX 8      // highest cost cap taken
#define FT1_WS_INTS (TRGT_WFA_META * (FT1_SMAX + 1) + 3 * FT1_WMAX * (FT1_SMAX + 1))

// flank_seed_band_indexed by one lane.  Two loops, so that the lanes of a warp (each on its own pair)
// stay together: first every probe's candidates are collected, then they are verified.
#define FT1_CANDS 12  // candidates a lane keeps; more (repetitive pieces) hands the pair on
TRGT_HD int flank_seed_band_thread(const KmerIndex &idx, const WfaProb &pr, int S, int *klo, int *khi) {
  const int nb = flank_seed_blocks(pr, S);
  if (nb > 32) return 0;
  const int blen = pr.P / nb;
  if (blen < TRGT_SEED_BLEN_MIN || pr.T < blen) return 0;
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  const int step = blen - 7;
  const int n_probes = (pr.T - 7) / step;
  int cand[FT1_CANDS];  // block start in the text << 5 | block
  int n = 0;
  for (int ib = 0; ib < n_probes; ib += 4) {  // four probes per step: their loads go out together
    uint64_t key[4];
#pragma unroll
    for (int u = 0; u < 4; u++) key[u] = ib + u < n_probes ? wfa_ld64u(pr.t + (ib + u + 1) * step - 1) : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (ib + u >= n_probes) break;
      const int j = (ib + u + 1) * step - 1;
      const uint64_t mixed = kidx_mix(key[u]);
      const uint32_t fp = kidx_fp(mixed);
      uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
        if ((v >> 9) != fp) continue;
        const int d = (int)(v & 511u), b = d / blen, q = j - (d - b * blen);
        if (b >= nb || d - b * blen + 8 > blen || q < 0 || q + blen > pr.T) continue;
        if (n == FT1_CANDS) return -2;
        cand[n++] = (q << 5) | b;
      }
    }
  }
  int kmin = INT_MAX, kmax = INT_MIN;
  for (int c = 0; c < n; c++) {
    const int q = cand[c] >> 5, b = cand[c] & 31, k = q - b * blen;
    if (k >= kmin && k <= kmax) continue;  // would not move the band
    if (wfa_match_len(pr.p + b * blen, pr.t + q, blen) == blen) {
      kmin = wfa_imin(kmin, k);
      kmax = wfa_imax(kmax, k);
    }
  }
  if (kmin == INT_MAX) return 0;
  *klo = wfa_imax(-pr.P, kmin - R);
  *khi = wfa_imin(pr.T, kmax + R);
  if (*khi + pr.P >= pr.T) return 0;  // see flank_seed_band
  return 1;
}

// wfa_forward_band_hist_narrow by one lane (it walks the band's diagonals itself)
TRGT_HD WfaEnd wfa_forward_band_hist_narrow_thread(const WfaProb &pr, int s_cap, int *ws, size_t cap_ints) {
  WfaEnd out;
  out.status = TRGT_WFA_OK; out.s = 0; out.k = 0; out.off = 0;
  const int W = pr.bhi - pr.blo + 1;
  const size_t hdr = (size_t)TRGT_WFA_META * ((size_t)s_cap + 1);
  unsigned long long present = 0;
  for (int s = 0; s <= s_cap; s++) {
    const int sx = s - pr.x, so = s - pr.oe, se = s - pr.e;
    const bool px = sx >= 0 && ((present >> sx) & 1ull), po = so >= 0 && ((present >> so) & 1ull);
    const bool pe = se >= 1 && ((present >> se) & 1ull);
    const bool live = s == 0 || px || po || pe;
    const size_t base = hdr + (size_t)3 * W * s;
    if (live && base + (size_t)3 * W > cap_ints) { out.status = TRGT_WFA_OOM; return out; }
    int *meta = ws + (size_t)TRGT_WFA_META * s;
    meta[0] = meta[2] = live ? pr.blo : 1;
    meta[1] = meta[3] = live ? pr.bhi : 0;
    meta[4] = (int)(unsigned)(base & 0xffffffffu);
    meta[5] = (int)(unsigned)(base >> 32);
    if (!live) continue;
    const int *mo = ws + hdr + (size_t)3 * W * (so > 0 ? so : 0);
    const int *me = ws + hdr + (size_t)3 * W * (se > 0 ? se : 0);
    const int *mxs = ws + hdr + (size_t)3 * W * (sx > 0 ? sx : 0);
    int endk = INT_MAX;
    for (int idx = 0; idx < W; idx++) {
      const int k = pr.blo + idx;
      int mx = TRGT_WFA_NULL;
      if (s == 0) {
        if (k >= -pr.pbf && k <= pr.tbf) mx = k >= 0 ? k : 0;
      } else {
        const int o_l = (po && idx > 0) ? mo[idx - 1] : TRGT_WFA_NULL;
        const int o_r = (po && idx + 1 < W) ? mo[idx + 1] : TRGT_WFA_NULL;
        const int i_l = (pe && idx > 0) ? me[W + idx - 1] : TRGT_WFA_NULL;
        const int d_r = (pe && idx + 1 < W) ? me[2 * W + idx + 1] : TRGT_WFA_NULL;
        const int i1 = wfa_imax(o_l, i_l) + 1;
        const int d1 = wfa_imax(o_r, d_r);
        const int mm = (px ? mxs[idx] : TRGT_WFA_NULL) + 1;
        mx = wfa_imax(mm, wfa_imax(i1, d1));
        const int h = mx, v = mx - k;
        if (mx < 0 || h > pr.T || v > pr.P || v < 0) mx = TRGT_WFA_NULL;
        ws[base + W + idx] = i1;
        ws[base + 2 * W + idx] = d1;
      }
      if (mx >= 0) {  // match extension along the diagonal
        const int v = mx - k, n = wfa_imin(pr.P - v, pr.T - mx);
        if (n > 0) mx += wfa_match_len(pr.p + v, pr.t + mx, n);
      }
      ws[base + idx] = mx;
      if (endk == INT_MAX && mx >= 0) {  // end condition, lowest diagonal first
        const int h = mx, v = mx - k;
        if (v >= 0 && v <= pr.P && h <= pr.T &&
            ((h >= pr.T && pr.P - v <= pr.pef) || (v >= pr.P && pr.T - h <= pr.tef))) endk = k;
      }
    }
    present |= 1ull << s;
    if (endk != INT_MAX) {
      out.s = s; out.k = endk; out.off = ws[base + (endk - pr.blo)];
      return out;
    }
  }
  out.status = TRGT_WFA_MAX_STEPS;
  return out;
}

// 0 and *hit filled, or 1 if the pair has to go on to the next tier.  ws: FT1_WS_INTS ints of the lane.
TRGT_HD int flank_locate_tier1_thread(const WfaProb &pr, int S, double min_flank_id_frac, int *ws, FlankHit *hit,
                                      const KmerIndex &idx) {
  const int cap = wfa_imin(S, wfa_imax(pr.x, pr.oe));
  if (cap > FT1_SMAX) return 1;
  int klo, khi;
  if (flank_seed_band_thread(idx, pr, cap, &klo, &khi) != 1) return 1;
  if (khi - klo + 1 > FT1_WMAX) return 1;
  WfaProb bp = pr;
  bp.blo = klo; bp.bhi = khi;
  const WfaEnd end = wfa_forward_band_hist_narrow_thread(bp, cap, ws, FT1_WS_INTS);
  if (end.status != TRGT_WFA_OK) return 1;
  WfaFlankSink sink(pr.T);
  wfa_backtrace(pr, end.s, end.k, end.off, ws, sink);
  hit->matches = sink.matches;
  hit->score = -end.s;
  if ((double)sink.matches >= (double)pr.P * min_flank_id_frac) {  // span_locater.rs:19-25, :46
    hit->via = 2; hit->start = sink.ystart(); hit->end = sink.yend();
  } else {
    hit->via = 3; hit->start = 0; hit->end = 0;
  }
  return 0;
}

// ---------------------------------------------------------------- first cost tier in two passes ---
//
// flank_locate_tier1_thread above walks a chain of dependent loads per pair: probe -> table -> block
// verification (8-byte steps over the read in HBM) -> match extensions (the same) -> history in local memory.
// The two routines below do the same computation with two global round trips per pair:
//   seed pass   every probe is loaded up front; the band is the hull of the diagonals of ALL index hits
//               (checked on 16 bases only, not on the whole block).  Every alignment of cost <= S contains an intact block, that block's 8-mer
//               at its probe is in the table, so the hull of all hits contains the hull of the true block
//               occurrences: a superset band, and the banded computation equals the full one on any band
//               that contains every alignment of cost <= S (see flank_seed_band).  A chance hit elsewhere
//               only makes the hull too wide for this tier; those pairs take the verifying filter.
//   band pass   the text window the band can touch ([klo, khi + P], <= 304 bytes) is staged on chip in one
//               bulk copy, the history lives on chip as 16-bit offsets relative to the band's lowest
//               diagonal, and the back-trace reads it through the same routine as every other path.

TRGT_HD int ft1_ld(int16_t v, int blo) { return v < 0 ? TRGT_WFA_NULL : (int)v + blo; }

// wfa_ld64u with the address space spelled out: the piece through the read-only path, the window from shared memory
TRGT_HD uint64_t ft1_ld64_piece(const uint8_t *a) {
#if defined(__CUDA_ARCH__)
  const uintptr_t addr = (uintptr_t)a;
  const uint32_t *w = (const uint32_t *)(addr & ~(uintptr_t)3);
  const unsigned sh = (unsigned)(addr & 3u) * 8u;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
#else
  return wfa_ld64u(a);
#endif
}
TRGT_HD uint64_t ft1_ld64_win(const uint8_t *a) {
#if defined(__CUDA_ARCH__)
  const unsigned addr = (unsigned)__cvta_generic_to_shared(a);
  const unsigned base = addr & ~3u, sh = (addr & 3u) * 8u;
  uint32_t w0, w1, w2;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(base));
  asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(base));
  return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
#else
  return wfa_ld64u(a);
#endif
}
// wfa_match_len(piece, window, n)
TRGT_HD int ft1_match_len(const uint8_t *p, const uint8_t *w, int n) {
  int i = 0;
  while (i < n) {
    const uint64_t x = ft1_ld64_piece(p + i) ^ ft1_ld64_win(w + i);
    if (x) {
      i += wfa_ctz64(x) >> 3;
      return i < n ? i : n;
    }
    i += 8;
  }
  return n;
}

// floor(d / blen) for d < 512 by multiplication: inv = ceil(2^20 / blen)
TRGT_HD uint32_t ft1_div_magic(int blen) { return ((1u << 20) + (uint32_t)blen - 1u) / (uint32_t)blen; }

// 1 and [klo, khi] (hull of every index hit +- R), or 0 if there is no hit
TRGT_HD int flank_seed_hull_thread(const KmerIndex &idx, const WfaProb &pr, int S, int *klo, int *khi) {
  const int nb = flank_seed_blocks(pr, S);
  if (nb > 32) return 0;
  const int blen = pr.P / nb;
  if (blen < TRGT_SEED_BLEN_MIN || pr.T < blen) return 0;
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  const int step = blen - 7;
  const int n_probes = (pr.T - 7) / step;
  const uint32_t inv = ft1_div_magic(blen);
  int kmin = INT_MAX, kmax = INT_MIN;
  for (int ib = 0; ib < n_probes; ib += 6) {  // six probes per step: their loads go out together
    uint64_t key[6];
#pragma unroll
    for (int u = 0; u < 6; u++) key[u] = ib + u < n_probes ? wfa_ld64u(pr.t + (ib + u + 1) * step - 1) : 0;
#pragma unroll
    for (int u = 0; u < 6; u++) {
      if (ib + u >= n_probes) break;
      const int j = (ib + u + 1) * step - 1;
      const uint64_t mixed = kidx_mix(key[u]);
      const uint32_t fp = kidx_fp(mixed);
      uint32_t v;
      TRGT_KIDX_FOR(idx, mixed, v) {
        if ((v >> 9) != fp) continue;
        const int d = (int)(v & 511u), b = (int)(((uint32_t)d * inv) >> 20), r = d - b * blen, q = j - r;
        if (b >= nb || r + 8 > blen || q < 0 || q + blen > pr.T) continue;
        // a fingerprint clash or a chance 8-mer would push the pair onto the verifying path: compare the 8-mer itself
        // and eight more bases of the block (a true block occurrence passes both)
        if (wfa_ld64u(pr.p + d) != key[u]) continue;
        if (r + 16 <= blen) { if (wfa_ld64u(pr.p + d + 8) != wfa_ld64u(pr.t + j + 8)) continue; }
        else if (r >= 8) { if (wfa_ld64u(pr.p + d - 8) != wfa_ld64u(pr.t + j - 8)) continue; }
        const int k = j - d;
        kmin = wfa_imin(kmin, k);
        kmax = wfa_imax(kmax, k);
      }
    }
  }
  if (kmin == INT_MAX) return 0;
  *klo = wfa_imax(-pr.P, kmin - R);
  *khi = wfa_imin(pr.T, kmax + R);
  return 1;
}

// Seed pass of one pair: 1 and the band if the band pass can take it, 0 if the pair goes to the next tier.
// pr.p must be readable (only the rare verifying path reads it).
TRGT_HD int flank_tier1_seed_thread(const KmerIndex &idx, const WfaProb &pr, int S, int *klo, int *khi) {
  const int cap = wfa_imin(S, wfa_imax(pr.x, pr.oe));
  if (cap > FT1_SMAX) return 0;
  if (flank_seed_hull_thread(idx, pr, cap, klo, khi) != 1) return 0;  // no hit at all: no verified one either
  if (*khi - *klo + 1 <= FT1_WMAX && *khi + pr.P < pr.T) return 1;  // (the text-end rule of flank_seed_band)
  if (flank_seed_band_thread(idx, pr, cap, klo, khi) != 1) return 0;  // a hit off the alignment: verify them
  return *khi - *klo + 1 <= FT1_WMAX ? 1 : 0;
}

#define FT1_WIN_BYTES 288  // text window of a pair: <= 256 of piece + 3 of band + 15 of alignment + 12 of over-read, in 16-byte chunks
#define FT1_WIN_STRIDE 304  // bytes between the windows of two lanes: an 8-byte read of the window's last bytes touches the 12 after them
#define FT1_PMAX 256       // longest piece the band pass takes
#define FT1_HIST_HALFS ((FT1_SMAX + 1) * 3 * FT1_WMAX)  // 16-bit cells of one pair's history (216 bytes)

// History of the band pass: cell (s, component, k) at h[((row(s) * 3 + c) * WM + k - blo) * LS] as offset - blo
// (never negative for a reachable cell), -1 = null.  LS = distance between a lane's consecutive cells (cells of
// the lanes of a warp interleaved: every lane of a converged access hits its own bank).
// wfa_live_scores for the first tier (s_cap <= FT1_SMAX)
TRGT_HD unsigned ft1_live_scores(int x, int oe, int e, int s_cap, unsigned *live_gap) {
  unsigned long long lg = 0;
  const unsigned long long lm = wfa_live_scores(x, oe, e, s_cap, &lg);
  if (live_gap) *live_gap = (unsigned)lg;
  return (unsigned)lm;
}
// the history keeps a row only for those scores: row of score s
TRGT_HD int ft1_row(unsigned live_m, int s) {
#if defined(__CUDA_ARCH__)
  return __popc(live_m & ((1u << s) - 1u));
#else
  return __builtin_popcount(live_m & ((1u << s) - 1u));
#endif
}

template <int LS, int WM = FT1_WMAX>
struct Ft1Hist {
  int16_t *h;
  int blo, W;
  unsigned present;  // bit s: score s has a wavefront
  unsigned live_m;   // ft1_live_scores of the scoring (row numbering)
  TRGT_HD int cell(int s, int c, int k) const {
    if (s < 0 || !((present >> s) & 1u) || k < blo || k >= blo + W) return TRGT_WFA_NULL;
    const int v = h[((ft1_row(live_m, s) * 3 + c) * WM + (k - blo)) * LS];
    return v < 0 ? TRGT_WFA_NULL : v + blo;
  }
  struct Row {
    const Ft1Hist *h;
    int s;
    TRGT_HD int m(int k) const { return h->cell(s, 0, k); }
    TRGT_HD int i(int k) const { return s < 1 ? TRGT_WFA_NULL : h->cell(s, 1, k); }
    TRGT_HD int d(int k) const { return s < 1 ? TRGT_WFA_NULL : h->cell(s, 2, k); }
  };
  TRGT_HD Row row(int s) const { Row r; r.h = this; r.s = s; return r; }
};

// wfa_forward_band_hist_narrow_thread on the staged window: text offset h lives at win[h - a0].
// The lanes of a warp, each on its own pair, walk the SAME sequence of cells (which scores have a wavefront
// depends only on the scoring; the band is FT1_WMAX diagonals for everybody, the ones past a lane's own band
// masked), so the recurrences, the first eight bytes of every match extension and all history traffic run
// with the whole warp.  What differs between lanes is where the long extensions are (the error-free stretches
// of the piece, ~P bases in all): a cell that is not settled by its first eight bytes is parked, and after the
// cells of a score all parked extensions are finished in one loop, sixteen bytes a step, that the lanes leave
// together.  (With a loop per cell every cell costs the warp its longest extension among 32 pairs.)
// lanes: the lanes of the warp that call this together (0 = a lane on its own); idle: this lane has no pair
// and only keeps company.  live_m / live_g: ft1_live_scores of the scoring (the same for every pair of a launch,
// so the caller computes them once).  An all-null wavefront reads exactly like an absent one, so a score that
// only has null predecessors may be computed (here: whenever the scoring allows a wavefront) or skipped alike.
// WM: diagonals a row has room for (the band may be narrower); SHARED: the text (win) is in shared memory, else in
// global memory like the pattern.
template <int LS, int WM = FT1_WMAX, bool SHARED = true>
TRGT_HD WfaEnd ft1_forward(const WfaProb &pr, int s_cap, const uint8_t *win, int a0, int16_t *hist, unsigned *present_out,
                           unsigned lanes, bool idle, unsigned live_m, unsigned live_g) {
  WfaEnd out;
  out.status = TRGT_WFA_OK; out.s = 0; out.k = 0; out.off = 0;
  const int W = idle ? 0 : pr.bhi - pr.blo + 1;
  unsigned present = 0;
  bool done = idle;
#define FT1_LD(s_, c_, i_) ft1_ld(hist[((ft1_row(live_m, (s_)) * 3 + (c_)) * WM + (i_)) * LS], pr.blo)
#define FT1_ST(s_, c_, i_, val) hist[((ft1_row(live_m, (s_)) * 3 + (c_)) * WM + (i_)) * LS] = (int16_t)((val) < 0 ? -1 : (val) - pr.blo)
  for (int s = 0; s <= s_cap; s++) {
    if (!((live_m >> s) & 1u)) continue;  // the same for every lane
    if (!TRGT_ANY(lanes, !done)) break;
    const int sx = s - pr.x, so = s - pr.oe, se = s - pr.e;
    const bool px = sx >= 0 && ((live_m >> sx) & 1u), po = so >= 0 && ((live_m >> so) & 1u);
    const bool pe = se >= 1 && ((live_g >> se) & 1u);
    unsigned parked = 0;  // bit idx: the extension of cell (s, idx) is not finished
#pragma unroll
    for (int idx = 0; idx < WM; idx++) {
      if (done || idx >= W) continue;
      const int k = pr.blo + idx;
      int mx = TRGT_WFA_NULL;
      if (s == 0) {
        if (k >= -pr.pbf && k <= pr.tbf) mx = k >= 0 ? k : 0;
      } else {
        const int o_l = (po && idx > 0) ? FT1_LD(so, 0, idx - 1) : TRGT_WFA_NULL;
        const int o_r = (po && idx + 1 < W) ? FT1_LD(so, 0, idx + 1) : TRGT_WFA_NULL;
        const int i_l = (pe && idx > 0) ? FT1_LD(se, 1, idx - 1) : TRGT_WFA_NULL;
        const int d_r = (pe && idx + 1 < W) ? FT1_LD(se, 2, idx + 1) : TRGT_WFA_NULL;
        const int i1 = wfa_imax(o_l, i_l) + 1;
        const int d1 = wfa_imax(o_r, d_r);
        const int mm = (px ? FT1_LD(sx, 0, idx) : TRGT_WFA_NULL) + 1;
        mx = wfa_imax(mm, wfa_imax(i1, d1));
        const int h = mx, v = mx - k;
        if (mx < 0 || h > pr.T || v > pr.P || v < 0) mx = TRGT_WFA_NULL;
        FT1_ST(s, 1, idx, i1);
        FT1_ST(s, 2, idx, d1);
      }
      if (mx >= 0) {  // the first eight bytes of the match extension along the diagonal
        const int v = mx - k, n = wfa_imin(pr.P - v, pr.T - mx);
        if (n > 0) {
          const uint64_t x = ft1_ld64_piece(pr.p + v) ^ (SHARED ? ft1_ld64_win(win + (mx - a0)) : ft1_ld64_piece(win + (mx - a0)));
          int adv = x ? (wfa_ctz64(x) >> 3) : 8;
          if (adv > n) adv = n;
          mx += adv;
          if (!x && n > 8) parked |= 1u << idx;
        }
      }
      FT1_ST(s, 0, idx, mx);
    }
    while (TRGT_ANY(lanes, parked != 0)) {  // the parked extensions, sixteen bytes a step
      if (parked) {
#if defined(__CUDA_ARCH__)
        const int idx = __ffs((int)parked) - 1;
#else
        const int idx = __builtin_ctz(parked);
#endif
        const int k = pr.blo + idx;
        int mx = FT1_LD(s, 0, idx);
        const int v = mx - k, n = wfa_imin(pr.P - v, pr.T - mx);
        const uint64_t x0 = ft1_ld64_piece(pr.p + v) ^ (SHARED ? ft1_ld64_win(win + (mx - a0)) : ft1_ld64_piece(win + (mx - a0)));
        const uint64_t x1 = n > 8 ? ft1_ld64_piece(pr.p + v + 8) ^ (SHARED ? ft1_ld64_win(win + (mx - a0) + 8) : ft1_ld64_piece(win + (mx - a0) + 8)) : 0;  // (stays inside the padding)
        int adv = x0 ? (wfa_ctz64(x0) >> 3) : (x1 ? 8 + (wfa_ctz64(x1) >> 3) : 16);
        if (adv > n) adv = n;
        mx += adv;
        FT1_ST(s, 0, idx, mx);
        if (x0 || x1 || n <= 16) parked &= ~(1u << idx);
      }
    }
    if (!done) {  // end condition, lowest diagonal first
      present |= 1u << s;
      for (int idx = 0; idx < W; idx++) {
        const int k = pr.blo + idx;
        const int mx = FT1_LD(s, 0, idx);
        if (mx < 0) continue;
        const int h = mx, v = mx - k;
        if (v >= 0 && v <= pr.P && h <= pr.T &&
            ((h >= pr.T && pr.P - v <= pr.pef) || (v >= pr.P && pr.T - h <= pr.tef))) {
          out.s = s; out.k = k; out.off = mx;
          done = true;
          break;
        }
      }
    }
  }
#undef FT1_LD
#undef FT1_ST
  if (!done) out.status = TRGT_WFA_MAX_STEPS;
  *present_out = present;
  return out;
}

// Band pass of one pair the seed pass listed with band [klo, khi]: 0 and *hit filled, or 1 if the pair goes to
// the next tier.  pr: the pair as the reference poses it (pr.t is not dereferenced: the text is read from the
// window, win[h - a0] = text[h] for every h in [max(klo, 0), min(T, khi + P) + 11]).
template <int LS>
TRGT_HD int flank_tier1_band_thread(const WfaProb &pr, int klo, int khi, int S, double min_flank_id_frac,
                                    const uint8_t *win, int a0, int16_t *hist, FlankHit *hit, unsigned lanes = 0,
                                    bool idle = false, unsigned live_m = 0, unsigned live_g = 0) {
  const int cap = wfa_imin(S, wfa_imax(pr.x, pr.oe));
  const bool skip = idle || cap > FT1_SMAX || khi - klo + 1 > FT1_WMAX;
  if (live_m == 0 && !skip) live_m = ft1_live_scores(pr.x, pr.oe, pr.e, cap, &live_g);  // (bit 0 is always set)
  WfaProb bp = pr;
  bp.blo = klo; bp.bhi = khi;
  unsigned present = 0;
  const WfaEnd end = ft1_forward<LS>(bp, cap, win, a0, hist, &present, lanes, skip, live_m, live_g);
  if (skip || end.status != TRGT_WFA_OK) return 1;
  const Ft1Hist<LS> H{hist, klo, khi - klo + 1, present, live_m};
  WfaFlankSink sink(pr.T);
  wfa_backtrace_h(pr, end.s, end.k, end.off, H, sink);
  hit->matches = sink.matches;
  hit->score = -end.s;
  if ((double)sink.matches >= (double)pr.P * min_flank_id_frac) {  // span_locater.rs:19-25, :46
    hit->via = 2; hit->start = sink.ystart(); hit->end = sink.yend();
  } else {
    hit->via = 3; hit->start = 0; hit->end = 0;
  }
  return 0;
}

// End-to-end alignment of a short pair by one lane with the same machinery (phase B's members that differ from their
// backbone): with cost cap S nothing is reachable outside |k| <= (S - o) / e (wfa_e2e_narrow's argument), so that band
// is the whole computation.  Both sequences are read where they lie (a few dozen bases each); the history is the
// lane's 16-bit rows.  Fills *end (status OK / MAX_STEPS); on OK the back-trace hands the CIGAR to `sink`.
#define E2L_WMAX 8   // room for |k| <= 3: cost cap 8 under the consensus scoring 2 / 5 / 1
template <int LS, class Sink>
TRGT_HD void e2e_narrow_lane(const WfaProb &pr, int S, int16_t *hist, unsigned lanes, bool idle, unsigned live_m,
                             unsigned live_g, WfaEnd *end, Sink &sink) {
  const int o = pr.oe - pr.e;
  const int R = S > o ? (S - o) / pr.e : 0;
  WfaProb bp = pr;
  bp.blo = wfa_imax(-pr.P, -R);
  bp.bhi = wfa_imin(pr.T, R);
  const bool skip = idle || bp.bhi - bp.blo + 1 > E2L_WMAX || S > FT1_SMAX;
  unsigned present = 0;
  *end = ft1_forward<LS, E2L_WMAX, false>(bp, S, pr.t, 0, hist, &present, lanes, skip, live_m, live_g);
  if (skip) { end->status = TRGT_WFA_MAX_STEPS; return; }
  if (end->status != TRGT_WFA_OK) return;
  const Ft1Hist<LS, E2L_WMAX> H{hist, bp.blo, bp.bhi - bp.blo + 1, present, live_m};
  wfa_backtrace_h(pr, end->s, end->k, end->off, H, sink);
}

// ---------------------------------------------------------------- unit-cost edit distance ------

// Levenshtein distance, one lane per pair: bit-vector DP (Myers 1999, global variant of Hyyro 2003)
// over the shorter sequence, which get_dist's MAX_OPS guard (genotype_cluster.rs:237-243) bounds
// by 100 symbols, so one 128-bit word holds the column.  Score-only, so any exact algorithm
// returns what the reference's WFA edit aligner returns.
TRGT_HD int edit_distance_128(const uint8_t *a, int la, const uint8_t *b, int lb) {
  typedef unsigned __int128 u128;
  const uint8_t *pat = a; int m = la;
  const uint8_t *txt = b; int n = lb;
  if (m > n) { pat = b; m = lb; txt = a; n = la; }
  if (m == 0) return n;
  u128 peqA = 0, peqC = 0, peqG = 0, peqT = 0;
  for (int i = 0; i < m; i++) {
    const u128 bit = (u128)1 << i;
    const uint8_t c = pat[i];
    if (c == 'A') peqA |= bit; else if (c == 'C') peqC |= bit; else if (c == 'G') peqG |= bit; else if (c == 'T') peqT |= bit;
  }
  const u128 ones = (m == 128) ? ~(u128)0 : (((u128)1 << m) - 1);
  const u128 top = (u128)1 << (m - 1);
  u128 Pv = ones, Mv = 0;
  int score = m;
  for (int j = 0; j < n; j++) {
    const uint8_t c = txt[j];
    u128 Eq;
    if (c == 'A') Eq = peqA; else if (c == 'C') Eq = peqC; else if (c == 'G') Eq = peqG; else if (c == 'T') Eq = peqT;
    else {
      Eq = 0;
      for (int i = 0; i < m; i++) if (pat[i] == c) Eq |= (u128)1 << i;
    }
    const u128 Xv = Eq | Mv;
    const u128 Xh = ((((Eq & Pv) + Pv) ^ Pv) | Eq) & ones;
    u128 Ph = (Mv | ~(Xh | Pv)) & ones;
    u128 Mh = Pv & Xh;
    if (Ph & top) score++;
    else if (Mh & top) score--;
    Ph = ((Ph << 1) | 1) & ones;
    Mh = (Mh << 1) & ones;
    Pv = (Mh | ~(Xv | Ph)) & ones;
    Mv = Ph & Xv;
  }
  return score;
}

}  // namespace trgt
