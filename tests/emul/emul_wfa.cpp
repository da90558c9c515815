// emul_wfa.cpp -- TEST INFRASTRUCTURE ONLY.  Serial instantiation of trgt_b200/csrc/wfa_core.h
// (see emul_hmm.cpp).  Never linked into libtrgt_b200.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../trgt_b200/csrc/wfa_core.h"
#include "lanes.h"

using namespace trgt;

extern "C" {

// Two-pass alignment exactly as the device runs it: ring score pass, cone trace pass.
// out[0]=status out[1]=score(-cost) out[2]=k out[3]=off out[4]=matches out[5]=ystart out[6]=yend
// out[7]=n_words out[8]=trace ints used (bound)
int emu_wfa_align(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int pbf, int pef,
                  int tbf, int tef, int *out, uint32_t *words, uint32_t words_cap) {
  // the cores read 8 bytes at a time: give them the 16-byte padding every engine buffer has
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 16, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  const uint8_t *p = pbuf.data(), *t = tbuf.data();
  WfaProb pr;
  pr.p = p; pr.P = P; pr.t = t; pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = pbf; pr.pef = pef; pr.tbf = tbf; pr.tef = tef;
  wfa_unband(pr);
  SerialGroup g;
  std::vector<int> ring(wfa_ring_ints(pr) + 1, 0x7ead);
  const WfaEnd end = wfa_score_ring(g, pr, ring.data(), wfa_score_cap(pr));
  out[0] = end.status; out[1] = -end.s; out[2] = end.k; out[3] = end.off;
  if (end.status != TRGT_WFA_OK) return end.status;
  const size_t need = wfa_trace_ints(pr, end.s);
  std::vector<int> ws(need + 1, 0x7ead);
  const int rc = wfa_trace_forward(g, pr, end.s, end.k, ws.data(), need);
  if (rc != 0) { out[0] = rc; return rc; }
  // the cone must reproduce the terminating cell
  {
    WfaView v = wfa_hist_view(ws.data(), end.s);
    if (wfa_at(v.m, v, end.k) != end.off) { out[0] = -999; return -999; }
  }
  WfaFlankSink fs(T);
  wfa_backtrace(pr, end.s, end.k, end.off, ws.data(), fs);
  out[4] = fs.matches; out[5] = fs.ystart(); out[6] = fs.yend();
  WfaCigarSink cs(words, words_cap);
  wfa_backtrace(pr, end.s, end.k, end.off, ws.data(), cs);
  out[7] = (int)cs.finish();
  if (cs.overflow) { out[0] = -998; return -998; }
  out[8] = (int)need;
  return 0;
}

int emu_flank_scan(const uint8_t *piece, int P, const uint8_t *t, int T) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), piece, P);
  memcpy(tbuf.data(), t, T);
  SerialGroup g;
  return flank_scan(g, pbuf.data(), P, tbuf.data(), T);
}

// flank fallback through the seed filter + banded pass + cone trace.
// out: [0]=rc (0 resolved, 1 deferred) [1]=via [2]=matches [3]=score [4]=start [5]=end
int emu_flank_banded(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S, double frac,
                     int ws_ints, int *out) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 16, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = 0; pr.pef = 0; pr.tbf = T; pr.tef = T;
  wfa_unband(pr);
  SerialGroup g;
  std::vector<int> ws(ws_ints + 1, 0x7ead);
  uint64_t keys[32];
  FlankHit hit = {0, 0, 0, 0, 0};
  out[0] = flank_locate_banded(g, pr, S, frac, keys, ws.data(), (size_t)ws_ints, &hit);
  out[1] = hit.via; out[2] = hit.matches; out[3] = hit.score; out[4] = hit.start; out[5] = hit.end;
  return out[0];
}

int emu_edit_distance(const uint8_t *a, int la, const uint8_t *b, int lb) {
  return edit_distance_128(a, la, b, lb);
}

}  // extern "C"

// ---- the same entry points with N lock-step host lanes instead of one --------------------------

extern "C" {

int emu_wfa_align_lanes(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int pbf, int pef,
                        int tbf, int tef, int *out, uint32_t *words, uint32_t words_cap, int lanes) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 16, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = pbf; pr.pef = pef; pr.tbf = tbf; pr.tef = tef;
  wfa_unband(pr);
  std::vector<int> ring(wfa_ring_ints(pr) + 1, 0x7ead);
  WfaEnd end{};
  trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
    const WfaEnd e2 = wfa_score_ring(g, pr, ring.data(), wfa_score_cap(pr));
    if (g.lane() == 0) end = e2;
  });
  out[0] = end.status; out[1] = -end.s; out[2] = end.k; out[3] = end.off;
  if (end.status != TRGT_WFA_OK) return end.status;
  const size_t need = wfa_trace_ints(pr, end.s);
  std::vector<int> ws(need + 1, 0x7ead);
  int rc = 0;
  trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
    const int r = wfa_trace_forward(g, pr, end.s, end.k, ws.data(), need);
    if (g.lane() == 0) rc = r;
  });
  if (rc != 0) { out[0] = rc; return rc; }
  WfaFlankSink fs(T);
  wfa_backtrace(pr, end.s, end.k, end.off, ws.data(), fs);
  out[4] = fs.matches; out[5] = fs.ystart(); out[6] = fs.yend();
  WfaCigarSink cs(words, words_cap);
  wfa_backtrace(pr, end.s, end.k, end.off, ws.data(), cs);
  out[7] = (int)cs.finish();
  return 0;
}

int emu_flank_scan_lanes(const uint8_t *piece, int P, const uint8_t *t, int T, int lanes) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), piece, P);
  if (T > 0) memcpy(tbuf.data(), t, T);
  int res = -2;
  trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
    const int r = flank_scan(g, pbuf.data(), P, tbuf.data(), T);
    if (g.lane() == 0) res = r;
  });
  return res;
}

int emu_flank_banded_lanes(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S,
                           double frac, int ws_ints, int *out, int lanes) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 16, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = 0; pr.pef = 0; pr.tbf = T; pr.tef = T;
  wfa_unband(pr);
  std::vector<int> ws(ws_ints + 1, 0x7ead);
  uint64_t keys[32];
  FlankHit hit = {0, 0, 0, 0, 0};
  int rc = -1;
  trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
    FlankHit h = {0, 0, 0, 0, 0};
    const int r = flank_locate_banded(g, pr, S, frac, keys, ws.data(), (size_t)ws_ints, &h);
    if (g.lane() == 0) { rc = r; hit = h; }
  });
  out[0] = rc;
  out[1] = hit.via; out[2] = hit.matches; out[3] = hit.score; out[4] = hit.start; out[5] = hit.end;
  return rc;
}

}  // extern "C"

// ---- index-based exact search and seed filter ---------------------------------------------------

extern "C" {

// lanes == 0: serial.  out[0] = flank_scan_indexed result; if it is -1 (no exact hit) also runs the
// banded fallback through the index: out[1]=rc out[2]=via out[3]=matches out[4]=score out[5]=start out[6]=end
int emu_flank_indexed(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S, double frac,
                      int ws_ints, int *out, int lanes) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = 0; pr.pef = 0; pr.tbf = T; pr.tef = T;
  wfa_unband(pr);
  std::vector<uint16_t> islot(TRGT_KIDX_SLOTS);
  KmerIndex idx{islot.data()};
  std::vector<int> ws(ws_ints + 1, 0x7ead), cand(TRGT_CAND_CAP + 1);
  uint64_t keys[32];
  FlankHit hit = {0, 0, 0, 0, 0};
  int pos = -3, rc = -3;
  auto body = [&](const auto &g) {
    kidx_build(g, idx, pr.p, P);
    int ps = flank_scan_indexed(g, idx, pr.p, P, pr.t, T, cand.data());
    if (ps == -2) ps = flank_scan(g, pr.p, P, pr.t, T);
    int r = -3;
    FlankHit h = {0, 0, 0, 0, 0};
    if (ps == -1) r = flank_locate_banded(g, pr, S, frac, keys, ws.data(), (size_t)ws_ints, &h, &idx, cand.data());
    if (g.lane() == 0) { pos = ps; rc = r; hit = h; }
  };
  if (lanes == 0) { SerialGroup g; body(g); }
  else trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) { body(g); });
  out[0] = pos; out[1] = rc;
  out[2] = hit.via; out[3] = hit.matches; out[4] = hit.score; out[5] = hit.start; out[6] = hit.end;
  return 0;
}

}  // extern "C"

extern "C" {

// end-to-end pair through wfa_e2e_narrow + back-trace with `lanes` lock-step lanes.
// out[0]=status out[1]=score out[2]=n_words
int emu_e2e_narrow(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S, int ws_ints,
                   int *out, uint32_t *words, uint32_t words_cap, int lanes) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 16, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = pr.pef = pr.tbf = pr.tef = 0;
  wfa_unband(pr);
  std::vector<int> ws(ws_ints + 1, 0x7ead);
  WfaEnd end{};
  trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
    const WfaEnd e2 = wfa_e2e_narrow(g, pr, S, ws.data(), (size_t)ws_ints);
    if (g.lane() == 0) end = e2;
  });
  out[0] = end.status; out[1] = -end.s; out[2] = 0;
  if (end.status != TRGT_WFA_OK) return end.status;
  WfaCigarSink cs(words, words_cap);
  wfa_backtrace(pr, end.s, end.k, end.off, ws.data(), cs);
  out[2] = (int)cs.finish();
  return 0;
}

// end-to-end alignment of a short pair by one lane on 16-bit history rows (e2e_narrow_lane): out[0] status, out[1] score,
// out[2] number of CIGAR words
int emu_e2e_lane(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S, int *out,
                 uint32_t *words, uint32_t words_cap) {
  std::vector<uint8_t> pbuf(P + 32, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = pr.pef = pr.tbf = pr.tef = 0;
  wfa_unband(pr);
  int16_t hist[(FT1_SMAX + 1) * 3 * E2L_WMAX];
  for (auto &h : hist) h = 0x7ead;
  unsigned live_g = 0;
  const unsigned live_m = S <= FT1_SMAX ? ft1_live_scores(pr.x, pr.oe, pr.e, S, &live_g) : 1u;
  WfaEnd end{};
  WfaCigarSink cs(words, words_cap);
  e2e_narrow_lane<1>(pr, S, hist, 0u, false, live_m, live_g, &end, cs);
  out[0] = end.status; out[1] = -end.s; out[2] = 0;
  if (end.status != TRGT_WFA_OK) return end.status;
  out[2] = (int)cs.finish();
  return 0;
}

// exact search by one lane per pair (flank_exact_thread): copies and index built by `lanes` lanes;
// text must carry 16 readable bytes on both sides
int emu_flank_exact_thread(const uint8_t *p, int P, const uint8_t *t, int T, int lanes) {
  std::vector<uint16_t> slot(TRGT_KIDX_SLOTS);
  alignas(16) static thread_local uint8_t copies[FXT_COPIES * FXT_STRIDE];
  memset(copies, 0xEE, sizeof copies);
  const KmerIndex idx{slot.data()};
  if (lanes == 0) {
    SerialGroup g;
    fxt_build_copies(g, p, P, copies);
    kidx_build(g, idx, copies + 8, P);
  } else {
    uint8_t *cp = copies;
    trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
      fxt_build_copies(g, p, P, cp);
      kidx_build(g, idx, cp + 8, P);
    });
  }
  return flank_exact_thread(idx, copies, P, t, T);
}

// first cost tier by one lane (flank_locate_tier1_thread): out[0]=rc (0 settled, 1 handed on) out[1]=via
// out[2]=matches out[3]=score out[4]=start out[5]=end
int emu_flank_tier1_thread(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S,
                           double frac, int *out) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = 0; pr.pef = 0; pr.tbf = T; pr.tef = T;
  wfa_unband(pr);
  std::vector<uint16_t> islot(TRGT_KIDX_SLOTS);
  KmerIndex idx{islot.data()};
  SerialGroup g;
  kidx_build(g, idx, pr.p, P);
  int ws[FT1_WS_INTS];
  for (int i = 0; i < FT1_WS_INTS; i++) ws[i] = 0x7ead;
  FlankHit hit = {0, 0, 0, 0, 0};
  out[0] = flank_locate_tier1_thread(pr, S, frac, ws, &hit, idx);
  out[1] = hit.via; out[2] = hit.matches; out[3] = hit.score; out[4] = hit.start; out[5] = hit.end;
  return 0;
}

// first cost tier in two passes (flank_tier1_seed_thread + flank_tier1_band_thread on a staged window with the
// 16-bit on-chip history): same outputs as emu_flank_tier1_thread; `slack` (0..15) = how far the window starts
// before the first byte the band can touch (on the device: the distance to the previous 16-byte boundary)
int emu_flank_tier1_split(const uint8_t *p_in, int P, const uint8_t *t_in, int T, int x, int o, int e, int S,
                          double frac, int slack, int *out) {
  std::vector<uint8_t> pbuf(P + 16, 0), tbuf(T + 32, 0);
  memcpy(pbuf.data(), p_in, P);
  memcpy(tbuf.data(), t_in, T);
  WfaProb pr;
  pr.p = pbuf.data(); pr.P = P; pr.t = tbuf.data(); pr.T = T; pr.x = x; pr.oe = o + e; pr.e = e;
  pr.pbf = 0; pr.pef = 0; pr.tbf = T; pr.tef = T;
  wfa_unband(pr);
  std::vector<uint16_t> islot(TRGT_KIDX_SLOTS);
  KmerIndex idx{islot.data()};
  SerialGroup g;
  kidx_build(g, idx, pr.p, P);
  FlankHit hit = {0, 0, 0, 0, 0};
  out[0] = 1; out[1] = out[2] = out[3] = out[4] = out[5] = 0;
  int klo = 0, khi = 0;
  if (P > FT1_PMAX || flank_tier1_seed_thread(idx, pr, S, &klo, &khi) != 1) return 0;
  const int first = klo > 0 ? klo : 0;
  const int a0 = first - slack;  // may be negative: bytes before the text are never read
  uint8_t win[FT1_WIN_STRIDE];
  memset(win, 0xEE, sizeof win);
  for (int i = 0; i < FT1_WIN_BYTES; i++) {
    const int h = a0 + i;
    if (h >= 0 && h < T + 16) win[i] = tbuf[h];  // the sequence buffers of the engine carry 16 bytes of padding
  }
  int16_t hist[FT1_HIST_HALFS];
  for (int i = 0; i < FT1_HIST_HALFS; i++) hist[i] = 0x7ead;
  WfaProb q = pr;
  q.t = nullptr;  // must not be read
  out[0] = flank_tier1_band_thread<1>(q, klo, khi, S, frac, win, a0, hist, &hit);
  out[1] = hit.via; out[2] = hit.matches; out[3] = hit.score; out[4] = hit.start; out[5] = hit.end;
  return 0;
}

}  // extern "C"
