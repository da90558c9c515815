// emul_hmm.cpp -- TEST INFRASTRUCTURE ONLY.  Instantiates the HMM core (trgt_b200/csrc/hmm_core.h)
// with the one-lane SerialGroup so that its index arithmetic and tie-breaking can be compared with
// the oracle on a machine without a GPU.  Never linked into libtrgt_b200.so.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../trgt_b200/csrc/hmm_core.h"
#include "../../trgt_b200/csrc/hmm_host.h"
#include "lanes.h"

using namespace trgt;

extern "C" {

// returns n_spans (collapsed) or <0; path_out gets Hmm::label (forward order), *path_len its length
long emu_hmm_annotate_lanes(const uint8_t *motifs, const uint64_t *moff, int nm, const uint8_t *allele, int L,
                            uint32_t *mc, HmmSpan *spans, uint32_t span_cap, double *purity,
                            uint32_t *path_out, uint64_t path_cap, uint64_t *path_len, int *S_out, int lanes);

long emu_hmm_annotate(const uint8_t *motifs, const uint64_t *moff, int nm, const uint8_t *allele, int L,
                      uint32_t *mc, HmmSpan *spans, uint32_t span_cap, double *purity,
                      uint32_t *path_out, uint64_t path_cap, uint64_t *path_len, int *S_out) {
  return emu_hmm_annotate_lanes(motifs, moff, nm, allele, L, mc, spans, span_cap, purity, path_out, path_cap,
                                path_len, S_out, 0);
}

// lanes == 0: one serial lane (SerialGroup); otherwise that many lock-step host threads
long emu_hmm_annotate_lanes(const uint8_t *motifs, const uint64_t *moff, int nm, const uint8_t *allele, int L,
                            uint32_t *mc, HmmSpan *spans, uint32_t span_cap, double *purity,
                            uint32_t *path_out, uint64_t path_cap, uint64_t *path_len, int *S_out, int lanes) {
  static HmmJumpTable jt;
  int max_len = 0;
  int mbytes = 0;
  for (int b = 0; b < nm; b++) {
    int n = (int)(moff[b + 1] - moff[b]);
    if (n > max_len) max_len = n;
    mbytes += n;
  }
  jt.ensure(max_len);
  const HmmConsts c = hmm_make_consts();
  const int nb = nm + 1;
  std::vector<uint8_t> o_bytes(mbytes + 1);
  std::vector<uint32_t> o_moff(nb), o_mmoff(nb);
  std::vector<uint16_t> o_n(nb), o_ms(nb);
  int S_guess = 7;
  for (int b = 0; b < nm; b++) S_guess += 3 * (int)(moff[b + 1] - moff[b]) + 1;
  std::vector<uint16_t> o_stblk(S_guess + 8);
  HmmModel model;
  SerialGroup g;
  int S = 0;
  if (lanes <= 0) {
    S = hmm_model_build(g, motifs, moff, nm, jt.off.data(), o_bytes.data(), o_moff.data(), o_mmoff.data(),
                        o_n.data(), o_ms.data(), o_stblk.data(), &model);
  } else {
    trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &lg) {
      HmmModel m2;
      const int s2 = hmm_model_build(lg, motifs, moff, nm, jt.off.data(), o_bytes.data(), o_moff.data(),
                                     o_mmoff.data(), o_n.data(), o_ms.data(), o_stblk.data(), &m2);
      if (lg.lane() == 0) { S = s2; model = m2; }
    });
  }
  if (S < 0) return -400;
  if (S != S_guess) return -401;
  if (S_out) *S_out = S;
  for (int b = 0; b < nm; b++) mc[b] = 0;
  if (L == 0) {
    *purity = NAN;
    if (path_len) *path_len = 0;
    return 0;
  }
  if (lanes == -2) {  // single-motif fast path: registers + one packed word per column (hmm_viterbi_lane)
    const int n = nm == 1 ? (int)(moff[1] - moff[0]) : 0;
    if (n < 1 || n > HMM_LANE_NMAX) return -405;
    const int stride = 3;  // any stride: the device uses 32
    std::vector<uint32_t> words((size_t)L * stride + 1, 0xEEEEEEEEu);
    const uint64_t mb = hmm_pack_motif(motifs + moff[0], n);
    const double *jump = jt.lp.data() + jt.off[n];
    HmmAnnot a1, a2;
    uint64_t plen = 0;
    std::vector<uint32_t> rev(path_cap ? path_cap : 1);
#define EMU_LANE(N)                                                                                              \
  case N: {                                                                                                      \
    hmm_viterbi_lane<N>(c, jump, mb, allele, L, words.data(), stride);                                           \
    const HmmModelSingle<N> m1{mb};                                                                              \
    HmmBpWords<N> bw(words.data(), stride, L);                                                             \
    std::vector<uint32_t> mc_tmp(2, 0);                                                                          \
    a1 = hmm_annotate_bp(m1, allele, L, bw, 6, mc_tmp.data(), HmmSpanArray{nullptr}, 0, rev.data(), path_cap, 0, &plen); \
    if (a1.status < 0) return -402;                                                                              \
    if (a1.n_spans > span_cap) return -2;                                                                        \
    a2 = hmm_annotate_bp(m1, allele, L, bw, 6, mc, HmmSpanArray{spans}, a1.n_spans, (uint32_t *)nullptr, 0, 0, (uint64_t *)nullptr); \
    {  /* the table-driven walk the device uses for counting and spans must agree with it */                     \
      std::vector<HmmLaneEntry> tab(32);                                                                         \
      for (int st_ = 0; st_ < 32; st_++) tab[st_] = hmm_lane_table_entry(N, st_);                                \
      std::vector<uint32_t> mc3(2, 0);                                                                           \
      std::vector<HmmSpan> sp3(a1.n_spans + 1);                                                                  \
      uint64_t plen3 = 0;                                                                                        \
      const HmmAnnot a3 = hmm_walk_table(tab.data(), N, mb, L, words.data(), stride, 6, mc3.data(),      \
                                         HmmSpanArray{sp3.data()}, a1.n_spans, &plen3);                          \
      if (a3.status != 0 || a3.n_spans != a1.n_spans || plen3 != plen || mc3[0] != mc[0]) return -406;           \
      if (!(a3.purity == a1.purity || (a3.purity != a3.purity && a1.purity != a1.purity))) return -407;          \
      for (uint32_t q_ = 0; q_ < a1.n_spans; q_++)                                                               \
        if (sp3[q_].motif_index != spans[q_].motif_index || sp3[q_].start != spans[q_].start || sp3[q_].end != spans[q_].end) return -408; \
    }                                                                                                            \
    break;                                                                                                       \
  }
    switch (n) {
      EMU_LANE(1) EMU_LANE(2) EMU_LANE(3) EMU_LANE(4) EMU_LANE(5) EMU_LANE(6) EMU_LANE(7)
      default: return -405;
    }
#undef EMU_LANE
    if (a2.n_spans != a1.n_spans) return -403;
    *purity = a1.purity;
    if (path_len) *path_len = plen;
    if (path_out) {
      const uint64_t np = plen < path_cap ? plen : path_cap;
      for (uint64_t i = 0; i < np; i++) path_out[i] = rev[np - 1 - i];
    }
    return (long)a1.n_spans;
  }
  std::vector<double> sc0(S), sc1(S);
  std::vector<uint8_t> bp((size_t)(L + 2) * S, 0xEE);
  if (lanes == -1) {  // the one-thread-per-allele variant, with a stride as on the device
    const int stride = 5;
    std::vector<double> t0((size_t)S * stride, 0.0), t1((size_t)S * stride, 0.0);
    const HmmModelScan scan0 = hmm_model_scan(motifs, moff, nm);
    if (scan0.S <= HMM_THREAD_S)  // as on the device: the whole model in registers
      hmm_viterbi_thread(hmm_model_pack(scan0, jt.off.data()), c, jt.off.data(), jt.lp.data(), allele, L, t0.data() + 2, t1.data() + 2,
                         stride, bp.data());
    else
      hmm_viterbi_thread(scan0, c, jt.off.data(), jt.lp.data(), allele, L, t0.data() + 2, t1.data() + 2, stride, bp.data());
  } else if (lanes == 0) {
    hmm_viterbi(g, model, c, jt.lp.data(), allele, L, sc0.data(), sc1.data(), bp.data());
  } else {
    trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &lg) {
      hmm_viterbi(lg, model, c, jt.lp.data(), allele, L, sc0.data(), sc1.data(), bp.data());
    });
  }
  // counting walk, then writing walk (as the device does)
  std::vector<uint32_t> mc_tmp(nm + 1, 0);
  std::vector<uint32_t> rev(path_cap ? path_cap : 1);
  uint64_t plen = 0;
  // the walks use the table-free model, as the device's one-thread-per-allele kernels do
  const HmmModelScan scan = hmm_model_scan(motifs, moff, nm);
  if (scan.S != S) return -404;
  HmmAnnot a = hmm_annotate(scan, allele, L, bp.data(), 6, mc_tmp.data(), nullptr, 0, rev.data(),
                            path_cap, 0, &plen);
  if (a.status < 0) return -402;
  if (a.n_spans > span_cap) return -2;
  // the writing walk through the register-packed model where the device uses it (small models, thread variant)
  HmmAnnot b2 = (lanes == -1 && scan.S <= HMM_THREAD_S)
                    ? hmm_annotate(hmm_model_pack(scan), allele, L, bp.data(), 6, mc, spans, a.n_spans, nullptr, 0, 0, nullptr)
                    : hmm_annotate(scan, allele, L, bp.data(), 6, mc, spans, a.n_spans, nullptr, 0, 0, nullptr);
  if (b2.n_spans != a.n_spans) return -403;
  *purity = a.purity;
  if (path_len) *path_len = plen;
  if (path_out) {
    const uint64_t n = plen < path_cap ? plen : path_cap;
    for (uint64_t i = 0; i < n; i++) path_out[i] = rev[n - 1 - i];
  }
  return (long)a.n_spans;
}

}  // extern "C"
