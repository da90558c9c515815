// emul_consensus.cpp -- TEST INFRASTRUCTURE ONLY.  Serial / lock-step multi-lane run of
// trgt_b200/csrc/consensus_core.h.  Never linked into libtrgt_b200.so.
#include <string.h>

#include <vector>

#include "../../trgt_b200/csrc/consensus_core.h"
#include "lanes.h"

using namespace trgt;

extern "C" {

// one group: returns the consensus length (or <0), bytes in out (cap)
long emu_consensus(int B, const uint8_t *seqs, const uint64_t *seq_off, uint32_t n, const uint32_t *words,
                   const unsigned long long *word_off, uint8_t *out, uint64_t cap, int lanes) {
  ConsGroup gr;
  gr.B = B; gr.s0 = 0; gr.n = n; gr.seqs = seqs; gr.seq_off = seq_off; gr.words = words; gr.word_off = word_off;
  const uint32_t rec_cap = (uint32_t)(word_off[n] + 1);
  std::vector<int> counts((size_t)6 * B + 1);
  std::vector<ConsRec> recs(rec_cap);
  int shared[2];
  long long len1 = -9, len2 = -9;
  std::vector<uint8_t> buf;
  auto pass = [&](uint8_t *o, long long *res) {
    if (lanes == 0) { SerialGroup g; *res = consensus_vote(g, gr, counts.data(), recs.data(), rec_cap, shared, o); }
    else trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
      const long long r = consensus_vote(g, gr, counts.data(), recs.data(), rec_cap, shared, o);
      if (g.lane() == 0) *res = r;
    });
  };
  pass(nullptr, &len1);
  if (len1 < 0) return (long)len1;
  buf.assign((size_t)len1 + 1, 0);
  pass(buf.data(), &len2);
  if (len2 != len1) return -77;
  if ((uint64_t)len1 > cap) return -2;
  memcpy(out, buf.data(), (size_t)len1);
  return (long)len1;
}

}  // extern "C"
