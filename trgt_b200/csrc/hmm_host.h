// hmm_host.h -- host-side constants of the motif HMM.
//
// Every ln() the model needs is taken here with the host libm, from the same f64 expressions
// the reference evaluates (src/hmm/builder.rs:84-93, hmm_model.rs:44-52), and shipped to the
// device; the device never calls log(), so transition/emission terms are bit-identical to the
// reference's.
#pragma once
#include <math.h>

#include <vector>

#include "hmm_core.h"

namespace trgt {

inline HmmConsts hmm_make_consts() {
  HmmConsts c;
  const double match_prob = 0.90;
  const double ins_to_ins = 0.25;
  const double match_to_indel = (1.00 - match_prob) / 2.00;
  const double del_to_match = 0.50;
  c.lp_match = log(match_prob);
  c.lp_ins_exit = log(1.0 - ins_to_ins);
  c.lp_half = log(del_to_match);
  c.lp_ins_loop = log(ins_to_ins);
  c.lp_indel_open = log(match_to_indel);
  c.lp_end = log(0.10);
  c.lp_one = log(1.00);
  c.em_hi = log(0.90);
  c.em_lo = log(0.03);
  c.em_quarter = log(0.25);
  c.em_one = log(1.00);
  return c;
}

// Jump-in table: for motif length n, entry i (1 <= i < n) is ln(seed * (n - i)) with
// seed = 2 (1 - 0.9) / (n (n - 1))          builder.rs:93,100-101,110-111
// Layout: mm_off[n] is the offset of length n's n entries in mm_lp (entry 0 unused).
struct HmmJumpTable {
  std::vector<uint32_t> off;  // [max_len + 1]
  std::vector<double> lp;
  void ensure(int max_len) {
    if ((int)off.size() > max_len) return;
    const double match_prob = 0.90;
    int n0 = (int)off.size();
    off.resize((size_t)max_len + 1, 0);
    for (int n = n0; n <= max_len; n++) {
      off[n] = (uint32_t)lp.size();
      if (n == 0) continue;
      const double seed = 2.00 * (1.00 - match_prob) / (double)((size_t)n * (size_t)(n - 1));
      for (int i = 0; i < n; i++) {
        const double p = seed * (double)(n - i);
        lp.push_back(i == 0 ? 0.0 : log(p));
      }
    }
  }
};

}  // namespace trgt
