// emul_cluster.cpp -- TEST INFRASTRUCTURE ONLY.  Instantiates trgt_b200/csrc/cluster_core.h with the one-lane
// SerialGroup or N lock-step host lanes so that it can be compared with the oracle without a GPU.
#include <string.h>

#include <vector>

#include "../../trgt_b200/csrc/cluster_core.h"
#include "lanes.h"

using namespace trgt;

extern "C" {

// dists (condensed, n(n-1)/2) is modified in place as the reference's linkage call does.
// sel[n], central[2]; returns the number of groups.
int emu_cluster_locus(double *dists, uint32_t n, int32_t *sel, uint32_t *central, int lanes) {
  std::vector<uint64_t> buf(cl_ws_bytes(n) / 8 + 2);
  const ClusterWs w = cl_carve(buf.data(), n);
  int ng = 0;
  if (lanes <= 0) {
    SerialGroup g;
    ng = cl_cluster_locus(g, dists, n, w, sel, central);
  } else {
    trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &lg) {
      const int r = cl_cluster_locus(lg, dists, n, w, sel, central);
      if (lg.lane() == 0) ng = r;
    });
  }
  return ng;
}

}  // extern "C"
