// emul_clip.cpp -- TEST INFRASTRUCTURE ONLY.  Serial / lock-step multi-lane run of
// trgt_b200/csrc/clip_core.h.  Never linked into libtrgt_b200.so.
#include <string.h>

#include <vector>

#include "../../trgt_b200/csrc/clip_core.h"
#include "../../trgt_b200/csrc/vcf_core.h"
#include "lanes.h"

using namespace trgt;

extern "C" {

void emu_clip_cigar(const uint32_t *ops, uint32_t n_ops, long long ref_start, long long region_start,
                    long long region_end, trgt_clip_t *out) {
  *out = clip_cigar_one(ops, n_ops, ref_start, region_start, region_end);
}

// decode n reads into out (CSR offsets out_off[n+1] given); data must carry 16 bytes of padding on both
// sides, as the engine's device buffer does; out must be 16-byte aligned
void emu_seq4_unpack(const uint8_t *data, const uint64_t *starts, const uint32_t *lengths, const uint64_t *out_off,
                     uint32_t n, uint8_t *out, int lanes) {
  for (uint32_t r = 0; r < n; r++) {
    if (lanes == 0) { SerialGroup g; seq4_unpack_read(g, data, starts[r], lengths[r], out, out_off[r]); }
    else trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
      seq4_unpack_read(g, data, starts[r], lengths[r], out, out_off[r]);
    });
  }
}

// format!("{:.6}", v) of vcf_core.h into out (cap >= 40); returns the length, -1 if the value is not taken
int emu_vcf_fixed6(double v, char *out) {
  VcfWriter w;
  w.out = (uint8_t *)out; w.n = 0;
  if (!vcf_put_fixed6(w, v)) return -1;
  return (int)w.n;
}

// clip_bases for the BAMlet (bamlet_clip_read) by `lanes` lanes (0 = one serial lane)
void emu_bamlet_clip(const uint8_t *bases, uint32_t len, int found, uint32_t span_start, uint32_t span_end,
                     uint32_t flank_len, const uint32_t *ops, uint32_t n_ops, long long ref_pos, int lanes,
                     trgt_bamlet_clip_t *out) {
  if (lanes == 0) {
    SerialGroup g;
    *out = bamlet_clip_read(g, bases, len, found != 0, span_start, span_end, flank_len, ops, n_ops, ref_pos);
  } else {
    trgt_test::run_lanes(lanes, [&](const trgt_test::LaneGroup &g) {
      const trgt_bamlet_clip_t c = bamlet_clip_read(g, bases, len, found != 0, span_start, span_end, flank_len, ops, n_ops, ref_pos);
      if (g.lane() == 0) *out = c;
    });
  }
}

}  // extern "C"
