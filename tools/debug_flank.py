"""GPU debugging aid: phase A on a synthetic workload against the oracle, printing mismatching pairs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trgt_b200
from harness import workload
from oracle import oracle as orc

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
w = workload.generate(n_loci, 30)
exp_spans, exp_hits = orc.flank_batch(w.left, w.right, w.reads, w.locus_read_off, n_threads=os.cpu_count())
eng = trgt_b200.Engine(0)
for budget in (20, 0):
    eng.set_flank_band_budget(budget)
    spans, hits = eng.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off)
    bad = np.nonzero((hits["via"] != exp_hits["via"]) | (hits["matches"] != exp_hits["matches"]) |
                     (hits["start"] != exp_hits["start"]) | (hits["end"] != exp_hits["end"]))[0]
    print(f"budget {budget}: {bad.size} mismatching (read, side) pairs of {hits.size}; via histogram {np.bincount(hits['via'], minlength=4)}"
          f" expected {np.bincount(exp_hits['via'], minlength=4)}")
    for i in bad[:12]:
        r, side = i // 2, i % 2
        off = int(w.reads.offsets[r]); T = int(w.reads.offsets[r + 1]) - off
        l = r // 30
        poff = int(w.left.offsets[l])
        print("  pair", i, "read", r, "side", side, "T", T, "read_off%16", off % 16, "piece_off%16", poff % 16,
              "got", hits[i], "exp", exp_hits[i])
eng.close()
