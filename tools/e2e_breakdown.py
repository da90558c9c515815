"""GPU aid: wall-clock breakdown of the end-to-end pass over ONE chunk (host buffers through the C ABI, one
engine, one host thread): where a host thread of bench.py's e2e driver spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trgt_b200
from harness import workload
from harness.pipeline import HotPath
from harness.workload import GlueContext, genotype_glue

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7813
eng = trgt_b200.Engine(0)
w = workload.generate(n, 30, alloc_reads=eng.pinned_array)
w.pack_seq4(alloc=eng.pinned_array)
hp = HotPath(eng, w, use_seq4=True)
ctx = GlueContext(eng.lib)
for it in range(5):
    t = [time.perf_counter()]
    spans, _ = eng.flank_spans_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                                    want_hits=False, spans_out=hp._spans)
    t.append(time.perf_counter())
    trs = eng.flank_trs()
    t.append(time.perf_counter())
    glue = genotype_glue(w, spans, threads=2, ctx=ctx, trs=trs)
    t.append(time.perf_counter())
    cig = eng.align_packed(glue.backbones, glue.seqs, glue.group_seq_off, copy=True)
    t.append(time.perf_counter())
    ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus, copy=True)
    t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"iter {it}: flank_spans_seq4 {d[0]:.2f}  flank_trs {d[1]:.2f}  glue {d[2]:.2f}  align {d[3]:.2f}  hmm {d[4]:.2f}  "
          f"total {sum(d):.2f} ms")
nb = w.reads4.data.nbytes
print(f"chunk: {n} loci, {w.n_reads} reads, {nb/1e6:.1f} MB packed ({nb/55.5e9*1e3:.2f} ms at 55.5 GB/s)")
eng.reset_stats(); eng.set_profiling(True)
eng.flank_spans_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac, want_hits=False, spans_out=hp._spans)
eng.flank_trs(); eng.align_packed(glue.backbones, glue.seqs, glue.group_seq_off); eng.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus)
print("kernel ms:", {k: round(v[1], 3) for k, v in eng.kernel_stats().items()}, "sum", round(sum(v[1] for v in eng.kernel_stats().values()), 2))
