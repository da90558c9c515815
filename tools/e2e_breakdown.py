"""GPU aid: wall-clock breakdown of one end-to-end pass (host buffers through the C ABI)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trgt_b200
from trgt_b200 import workload
from trgt_b200.pipeline import HotPath
from trgt_b200.workload import genotype_glue

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
eng = trgt_b200.Engine(0)
w = workload.generate(n, 30, alloc_reads=eng.pinned_array)
hp = HotPath(eng, w)
for it in range(3):
    t = [time.perf_counter()]
    spans, hits = eng.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                                         want_hits=False, spans_out=hp._spans)
    t.append(time.perf_counter())
    glue = genotype_glue(w, spans)
    t.append(time.perf_counter())
    cig = eng.align_packed(glue.backbones, glue.seqs, glue.group_seq_off)
    t.append(time.perf_counter())
    ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus)
    t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"iter {it}: flank {d[0]:.1f} ms  glue {d[1]:.1f} ms  align {d[2]:.1f} ms  hmm {d[3]:.1f} ms  total {sum(d):.1f} ms")
# raw H2D bandwidth of the pinned read buffer
import ctypes as C
b = eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off)
t0 = time.perf_counter(); b2 = eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off); dt = time.perf_counter() - t0
print(f"flank_upload of {w.reads.data.nbytes/1e9:.2f} GB: {dt*1e3:.1f} ms = {w.reads.data.nbytes/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); eng.flank_run(b2); eng.sync(); print(f"flank_run {1e3*(time.perf_counter()-t0):.1f} ms")
t0 = time.perf_counter(); eng.flank_download(b2, w.n_reads, want_hits=False); print(f"flank_download {1e3*(time.perf_counter()-t0):.1f} ms")
