"""GPU aid: is the end-to-end driver held back by Python?  Times ChunkedHotPath.run_e2e (8 host threads, 16 chunks of
the config-4 shard) as the bench does, and again with the result copies switched off (copy=False: results alias the
engines' pinned buffers, so the numbers are only good for timing) and with the per-chunk host glue replaced by a
precomputed one (the phases then only move data and launch kernels)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trgt_b200
from harness import workload, pipeline
from harness.pipeline import ChunkedHotPath, HotPath

n_loci, threads = 125000, 8
engines = [trgt_b200.Engine(0) for _ in range(threads)]
w = workload.generate(n_loci, 30, alloc_reads=engines[0].pinned_array)
w.pack_seq4(alloc=engines[0].pinned_array)
chp = ChunkedHotPath(engines, w, chunk_loci=-(-n_loci // (2 * threads)), glue_threads=2, use_seq4=True, upload_slots=0)

def timeit(label, n=5):
    for _ in range(2):
        chp.run_e2e()
    t0 = time.perf_counter()
    for _ in range(n):
        chp.run_e2e()
    print(f"{label}: {(time.perf_counter() - t0) / n * 1e3:.2f} ms per step", flush=True)

timeit("as benched (copy=True)")
orig = HotPath.run_e2e
HotPath.run_e2e = lambda self, copy=False, eng=None, uploaded=None: orig(self, copy=False, eng=eng, uploaded=uploaded)
timeit("copy=False")
# flank phase only (uploads + kernels + spans back), no glue / align / hmm
def flank_only(self, copy=False, eng=None, uploaded=None):
    w, e = self.w, (eng or self.eng)
    e.flank_spans_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac, want_hits=False,
                       spans_out=self._spans, hits_out=None)
    e.flank_trs()
    return None
HotPath.run_e2e = flank_only
timeit("phase A only (flank_spans_seq4 + flank_trs)")

from harness.workload import genotype_glue
def upto(stage):
    def f(self, copy=False, eng=None, uploaded=None):
        w, e = self.w, (eng or self.eng)
        spans, _ = e.flank_spans_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                                      want_hits=False, spans_out=self._spans, hits_out=None)
        trs = e.flank_trs()
        if stage == "A":
            return None
        glue = genotype_glue(w, spans, threads=self.glue_threads, ctx=self._glue_ctx, trs=trs)
        if stage == "glue":
            return None
        e.align_packed(glue.backbones, glue.seqs, glue.group_seq_off, copy=False)
        if stage == "align":
            return None
        e.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus, copy=False)
        return None
    return f
for st in ("glue", "align", "hmm"):
    HotPath.run_e2e = upto(st)
    timeit("phases up to " + st)
# the same with the glue done by one thread per worker
for hp in chp.paths:
    hp.glue_threads = 1
HotPath.run_e2e = upto("hmm")
timeit("all phases, glue_threads=1")
