"""GPU aid: who does what when in one end-to-end step (8 host threads, 16 chunks)?  Prints one row per host thread,
one character per 0.5 ms: F = inside the phase-A call (upload + kernels + spans), t = repeat sequences, g = glue,
A = align call, H = HMM call, . = idle / finished."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trgt_b200
from harness import workload
from harness.pipeline import ChunkedHotPath, HotPath
from harness.workload import genotype_glue

n_loci, threads = 125000, 8
engines = [trgt_b200.Engine(0) for _ in range(threads)]
w = workload.generate(n_loci, 30, alloc_reads=engines[0].pinned_array)
w.pack_seq4(alloc=engines[0].pinned_array)
chp = ChunkedHotPath(engines, w, chunk_loci=-(-n_loci // (2 * threads)), glue_threads=2, use_seq4=True, upload_slots=0)
log = []
def traced(self, copy=False, eng=None, uploaded=None):
    w_, e = self.w, (eng or self.eng)
    tid = threading.get_ident()
    t = [time.perf_counter()]
    spans, _ = e.flank_spans_seq4(w_.left, w_.right, w_.reads4, w_.locus_read_off, w_.scoring, w_.min_flank_id_frac,
                                  want_hits=False, spans_out=self._spans, hits_out=None)
    t.append(time.perf_counter())
    trs = e.flank_trs()
    t.append(time.perf_counter())
    glue = genotype_glue(w_, spans, threads=self.glue_threads, ctx=self._glue_ctx, trs=trs)
    t.append(time.perf_counter())
    e.align_packed(glue.backbones, glue.seqs, glue.group_seq_off, copy=True)
    t.append(time.perf_counter())
    e.hmm_label_packed(w_.motifs, w_.locus_motif_off, glue.backbones, glue.group_locus, copy=True)
    t.append(time.perf_counter())
    log.append((tid, t))
    return None
HotPath.run_e2e = traced
for _ in range(3):
    chp.run_e2e()
log.clear()
t_start = time.perf_counter()
chp.run_e2e()
t_end = time.perf_counter()
print("step: %.2f ms" % ((t_end - t_start) * 1e3))
tids = sorted(set(t for t, _ in log))
cell = 0.5e-3
n = int((t_end - t_start) / cell) + 1
for tid in tids:
    row = ["."] * n
    for t, ts in log:
        if t != tid:
            continue
        for ch, a, b in zip("FtgAH", ts[:-1], ts[1:]):
            for i in range(int((a - t_start) / cell), min(n, int((b - t_start) / cell) + 1)):
                row[i] = ch
    print("".join(row))
