"""One pass of the hot path over a chunk of loci, phase-structured as the reference's host would
drive it: A = flank spans for every read, host genotype glue, B = consensus
alignments, C = motif HMM annotation.  Used by bench.py, __graft_entry__.smoke() and the tests.

    HotPath.run_e2e()        host buffers in, host buffers out, through the one-shot C-ABI calls
    HotPath.run_resident()   the same three phases on batches already resident in HBM
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from trgt_b200.engine import AnnotationBatch, CigarBatch, Engine, HIT_DTYPE, SPAN_DTYPE
import threading
import time

from .workload import GenotypeGlue, GlueContext, Workload, genotype_glue


@dataclass
class HotPathResult:
    spans: np.ndarray                 # SPAN_DTYPE [n_reads]
    hits: Optional[np.ndarray]        # HIT_DTYPE [2*n_reads]
    glue: GenotypeGlue
    cigars: CigarBatch
    annotations: AnnotationBatch


class HotPath:
    def __init__(self, engine: Engine, w: Workload, want_hits: bool = False, pinned_outputs: bool = True,
                 glue_threads: int = 0, use_seq4: bool = False, align_by_index: bool = False):
        self.eng = engine
        self.w = w
        # end to end, phase B names backbones and members by read index (trgt_align_trs: the device gathers the repeat
        # sequences it cut itself) instead of sending their bases up again; same CIGARs
        self.align_by_index = align_by_index
        # end-to-end input of phase A: the reads as BAM 4-bit bases (w.reads4) instead of ASCII
        self.use_seq4 = bool(use_seq4 and w.reads4 is not None)
        self.want_hits = want_hits
        self.glue_threads = glue_threads  # 0: all cores
        self.timing = {"flank": 0.0, "glue": 0.0, "align": 0.0, "hmm": 0.0}
        n = w.n_reads
        if pinned_outputs:
            self._spans = engine.pinned_array(max(1, n) * SPAN_DTYPE.itemsize)[:n * SPAN_DTYPE.itemsize].view(SPAN_DTYPE)
            self._hits = (engine.pinned_array(max(1, 2 * n) * HIT_DTYPE.itemsize)[:2 * n * HIT_DTYPE.itemsize]
                          .view(HIT_DTYPE)) if want_hits else None
        else:
            self._spans = np.zeros(n, dtype=SPAN_DTYPE)
            self._hits = np.zeros(2 * n, dtype=HIT_DTYPE) if want_hits else None
        self._fb = self._ab = self._hb = None
        self.glue: Optional[GenotypeGlue] = None
        self._glue_ctx = GlueContext(engine.lib) if pinned_outputs else None  # pinned, reused every pass

    # -- bytes that cross PCIe in one e2e step ------------------------------------------------
    def h2d_bytes(self, glue: GenotypeGlue) -> int:
        w = self.w
        n = 0
        if self.use_seq4:
            n += w.reads4.data.nbytes + w.reads4.starts.nbytes + w.reads4.lengths.nbytes
        else:
            n += w.reads.data.nbytes + w.reads.offsets.nbytes
        for s in (w.left, w.right, w.motifs):
            n += s.data.nbytes + s.offsets.nbytes
        if self.align_by_index and self.use_seq4:   # only the read indices: lengths, offsets and bytes are produced on the device
            n += 4 * (len(glue.backbones) + len(glue.seqs))
        else:
            for s in (glue.backbones, glue.seqs):
                n += s.data.nbytes + s.offsets.nbytes
        n += glue.backbones.data.nbytes + glue.backbones.offsets.nbytes  # alleles = backbones, sent again for phase C
        n += w.locus_read_off.nbytes + glue.group_seq_off.nbytes + w.locus_motif_off.nbytes + glue.group_locus.nbytes
        n += 2 * 8 * (len(glue.backbones) + 1)  # bp / mc offsets of the HMM batch
        return int(n)

    def d2h_bytes(self, res: HotPathResult) -> int:
        n = res.spans.nbytes + (res.hits.nbytes if res.hits is not None else 0)
        if self.use_seq4:  # trgt_flank_trs: offsets + bytes of the repeat sequences
            n += 8 * (self.w.n_reads + 1) + int(res.glue.seqs.data.nbytes)
        c, a = res.cigars, res.annotations
        n += c.offsets.nbytes + c.words.nbytes + c.scores.nbytes + c.status.nbytes
        n += a.motif_counts.nbytes + a.span_offsets.nbytes + a.spans.nbytes + a.purity.nbytes + a.status.nbytes
        return int(n)

    # -- end to end through the one-shot C-ABI calls ------------------------------------------
    def run_e2e(self, copy: bool = False, eng=None, uploaded=None) -> HotPathResult:
        """copy=False: results alias pinned buffers (the engine's, this object's) until the next pass.
        eng: the engine to call (default: the one this object was built on; its pinned buffers serve any engine of
        the process); uploaded: called once the phase-A call -- the one that moves the reads over PCIe -- has returned."""
        w = self.w
        eng = eng or self.eng
        t0 = time.perf_counter()
        try:
            if self.use_seq4:
                spans, hits = eng.flank_spans_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring,
                                                   w.min_flank_id_frac, want_hits=self.want_hits,
                                                   spans_out=self._spans, hits_out=self._hits)
            else:
                spans, hits = eng.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring,
                                                     w.min_flank_id_frac, want_hits=self.want_hits,
                                                     spans_out=self._spans, hits_out=self._hits)
        finally:
            if uploaded is not None:
                uploaded()
        trs = eng.flank_trs() if self.use_seq4 else None  # the host holds no ASCII reads to cut them from
        t1 = time.perf_counter()
        glue = genotype_glue(w, spans, threads=self.glue_threads, ctx=self._glue_ctx, trs=trs)
        t2 = time.perf_counter()
        if self.align_by_index and self.use_seq4:
            cigars = eng.align_trs(None, glue.seq_read[glue.group_seq_off[:-1]], glue.seq_read, glue.group_seq_off, copy=copy)
        else:
            cigars = eng.align_packed(glue.backbones, glue.seqs, glue.group_seq_off, copy=copy)
        t3 = time.perf_counter()
        ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus, copy=copy)
        t4 = time.perf_counter()
        for k, v in (("flank", t1 - t0), ("glue", t2 - t1), ("align", t3 - t2), ("hmm", t4 - t3)):
            self.timing[k] += v
        return HotPathResult(spans, hits, glue, cigars, ann)

    # -- end to end, phase A's upload done by somebody else --------------------------------------
    def upload(self, eng):
        """phase A's inputs of this chunk into a resident batch (blocks until they have landed)"""
        w = self.w
        if self.use_seq4:
            return eng.flank_upload_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
        return eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)

    def run_uploaded(self, eng, fb, copy: bool = False) -> HotPathResult:
        """the rest of the end-to-end pass on a batch `upload` produced (frees it)"""
        w = self.w
        t0 = time.perf_counter()
        try:
            eng.flank_run(fb)
            spans, hits = eng.flank_download(fb, w.n_reads, want_hits=self.want_hits, spans_out=self._spans,
                                             hits_out=self._hits)
            trs = eng.flank_trs(fb) if self.use_seq4 else None
            t1 = time.perf_counter()
            glue = genotype_glue(w, spans, threads=self.glue_threads, ctx=self._glue_ctx, trs=trs)
        finally:
            eng.flank_free(fb)  # (the repeat sequences were views of the batch's pinned buffers: the glue has copied them)
        t2 = time.perf_counter()
        cigars = eng.align_packed(glue.backbones, glue.seqs, glue.group_seq_off, copy=copy)
        t3 = time.perf_counter()
        ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus, copy=copy)
        t4 = time.perf_counter()
        for k, v in (("flank", t1 - t0), ("glue", t2 - t1), ("align", t3 - t2), ("hmm", t4 - t3)):
            self.timing[k] += v
        return HotPathResult(spans, hits, glue, cigars, ann)

    # -- resident batches ---------------------------------------------------------------------
    def prepare_resident(self) -> GenotypeGlue:
        """Upload phase A's inputs, run it once, derive phases B/C inputs on the host and upload them."""
        w, eng = self.w, self.eng
        self.free_resident()
        self._fb = eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
        eng.flank_run(self._fb)
        spans, _ = eng.flank_download(self._fb, w.n_reads, want_hits=False)
        self.glue = genotype_glue(w, spans)
        self._ab = eng.align_upload(self.glue.backbones, self.glue.seqs, self.glue.group_seq_off)
        self._hb = eng.hmm_upload(w.motifs, w.locus_motif_off, self.glue.backbones, self.glue.group_locus)
        return self.glue

    def run_resident(self, sync: bool = True):
        eng = self.eng
        eng.flank_run(self._fb)
        eng.align_run(self._ab)
        eng.hmm_run(self._hb)
        if sync:
            eng.sync()

    def download_resident(self) -> HotPathResult:
        eng = self.eng
        spans, hits = eng.flank_download(self._fb, self.w.n_reads, want_hits=self.want_hits)
        return HotPathResult(spans, hits, self.glue, eng.align_download(self._ab), eng.hmm_download(self._hb))

    def fallback_counts(self):
        return self.eng.flank_fallback_counts(self._fb) if self._fb else (0, 0, 0)

    def n_wfa(self) -> int:
        return self.eng.flank_n_wfa(self._fb) if self._fb else 0

    def free_resident(self):
        if self._fb:
            self.eng.flank_free(self._fb)
        if self._ab:
            self.eng.align_free(self._ab)
        if self._hb:
            self.eng.hmm_free(self._hb)
        self._fb = self._ab = self._hb = None


class ChunkedHotPath:
    """The end-to-end pass over a shard in chunks of loci, the way the reference's host drives its
    worker pool: each host thread owns one engine (its own CUDA streams and pinned buffers) and
    runs phases A, glue, B, C per chunk through the blocking C ABI, so one chunk's PCIe transfers
    overlap another chunk's kernels and host glue."""

    def __init__(self, engines, w: Workload, chunk_loci: int = 16384, glue_threads: int = 0, use_seq4: bool = False,
                 upload_slots: int = 2, uploaders: int = 0, max_inflight: int = 4, guided: bool = False,
                 min_chunk_loci: int = 1500, align_by_index: bool = False):
        """upload_slots: how many chunks may be inside their phase-A call (the one that moves the reads over PCIe)
        at a time; 0 = no limit and chunk i statically on thread i mod threads.  Without a limit all threads
        upload together, then all compute together while the link idles (measured: 52 ms per 125 k-locus shard at
        8 threads where the copies alone take 40); with one or two slots the chunks go up one after the other, in
        order, and every other phase of a chunk overlaps the uploads of the next ones."""
        self.engines = list(engines)
        self.upload_slots = upload_slots
        # uploaders > 0: that many of the host threads do nothing but phase A's uploads (trgt_flank_upload*), chunk after
        # chunk in order, at most max_inflight chunks ahead; the others take the uploaded batches through
        # trgt_flank_run / _download / _trs, the glue and phases B and C.  The link never waits for a kernel or the host.
        self.uploaders = min(uploaders, max(0, len(self.engines) - 1))
        self.max_inflight = max_inflight
        self.upload_s = 0.0
        self.w = w
        self.guided = guided
        if guided:
            # guided self-scheduling: chunks of chunk_loci while plenty is left, then ever smaller ones (what is left /
            # threads, at least min_chunk_loci), handed to whichever thread is free.  Phase A saturates the link for as
            # long as there are reads to send; what is left after the last upload -- glue, phases B and C of the chunks
            # still in flight -- is proportional to the size of the last chunks (measured: 51.6 ms with 16 equal chunks,
            # 40.1 ms for their phase A alone).
            self.bounds = []
            l0 = 0
            while l0 < w.n_loci:
                left = w.n_loci - l0
                size = max(min_chunk_loci, min(chunk_loci, -(-left // max(1, len(self.engines)))))
                self.bounds.append((l0, min(l0 + size, w.n_loci)))
                l0 += size
        else:
            self.bounds = [(l0, min(l0 + chunk_loci, w.n_loci)) for l0 in range(0, w.n_loci, chunk_loci)]
        # every chunk's arrays (the rebased offsets are fresh copies) in pinned memory, as a host that packs into
        # trgt_host_alloc buffers has them
        self.paths = [HotPath(self.engines[i % len(self.engines)],
                              w.slice(l0, l1).pinned(self.engines[i % len(self.engines)].pinned_array),
                              glue_threads=glue_threads, use_seq4=use_seq4, align_by_index=align_by_index)
                      for i, (l0, l1) in enumerate(self.bounds)]

    def timing(self) -> dict:
        """Seconds spent per phase, summed over chunks and host threads since construction."""
        out = {"flank": 0.0, "glue": 0.0, "align": 0.0, "hmm": 0.0}
        for p in self.paths:
            for k in out:
                out[k] += p.timing[k]
        return out

    def _run_pipelined(self):
        n, engines, U = len(self.paths), self.engines, self.uploaders
        ready = [threading.Event() for _ in range(n)]
        fbs = [None] * n
        results = [None] * n
        errors = []
        inflight = threading.Semaphore(self.max_inflight)
        lock = threading.Lock()
        nxt = {"up": 0, "wk": 0}

        def take(kind):
            with lock:
                i = nxt[kind]
                nxt[kind] += 1
            return i

        def uploader(k):
            try:
                while True:
                    inflight.acquire()
                    i = take("up")
                    if i >= n or errors:
                        inflight.release()
                        return
                    t0 = time.perf_counter()
                    fbs[i] = self.paths[i].upload(engines[k])
                    self.upload_s += time.perf_counter() - t0
                    ready[i].set()
            except Exception as exc:
                errors.append(exc)
                for ev in ready:
                    ev.set()

        def worker(k):
            try:
                while True:
                    i = take("wk")
                    if i >= n:
                        return
                    ready[i].wait()
                    if errors:
                        return
                    results[i] = self.paths[i].run_uploaded(engines[k], fbs[i], copy=True)
                    fbs[i] = None
                    inflight.release()
            except Exception as exc:
                errors.append(exc)
                inflight.release()

        threads = [threading.Thread(target=uploader, args=(k,)) for k in range(U)]
        threads += [threading.Thread(target=worker, args=(k,)) for k in range(U, len(engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def run_e2e(self):
        if self.uploaders > 0:
            return self._run_pipelined()
        n_eng = len(self.engines)
        results = [None] * len(self.paths)
        errors = []

        gate = threading.Semaphore(self.upload_slots) if self.upload_slots > 0 else None
        take = threading.Lock()
        nxt = [0]

        def work(k):
            try:
                if gate is None and not self.guided:
                    for i in range(k, len(self.paths), n_eng):
                        results[i] = self.paths[i].run_e2e(copy=True)
                    return
                if gate is None:   # chunks in order to whichever thread is free
                    while True:
                        with take:
                            i = nxt[0]
                            nxt[0] += 1
                        if i >= len(self.paths):
                            return
                        results[i] = self.paths[i].run_e2e(copy=True, eng=self.engines[k])
                while True:
                    gate.acquire()          # a slot on the link first, then the next chunk: chunks go up in order
                    with take:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= len(self.paths):
                        gate.release()
                        return
                    results[i] = self.paths[i].run_e2e(copy=True, eng=self.engines[k], uploaded=gate.release)
            except Exception as exc:  # surfaced to the caller below
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(n_eng)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def h2d_bytes(self, results) -> int:
        return sum(p.h2d_bytes(r.glue) for p, r in zip(self.paths, results))

    def d2h_bytes(self, results) -> int:
        return sum(p.d2h_bytes(r) for p, r in zip(self.paths, results))


def concat_results(results):
    """(spans, cigar words, cigar scores, motif counts, HMM spans, purity) of chunk results, in locus order."""
    return (np.concatenate([r.spans for r in results]), np.concatenate([r.cigars.words for r in results]),
            np.concatenate([r.cigars.scores for r in results]),
            np.concatenate([r.annotations.motif_counts for r in results]),
            np.concatenate([r.annotations.spans for r in results]),
            np.concatenate([r.annotations.purity for r in results]))


def oracle_pass(orc, w: Workload, n_threads: int):
    """The same pass on the CPU oracle (TEST / BASELINE USE ONLY: callers are tests/, smoke() and
    bench.py's cpu_baseline and --impl reference legs).  `orc` is the oracle.oracle module."""
    spans, _ = orc.flank_batch(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                               n_threads=n_threads, want_hits=False)
    glue = genotype_glue(w, spans)
    offs, words, scores = orc.align_batch(glue.backbones, glue.seqs, glue.group_seq_off, n_threads=n_threads)
    hmm = orc.hmm_batch(w.motifs, w.locus_motif_off, glue.backbones, glue.group_locus, n_threads=n_threads)
    return spans, glue, (offs, words, scores), hmm


def compare_with_oracle(res: HotPathResult, ref) -> None:
    """Bit-exact comparison of a HotPathResult with oracle_pass output (raises AssertionError)."""
    spans, glue, (offs, words, scores), (mc_off, mc, span_off, hspans, purity, status) = ref
    assert np.array_equal(res.spans["found"], spans["found"]), "span found flags differ"
    f = spans["found"] != 0
    assert np.array_equal(res.spans["start"][f], spans["start"][f]) and np.array_equal(res.spans["end"][f], spans["end"][f]), "spans differ"
    assert np.array_equal(res.glue.seq_read, glue.seq_read) and np.array_equal(res.glue.group_locus, glue.group_locus)
    assert np.array_equal(res.cigars.offsets, offs) and np.array_equal(res.cigars.words, words), "CIGARs differ"
    assert np.array_equal(res.cigars.scores, scores), "alignment scores differ"
    assert not res.cigars.status.any() and not res.annotations.status.any() and not status.any()
    a = res.annotations
    assert np.array_equal(a.motif_counts, mc), "MC differs"
    assert np.array_equal(a.span_offsets, span_off) and np.array_equal(a.spans, hspans), "MS differs"
    assert np.array_equal(a.purity, purity, equal_nan=True), "AP differs"
