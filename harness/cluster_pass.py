"""The pass of the cluster genotyper (`--genotyper cluster`, BASELINE config 5: long expansions) over a chunk of
loci, phase-structured like harness/pipeline.py but with the reference's own grouping logic between the phases:

    A   flank spans of every read                                  find_tr_spans        span_locater.rs:32-68
        spanning reads with long enough flanks, by repeat length   get_spanning_reads   tr.rs:111-165
    B1  get_dist_matrix -> cluster() -> group1 / group2 -> central_read            genotype_cluster.rs:57-72
    B2  make_consensus of both groups (align + repair_consensus)                    genotype_cluster.rs:41-55
        outlier rule, allele order                                                  genotype_cluster.rs:84-152
    C   label_with_hmm on the two alleles                                           tr.rs:454-492

`engine_cluster_pass` keeps the reads resident in HBM from phase A on: B1 and B2 name the repeat sequences by
read index (trgt_cluster_trs / trgt_consensus_trs), so 5-50 kb repeat sequences never travel back to the host.
`oracle_cluster_pass` is the same pass on the CPU oracle (TEST / BASELINE USE ONLY).  The glue in between is
numpy on spans and index lists, shared by both.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from trgt_b200.engine import AnnotationBatch, PackedSeqs

from .workload import Workload


@dataclass
class ClusterPassResult:
    spans: np.ndarray             # SPAN_DTYPE [n_reads]
    sel_reads: np.ndarray         # uint32: spanning reads in genotyping order, locus after locus
    sel_off: np.ndarray           # uint32 [n_loci+1]
    group: np.ndarray             # int32 per selected read: 0 group1, 1 group2, 2 neither
    central: np.ndarray           # uint32 [n_loci, 2]
    alleles: PackedSeqs           # two per genotyped locus (shorter first), none for a locus without spanning reads
    allele_locus: np.ndarray      # uint32 per allele
    redone: np.ndarray            # loci that took the outlier rule (alternate split)
    annotations: Optional[AnnotationBatch] = None
    oracle_annotations: Optional[list] = None


def spanning_order(w: Workload, spans: np.ndarray, search_flank_len: int = 250, max_depth: int = 250):
    """get_spanning_reads (tr.rs:111-165) on the spans alone: reads with a span and >= search_flank_len bases on
    both sides of it, stable-sorted by repeat length within their locus (down-sampling to max_depth is never
    reached at the depths of the configs and is not restated) -> (read indices, per-locus offsets)"""
    read_len = np.diff(w.reads.offsets.astype(np.int64))
    start, end = spans["start"].astype(np.int64), spans["end"].astype(np.int64)
    ok = (spans["found"] != 0) & (start >= search_flank_len) & (read_len - end >= search_flank_len)
    read_locus = np.repeat(np.arange(w.n_loci, dtype=np.int64), np.diff(w.locus_read_off.astype(np.int64)))
    idx = np.nonzero(ok)[0]
    order = np.lexsort(((end - start)[idx], read_locus[idx]))   # stable: by locus, then by repeat length
    sel = idx[order].astype(np.uint32)
    counts = np.bincount(read_locus[sel], minlength=w.n_loci)
    assert counts.max(initial=0) <= max_depth
    off = np.zeros(w.n_loci + 1, dtype=np.uint32)
    np.cumsum(counts, out=off[1:])
    return sel, off


def consensus_groups(sel_reads, sel_off, group, central):
    """The make_consensus calls of genotype() (genotype_cluster.rs:57-72): per locus with spanning reads group1 and,
    when there is one, group2 -> (backbone read, member reads, group offsets, locus of each group)"""
    bb, members, goff, gl = [], [], [0], []
    for l in range(sel_off.size - 1):
        a, b = int(sel_off[l]), int(sel_off[l + 1])
        if b == a:
            continue
        for which in (0, 1):
            c = int(central[l, which])
            if c == 0xFFFFFFFF:
                continue
            m = sel_reads[a:b][group[a:b] == which]
            bb.append(int(sel_reads[a + c]))
            members.append(m)
            goff.append(goff[-1] + m.size)
            gl.append(l)
    return (np.array(bb, dtype=np.uint32), np.concatenate(members) if members else np.zeros(0, dtype=np.uint32),
            np.array(goff, dtype=np.uint32), np.array(gl, dtype=np.uint32))


def alternate_groups(sel_reads, sel_off, loci):
    """the homozygous redo of genotype() (:98-103): reads 0, 2, 4 .. and 1, 3, 5 .. of the locus, backbone = the
    first read of each (central_read of <= 2 reads; for larger groups the reference calls central_read again --
    the caller passes those through the clustering entry with the split fixed, see redo_central)"""
    bb, members, goff, gl = [], [], [0], []
    for l in loci:
        a, b = int(sel_off[l]), int(sel_off[l + 1])
        for which in (0, 1):
            m = sel_reads[a + which:b:2]
            bb.append(0)  # filled by the caller
            members.append(m)
            goff.append(goff[-1] + m.size)
            gl.append(l)
    return members, np.array(goff, dtype=np.uint32), np.array(gl, dtype=np.uint32)


def small_group_is_outlier(len1, len2, cov1, cov2) -> bool:
    """genotype_cluster.rs:86-93"""
    return abs(len1 - len2) < 100 and min(cov1, cov2) * 4 < max(cov1, cov2)


def _order_alleles(cons: PackedSeqs, gl: np.ndarray, n_loci: int):
    """two alleles per genotyped locus, the shorter first (:146-151); a single group fills both (:61-68)"""
    seqs, locus = [], []
    by = {}
    for g, l in enumerate(gl.tolist()):
        by.setdefault(l, []).append(cons.get(g))
    for l in sorted(by):
        a = by[l]
        if len(a) == 1:
            a = [a[0], a[0]]
        if len(a[0]) > len(a[1]):
            a = [a[1], a[0]]
        seqs += a
        locus += [l, l]
    return PackedSeqs.from_list(seqs), np.array(locus, dtype=np.uint32)


def _central_of_fixed_group(dmat_after_linkage, n, members):
    """central_read (:12-39) of a fixed group on the matrix the linkage left behind (host side of the redo)"""
    if len(members) <= 2:
        return members[0]
    sums = [0.0] * len(members)
    for i in range(len(members) - 1):
        for j in range(i + 1, len(members)):
            i1, i2 = members[i], members[j]
            v = dmat_after_linkage[n * i1 - i1 * (i1 + 3) // 2 + i2 - 1]
            sums[i] += v
            sums[j] += v
    return members[min(range(len(members)), key=lambda k: (sums[k], k))]


def engine_cluster_pass(eng, w: Workload, orc_for_redo=None, use_seq4: bool = False, resident_batch=None) -> ClusterPassResult:
    """phases A, B1, B2, C on the CUDA engine; the reads stay resident from phase A on (resident_batch: a flank
    batch the caller has uploaded and keeps)"""
    if resident_batch is not None:
        fb = resident_batch
    elif use_seq4 and w.reads4 is not None:
        fb = eng.flank_upload_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    else:
        fb = eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    try:
        eng.flank_run(fb)
        spans, _ = eng.flank_download(fb, w.n_reads, want_hits=False)
        sel, off = spanning_order(w, spans)
        group, central, _ = eng.cluster_trs(fb, sel, off)
        bb, members, goff, gl = consensus_groups(sel, off, group, central)
        cons, status = eng.consensus_trs(fb, bb, members, goff)
        assert not status.any(), "consensus failed"
        cons = PackedSeqs(cons.data.copy(), cons.offsets.copy())
        redone = _outlier_loci(cons, gl, goff)
        if redone.size:
            cons, gl = _redo(redone, sel, off, cons, gl, w, spans,
                             lambda b_, m_, g_: eng.consensus_trs(fb, b_, m_, g_)[0], orc_for_redo)
        alleles, allele_locus = _order_alleles(cons, gl, w.n_loci)
        ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, alleles, allele_locus)
        return ClusterPassResult(spans, sel, off, group, central, alleles, allele_locus, redone, annotations=ann)
    finally:
        if resident_batch is None:
            eng.flank_free(fb)


def _outlier_loci(cons: PackedSeqs, gl: np.ndarray, goff: np.ndarray) -> np.ndarray:
    lens = np.diff(cons.offsets.astype(np.int64))
    cov = np.diff(goff.astype(np.int64))
    out = []
    for g in range(gl.size - 1):
        if gl[g] == gl[g + 1] and small_group_is_outlier(int(lens[g]), int(lens[g + 1]), int(cov[g]), int(cov[g + 1])):
            out.append(int(gl[g]))
    return np.array(out, dtype=np.int64)


def _redo(redone, sel, off, cons, gl, w, spans, consensus_fn, orc):
    """genotype() :94-111 for the loci the outlier rule sends back: alternate split, central_read on the matrix the
    first linkage left behind (recomputed with the oracle's restatement: a handful of loci, host logic), consensus"""
    assert orc is not None, "the outlier redo needs the distance matrices (pass orc_for_redo)"
    members, goff2, gl2 = alternate_groups(sel, off, redone.tolist())
    bb2 = []
    for k, l in enumerate(gl2.tolist()):
        a, b = int(off[l]), int(off[l + 1])
        n = b - a
        trs = [w.reads.get(int(r))[int(spans[int(r)]["start"]):int(spans[int(r)]["end"])] for r in sel[a:b]]
        _, dm = orc.ward_linkage(orc.get_dist_matrix(trs), n) if n >= 3 else ([], None)
        pos = list(range(k % 2, n, 2))
        bb2.append(int(sel[a + (_central_of_fixed_group(dm, n, pos) if dm is not None else pos[0])]))
    cons2 = consensus_fn(np.array(bb2, dtype=np.uint32), np.concatenate(members), goff2)
    seqs, locus = [], []
    redo_set = set(redone.tolist())
    for g, l in enumerate(gl.tolist()):
        if l not in redo_set:
            seqs.append(cons.get(g))
            locus.append(l)
    for g, l in enumerate(gl2.tolist()):
        seqs.append(cons2.get(g))
        locus.append(l)
    order = np.argsort(np.array(locus), kind="stable")
    return PackedSeqs.from_list([seqs[i] for i in order]), np.array([locus[i] for i in order], dtype=np.uint32)


def oracle_cluster_pass(orc, w: Workload, n_threads: int = 1, annotate: bool = True) -> ClusterPassResult:
    """the same pass on the CPU oracle (TEST / BASELINE USE ONLY)"""
    spans, _ = orc.flank_batch(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                               n_threads=n_threads, want_hits=False)
    sel, off = spanning_order(w, spans)

    def tr(r):
        return w.reads.get(int(r))[int(spans[int(r)]["start"]):int(spans[int(r)]["end"])]

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max(1, n_threads))   # the oracle's C calls release the GIL: one task per locus / group
    group = np.zeros(sel.size, dtype=np.int32)
    central = np.full((w.n_loci, 2), 0xFFFFFFFF, dtype=np.uint32)

    def cluster_one(l):
        a, b = int(off[l]), int(off[l + 1])
        if b == a:
            return
        trs = [tr(r) for r in sel[a:b]]
        s, c, _ = orc.cluster_locus(orc.get_dist_matrix(trs) if b - a >= 2 else [], b - a)
        group[a:b] = s
        central[l] = [0xFFFFFFFF if x is None else x for x in c]

    list(pool.map(cluster_one, range(w.n_loci)))
    bb, members, goff, gl = consensus_groups(sel, off, group, central)

    def consensus_fn(b_, m_, g_):
        return PackedSeqs.from_list(list(pool.map(
            lambda g: orc.repair_consensus(tr(b_[g]), [tr(r) for r in m_[int(g_[g]):int(g_[g + 1])]]), range(len(b_)))))

    cons = consensus_fn(bb, members, goff)
    redone = _outlier_loci(cons, gl, goff)
    if redone.size:
        cons, gl = _redo(redone, sel, off, cons, gl, w, spans, consensus_fn, orc)
    alleles, allele_locus = _order_alleles(cons, gl, w.n_loci)
    ann = None
    if annotate:
        def annotate_one(i):
            h = orc.Hmm([orc.replace_invalid_bases(m, b"ATCGN") for m in w.locus_motifs(int(allele_locus[i]))])
            return h.annotate(alleles.get(i))
        ann = list(pool.map(annotate_one, range(len(alleles))))
    pool.shutdown()
    return ClusterPassResult(spans, sel, off, group, central, alleles, allele_locus, redone, oracle_annotations=ann)


def compare_cluster_pass(got: ClusterPassResult, ref: ClusterPassResult) -> None:
    """bit-exact comparison (raises AssertionError)"""
    assert np.array_equal(got.spans["found"], ref.spans["found"]), "span found flags differ"
    f = ref.spans["found"] != 0
    assert np.array_equal(got.spans["start"][f], ref.spans["start"][f]) and np.array_equal(got.spans["end"][f], ref.spans["end"][f])
    assert np.array_equal(got.sel_reads, ref.sel_reads) and np.array_equal(got.sel_off, ref.sel_off)
    assert np.array_equal(got.group, ref.group), "cluster groups differ"
    assert np.array_equal(got.central, ref.central), "central reads differ"
    assert np.array_equal(got.redone, ref.redone)
    assert np.array_equal(got.allele_locus, ref.allele_locus)
    assert np.array_equal(got.alleles.offsets, ref.alleles.offsets) and np.array_equal(got.alleles.data, ref.alleles.data), "alleles differ"
    if got.annotations is not None and ref.oracle_annotations is not None:
        a = got.annotations
        assert not a.status.any()
        for i, (mc, sp, pur) in enumerate(ref.oracle_annotations):
            x = a.annotation(i)
            assert x.motif_counts == mc and (x.labels or []) == sp, ("MC/MS differ", i)
            assert x.purity == pur or (np.isnan(x.purity) and np.isnan(pur)), ("AP differs", i)
