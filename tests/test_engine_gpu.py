"""GPU parity tests: the CUDA path (libtrgt_b200.so through its C ABI) against the oracle on the same
seeded inputs.  Integer / index results must be bit-exact; AP (purity) is an integer ratio evaluated in
f64 on both sides and must be identical."""
import json
import math
import os
import random

import numpy as np
import pytest

from tests.test_cores_serial import mutate, noisy_repeat, rnd

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ phase C: HMM ---------

def _check_annotations(oracle, loci, got, paths=None):
    k = 0
    for li, (motifs, alleles) in enumerate(loci):
        h = oracle.Hmm([oracle.replace_invalid_bases(m, b"ATCGN") for m in motifs])
        for ai, allele in enumerate(alleles):
            exp_mc, exp_sp, exp_pur = h.annotate(allele)
            ann = got[li][ai]
            assert ann.motif_counts == exp_mc, (motifs, allele)
            assert (ann.labels or []) == exp_sp, (motifs, allele)
            assert (math.isnan(ann.purity) and math.isnan(exp_pur)) or ann.purity == exp_pur, (motifs, allele)
            if paths is not None:
                exp_path = h.label(oracle.replace_invalid_bases(allele, b"ATCG")) if allele else []
                assert paths[k] == exp_path, (motifs, allele)
            k += 1


def test_hmm_reference_goldens(engine, oracle):
    # src/hmm/builder.rs:208-217, docs/tutorial.md:44, src/hmm/purity.rs:48-96
    loci = [
        ([b"CAG", b"A"], [b"CAGCAGCAGCAGAAAAA"]),
        ([b"CAG"], [b"CAG" * 11, b""]),
        ([b"CAG", b"CCG"], [b"CAGCAGCAGCCGCCGCCG", b"CAGCAGCATCAGCCGCCG"]),
        ([b"GCN"], [b"GCAGCGGCTGCC"]),
    ]
    got = engine.label_with_hmm(loci)
    assert got[0][0].labels == [(0, 0, 12), (1, 12, 17)] and got[0][0].motif_counts == [4, 5]
    assert got[1][0].motif_counts == [11] and got[1][0].labels == [(0, 0, 33)] and got[1][0].purity == 1.0
    assert got[1][1].labels is None and math.isnan(got[1][1].purity) and got[1][1].motif_counts == [0]
    _check_annotations(oracle, loci, got)


def test_hmm_random_parity_with_paths(engine, oracle):
    from trgt_b200 import PackedSeqs
    rng = random.Random(11)
    loci = []
    for _ in range(400):
        k = rng.choice([1, 1, 1, 2, 3, 5])
        motifs = [rnd(rng, rng.choice([1, 2, 2, 3, 4, 5, 6, 7, 12]), "ACGTN" if rng.random() < 0.2 else "ACGT")
                  for _ in range(k)]
        alleles = []
        for _ in range(rng.randint(1, 3)):
            a = noisy_repeat(rng, motifs, rng.choice([4, 12, 40]))
            if rng.random() < 0.1:
                a = a[:3] + b"N" + a[3:] + b"X"
            alleles.append(a)
        loci.append((motifs, alleles))
    got = engine.label_with_hmm(loci)
    motifs = PackedSeqs.from_list([m for ms, _ in loci for m in ms])
    lmo = np.cumsum([0] + [len(ms) for ms, _ in loci]).astype(np.uint32)
    alleles = PackedSeqs.from_list([a for _, als in loci for a in als])
    al = np.array([i for i, (_, als) in enumerate(loci) for _ in als], dtype=np.uint32)
    res = engine.hmm_label_packed(motifs, lmo, alleles, al, want_paths=True)
    paths = [res.path(i) for i in range(len(alleles))]
    _check_annotations(oracle, loci, got, paths)


def test_hmm_pathogenic_catalog_shapes(engine, oracle):
    """BASELINE config 2 shapes: the 56 motif sets of repeats/pathogenic_repeats.hg38.bed (up to 10
    motifs, 170 states, N allowed)."""
    from harness import workload
    sets = workload.pathogenic_motif_sets()
    w = workload.generate(len(sets), 4, motif_sets=sets, tr_len_median=60.0)
    rng = random.Random(5)
    loci = []
    for l in range(w.n_loci):
        als = [w.alleles.get(2 * l), mutate(rng, w.alleles.get(2 * l + 1), 0.05)]
        loci.append((w.locus_motifs(l), als))
    _check_annotations(oracle, loci, engine.label_with_hmm(loci))


def test_hmm_long_allele_and_waves(engine, oracle):
    rng = random.Random(13)
    motifs = [b"CAG", b"CCG"]
    long_allele = mutate(rng, b"CAG" * 5000 + b"CCG" * 2000, 0.02)
    loci = [(motifs, [long_allele, b"CAGCAG"]), ([b"AT"], [b"AT" * 9000, mutate(rng, b"AT" * 5000, 0.05)])]
    engine.set_workspace_budget(1 << 20)  # forces several back-pointer waves
    try:
        got = engine.label_with_hmm(loci)
    finally:
        engine.set_workspace_budget(24 << 30)
    _check_annotations(oracle, loci, got)


def test_hmm_lane_path_single_motif_loci(engine, oracle):
    """Single-motif loci (motif of 1..7 bases) run through k_hmm_lane_* (score column in registers, one packed
    back-pointer word per column, slots sorted by motif and allele length): MC / MS / AP and the state paths
    must be the oracle's, with the lane path on and off, in one wave and in several, mixed with loci the
    generic kernels take (several motifs, 12-base motifs, empty alleles)."""
    from trgt_b200 import PackedSeqs
    rng = random.Random(2026)
    loci = []
    for it in range(900):
        if rng.random() < 0.85:
            motifs = [rnd(rng, rng.choice([1, 2, 2, 2, 3, 4, 4, 5, 6, 7, 8]), "ACGTN" if rng.random() < 0.1 else "ACGT")]
        else:
            motifs = [rnd(rng, rng.choice([2, 3, 12])) for _ in range(rng.choice([1, 2, 3]))]
        alleles = []
        for _ in range(rng.randint(1, 3)):
            a = noisy_repeat(rng, motifs, rng.choice([4, 12, 40]))
            if rng.random() < 0.1:
                a = a[:3] + b"N" + a[3:] + b"X"
            if rng.random() < 0.03:
                a = b""
            if it % 150 == 0:
                a = (a or b"AC") * 60   # long alleles: the last length bin, sorted by length
            alleles.append(a)
        loci.append((motifs, alleles))
    motifs = PackedSeqs.from_list([m for ms, _ in loci for m in ms])
    lmo = np.cumsum([0] + [len(ms) for ms, _ in loci]).astype(np.uint32)
    alleles = PackedSeqs.from_list([a for _, als in loci for a in als])
    al = np.array([i for i, (_, als) in enumerate(loci) for _ in als], dtype=np.uint32)
    results = {}
    try:
        for lane in (True, False):
            engine.set_hmm_lane_path(lane)
            engine.reset_stats()
            got = engine.label_with_hmm(loci)
            stats = engine.kernel_stats()
            assert ("k_hmm_lane_viterbi" in stats) == lane and "k_hmm_viterbi_thread" in stats and "k_hmm_viterbi" in stats
            res = engine.hmm_label_packed(motifs, lmo, alleles, al, want_paths=True)
            paths = [res.path(i) for i in range(len(alleles))]
            _check_annotations(oracle, loci, got, paths)
            results[lane] = res
        engine.set_hmm_lane_path(True)
        engine.set_workspace_budget(1 << 20)
        small = [l for l in loci if sum(len(a) for a in l[1]) < 400][:600] + [([b"AT"], [b"AT" * 90000, b"AT" * 7])]
        _check_annotations(oracle, small, engine.label_with_hmm(small))
    finally:
        engine.set_hmm_lane_path(True)
        engine.set_workspace_budget(24 << 30)
    a, b = results[True], results[False]
    assert np.array_equal(a.motif_counts, b.motif_counts) and np.array_equal(a.spans, b.spans)
    assert np.array_equal(a.purity, b.purity, equal_nan=True) and np.array_equal(a.paths, b.paths)


def test_filter_impure_trs_targeted_preset(engine, oracle):
    """a14, filter_impure_trs (tr.rs:400-452, targeted preset: min_read_qual < 0.9): the HMM runs on every READ's
    repeat sequence (D sequences per locus instead of two alleles) and calc_purity decides which reads are dropped.
    trgt_hmm_label serves it with the reads' repeat sequences as `alleles`; purities must be the oracle's and the
    filter (host logic, oracle/host.py) must keep the same reads."""
    from oracle import host
    from harness import workload
    from trgt_b200 import PackedSeqs
    rng = random.Random(314)
    w = workload.generate(60, 30, seed=99, sub_rate=0.002, unit_indel_rate=0.02)
    spans, _ = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    trs, read_locus = [], []
    for l in range(w.n_loci):
        for r in range(int(w.locus_read_off[l]), int(w.locus_read_off[l + 1])):
            if not spans[r]["found"]:
                continue
            tr = w.reads.get(r)[int(spans[r]["start"]):int(spans[r]["end"])]
            if rng.random() < 0.15:   # an impure read: part of its repeat scrambled
                k = max(1, len(tr) // 3)
                tr = tr[:k] + rnd(rng, rng.randint(1, 2 * k)) + tr[2 * k:]
            if rng.random() < 0.02:
                tr = b""
            trs.append(tr)
            read_locus.append(l)
    engine.reset_stats()
    res = engine.hmm_label_packed(w.motifs, w.locus_motif_off, PackedSeqs.from_list(trs), np.array(read_locus, dtype=np.uint32))
    assert not res.status.any() and "k_hmm_lane_walk" in engine.kernel_stats()
    exp = []
    hmms = {}
    for tr, l in zip(trs, read_locus):
        h = hmms.setdefault(l, oracle.Hmm([oracle.replace_invalid_bases(m, b"ATCGN") for m in w.locus_motifs(l)]))
        exp.append(h.annotate(tr)[2])
    assert np.array_equal(res.purity, np.array(exp), equal_nan=True)
    assert (res.purity[~np.isnan(res.purity)] < 0.9).sum() > 20
    # the filter itself, per locus, on the engine's purities and on the oracle's: same reads kept, same order
    read_locus = np.array(read_locus)
    for l in range(w.n_loci):
        idx = np.nonzero(read_locus == l)[0]
        rq = [None if rng.random() < 0.3 else rng.choice([0.85, 0.95, 0.999]) for _ in idx]
        assert host.filter_impure_trs(res.purity[idx].tolist(), rq) == host.filter_impure_trs([exp[i] for i in idx], rq)


# ------------------------------------------------------------------ cluster-genotyper glue ----# ------------------------------------------------------------------ cluster-genotyper glue ----

def test_cluster_matches_oracle(engine, oracle):
    """trgt_cluster (get_dist_matrix -> Ward linkage -> cluster() -> group1 / group2 -> central_read) against the
    oracle on short repeat sequences (real edit distances, many ties), on long alleles (sqrt of the length
    difference, BASELINE config 5 shape), on tiny loci and at max_depth (250 sequences)."""
    rng = random.Random(97)
    loci = []
    for it in range(160):
        kind = rng.random()
        if kind < 0.5:      # two haplotypes of a short repeat, noisy copies
            unit = rnd(rng, rng.randint(2, 6))
            a, b = unit * rng.randint(3, 12), unit * rng.randint(3, 12)
            n = rng.choice([1, 2, 3, 5, 12, 30, 40])
            trs = [mutate(rng, a if rng.random() < 0.5 else b, rng.choice([0.0, 0.02, 0.08])) or b"A" for _ in range(n)]
        elif kind < 0.9:    # long alleles: no alignment at all (len1 * len2 > 10000)
            la, lb = rng.randint(300, 3000), rng.randint(300, 3000)
            n = rng.choice([4, 20, 40, 41])
            trs = [rnd(rng, (la if rng.random() < 0.6 else lb) + rng.randint(-6, 6)) for _ in range(n)]
        else:
            trs = [rnd(rng, rng.randint(1, 30)) for _ in range(rng.choice([100, 250]))]
        loci.append(trs)
    loci.append([])
    got = engine.cluster(loci)
    for trs, (sel, central, ng) in zip(loci, got):
        n = len(trs)
        if n == 0:
            assert (sel, central, ng) == ([], (None, None), 0)
            continue
        d = oracle.get_dist_matrix(trs) if n >= 2 else []
        exp = oracle.cluster_locus(d, n)
        assert (sel, central, ng) == exp, (n, [len(t) for t in trs])
    assert "k_cluster_ward" in engine.kernel_stats()


# ------------------------------------------------------------------ phase A: flanks --------

def _check_flanks(oracle, w, spans, hits, scoring, frac):
    n_wfa = 0
    for l in range(w.n_loci):
        lp, rp = w.left.get(l), w.right.get(l)
        for r in range(int(w.locus_read_off[l]), int(w.locus_read_off[l + 1])):
            read = w.reads.get(r)
            sides = []
            for side, piece in enumerate((lp, rp)):
                exp, via, nm = oracle.find_span(piece, read, scoring, len(piece) * frac)
                h = hits[2 * r + side]
                assert int(h["via"]) == via, (l, r, side)
                assert int(h["matches"]) == nm, (l, r, side)
                if exp is not None:
                    assert (int(h["start"]), int(h["end"])) == exp, (l, r, side)
                n_wfa += via >= 2
                sides.append(exp)
            exp_span = None
            if sides[0] is not None and sides[1] is not None and sides[0][1] <= sides[1][0]:
                exp_span = (sides[0][1], sides[1][0])
            s = spans[r]
            got = (int(s["start"]), int(s["end"])) if s["found"] else None
            assert got == exp_span, (l, r)
    return n_wfa


@pytest.mark.parametrize("band_budget", [20, 6, 0])
def test_flank_spans_synthetic_hifi(engine, oracle, band_budget):
    """band_budget 20: misses settled by the on-chip banded path; 0: all by the full-width kernels;
    6: a mix.  All three must give the reference's answer."""
    from harness import workload
    w = workload.generate(40, 12, seed=99)
    engine.set_flank_band_budget(band_budget)
    try:
        engine.reset_stats()
        spans, hits = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring,
                                                w.min_flank_id_frac)
        stats = engine.kernel_stats()
    finally:
        engine.set_flank_band_budget(20)
    n_wfa = _check_flanks(oracle, w, spans, hits, w.scoring, w.min_flank_id_frac)
    assert n_wfa > 20  # the WFA fallback was exercised
    assert "k_flank_exact_t" in stats
    if band_budget > 0:   # the first cost tier ran in its two passes (seed pass, band pass on bulk-copied windows)
        assert stats["k_flank_seed"][0] >= 1 and stats["k_flank_band1"][0] >= 1 and "k_flank_band2" in stats
    else:
        assert "k_flank_seed" not in stats and "k_wfa_score_block" in stats


def test_flank_spans_long_reads_and_repetitive_flanks(engine, oracle):
    """Reads too long for the staged on-chip copy, and low-complexity flanks with many seed hits."""
    rng = random.Random(17)
    loci = []
    for kind in range(6):
        if kind % 2 == 0:
            lf, rf = rnd(rng, 250), rnd(rng, 250)
        else:
            u1, u2 = rnd(rng, rng.randint(2, 9)), rnd(rng, rng.randint(2, 9))
            lf = mutate(rng, (u1 * 200)[:260], 0.03)[:250].ljust(250, b"A")
            rf = mutate(rng, (u2 * 200)[:260], 0.03)[:250].ljust(250, b"C")
        reads = []
        for _ in range(6):
            ctx = rng.choice([300, 1200, 3000])
            read = rnd(rng, ctx) + mutate(rng, lf, 0.01) + b"CAG" * rng.randint(3, 400) + mutate(rng, rf, 0.01) + rnd(rng, ctx)
            reads.append(read)
        loci.append((lf, rf, reads))
    got = engine.find_tr_spans(loci)
    for (lf, rf, reads), g in zip(loci, got):
        assert g == oracle.find_tr_spans(lf, rf, reads)


def test_flank_spans_noisy_targeted_scoring(engine, oracle):
    """--preset targeted scoring (1,0,1), 0.8 identity, 200-bp pieces (cli.rs:271-302); noisy reads so
    that accepted, rejected and discordant cases all occur."""
    from harness import workload
    w = workload.generate(24, 10, seed=5, context=260, piece=200, sub_rate=0.1, ins_rate=0.1, del_rate=0.1)
    spans, hits = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, (1, 0, 1), 0.8)
    _check_flanks(oracle, w, spans, hits, (1, 0, 1), 0.8)
    vias = set(int(v) for v in hits["via"])
    assert 2 in vias and 3 in vias


@pytest.mark.parametrize("scoring", [(3, 3, 1), (4, 6, 2), (1, 0, 1), (2, 1, 1), (6, 5, 1)])
def test_flank_spans_custom_scorings_with_decoys(engine, oracle, scoring):
    """--aln-scoring other than the presets through every tier (a cheap gap open moves the seed filter's block count,
    flank_seed_blocks; the first tier's cost cap and live scores follow the scoring), HiFi-like noise so that most
    misses are one edit, plus reads that carry a second, damaged copy of a flank piece (an index hit off the
    alignment: the first tier's hull has to be verified or handed on)."""
    from harness import workload
    from trgt_b200 import PackedSeqs
    w = workload.generate(30, 10, seed=11 + scoring[0], sub_rate=1e-3, ins_rate=1e-3, del_rate=1e-3)
    rng = random.Random(scoring[1])
    reads = []
    for r in range(w.n_reads):
        read = w.reads.get(r)
        if r % 7 == 3:    # a decoy: 60 bases of the left piece, one of them changed, ahead of the read
            l = int(np.searchsorted(w.locus_read_off, r, side="right") - 1)
            bit = bytearray(w.left.get(l)[40:100])
            bit[30] = ord("A") if bit[30] != ord("A") else ord("C")
            read = bytes(bit) + rnd(rng, 40) + read
        reads.append(read)
    packed = PackedSeqs.from_list(reads)
    w.reads = packed
    spans, hits = engine.flank_spans_packed(w.left, w.right, packed, w.locus_read_off, scoring, 0.7)
    n_wfa = _check_flanks(oracle, w, spans, hits, scoring, 0.7)
    assert n_wfa > 10


def test_flank_spans_edge_cases(engine, oracle):
    rng = random.Random(21)
    lf = rnd(rng, 300)
    rf = rnd(rng, 300)
    tr = b"CAG" * 10
    good = lf[-250:] + tr + rf[:250]
    reads = [
        good,
        good[100:],                      # left flank truncated
        rnd(rng, 600),                   # unrelated read
        rf[:250] + tr + lf[-250:],       # flanks in the wrong order (discordant)
        lf[-250:] + rf[:250],            # empty repeat
        lf[-250:][:120],                 # read shorter than the piece
        b"A",
    ]
    out, hits = engine.find_tr_spans([(lf, rf, reads)], return_hits=True)
    exp = oracle.find_tr_spans(lf, rf, reads)
    assert out[0] == exp
    assert out[0][0] == (250, 280) and out[0][4] == (250, 250) and out[0][2] is None and out[0][3] is None
    # ragged batch: a locus without reads between two with reads
    out2 = engine.find_tr_spans([(lf, rf, [good]), (lf, rf, []), (lf, rf, [good, good[100:]])])
    assert out2 == [[(250, 280)], [], [(250, 280), exp[1]]]
    assert engine.find_tr_spans([]) == []


# ------------------------------------------------------------------ phase B ------------------

def test_align_parity_short(engine, oracle):
    rng = random.Random(31)
    groups = []
    for _ in range(150):
        bb = rnd(rng, rng.randint(1, 80))
        seqs = []
        for _ in range(rng.randint(0, 6)):
            r = rng.random()
            seqs.append(bb if r < 0.3 else (mutate(rng, bb, rng.choice([0.02, 0.1, 0.4])) or b"A") if r < 0.9
                        else rnd(rng, rng.randint(1, 80)))
        groups.append((bb, seqs))
    got = engine.align(groups)
    for (bb, seqs), res in zip(groups, got):
        assert res == oracle.align(bb, seqs)


def test_align_parity_long_and_scores(engine, oracle):
    from trgt_b200 import PackedSeqs
    rng = random.Random(37)
    bbs, seqs, offs = [], [], [0]
    for L in (300, 1500, 6000):
        bb = rnd(rng, L)
        bbs.append(bb)
        seqs += [mutate(rng, bb, 0.01), mutate(rng, bb, 0.03), bb]
        offs.append(len(seqs))
    res = engine.align_packed(PackedSeqs.from_list(bbs), PackedSeqs.from_list(seqs), np.array(offs, dtype=np.uint32))
    assert not res.status.any()
    for i, s in enumerate(seqs):
        words, score = oracle.align_words(bbs[i // 3], s)
        assert res.cigar(i) == words and int(res.scores[i]) == score


def test_consensus_parity(engine, oracle):
    """next row: utils::align + repair_consensus on the device vs the oracle."""
    from trgt_b200 import TrgtError
    rng = random.Random(53)
    groups = []
    for it in range(200):
        unit = rnd(rng, rng.randint(2, 6))
        truth = unit * rng.randint(2, 25 if it % 10 else 400)
        pos = rng.randint(0, len(truth))
        truth2 = truth[:pos] + rnd(rng, rng.randint(1, 4)) + truth[pos:]
        seqs = [(mutate(rng, truth2 if rng.random() < 0.6 else truth, rng.choice([0, 0.02, 0.1])) or b"A")
                for _ in range(rng.randint(1, 40))]
        groups.append((rng.choice(seqs) if rng.random() < 0.7 else truth, seqs))
    groups.append((b"ACGT", []))  # a group without members: every column ties at 0 -> '-' -> empty consensus
    got = engine.repair_consensus(groups)
    for (bb, seqs), g in zip(groups, got):
        assert g == oracle.repair_consensus(bb, seqs)
    assert engine.repair_consensus([]) == []
    with pytest.raises(TrgtError):  # consensus.rs:81 panics on a base outside ACGT
        engine.repair_consensus([(b"ACGT", [b"ACNT", b"ACGT"])])


def test_edit_dist_parity(engine, oracle):
    rng = random.Random(41)
    loci = []
    for _ in range(40):
        base = rnd(rng, rng.randint(1, 90))
        n = rng.choice([0, 1, 2, 5, 12])
        trs = [mutate(rng, base, rng.choice([0, 0.05, 0.3])) or b"G" for _ in range(n)]
        if n > 2 and rng.random() < 0.3:
            trs[1] = rnd(rng, 400)  # len1*len2 > MAX_OPS -> length difference
        loci.append(trs)
    got = engine.get_dist_matrix(loci)
    for trs, g in zip(loci, got):
        assert g == oracle.get_dist_matrix(trs)


# ------------------------------------------------------------------ boundary behaviour -------

def test_error_reporting(engine):
    from trgt_b200 import PackedSeqs, TrgtError
    one = PackedSeqs.from_list([b"ACGT"])
    with pytest.raises(TrgtError):  # one piece per locus is required
        engine.flank_spans_packed(one, PackedSeqs.from_list([]), one, np.array([0, 1], dtype=np.uint32))
    with pytest.raises(TrgtError):  # gap_extend must be >= 1
        engine.flank_spans_packed(one, one, one, np.array([0, 1], dtype=np.uint32), scoring=(2, 5, 0))
    with pytest.raises(TrgtError):  # allele pointing at a locus that does not exist
        engine.hmm_label_packed(one, np.array([0, 1], dtype=np.uint32), one, np.array([3], dtype=np.uint32))
    # the engine is still usable afterwards
    assert engine.label_with_hmm([([b"CAG"], [b"CAGCAG"])])[0][0].motif_counts == [2]


def test_kernels_actually_launched(engine):
    engine.reset_stats()
    engine.label_with_hmm([([b"CAG"], [b"CAGCAGCAG"])])
    stats = engine.kernel_stats()
    assert stats["k_hmm_lane_viterbi"][0] >= 1 and stats["k_hmm_lane_walk"][0] >= 1 and engine.launches() >= 2


# ------------------------------------------------------------------ whole pass, other BASELINE configs ---

def _pass_and_compare(engine, oracle, w):
    from harness.pipeline import HotPath, compare_with_oracle, oracle_pass
    hp = HotPath(engine, w, want_hits=False, pinned_outputs=False)
    res = hp.run_e2e(copy=True)
    ref = oracle_pass(oracle, w, 4)
    compare_with_oracle(res, ref)
    # the same pass with the reads handed over as BAM 4-bit bases and the repeat sequences cut on the device
    w.pack_seq4()
    hp4 = HotPath(engine, w, want_hits=False, pinned_outputs=False, use_seq4=True)
    assert hp4.use_seq4
    compare_with_oracle(hp4.run_e2e(copy=True), ref)
    # the resident-batch interface must give the same answers
    hp.prepare_resident()
    hp.run_resident()
    res2 = hp.download_resident()
    hp.free_resident()
    assert np.array_equal(res.spans, res2.spans) and np.array_equal(res.cigars.words, res2.cigars.words)
    assert np.array_equal(res.annotations.spans, res2.annotations.spans)
    assert np.array_equal(res.annotations.motif_counts, res2.annotations.motif_counts)
    return res


def test_pass_pathogenic_catalog_config2(engine, oracle):
    """BASELINE config 2: the 56 loci of repeats/pathogenic_repeats.hg38.bed (motif sets from the committed
    fixture; up to 10 motifs and 170 HMM states per locus, N in motifs), synthetic 30x HiFi."""
    from harness import workload
    sets = workload.pathogenic_motif_sets()
    w = workload.generate(len(sets), 30, motif_sets=sets, tr_len_median=60.0, seed=2)
    res = _pass_and_compare(engine, oracle, w)
    assert res.spans["found"].mean() > 0.9


def test_pass_long_expansions_config5_shape(engine, oracle):
    """BASELINE config 5 shape at test size: alleles of several kb (reads too long for the staged on-chip
    path, consensus pairs through the CTA-per-pair kernels, Viterbi over thousands of columns)."""
    from harness import workload
    w = workload.generate(5, 6, tr_len_median=3000.0, tr_len_sigma=0.5, tr_len_min=1500, tr_len_max=9000, seed=8)
    assert int(np.diff(w.reads.offsets.astype(np.int64)).max()) > 4000
    _pass_and_compare(engine, oracle, w)


def test_pass_long_expansions_config5_at_size(engine, oracle):
    """BASELINE config 5 at its own sizes: 16 loci, 40x, alleles of 20-50 kb (heterozygous loci draw two
    independent lengths), the cluster-genotyper pass -- flank spans on 21-51 kb reads, spanning order, distance
    matrices + Ward clusters + central reads, consensus of both groups (20 alignments of ~35 kb each, repaired),
    HMM on 20-50 kb alleles -- against the same pass on the oracle.  Run twice: reads resident from ASCII with the
    lane HMM kernels, and from BAM 4-bit bases with the generic HMM kernels in several back-pointer waves."""
    from harness import workload
    from harness.cluster_pass import compare_cluster_pass, engine_cluster_pass, oracle_cluster_pass
    w = workload.generate(16, 40, seed=505, tr_len_dist="loguniform", tr_len_min=20000, tr_len_max=50000,
                          het_independent=True)
    assert int(np.diff(w.alleles.offsets.astype(np.int64)).min()) >= 19000
    ref = oracle_cluster_pass(oracle, w, n_threads=os.cpu_count() or 1)
    assert ref.sel_off[-1] > 0.8 * w.n_reads and len(ref.alleles) == 32
    engine.reset_stats()
    compare_cluster_pass(engine_cluster_pass(engine, w, orc_for_redo=oracle), ref)
    stats = engine.kernel_stats()
    for k in ("k_cluster_ward", "k_trs_gather", "k_wfa_score_block", "k_wfa_trace", "k_consensus_vote_write", "k_hmm_lane_viterbi"):
        assert k in stats, (k, sorted(stats))
    w.pack_seq4()
    engine.set_hmm_lane_path(False)
    engine.set_workspace_budget(8 << 20)   # one wavefront ring of a 50 kb pair is 4.3 MB; the back-pointers are 19 MB
    try:
        engine.reset_stats()
        compare_cluster_pass(engine_cluster_pass(engine, w, orc_for_redo=oracle, use_seq4=True), ref)
        stats = engine.kernel_stats()
        assert stats["k_hmm_viterbi_thread"][0] >= 3 and "k_unpack_seq4" in stats
    finally:
        engine.set_hmm_lane_path(True)
        engine.set_workspace_budget(24 << 30)


def test_cluster_pass_outlier_rule_and_small_loci(engine, oracle):
    """the cluster pass on short repeats at low depth: loci with 0 / 1 / 2 spanning reads, real edit distances,
    and the outlier rule (a group 4x smaller and within 100 bp -> alternate split, genotype_cluster.rs:84-111)"""
    from harness import workload
    from harness.cluster_pass import compare_cluster_pass, engine_cluster_pass, oracle_cluster_pass
    seen_redo = 0
    for seed, depth in ((3, 1), (4, 2), (5, 3), (6, 12), (7, 30)):
        w = workload.generate(48, depth, seed=seed, tr_len_median=30.0, unit_indel_rate=0.05, sub_rate=0.004)
        ref = oracle_cluster_pass(oracle, w, n_threads=4)
        compare_cluster_pass(engine_cluster_pass(engine, w, orc_for_redo=oracle), ref)
        seen_redo += ref.redone.size
    assert seen_redo > 0


def test_hmm_locus_without_motifs_and_single_base_motif(engine, oracle):
    loci = [([], [b"ACGTACGT", b"A"]), ([b"A"], [b"AAAAAAA", b"AAACAAA", b""]), ([b"N"], [b"ACGT"])]
    _check_annotations(oracle, loci, engine.label_with_hmm(loci))


@pytest.mark.parametrize("piece_len", [10, 16, 40, 256, 257, 500])
def test_flank_spans_unindexed_piece_lengths(engine, oracle, piece_len):
    """Pieces too short or too long for the 8-mer index: linear exact scan, and misses go all the way
    down to the wide-band / full-width kernels."""
    rng = random.Random(piece_len)
    loci = []
    for _ in range(6):
        lf, rf = rnd(rng, piece_len + 20), rnd(rng, piece_len + 20)
        reads = []
        for _ in range(8):
            rate = rng.choice([0.0, 0.0, 0.01, 0.05])
            reads.append(rnd(rng, rng.randint(0, 200)) + mutate(rng, lf[-piece_len:], rate) + b"CAG" * rng.randint(0, 30) +
                         mutate(rng, rf[:piece_len], rate) + rnd(rng, rng.randint(piece_len + 10, piece_len + 300)))
        loci.append((lf, rf, reads))
    got = engine.find_tr_spans(loci, search_flank_len=piece_len)
    for (lf, rf, reads), g in zip(loci, got):
        assert g == oracle.find_tr_spans(lf, rf, reads, search_flank_len=piece_len)


# ------------------------------------------------------------------ producer row: clip + BAM bases ---

def test_clip_reads_reference_vectors(engine, oracle):
    """src/trgt/reads/clip_region.rs:214-299 through trgt_clip_reads (one locus per region)"""
    cases = [("3=2D2=1X2=5I3=", (0, 10)), ("3=2D2=1X2=5I3=", (23, 33)), ("5S3=2D2=1X2=5I3=10S", (9, 23)),
             ("3=2D2=1X2=5I3=", (0, 15)), ("3=2D2=1X2=5I3=", (12, 17)), ("3=2D2=1X2=5I3=", (21, 22)),
             ("3=2D2=1X2=5I3=", (0, 17))]
    ops, offs = [], [0]
    for cig, _ in cases:
        ops += oracle.encode_bam_cigar(cig)
        offs.append(len(ops))
    clips = engine.clip_reads(np.array(ops, dtype=np.uint32), np.array(offs, dtype=np.uint64),
                              np.full(len(cases), 10, dtype=np.int64),
                              np.array([r for _, r in cases], dtype=np.int64),
                              np.arange(len(cases) + 1, dtype=np.uint32))
    exp = [None, None, (10, 0, 31, "5S3=2D2=1X2=5I3=10S"), (10, 0, 3, "3=2D"), (12, 2, 5, "1=2D2="),
           (21, 14, 15, "1="), (10, 0, 5, "3=2D2=")]
    for i, (c, e) in enumerate(zip(clips, exp)):
        if e is None:
            assert c["status"] == 0
            continue
        o0 = offs[i]
        words = [int(c["first_word"]) if k == 0 else int(c["last_word"]) if k == c["n_ops"] - 1
                 else ops[o0 + int(c["first_op"]) + k] for k in range(int(c["n_ops"]))]
        text = "".join(f"{w >> 4}{oracle.BAM_OPS[w & 15]}" for w in words)
        assert c["status"] == 1 and (int(c["ref_start"]), int(c["query_start"]), int(c["query_end"]), text) == e


def test_clip_reads_random_parity(engine, oracle):
    from tests.test_cores_serial import random_bam_cigar
    rng = random.Random(123)
    n_loci = 300
    regions, lro, ops, offs, refs = [], [0], [], [0], []
    for _ in range(n_loci):
        a = rng.randint(0, 400)
        regions.append((a, a + rng.randint(0, 300)))
        for _ in range(rng.choice([0, 1, 5, 30])):
            o = random_bam_cigar(rng, rng.randint(0, 40))
            ops += o
            offs.append(len(ops))
            refs.append(rng.randint(0, 500))
        lro.append(len(refs))
    clips = engine.clip_reads(np.array(ops, dtype=np.uint32), np.array(offs, dtype=np.uint64),
                              np.array(refs, dtype=np.int64), np.array(regions, dtype=np.int64),
                              np.array(lro, dtype=np.uint32))
    n_hit = 0
    for l in range(n_loci):
        for r in range(lro[l], lro[l + 1]):
            o = ops[offs[r]:offs[r + 1]]
            exp = oracle.clip_cigar(o, refs[r], regions[l])
            c = clips[r]
            if exp is None:
                assert c["status"] == 0
                continue
            n_hit += 1
            words = [int(c["first_word"]) if k == 0 else int(c["last_word"]) if k == c["n_ops"] - 1
                     else o[int(c["first_op"]) + k] for k in range(int(c["n_ops"]))]
            assert c["status"] == 1 and (int(c["ref_start"]), int(c["query_start"]), int(c["query_end"]), words) == exp
    assert n_hit > 500
    assert "k_clip_cigar" in engine.kernel_stats()


def test_bamlet_clip_parity(engine, oracle):
    """trgt_bamlet_clip (write_bam.rs:72-92 -> clip_bases.rs:9-119) on the resident reads and spans of a phase-A
    pass: every read against the oracle, with CIGARs of the read's query length, reads without CIGAR, alignments
    that are too short (the reference's assert) and flank lengths from 0 to longer than the flanks"""
    from harness import workload
    from tests.test_cores_serial import _bamlet_expect, bamlet_got
    rng = random.Random(77)
    w = workload.generate(60, 12, seed=99)
    spans, _ = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    for flank in (0, 50, 250, 600):
        ops, offs, refs, per_read = [], [0], [], []
        for r in range(w.n_reads):
            n = len(w.reads.get(r))
            o = []
            if rng.random() < 0.9:
                left = n
                if rng.random() < 0.4:
                    k = rng.randint(1, 30); o.append((k << 4) | 4); left -= k
                while left > 0:
                    op = rng.choice([7, 7, 7, 8, 1, 2, 0])
                    k = rng.randint(1, min(120, left)) if op != 2 else rng.randint(1, 9)
                    o.append((k << 4) | op)
                    if op != 2:
                        left -= k
                if rng.random() < 0.03 and len(o) > 1:
                    o = o[:-1]
            ops += o
            offs.append(len(ops))
            refs.append(rng.randint(0, 10 ** 7))
            per_read.append(o)
        clips = engine.bamlet_clip(None, np.array(ops, dtype=np.uint32), np.array(offs, dtype=np.uint64),
                                   np.array(refs, dtype=np.int64), flank)
        seen = {1: 0, 0: 0, -500: 0}
        for r in range(w.n_reads):
            span = (int(spans[r]["start"]), int(spans[r]["end"])) if spans[r]["found"] else None
            exp = _bamlet_expect(oracle, w.reads.get(r), per_read[r], refs[r], span, flank)
            assert bamlet_got(clips[r], per_read[r]) == exp, (flank, r)
            seen[exp[0]] += 1
        assert seen[1 if flank <= 250 else 0] > 300, (flank, seen)
    assert "k_bamlet_clip" in engine.kernel_stats()


def test_seq4_decode_parity(engine, oracle):
    from trgt_b200 import PackedSeq4
    rng = random.Random(9)
    seqs = [rnd(rng, n, "ACGTN=MRSVWYHKDB") for n in
            [0, 1, 2, 15, 16, 17, 31, 32, 33, 1024, 1023, 1025, 0, 5000] + [rng.randint(0, 1500) for _ in range(300)]]
    for odd in (True, False):
        p4 = PackedSeq4.from_ascii(seqs, odd_starts=odd)
        got = engine.seq4_decode(p4)
        assert [got.get(i) for i in range(len(seqs))] == seqs
        for i in (0, 3, 9, 13, 50):
            assert oracle.decode_seq4(p4.data.tobytes(), int(p4.starts[i]), int(p4.lengths[i])) == seqs[i]
    empty = engine.seq4_decode(PackedSeq4.from_ascii([]))
    assert len(empty) == 0
    assert "k_unpack_seq4" in engine.kernel_stats()


def test_flank_spans_seq4_matches_ascii_path(engine, oracle):
    """BAM 4-bit input gives the ASCII path's result bit for bit (and hence the oracle's: the ASCII path is
    checked against it above), one-shot and resident"""
    from trgt_b200 import PackedSeq4
    from harness import workload
    w = workload.generate(60, 14, seed=31)
    p4 = PackedSeq4.from_ascii([w.reads.get(r) for r in range(w.n_reads)])
    spans_a, hits_a = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    spans_p, hits_p = engine.flank_spans_seq4(w.left, w.right, p4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    assert np.array_equal(spans_a, spans_p) and np.array_equal(hits_a, hits_p)
    n_wfa = _check_flanks(oracle, w, spans_p, hits_p, w.scoring, w.min_flank_id_frac)
    assert n_wfa > 20
    b = engine.flank_upload_seq4(w.left, w.right, p4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    try:
        engine.flank_run(b)
        spans_r, hits_r = engine.flank_download(b, w.n_reads)
    finally:
        engine.flank_free(b)
    assert np.array_equal(spans_a, spans_r) and np.array_equal(hits_a, hits_r)
    # ragged: loci without reads, an empty batch
    lro = np.array([0, 0, 3, 3, w.locus_read_off[1]], dtype=np.uint32)
    from trgt_b200 import PackedSeqs
    l4 = PackedSeqs.from_list([w.left.get(0)] * 4)
    r4 = PackedSeqs.from_list([w.right.get(0)] * 4)
    n0 = int(w.locus_read_off[1])
    sub = PackedSeq4.from_ascii([w.reads.get(r) for r in range(n0)])
    sp, _ = engine.flank_spans_seq4(l4, r4, sub, lro, w.scoring, w.min_flank_id_frac)
    assert np.array_equal(sp, spans_a[:n0])
    sp0, _ = engine.flank_spans_seq4(PackedSeqs.from_list([]), PackedSeqs.from_list([]), PackedSeq4.from_ascii([]),
                                     np.zeros(1, dtype=np.uint32), w.scoring, w.min_flank_id_frac)
    assert sp0.size == 0


def test_align_trs_matches_align_packed(engine, oracle):
    """trgt_align_trs (backbones and members named by read index, gathered on the device from the flank batch) gives the
    CIGARs of trgt_align_e2e on the same sequences, and the whole pass through it matches the oracle"""
    from harness import workload
    from harness.pipeline import HotPath, compare_with_oracle, oracle_pass
    w = workload.generate(80, 14, seed=123)
    w.pack_seq4()
    hp = HotPath(engine, w, want_hits=False, pinned_outputs=False, use_seq4=True, align_by_index=True)
    res = hp.run_e2e(copy=True)
    compare_with_oracle(res, oracle_pass(oracle, w, 4))
    ref = engine.align_packed(res.glue.backbones, res.glue.seqs, res.glue.group_seq_off)
    assert np.array_equal(ref.words, res.cigars.words) and np.array_equal(ref.offsets, res.cigars.offsets)
    assert "k_trs_gather" in engine.kernel_stats()


def test_flank_trs_are_the_span_slices(engine, oracle):
    """trgt_flank_trs = read.bases[span.0..span.1] per spanning read (tr.rs:58-62), for one-shot (ASCII and
    BAM 4-bit input) and resident batches; and the host glue fed with them builds the same phase B/C input."""
    from trgt_b200 import PackedSeq4
    from harness import workload
    w = workload.generate(50, 9, seed=77)
    p4 = PackedSeq4.from_ascii([w.reads.get(r) for r in range(w.n_reads)])

    def expect(spans):
        return [w.reads.get(r)[int(s["start"]):int(s["end"])] if s["found"] else b"" for r, s in enumerate(spans)]

    spans, _ = engine.flank_spans_seq4(w.left, w.right, p4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    trs = engine.flank_trs(copy=True)
    assert [trs.get(r) for r in range(w.n_reads)] == expect(spans) and spans["found"].sum() > 300
    g1 = workload.genotype_glue(w, spans)
    g2 = workload.genotype_glue(w, spans, trs=trs)
    for a, b in ((g1.seqs, g2.seqs), (g1.backbones, g2.backbones)):
        assert np.array_equal(a.data, b.data) and np.array_equal(a.offsets, b.offsets)
    assert np.array_equal(g1.seq_read, g2.seq_read) and np.array_equal(g1.group_locus, g2.group_locus)
    spans2, _ = engine.flank_spans_packed(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    trs2 = engine.flank_trs(copy=True)
    assert np.array_equal(trs2.data, trs.data) and np.array_equal(trs2.offsets, trs.offsets)
    b = engine.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    try:
        engine.flank_run(b)
        trs3 = engine.flank_trs(b, copy=True)
    finally:
        engine.flank_free(b)
    assert np.array_equal(trs3.data, trs.data) and np.array_equal(trs3.offsets, trs.offsets)
    # no spanning read at all / no reads
    from trgt_b200 import PackedSeqs
    junk = PackedSeqs.from_list([rnd(random.Random(1), 700) for _ in range(3)])
    engine.flank_spans_packed(PackedSeqs.from_list([w.left.get(0)]), PackedSeqs.from_list([w.right.get(0)]), junk,
                              np.array([0, 3], dtype=np.uint32), w.scoring, w.min_flank_id_frac)
    t0 = engine.flank_trs(copy=True)
    assert len(t0) == 3 and t0.data.size == 0


# ------------------------------------------------------------------ VCF sample fields (next row) ---

def test_vcf_fields_tutorial_and_random_parity(engine, oracle):
    """trgt_vcf_fields behind phase C against the oracle's restatement of write_vcf.rs:267-343; the tutorial
    locus must give the record of docs/tutorial.md:44"""
    rng = random.Random(88)
    loci = [([b"CAG"], [b"CAG" * 11, b"CAG" * 11])]
    for _ in range(300):
        k = rng.choice([1, 1, 2, 3])
        motifs = [rnd(rng, rng.choice([2, 3, 4, 5, 6, 9])) for _ in range(k)]
        alleles = [noisy_repeat(rng, motifs, rng.choice([1, 6, 30])) for _ in range(rng.choice([0, 1, 2, 2, 3]))]
        if rng.random() < 0.1 and alleles:
            alleles[0] = b""
        loci.append((motifs, alleles))
    ann = engine.label_with_hmm(loci)
    got = engine.vcf_fields()
    assert len(got) == len(loci)
    assert got[0] == (b"33,33", b"11,11", b"0(0-33),0(0-33)", b"1.000000,1.000000")
    n_none = 0
    for (motifs, alleles), a, g in zip(loci, ann, got):
        exp = oracle.vcf_fields([(len(s), x.motif_counts, x.labels, x.purity) for s, x in zip(alleles, a)])
        assert g == exp, (motifs, alleles)
        n_none += b"." in g[2].split(b",")
    assert n_none > 5
    stats = engine.kernel_stats()
    assert "k_vcf_fields_count" in stats and "k_vcf_fields_write" in stats
    # an invalid base: the reference panics; the allele's MC / MS / AP read '.'
    engine.label_with_hmm([([b"CAG"], [b"CAGCAG", b"CAGXAG"])])
    bad = engine.vcf_fields()
    assert bad[0][0] == b"6,6" and bad[0][1].count(b",") == 1
    # alleles not grouped by locus -> argument error, not a wrong answer
    from trgt_b200 import PackedSeqs, TrgtError
    m = PackedSeqs.from_list([b"CAG", b"AT"])
    al = PackedSeqs.from_list([b"CAGCAG", b"ATAT", b"CAGCAGCAG"])
    engine.hmm_label_packed(m, np.array([0, 1, 2], dtype=np.uint32), al, np.array([0, 1, 0], dtype=np.uint32))
    with pytest.raises(TrgtError):
        engine.vcf_fields()


def test_results_before_run_are_refused(engine):
    from trgt_b200 import TrgtError
    from harness import workload
    w = workload.generate(3, 4, seed=1)
    b = engine.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)
    try:
        with pytest.raises(TrgtError):
            engine.flank_download(b, w.n_reads)
        with pytest.raises(TrgtError):
            engine.flank_trs(b)
        engine.flank_run(b)
        assert len(engine.flank_trs(b, copy=True)) == w.n_reads
    finally:
        engine.flank_free(b)
