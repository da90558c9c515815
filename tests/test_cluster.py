"""Cluster-genotyper glue: the oracle's restatement of kodama's Ward linkage (NN-chain on
squared dissimilarities, un-vendored dependency) is pinned on scipy.cluster.hierarchy.linkage(method="ward"),
which implements the same fastcluster algorithm; cluster() / central_read() follow genotype_cluster.rs."""
import math
import random

import numpy as np
import pytest

from oracle import oracle as orc


def scipy_ward(d, n):
    from scipy.cluster.hierarchy import linkage
    return linkage(np.asarray(d, dtype=np.float64), method="ward")


def random_matrix(rng, n, kind):
    if kind == "lengths":      # get_dist of long alleles: sqrt(|len1 - len2|), genotype_cluster.rs:239-243
        a, b = rng.randint(5000, 50000), rng.randint(5000, 50000)
        lens = [(a if rng.random() < 0.5 else b) + rng.randint(-30, 30) for _ in range(n)]
        return [math.sqrt(abs(lens[i] - lens[j])) for i in range(n) for j in range(i + 1, n)]
    if kind == "edit":         # sqrt of small integers, many ties
        return [math.sqrt(rng.randint(0, 12)) for i in range(n) for j in range(i + 1, n)]
    pts = [(rng.random(), rng.random()) for _ in range(n)]   # generic euclidean, no ties
    return [math.dist(pts[i], pts[j]) for i in range(n) for j in range(i + 1, n)]


def test_ward_linkage_matches_scipy():
    rng = random.Random(17)
    for it in range(300):
        n = rng.choice([2, 3, 4, 5, 8, 13, 30, 40, 77])
        d = random_matrix(rng, n, "euclid")
        steps, _ = orc.ward_linkage(d, n)
        Z = scipy_ward(d, n)
        assert len(steps) == n - 1
        for (c1, c2, h, size), z in zip(steps, Z):
            assert (c1, c2) == (int(min(z[0], z[1])), int(max(z[0], z[1])))
            assert size == int(z[3])
            assert h == pytest.approx(z[2], rel=1e-9, abs=1e-12)


def test_ward_top_of_the_tree_with_ties_matches_scipy():
    """Long alleles (config 5) give sqrt(|length difference|): many tied dissimilarities, for which the order of
    the low merges is implementation-defined.  The top of the tree -- what cluster() cuts -- must still agree:
    the last merge joins the same two clusters at the same height."""
    rng = random.Random(18)
    for it in range(200):
        n = rng.choice([3, 5, 12, 30, 40])
        d = random_matrix(rng, n, "lengths")
        steps, _ = orc.ward_linkage(d, n)
        Z = scipy_ward(d, n)
        assert steps[-1][3] == n
        assert steps[-1][2] == pytest.approx(Z[-1, 2], rel=1e-9)
        sizes = lambda c, st: 1 if c < n else st[c - n]
        mine = sorted((sizes(steps[-1][0], [s[3] for s in steps]), sizes(steps[-1][1], [s[3] for s in steps])))
        ref = sorted((sizes(int(Z[-1, 0]), list(Z[:, 3])), sizes(int(Z[-1, 1]), list(Z[:, 3]))))
        assert mine == [int(x) for x in ref]


def test_cluster_groups_of_two_alleles():
    """Two well separated allele lengths (BASELINE config 5 shape): the reads split by allele, the largest
    group comes first, central reads belong to their groups."""
    rng = random.Random(19)
    for it in range(200):
        n = rng.choice([6, 20, 40, 41])
        a = rng.randint(5000, 40000)
        b = a + rng.randint(300, 9000)
        k = rng.randint(max(2, n // 4), n - max(2, n // 4))
        lens = [a + rng.randint(-8, 8) for _ in range(k)] + [b + rng.randint(-8, 8) for _ in range(n - k)]
        order = list(range(n))
        rng.shuffle(order)
        lens = [lens[i] for i in order]
        truth = [0 if order[i] < k else 1 for i in range(n)]
        d = [math.sqrt(abs(lens[i] - lens[j])) for i in range(n) for j in range(i + 1, n)]
        sel, central, ng = orc.cluster_locus(d, n)
        assert ng == 2 and set(sel) == {0, 1}
        big = 0 if k > n - k else 1          # allele with more reads = group1 (ties: the later group)
        for i in range(n):
            same = truth[i] == truth[sel.index(0)]
            assert (sel[i] == 0) == same
        if k != n - k:
            assert truth[sel.index(0)] == big
        assert sel[central[0]] == 0 and sel[central[1]] == 1


def test_cluster_small_and_homozygous():
    assert orc.cluster_locus([], 1) == ([0], (0, None), 1)
    sel, central, ng = orc.cluster_locus([3.0], 2)   # vec![vec![0], vec![1]]: group1 = [1], group2 = [0]
    assert (sel, central, ng) == ([1, 0], (1, 0), 2)
    n = 10                                            # all equal: no split of >= 2 vs >= 2 above cutoff 0 -> alternate
    sel, central, ng = orc.cluster_locus([0.0] * (n * (n - 1) // 2), n)
    assert ng == 2 and sel == [1, 0] * 5 or sel == [0, 1] * 5


# ---- the device core (trgt_b200/csrc/cluster_core.h) run on the CPU: serial and lock-step lanes -----------

@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_cluster_core_matches_oracle(lanes):
    import ctypes as C
    from tests.emul import build as eb
    emul = C.CDLL(eb.build())
    emul.emu_cluster_locus.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    rng = random.Random(23 + lanes)
    for it in range(120 if lanes else 400):
        n = rng.choice([1, 2, 3, 4, 7, 20, 40, 41, 100])
        kind = rng.choice(["lengths", "edit", "euclid", "zero"])
        d = [0.0] * (n * (n - 1) // 2) if kind == "zero" else random_matrix(rng, n, kind)
        sel, central, ng = orc.cluster_locus(d, n)
        ref_steps, ref_mat = orc.ward_linkage(d, n) if n >= 3 else ([], np.array(d))
        dd = np.array(d + [0.0], dtype=np.float64)
        got_sel = np.full(max(1, n), -1, dtype=np.int32)
        got_c = (C.c_uint32 * 2)()
        got_ng = emul.emu_cluster_locus(dd.ctypes.data, n, got_sel.ctypes.data, got_c, lanes)
        assert got_ng == ng
        assert got_sel[:n].tolist() == sel
        assert tuple(None if c == 0xFFFFFFFF else int(c) for c in got_c) == central
        if n >= 3:   # the matrix the linkage leaves behind (central_read reads it) is bit-identical
            assert np.array_equal(dd[:-1], ref_mat)


def _levenshtein(a: bytes, b: bytes) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_get_dist_matrix_correctness(oracle):  # genotype_cluster.rs:330-356, and each entry against a plain DP
    """The reference's own test: the condensed matrix equals pair-by-pair get_dist (10 random sequences of 50-150
    bases, 1e-9); on top of it every entry is sqrt(Levenshtein) or, past MAX_OPS (`:239-243`), sqrt(|len1 - len2|)."""
    import math
    import random
    rng = random.Random(330)
    seqs = [bytes(rng.choice(b"ACGT") for _ in range(rng.randint(50, 150))) for _ in range(10)]
    got = oracle.get_dist_matrix(seqs)
    assert len(got) == 45
    k = 0
    for i in range(10):
        for j in range(i + 1, 10):
            assert abs(got[k] - oracle.get_dist(seqs[i], seqs[j])) < 1e-9
            if len(seqs[i]) * len(seqs[j]) > 10000:
                assert got[k] == math.sqrt(abs(len(seqs[i]) - len(seqs[j])))
            else:
                assert got[k] == math.sqrt(_levenshtein(seqs[i], seqs[j]))
            k += 1
