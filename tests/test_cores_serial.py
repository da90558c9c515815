"""CPU checks of the kernel cores (trgt_b200/csrc/{hmm,wfa}_core.h) in their one-lane serial
instantiation against the oracle.  The same templates are what the CUDA kernels instantiate with
warp / CTA groups; the GPU parity tests proper are in test_engine_gpu.py."""
import ctypes as C
import json
import math
import os
import random

import numpy as np

import pytest

from tests.emul import build as emul_build

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _Span(C.Structure):
    _fields_ = [("m", C.c_uint32), ("s", C.c_uint32), ("e", C.c_uint32)]


@pytest.fixture(scope="module")
def emul():
    L = C.CDLL(emul_build.build())
    L.emu_hmm_annotate.restype = C.c_long
    return L


def emu_hmm(L, motifs, allele):
    data = b"".join(motifs)
    offs = [0]
    for m in motifs:
        offs.append(offs[-1] + len(m))
    moff = (C.c_uint64 * len(offs))(*offs)
    mc = (C.c_uint32 * (len(motifs) + 1))()
    cap = len(allele) + 2
    spans = (_Span * cap)()
    pur = C.c_double()
    pcap = (len(allele) + 2) * 64
    path = (C.c_uint32 * pcap)()
    plen, S = C.c_uint64(), C.c_int()
    n = L.emu_hmm_annotate(data, moff, len(motifs), allele, len(allele), mc, spans, cap, C.byref(pur), path,
                           C.c_uint64(pcap), C.byref(plen), C.byref(S))
    assert n >= 0, n
    return (list(mc[:len(motifs)]), [(spans[i].m, spans[i].s, spans[i].e) for i in range(n)], pur.value,
            list(path[:plen.value]), S.value)


def emu_wfa(L, p, t, x, o, e, ef=None):
    out = (C.c_int * 9)()
    cap = 2 * (len(p) + len(t)) + 16
    words = (C.c_uint32 * cap)()
    pbf, pef, tbf, tef = ef if ef else (0, 0, 0, 0)
    rc = L.emu_wfa_align(p, len(p), t, len(t), x, o, e, pbf, pef, tbf, tef, out, words, cap)
    return rc, list(out), list(words[:out[7]])


def rnd(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n)).encode()


def mutate(rng, s, rate):
    o = []
    for c in s:
        r = rng.random()
        if r < rate / 3:
            o.append(rng.choice(b"ACGT"))
        elif r < 2 * rate / 3:
            pass
        elif r < rate:
            o.append(c)
            o.append(rng.choice(b"ACGT"))
        else:
            o.append(c)
    return bytes(o)


def noisy_repeat(rng, motifs, max_units=12):
    parts = []
    for _ in range(rng.randint(0, max_units)):
        m = rng.choice(motifs).replace(b"N", rng.choice([b"A", b"C", b"G", b"T"]))
        r = rng.random()
        if r < 0.7:
            parts.append(m)
        elif r < 0.8:
            parts.append(rnd(rng, rng.randint(1, 5)))
        elif r < 0.9:
            parts.append(m[:-1])
        else:
            parts.append(m + rnd(rng, 1))
    return b"".join(parts)


def test_hmm_core_random(emul, oracle):
    rng = random.Random(1)
    for _ in range(1200):
        k = rng.choice([1, 1, 1, 2, 3, 5])
        motifs = [rnd(rng, rng.choice([1, 2, 2, 3, 4, 5, 6, 7, 12]), "ACGTN" if rng.random() < 0.2 else "ACGT")
                  for _ in range(k)]
        allele = noisy_repeat(rng, motifs)
        if rng.random() < 0.1:
            allele = allele[:3] + b"N" + allele[3:] + b"X"
        h = oracle.Hmm([oracle.replace_invalid_bases(m, b"ATCGN") for m in motifs])
        exp_mc, exp_sp, exp_pur = h.annotate(allele)
        exp_path = h.label(oracle.replace_invalid_bases(allele, b"ATCG")) if allele else []
        mc, sp, pur, path, S = emu_hmm(emul, motifs, allele)
        assert S == h.num_states
        assert path == exp_path
        assert mc == exp_mc and sp == exp_sp
        assert (math.isnan(pur) and math.isnan(exp_pur)) or pur == exp_pur


def test_hmm_core_reference_goldens(emul):
    # src/hmm/builder.rs:208-217 and :243-250 (spans after remove_imperfect_motifs are what annotate collapses)
    mc, sp, pur, _, _ = emu_hmm(emul, [b"CAG", b"A"], b"CAGCAGCAGCAGAAAAA")
    assert sp == [(0, 0, 12), (1, 12, 17)] and mc == [4, 5] and pur == 1.0
    # docs/tutorial.md:44: MC=11, MS=0(0-33), AP=1.000000
    mc, sp, pur, _, _ = emu_hmm(emul, [b"CAG"], b"CAG" * 11)
    assert mc == [11] and sp == [(0, 0, 33)] and pur == 1.0


def test_wfa_core_random(emul, oracle):
    rng = random.Random(7)
    for _ in range(1500):
        x, o, e = rng.choice([(2, 5, 1), (1, 0, 1), (4, 6, 2), (6, 4, 2), (1, 5, 1), (3, 1, 3)])
        mode = rng.random()
        if mode < 0.5:  # the production flank shape: pattern global, text free at both ends
            p = rnd(rng, rng.randint(1, 40))
            t = rnd(rng, rng.randint(0, 60)) + mutate(rng, p, rng.choice([0, 0.05, 0.2, 0.6])) + rnd(rng, rng.randint(0, 60))
            if rng.random() < 0.1:
                t = t[:rng.randint(0, len(t))]
            t = t or b"A"
            ef = (0, 0, len(t), len(t))
        elif mode < 0.8:  # end-to-end
            p = rnd(rng, rng.randint(1, 50))
            t = mutate(rng, p, rng.choice([0, 0.05, 0.2, 0.5])) if rng.random() < 0.8 else rnd(rng, rng.randint(1, 50))
            t = t or b"C"
            ef = None
        else:  # arbitrary ends-free allowances
            p, t = rnd(rng, rng.randint(1, 30)), rnd(rng, rng.randint(1, 30))
            ef = (rng.randint(0, len(p)), rng.randint(0, len(p)), rng.randint(0, len(t)), rng.randint(0, len(t)))
        a = oracle.wfa_align(p, t, oracle.AFFINE, x, o, e, ends_free=ef)
        rc, out, words = emu_wfa(emul, p, t, x, o, e, ef)
        assert rc == 0
        assert (out[1], out[2], out[3]) == (a.score, a.end_k, a.end_offset)
        assert out[4] == a.count_matches()
        if ef is not None:
            assert (out[5], out[6]) == a.alignment_span()[1]
        assert words == a.sam_cigar(True)


def test_wfa_core_reference_goldens(emul):
    P = b"AGCTAGTGTCAATGGCTACTTTTCAGGTCCT"       # src/wfaligner.rs:1133
    T = b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT"   # src/wfaligner.rs:1134

    def cig(words):
        return "".join(f"{w >> 4}{'MIDNSHP=X'[w & 15]}" for w in words).replace("=", "M")

    rc, out, words = emu_wfa(emul, P, T, 6, 4, 2)            # wfaligner.rs:1185-1198
    assert rc == 0 and out[1] == -40 and cig(words) == "1M1X3M1I5M2X8M3I1M1X9M"
    # wfaligner.rs:1795-1828: text free at both ends, the production flank call shape
    p = b"ACGTACGTACGTACG"
    t = b"T" * 10 + p + b"T" * 10
    rc, out, words = emu_wfa(emul, p, t, 2, 5, 1, (0, 0, len(t), len(t)))
    assert rc == 0 and out[1] == 0 and cig(words) == "10I15M10I" and (out[4], out[5], out[6]) == (15, 10, 25)


def test_flank_scan_core(emul):
    rng = random.Random(3)
    for _ in range(1500):
        P = rng.randint(1, 8)
        piece = rnd(rng, P) if rng.random() < 0.5 else b"A" * P
        t = rnd(rng, rng.randint(0, 70), "AC") if rng.random() < 0.5 else rnd(rng, rng.randint(0, 70))
        if rng.random() < 0.5 and len(t) >= P:
            i = rng.randint(0, len(t) - P)
            t = t[:i] + piece + t[i + P:]
        assert emul.emu_flank_scan(piece, P, t, len(t)) == t.find(piece)


def test_edit_distance_core(emul, oracle):
    rng = random.Random(5)

    def lev(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]

    for _ in range(800):
        a = rnd(rng, rng.randint(0, 128), "ACGTN")
        b = mutate(rng, a, rng.choice([0, 0.1, 0.5])) if rng.random() < 0.7 else rnd(rng, rng.randint(0, 200))
        if rng.random() < 0.1:
            b += b"XQ"
        if min(len(a), len(b)) > 128:
            continue
        got = emul.emu_edit_distance(a, len(a), b, len(b))
        assert got == lev(a, b)
        if a and b and len(a) * len(b) <= 10000:
            assert math.sqrt(got) == oracle.get_dist(a, b)


def test_flank_banded_core(emul, oracle):
    """Seed filter + banded score pass + cone trace must give what the full-width alignment gives
    whenever it claims to have settled the pair (rc 0), on clean, noisy and repetitive flanks."""
    emul.emu_flank_banded.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_int, C.POINTER(C.c_int)]
    rng = random.Random(77)
    resolved = 0
    for _ in range(2500):
        x, o, e = rng.choice([(2, 5, 1), (2, 5, 1), (1, 0, 1), (4, 6, 2), (3, 1, 3)])
        P = rng.choice([60, 120, 250])
        kind = rng.random()
        if kind < 0.3:  # low-complexity flank: many seed occurrences on many diagonals
            unit = rnd(rng, rng.randint(1, 7))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], 0.02)[:P]
        else:
            p = rnd(rng, P)
        pre, suf = rnd(rng, rng.randint(0, 300)), rnd(rng, rng.randint(0, 300))
        if kind < 0.3 and rng.random() < 0.5:
            pre += p[:len(p) // 2]
        body = mutate(rng, p, rng.choice([0.004, 0.01, 0.03, 0.08]))
        if rng.random() < 0.15:
            body = body[:len(body) // 2] + rnd(rng, rng.randint(1, 12)) + body[len(body) // 2:]
        t = pre + body + suf
        if rng.random() < 0.1:
            t = t[:rng.randint(len(pre) + 10, len(t))]
        S = rng.choice([8, 16, 20, 24])
        exp, via, nm = oracle.find_span(p, t, (x, o, e), len(p) * 0.7)
        if via == 1:
            continue
        out = (C.c_int * 6)()
        rc = emul.emu_flank_banded(p, len(p), t, len(t), x, o, e, S, 0.7, 8192, out)
        if rc == 0:
            resolved += 1
            assert (out[1], out[2]) == (via, nm)
            if exp is not None:
                assert (out[4], out[5]) == exp
    assert resolved > 300


def test_flank_banded_straddling_gaps(emul, oracle):
    """A scoring with a cheap gap open (3,3,1): two-base gaps that straddle block boundaries damage two seed
    blocks for o+2e < 2*min(x,o+e), so the block count of the seed filter must account for them
    (flank_seed_blocks); a decoy copy with one intact block must not capture the band."""
    emul.emu_flank_banded.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_int, C.POINTER(C.c_int)]
    rng = random.Random(5)
    x, o, e, S, P = 3, 3, 1, 24, 250
    nb_naive = S // min(x, o + e) + 1
    blen = P // nb_naive
    resolved = 0
    for _ in range(200):
        p = rnd(rng, P)
        true = bytearray(p)
        cuts = sorted(rng.sample(range(1, nb_naive), 4), reverse=True)
        touched = set()
        for b in cuts:
            touched |= {b - 1, b}
        mb = [b for b in range(nb_naive) if b not in touched][0]
        pos = mb * blen + blen // 2
        true[pos] = ord("A") if true[pos] != ord("A") else ord("C")
        for b in cuts:
            del true[b * blen - 1: b * blen + 1]
        decoy = bytearray(p)
        for b in range(8):
            q = b * blen + blen // 2
            decoy[q] = ord("A") if decoy[q] != ord("A") else ord("C")
        t = rnd(rng, 200) + bytes(true) + rnd(rng, 400) + bytes(decoy) + rnd(rng, 300)
        exp, via, nm = oracle.find_span(p, t, (x, o, e), P * 0.7)
        out = (C.c_int * 6)()
        rc = emul.emu_flank_banded(p, P, t, len(t), x, o, e, S, 0.7, 65536, out)
        if rc == 0:
            resolved += 1
            assert (out[1], out[2]) == (via, nm)
            if exp is not None:
                assert (out[4], out[5]) == exp
    assert resolved > 50


# ---- the same cores with several lock-step host lanes (tests/emul/lanes.h): leader election,
# ---- broadcasts and barrier placement are exercised, not just the index arithmetic --------------

@pytest.mark.parametrize("lanes", [4, 32])
def test_wfa_core_multilane(emul, oracle, lanes):
    rng = random.Random(100 + lanes)
    for it in range(120):
        x, o, e = rng.choice([(2, 5, 1), (1, 0, 1), (4, 6, 2)])
        if it % 2 == 0:
            p = rnd(rng, rng.randint(20, 120))
            t = rnd(rng, rng.randint(0, 200)) + mutate(rng, p, rng.choice([0, 0.02, 0.1])) + rnd(rng, rng.randint(0, 200))
            ef = (0, 0, len(t), len(t))
        else:
            p = rnd(rng, rng.randint(1, 150))
            t = mutate(rng, p, rng.choice([0, 0.03, 0.2])) or b"C"
            ef = None
        a = oracle.wfa_align(p, t, oracle.AFFINE, x, o, e, ends_free=ef)
        out = (C.c_int * 9)()
        cap = 2 * (len(p) + len(t)) + 16
        words = (C.c_uint32 * cap)()
        pbf, pef, tbf, tef = ef if ef else (0, 0, 0, 0)
        rc = emul.emu_wfa_align_lanes(p, len(p), t, len(t), x, o, e, pbf, pef, tbf, tef, out, words, cap, lanes)
        assert rc == 0
        assert (out[1], out[2], out[3], out[4]) == (a.score, a.end_k, a.end_offset, a.count_matches())
        assert list(words[:out[7]]) == a.sam_cigar(True)


@pytest.mark.parametrize("lanes", [4, 32])
def test_flank_locate_multilane(emul, oracle, lanes):
    emul.emu_flank_banded_lanes.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.c_int]
    rng = random.Random(200 + lanes)
    resolved = 0
    for it in range(150):
        P = rng.choice([120, 250])
        if it % 3 == 0:
            unit = rnd(rng, rng.randint(2, 7))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], 0.03)[:P]
        else:
            p = rnd(rng, P)
        t = rnd(rng, rng.randint(0, 400)) + mutate(rng, p, rng.choice([0.0, 0.004, 0.01, 0.03])) + rnd(rng, rng.randint(260, 500))
        assert emul.emu_flank_scan_lanes(p, len(p), t, len(t), lanes) == t.find(p)
        exp, via, nm = oracle.find_span(p, t, (2, 5, 1), len(p) * 0.7)
        if via == 1:
            continue
        out = (C.c_int * 6)()
        rc = emul.emu_flank_banded_lanes(p, len(p), t, len(t), 2, 5, 1, 20, 0.7, 1536, out, lanes)
        if rc == 0:
            resolved += 1
            assert (out[1], out[2]) == (via, nm)
            if exp is not None:
                assert (out[4], out[5]) == exp
    assert resolved > 15


@pytest.mark.parametrize("lanes", [3, 32])
def test_hmm_core_multilane(emul, oracle, lanes):
    emul.emu_hmm_annotate_lanes.restype = C.c_long
    rng = random.Random(300 + lanes)
    for _ in range(150):  # S <= 32 models take the one-state-per-lane variant when lanes == 32
        k = rng.choice([1, 1, 2, 5])
        motifs = [rnd(rng, rng.choice([1, 2, 3, 4, 6, 12]), "ACGTN" if rng.random() < 0.2 else "ACGT") for _ in range(k)]
        allele = noisy_repeat(rng, motifs) or b"A"
        h = oracle.Hmm([oracle.replace_invalid_bases(m, b"ATCGN") for m in motifs])
        exp_mc, exp_sp, exp_pur = h.annotate(allele)
        data = b"".join(motifs)
        offs = [0]
        for m in motifs:
            offs.append(offs[-1] + len(m))
        moff = (C.c_uint64 * len(offs))(*offs)
        mc = (C.c_uint32 * (len(motifs) + 1))()
        cap = len(allele) + 2
        spans = (_Span * cap)()
        pur, plen, S = C.c_double(), C.c_uint64(), C.c_int()
        n = emul.emu_hmm_annotate_lanes(data, moff, len(motifs), allele, len(allele), mc, spans, cap, C.byref(pur),
                                        None, C.c_uint64(0), C.byref(plen), C.byref(S), lanes)
        assert n >= 0
        assert list(mc[:len(motifs)]) == exp_mc
        assert [(spans[i].m, spans[i].s, spans[i].e) for i in range(n)] == exp_sp
        assert pur.value == exp_pur


@pytest.mark.parametrize("lanes", [0, 5, 32])
def test_flank_indexed_core(emul, oracle, lanes):
    """Exact search and seed filter through the per-piece 8-mer index (probe positions only), incl.
    repetitive pieces (long hash clusters), two copies of the piece, truncated reads, and scratch too
    small for the in-band history (ring pass + cone trace instead)."""
    emul.emu_flank_indexed.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.c_int, C.POINTER(C.c_int), C.c_int]
    rng = random.Random(500 + lanes)
    seen = {"exact": 0, "resolved": 0, "deferred": 0}
    for _ in range(1500 if lanes == 0 else 120):
        x, o, e = rng.choice([(2, 5, 1), (2, 5, 1), (1, 0, 1), (4, 6, 2)])
        P = rng.choice([60, 120, 200, 250, 250, 380])
        kind = rng.random()
        if kind < 0.3:
            unit = rnd(rng, rng.randint(1, 9))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], rng.choice([0, 0.02, 0.05]))[:P]
            if len(p) < 16:
                continue
        else:
            p = rnd(rng, P)
        pre, suf = rnd(rng, rng.randint(0, 600)), rnd(rng, rng.randint(0, 600))
        if kind < 0.3 and rng.random() < 0.5:
            pre += p[:len(p) // 2]
        body = p if rng.random() < 0.3 else mutate(rng, p, rng.choice([0.004, 0.01, 0.03, 0.08]))
        if rng.random() < 0.1:
            body += rnd(rng, 20) + p
        t = pre + body + suf
        if rng.random() < 0.05:
            t = t[:rng.randint(1, len(t))]
        S = rng.choice([8, 20, 20, 24])
        exp, via, nm = oracle.find_span(p, t, (x, o, e), len(p) * 0.7)
        out = (C.c_int * 7)()
        emul.emu_flank_indexed(p, len(p), t, len(t), x, o, e, S, 0.7, rng.choice([700, 1280, 8192]), out, lanes)
        assert out[0] == t.find(p)
        if out[0] >= 0:
            seen["exact"] += 1
            assert via == 1
        elif out[1] == 0:
            seen["resolved"] += 1
            assert (out[2], out[3]) == (via, nm)
            if exp is not None:
                assert (out[5], out[6]) == exp
        else:
            seen["deferred"] += 1
    assert seen["exact"] > 10 and seen["resolved"] > 10


def test_e2e_narrow_core(emul, oracle):
    """Short consensus pairs: the [-R, R] band with cost cap S is the full computation (nothing is
    reachable outside it), so its history back-traces to the reference's CIGAR."""
    rng = random.Random(909)
    resolved = 0
    for _ in range(250):
        unit = rnd(rng, rng.randint(2, 6))
        p = unit * rng.randint(1, 25) if rng.random() < 0.7 else rnd(rng, rng.randint(1, 80))
        r = rng.random()
        t = (mutate(rng, p, rng.choice([0.02, 0.05, 0.15])) if r < 0.7 else
             p + unit * rng.randint(1, 4) if r < 0.85 else p[:len(p) - len(unit) * rng.randint(0, 2)]) or b"A"
        words, score = oracle.align_words(p, t)
        out = (C.c_int * 3)()
        cap = len(p) + len(t) + 8
        w = (C.c_uint32 * cap)()
        rc = emul.emu_e2e_narrow(p, len(p), t, len(t), 2, 5, 1, 16, 1280, out, w, cap, 32)
        if rc == 0:
            resolved += 1
            assert out[1] == score and list(w[:out[2]]) == words
        else:
            assert -score > 16 or rc == -200
    assert resolved > 100


def test_flank_scan_core_reference_vectors(emul):
    """The reference's exact-search vectors (span_locater.rs:73-119) through the device core of the general exact
    search (flank_scan, as k_flank_exact runs it for pieces outside 16..256 bases), serial and with 7 lock-step lanes."""
    cases = [(b"ABCDEFG", b"CDE", 2), (b"ABCDEFG", b"XYZ", -1), (b"ABCABCABC", b"ABC", 0), (b"ABCDEFG", b"ABC", 0),
             (b"ABCDEFG", b"EFG", 4), (b"ABCDEFG", b"ABCDEFG", 0), (b"ABC", b"ABCDEFG", -1), (b"", b"ABC", -1),
             (b"ABCDEFG", b"A", 0), (b"ABCDEFG", b"G", 6), (b"ABCDEFG", b"D", 3), (b"AAAAA", b"AA", 0),
             (b"ACGTNACGT", b"N", 4)]
    for seq, piece, start in cases:
        assert emul.emu_flank_scan(piece, len(piece), seq, len(seq)) == start, (seq, piece)
        assert emul.emu_flank_scan_lanes(piece, len(piece), seq, len(seq), 7) == start, (seq, piece)


def test_e2e_lane_core(emul, oracle):
    """Phase B's lane routine (e2e_narrow_lane: cost cap 8 on |k| <= 3, 16-bit history rows only for the scores the
    scoring allows, cells in lockstep with parked extensions): whatever it settles is the reference's CIGAR and score,
    and it settles every pair whose cost is at most 8."""
    rng = random.Random(4711)
    resolved = 0
    for _ in range(1500):
        unit = rnd(rng, rng.randint(2, 6))
        p = unit * rng.randint(1, 25) if rng.random() < 0.7 else rnd(rng, rng.randint(1, 80))
        r = rng.random()
        t = (mutate(rng, p, rng.choice([0.01, 0.02, 0.05])) if r < 0.7 else
             p + unit * rng.randint(1, 2) if r < 0.85 else p[:len(p) - rng.randint(0, 3)]) or b"A"
        words, score = oracle.align_words(p, t)
        out = (C.c_int * 3)()
        cap = len(p) + len(t) + 8
        w = (C.c_uint32 * cap)()
        rc = emul.emu_e2e_lane(p, len(p), t, len(t), 2, 5, 1, 8, out, w, cap)
        if rc == 0:
            resolved += 1
            assert out[1] == score and list(w[:out[2]]) == words, (p, t)
        else:
            assert -score > 8, (p, t, score)
    assert resolved > 700


def test_hmm_core_thread_per_allele(emul, oracle):
    """hmm_viterbi_thread (one lane does the whole allele, strided score columns, table-free model)
    must give the oracle's state path, MC, MS and AP."""
    emul.emu_hmm_annotate_lanes.restype = C.c_long
    rng = random.Random(41)
    for _ in range(700):
        k = rng.choice([1, 1, 1, 2, 3, 5])
        motifs = [rnd(rng, rng.choice([1, 2, 2, 3, 4, 5, 6, 7, 12]), "ACGTN" if rng.random() < 0.2 else "ACGT")
                  for _ in range(k)]
        allele = noisy_repeat(rng, motifs)
        if rng.random() < 0.1:
            allele = allele[:3] + b"N" + allele[3:] + b"X"
        if not allele:
            continue
        h = oracle.Hmm([oracle.replace_invalid_bases(m, b"ATCGN") for m in motifs])
        exp_mc, exp_sp, exp_pur = h.annotate(allele)
        exp_path = h.label(oracle.replace_invalid_bases(allele, b"ATCG"))
        data = b"".join(motifs)
        offs = [0]
        for m in motifs:
            offs.append(offs[-1] + len(m))
        moff = (C.c_uint64 * len(offs))(*offs)
        mc = (C.c_uint32 * (len(motifs) + 1))()
        cap = len(allele) + 2
        spans = (_Span * cap)()
        pur, plen, S = C.c_double(), C.c_uint64(), C.c_int()
        pcap = (len(allele) + 2) * 64
        path = (C.c_uint32 * pcap)()
        n = emul.emu_hmm_annotate_lanes(data, moff, len(motifs), allele, len(allele), mc, spans, cap, C.byref(pur),
                                        path, C.c_uint64(pcap), C.byref(plen), C.byref(S), -1)
        assert n >= 0
        assert list(path[:plen.value]) == exp_path
        assert list(mc[:len(motifs)]) == exp_mc
        assert [(spans[i].m, spans[i].s, spans[i].e) for i in range(n)] == exp_sp and pur.value == exp_pur


def test_hmm_core_single_motif_lane(emul, oracle):
    """hmm_viterbi_lane<N> (single-motif loci: score column in registers, one packed back-pointer word per
    column, ms / skip_ms / rs folded into re) + the walk over the packed words must give the oracle's state
    path, MC, MS and AP for every motif length the fast path takes (1..8), 'N' motif bases, invalid allele
    bases, interruptions, and long alleles."""
    emul.emu_hmm_annotate_lanes.restype = C.c_long
    rng = random.Random(4242)
    seen = set()
    for it in range(1500):
        n = rng.choice([1, 2, 2, 2, 3, 3, 4, 4, 5, 6, 6, 7, 7])
        motif = rnd(rng, n, "ACGTN" if rng.random() < 0.15 else "ACGT")
        r = rng.random()
        if r < 0.75:
            allele = noisy_repeat(rng, [motif], max_units=rng.choice([3, 12, 40]))
        elif r < 0.85:   # unrelated sequence: the skip block carries it
            allele = rnd(rng, rng.randint(1, 60))
        else:            # repeat - interruption - repeat
            unit = oracle.replace_invalid_bases(motif, b"ATCG") if b"N" in motif else motif
            allele = unit * rng.randint(1, 9) + rnd(rng, rng.randint(1, 9)) + unit * rng.randint(1, 9)
        if rng.random() < 0.1:
            allele = allele[:3] + b"N" + allele[3:] + b"X"
        if it % 300 == 299:
            allele = (allele or b"A") * 40  # a few long ones
        if not allele:
            continue
        seen.add(n)
        h = oracle.Hmm([oracle.replace_invalid_bases(motif, b"ATCGN")])
        exp_mc, exp_sp, exp_pur = h.annotate(allele)
        exp_path = h.label(oracle.replace_invalid_bases(allele, b"ATCG"))
        moff = (C.c_uint64 * 2)(0, n)
        mc = (C.c_uint32 * 2)()
        cap = len(allele) + 2
        spans = (_Span * cap)()
        pur, plen, S = C.c_double(), C.c_uint64(), C.c_int()
        pcap = (len(allele) + 2) * 8 + 64
        path = (C.c_uint32 * pcap)()
        got = emul.emu_hmm_annotate_lanes(motif, moff, 1, allele, len(allele), mc, spans, cap, C.byref(pur),
                                          path, C.c_uint64(pcap), C.byref(plen), C.byref(S), -2)
        assert got >= 0, got
        assert list(path[:plen.value]) == exp_path
        assert list(mc[:1]) == exp_mc
        assert [(spans[i].m, spans[i].s, spans[i].e) for i in range(got)] == exp_sp
        assert pur.value == exp_pur or (math.isnan(pur.value) and math.isnan(exp_pur))
    assert seen == set(range(1, 8))


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_consensus_core(emul, oracle, lanes):
    """consensus_core.h (column vote + majority insertion) against the oracle's repair_consensus."""
    import numpy as np
    emul.emu_consensus.restype = C.c_long
    emul.emu_consensus.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_uint64, C.c_int]
    rng = random.Random(700 + lanes)
    for _ in range(120 if lanes == 0 else 40):
        unit = rnd(rng, rng.randint(2, 6))
        truth = unit * rng.randint(2, 20)
        pos = rng.randint(0, len(truth))
        truth2 = truth[:pos] + rnd(rng, rng.randint(1, 4)) + truth[pos:]
        seqs = [(mutate(rng, truth2 if rng.random() < 0.6 else truth, rng.choice([0, 0.02, 0.1])) or b"A")
                for _ in range(rng.randint(1, 12))]
        bb = rng.choice(seqs) if rng.random() < 0.7 else truth
        offs, words, woff = [0], [], [0]
        for s in seqs:
            offs.append(offs[-1] + len(s))
            w, _ = oracle.align_words(bb, s)
            words += w
            woff.append(len(words))
        so, wa, wo = np.array(offs, dtype=np.uint64), np.array(words or [0], dtype=np.uint32), np.array(woff, dtype=np.uint64)
        cap = len(bb) + sum(map(len, seqs)) + 8
        out = np.zeros(cap, dtype=np.uint8)
        n = emul.emu_consensus(len(bb), b"".join(seqs), so.ctypes.data, len(seqs), wa.ctypes.data, wo.ctypes.data,
                               out.ctypes.data, cap, lanes)
        assert n >= 0
        assert out[:n].tobytes() == oracle.repair_consensus(bb, seqs)


# ------------------------------------------------------------- producer row: clip + decode --

class _Clip(C.Structure):
    _fields_ = [("ref_start", C.c_int64), ("query_start", C.c_uint64), ("query_end", C.c_uint64),
                ("first_op", C.c_uint32), ("n_ops", C.c_uint32), ("first_word", C.c_uint32),
                ("last_word", C.c_uint32), ("status", C.c_int32)]


def random_bam_cigar(rng, n_ops):
    """HiFi-like CIGAR: optional soft clips at the ends, =/X/I/D/M/N inside"""
    ops = []
    if rng.random() < 0.3:
        ops.append((rng.randint(1, 30) << 4) | 4)
    for _ in range(n_ops):
        op = rng.choice([7, 7, 7, 8, 1, 2, 0, 3])
        ops.append((rng.randint(1, 40) << 4) | op)
    if rng.random() < 0.3:
        ops.append((rng.randint(1, 30) << 4) | 4)
    if rng.random() < 0.05:
        ops.insert(0, (5 << 4) | 5)
    return ops


def test_clip_core_random(emul, oracle):
    import numpy as np
    rng = random.Random(77)
    for it in range(3000):
        ops = random_bam_cigar(rng, rng.randint(0, 12))
        ref_pos = rng.randint(0, 200)
        a = rng.randint(0, 400)
        region = (a, a + rng.randint(0, 300))
        arr = np.array(ops if ops else [0], dtype=np.uint32)
        got = _Clip()
        emul.emu_clip_cigar(arr.ctypes.data_as(C.c_void_p), C.c_uint32(len(ops)), C.c_longlong(ref_pos),
                            C.c_longlong(region[0]), C.c_longlong(region[1]), C.byref(got))
        exp = oracle.clip_cigar(ops, ref_pos, region)
        if exp is None:
            assert got.status == 0, (it, ops, ref_pos, region)
            continue
        assert got.status == 1
        words = [got.first_word if i == 0 else got.last_word if i == got.n_ops - 1 else ops[got.first_op + i]
                 for i in range(got.n_ops)]
        assert (got.ref_start, got.query_start, got.query_end, words) == exp, (it, ops, ref_pos, region)


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_seq4_unpack_core(emul, oracle, lanes):
    import numpy as np
    rng = random.Random(5 + lanes)
    for it in range(40):
        n = rng.randint(0, 12)
        seqs = [rnd(rng, rng.choice([0, 1, 2, 15, 16, 17, 31, 33, rng.randint(0, 700)]), "ACGTN=MRSVWYHKDB")
                for _ in range(n)]
        packed = oracle.encode_seq4(b"")  # built below with arbitrary nibble starts
        nibs, starts = [], []
        for s_ in seqs:
            nibs += [0] * rng.randint(0, 5)
            starts.append(len(nibs))
            nibs += [oracle.SEQ4_ALPHABET.index(c) for c in s_]
        if len(nibs) & 1:
            nibs.append(0)
        packed = bytes((nibs[i] << 4) | nibs[i + 1] for i in range(0, len(nibs), 2))
        buf = np.zeros(len(packed) + 48, dtype=np.uint8)
        base = (-buf.ctypes.data) % 16 + 16      # 16 bytes of padding in front, data 16-byte aligned or not
        base += rng.randint(0, 7)                # the packed buffer itself may sit at any address
        buf[base:base + len(packed)] = np.frombuffer(packed, dtype=np.uint8)
        lens = np.array([len(s_) for s_ in seqs] + [0], dtype=np.uint32)
        st = np.array(starts + [0], dtype=np.uint64)
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens[:n], dtype=np.uint64)
        out = np.zeros(int(offs[n]) + 64, dtype=np.uint8)
        ob = (-out.ctypes.data) % 16
        out[:] = 0x7E
        emul.emu_seq4_unpack(C.c_void_p(buf.ctypes.data + base), st.ctypes.data_as(C.c_void_p),
                             lens.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                             C.c_void_p(out.ctypes.data + ob), C.c_int(lanes))
        got = out[ob:ob + int(offs[n])].tobytes()
        assert got == b"".join(seqs), (it, seqs)
        assert (out[ob + int(offs[n]):ob + int(offs[n]) + 16] == 0x7E).all() and (out[:ob] == 0x7E).all()
        for i, s_ in enumerate(seqs):
            assert oracle.decode_seq4(packed, starts[i], len(s_)) == s_


@pytest.mark.parametrize("lanes", [0, 32])
def test_flank_exact_thread_core(emul, lanes):
    """Exact search with one lane per (read, flank) pair: shifted piece copies + aligned 16-byte compares,
    at every text alignment, incl. repetitive pieces, several occurrences (the first one counts), an
    occurrence at the very start / end of the read and near misses."""
    import numpy as np
    emul.emu_flank_exact_thread.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    rng = random.Random(900 + lanes)
    hits = 0
    for it in range(1200 if lanes == 0 else 150):
        P = rng.choice([16, 17, 31, 60, 200, 250, 250, 250, 256])
        kind = rng.random()
        if kind < 0.3:
            unit = rnd(rng, rng.randint(1, 9))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], rng.choice([0, 0.02]))[:P]
            if len(p) < 16:
                continue
        else:
            p = rnd(rng, P)
        pre, suf = rnd(rng, rng.choice([0, 1, 7, rng.randint(0, 700)])), rnd(rng, rng.choice([0, 1, rng.randint(0, 700)]))
        r = rng.random()
        body = p if r < 0.6 else mutate(rng, p, 0.01) if r < 0.8 else p[:-1] + (b"A" if p[-1:] != b"A" else b"C")
        if rng.random() < 0.15:
            body += rnd(rng, rng.randint(0, 30)) + p
        if kind < 0.3 and rng.random() < 0.5:
            pre += p[len(p) // 2:]
        t = pre + body + suf
        if rng.random() < 0.05:
            t = t[:rng.randint(1, len(t))]
        pb = np.frombuffer(p + b"\0" * 16, dtype=np.uint8).copy()
        buf = np.full(len(t) + 64, 0x23, dtype=np.uint8)
        off = 16 + rng.randint(0, 15)
        buf[off:off + len(t)] = np.frombuffer(t, dtype=np.uint8)
        got = emul.emu_flank_exact_thread(pb.ctypes.data, len(p), buf.ctypes.data + off, len(t), lanes)
        assert got == t.find(p), (it, P, len(t), off)
        hits += got >= 0
    assert hits > 50


def test_flank_tier1_thread_core(emul, oracle):
    """First cost tier of the flank fallback by one lane per pair: whatever it settles must be the
    reference's answer (score, count_matches, span); what it hands on is checked by the other tiers' tests."""
    emul.emu_flank_tier1_thread.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_double, C.POINTER(C.c_int)]
    rng = random.Random(4242)
    seen = {"settled": 0, "handed_on": 0, "gap": 0, "rejected": 0}
    for it in range(2500):
        x, o, e = rng.choice([(2, 5, 1), (2, 5, 1), (2, 5, 1), (1, 0, 1), (4, 6, 2), (3, 1, 1)])
        P = rng.choice([60, 120, 200, 250, 250, 250, 380])
        kind = rng.random()
        if kind < 0.25:
            unit = rnd(rng, rng.randint(1, 9))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], rng.choice([0, 0.02, 0.05]))[:P]
            if len(p) < 16:
                continue
        else:
            p = rnd(rng, P)
        pre, suf = rnd(rng, rng.randint(0, 600)), rnd(rng, rng.randint(0, 600))
        if kind < 0.25 and rng.random() < 0.5:
            pre += p[:len(p) // 2]
        r = rng.random()
        if r < 0.5:      # exactly one edit: the bulk of what this tier is for
            i = rng.randrange(len(p))
            c = rng.random()
            if c < 0.3:
                body = p[:i] + bytes([rng.choice(b"ACGT")]) + p[i + 1:]
            elif c < 0.65:
                body = p[:i] + p[i + 1:]
            else:
                body = p[:i] + bytes([rng.choice(b"ACGT")]) + p[i:]
            seen["gap"] += c >= 0.3
        else:
            body = mutate(rng, p, rng.choice([0.002, 0.004, 0.01, 0.03]))
        if rng.random() < 0.1:
            body += rnd(rng, 20) + p[:-1]
        t = pre + body + suf
        if rng.random() < 0.05:
            t = t[:rng.randint(1, len(t))]
        if t.find(p) >= 0:
            continue  # exact hits never reach the fallback
        frac = rng.choice([0.7, 0.7, 0.999])
        out = (C.c_int * 6)()
        emul.emu_flank_tier1_thread(p, len(p), t, len(t), x, o, e, 20, frac, out)
        if out[0] != 0:
            seen["handed_on"] += 1
            continue
        seen["settled"] += 1
        exp, via, nm = oracle.find_span(p, t, (x, o, e), len(p) * frac)
        assert (out[1], out[2]) == (via, nm), (it, x, o, e)
        seen["rejected"] += via == 3
        if exp is not None:
            assert (out[4], out[5]) == exp, (it, x, o, e)
    assert seen["settled"] > 700 and seen["handed_on"] > 50 and seen["rejected"] > 5, seen


def test_flank_tier1_split_core(emul, oracle):
    """First cost tier in two passes (seed pass: hull of ALL index hits, unverified; band pass: staged text
    window + 16-bit on-chip history): whatever it settles is the reference's answer, and it settles at least
    every pair the one-pass routine settles with the same result.  Decoys (a second copy of a block elsewhere
    in the read, repetitive pieces) force the verifying path; windows start at every 16-byte phase."""
    emul.emu_flank_tier1_thread.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_double, C.POINTER(C.c_int)]
    emul.emu_flank_tier1_split.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int)]
    rng = random.Random(777)
    seen = {"settled": 0, "handed_on": 0, "rejected": 0, "decoy_settled": 0, "negative_k": 0}
    for it in range(4000):
        x, o, e = rng.choice([(2, 5, 1), (2, 5, 1), (2, 5, 1), (1, 0, 1), (4, 6, 2), (3, 1, 1), (3, 3, 1)])
        P = rng.choice([16, 60, 120, 200, 250, 250, 250, 256])
        kind = rng.random()
        if kind < 0.2:
            unit = rnd(rng, rng.randint(1, 9))
            p = mutate(rng, (unit * (P // len(unit) + 1))[:P], rng.choice([0, 0.02, 0.05]))[:P]
            if len(p) < 16:
                continue
        else:
            p = rnd(rng, P)
        pre, suf = rnd(rng, rng.choice([0, 0, 1, 3, rng.randint(0, 600)])), rnd(rng, rng.randint(0, 600))
        r = rng.random()
        if r < 0.6:
            i = rng.randrange(len(p))
            c = rng.random()
            if c < 0.3:
                body = p[:i] + bytes([rng.choice(b"ACGT")]) + p[i + 1:]
            elif c < 0.65:
                body = p[:i] + p[i + 1:]
            else:
                body = p[:i] + bytes([rng.choice(b"ACGT")]) + p[i:]
        else:
            body = mutate(rng, p, rng.choice([0.002, 0.004, 0.01, 0.03]))
        decoy = rng.random() < 0.25
        if decoy:   # a verbatim stretch of the piece somewhere else in the read: an index hit off the alignment
            a = rng.randrange(max(1, len(p) - 20))
            piece_bit = p[a:a + rng.choice([8, 12, 30, 70])]
            if rng.random() < 0.5:
                suf = suf + piece_bit + rnd(rng, rng.randint(0, 50))
            else:
                pre = piece_bit + rnd(rng, rng.randint(0, 50)) + pre
        t = pre + body + suf
        if rng.random() < 0.05:
            t = t[:rng.randint(1, len(t))]
        if t.find(p) >= 0:
            continue
        frac = rng.choice([0.7, 0.7, 0.999])
        one, two = (C.c_int * 6)(), (C.c_int * 6)()
        emul.emu_flank_tier1_thread(p, len(p), t, len(t), x, o, e, 20, frac, one)
        emul.emu_flank_tier1_split(p, len(p), t, len(t), x, o, e, 20, frac, it % 16, two)
        if one[0] == 0:
            assert two[0] == 0 and list(two) == list(one), (it, list(one), list(two))
        if two[0] != 0:
            seen["handed_on"] += 1
            continue
        seen["settled"] += 1
        seen["decoy_settled"] += decoy
        exp, via, nm = oracle.find_span(p, t, (x, o, e), len(p) * frac)
        assert (two[1], two[2]) == (via, nm), (it, x, o, e)
        seen["rejected"] += via == 3
        if exp is not None:
            assert (two[4], two[5]) == exp, (it, x, o, e)
            seen["negative_k"] += exp[0] == 0
    assert seen["settled"] > 1200 and seen["handed_on"] > 50 and seen["rejected"] > 5 and seen["decoy_settled"] > 100, seen


def _bamlet_expect(oracle, bases, ops, ref_pos, span, flank_len):
    """oracle.bamlet_clip -> the fields of trgt_bamlet_clip_t (status, base range, methylation range, ref_pos, words)"""
    if span is None:
        return (0,)
    try:
        res = oracle.bamlet_clip(bases, ops, ref_pos, span, flank_len)
    except ValueError:
        return (-500,)
    if res is None:
        return (0,)
    b0, b1, m0, m1, cig = res
    return (1, b0, b1, m0, m1) + ((cig[0], cig[1]) if cig is not None else (None, []))


def bamlet_got(c, ops):
    """trgt_bamlet_clip_t (numpy record or ctypes struct fields by name) -> the same tuple"""
    st = int(c["status"])
    if st != 1:
        return (st,)
    n = int(c["n_ops"])
    words = [int(c["first_word"]) if k == 0 else int(c["last_word"]) if k == n - 1 else int(ops[int(c["first_op"]) + k])
             for k in range(n)]
    has_cigar = len(ops) > 0
    return (1, int(c["base_start"]), int(c["base_end"]), int(c["meth_start"]), int(c["meth_end"]),
            int(c["ref_pos"]) if has_cigar else None, words)


def random_bamlet_case(rng):
    """a clipped read with a CIGAR whose query length is the read's, a span and a flank length"""
    n = rng.choice([2, 5, 31, 200, 1100])
    bases = bytes(rng.choice(b"ACGT" if rng.random() < 0.7 else b"CG") for _ in range(n))
    ops = []
    if rng.random() < 0.85:
        left = n
        if rng.random() < 0.5:
            k = rng.randint(1, max(1, min(20, left)))
            ops.append((k << 4) | 4); left -= k
        while left > 0:
            op = rng.choice([7, 7, 7, 8, 1, 2, 0, 3])
            k = rng.randint(1, max(1, min(60, left))) if op not in (2, 3) else rng.randint(1, 9)
            ops.append((k << 4) | op)
            if op not in (2, 3):
                left -= k
        if rng.random() < 0.3 and (ops[-1] & 15) in (7, 0):   # a soft clip at the end
            k = ops[-1] >> 4
            cut = rng.randint(1, k)
            ops[-1] = ((k - cut) << 4) | (ops[-1] & 15) if k > cut else (cut << 4) | 4
            if k > cut:
                ops.append((cut << 4) | 4)
        if rng.random() < 0.05:
            ops.append((3 << 4) | 7)    # alignment longer than the read: still fine for clip_bases
        if rng.random() < 0.05 and len(ops) > 1:
            ops = ops[:-1]               # alignment shorter than the read: the assert fires for long clips
    a = rng.randint(0, n)
    b = rng.randint(a, n)
    span = None if rng.random() < 0.1 else (a, b)
    flank = rng.choice([0, 1, 3, 50, 250, n // 3])
    return bases, ops, rng.randint(0, 10 ** 6), span, flank


@pytest.mark.parametrize("lanes", [0, 7, 32])
def test_bamlet_clip_core(emul, oracle, lanes):
    """clip_bases for the BAMlet (write_bam.rs:72-92, clip_bases.rs:9-119) by a group of lanes: the reference's own
    vectors (clip_bases.rs:147-230, with the span / flank length that asks for each clip) and random reads"""
    from trgt_b200 import BAMLET_CLIP_DTYPE
    emul.emu_bamlet_clip.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                     C.c_uint32, C.c_longlong, C.c_int, C.c_void_p]
    emul.emu_bamlet_clip.restype = None

    def run(bases, ops, ref_pos, span, flank):
        arr = np.array(ops if ops else [0], dtype=np.uint32)
        out = np.zeros(1, dtype=BAMLET_CLIP_DTYPE)
        emul.emu_bamlet_clip(bases, len(bases), span is not None, span[0] if span else 0, span[1] if span else 0, flank,
                             arr.ctypes.data, len(ops), ref_pos, lanes, out.ctypes.data)
        return bamlet_got(out[0], ops)

    read, cigar = b"AAAAACGCTCGTTAAATCACGAAAAAAAAAA", oracle.encode_bam_cigar("5S3=2D2=1X2=5I3=10S")
    text = lambda words: "".join(f"{w >> 4}{oracle.BAM_OPS[w & 15]}" for w in words)
    for (left, right), exp in [((3, 0), (3, 31, 0, 3, 10, "2S3=2D2=1X2=5I3=10S")), ((10, 0), (10, 31, 2, 3, 17, "1X2=5I3=10S")),
                               ((0, 15), (0, 16, 0, 2, 10, "5S3=2D2=1X2=3I")), ((8, 11), (8, 20, 1, 3, 13, "2D2=1X2=5I2=")),
                               ((13, 13), (13, 18, 2, 2, 20, "5I"))]:
        # span and flank length with span.0 - flank = left and len - span.1 - flank = right
        flank = 0
        got = run(read, cigar, 10, (left + flank, len(read) - right - flank), flank)
        assert got[:6] + (text(got[6]),) == (1,) + exp
    rng = random.Random(31 + lanes)
    seen = {1: 0, 0: 0, -500: 0}
    for _ in range(1500 if lanes == 0 else 150):
        bases, ops, ref_pos, span, flank = random_bamlet_case(rng)
        exp = _bamlet_expect(oracle, bases, ops, ref_pos, span, flank)
        assert run(bases, ops, ref_pos, span, flank) == exp
        seen[exp[0]] += 1
    assert seen[1] > 30 and seen[0] > 30, seen


def test_vcf_fixed6_core(emul):
    """{:.6} of the VCF writer (write_vcf.rs:339) by exact 128-bit integer arithmetic: identical to the
    correctly rounded decimal (Python's format, itself exact, ties to even) on purity-like ratios, on exact
    ties such as 1 - 1/128 and on random doubles"""
    import struct
    emul.emu_vcf_fixed6.argtypes = [C.c_double, C.c_char_p]
    rng = random.Random(6)
    vals = [0.0, 1.0, 0.5, 1 - 1 / 128, 1 / 128, 3 / 128, 0.9921875, 0.0078125, 1e-7, 4.9999995e-7, 5e-7, 0.1, 0.85,
            17 / 20, 18 / 26, 11 / 12, 123456.7890125, 2.5e-6 / 2, 1 - 2 ** -53, 2 ** -1074, 7.0, 999999.9999995,
            -0.5, -1e-9, 4503599627370495.5 / 1e6]
    vals += [a / b for b in range(1, 140) for a in range(0, b + 1, max(1, b // 7))]
    vals += [rng.random() for _ in range(3000)]
    vals += [struct.unpack("<d", struct.pack("<Q", rng.getrandbits(62) % (0x433 << 52)))[0] for _ in range(3000)]
    vals += [(2 * k + 1) / 2 ** rng.randint(7, 20) for k in range(400)]   # dyadic: exact decimal ties occur
    n = 0
    for v in vals:
        buf = C.create_string_buffer(64)
        ln = emul.emu_vcf_fixed6(v, buf)
        if ln < 0:
            assert abs(v) * 1e6 >= 2 ** 62
            continue
        assert buf.raw[:ln].decode() == format(v, ".6f"), v
        n += 1
    assert n > 6000
