"""Pins the CPU oracle against every golden vector the reference's own unit tests hold for the
hot path.  Each case cites the reference test it restates."""
import json
import math
import os

import pytest

PATTERN = b"AGCTAGTGTCAATGGCTACTTTTCAGGTCCT"          # src/wfaligner.rs:1133
TEXT = b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT"         # src/wfaligner.rs:1134
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ WFA --

def test_aligner_indel(oracle):  # wfaligner.rs:1137-1150
    a = oracle.wfa_align(PATTERN, TEXT, oracle.INDEL)
    assert a.status == 0 and a.score == 10
    assert a.cigar_string() == "1M1I1D3M1I5M2I2D8M1I1M1I1M1I9M"


def test_aligner_edit(oracle):  # wfaligner.rs:1153-1166
    a = oracle.wfa_align(PATTERN, TEXT, oracle.EDIT)
    assert a.status == 0 and a.score == 7
    assert a.cigar_string() == "1M1X3M1I5M2X8M1I1M1I1M1I9M"


def test_aligner_gap_linear(oracle):  # wfaligner.rs:1169-1182
    a = oracle.wfa_align(PATTERN, TEXT, oracle.LINEAR, x=6, e=2)
    assert a.score == -20
    assert a.cigar_string() == "1M1I1D3M1I5M2I2D8M1I1M1I1M1I9M"


def test_aligner_gap_affine(oracle):  # wfaligner.rs:1185-1198
    a = oracle.wfa_align(PATTERN, TEXT, oracle.AFFINE, 6, 4, 2)
    assert a.score == -40
    assert a.cigar_string() == "1M1X3M1I5M2X8M3I1M1X9M"


def test_aligner_score_only(oracle):  # wfaligner.rs:1201-1211
    a = oracle.wfa_align(PATTERN, TEXT, oracle.AFFINE, 6, 4, 2, score_only=True)
    assert a.score == -40 and a.cigar_string() == ""


def test_aligner_gap_affine_2pieces(oracle):  # wfaligner.rs:1214-1227
    a = oracle.wfa_align(PATTERN, TEXT, oracle.AFFINE2P, 6, 2, 2, 4, 1)
    assert a.score == -34
    assert a.cigar_string() == "1M1X3M1I5M2X8M1I1M1I1M1I9M"


def test_aligner_span_1(oracle):  # wfaligner.rs:1230-1243
    p = b"AATTTAAGTCTAGGCTACTTTC"
    t = b"CCGACTACTACGAAATTTAAGTATAGGCTACTTTCCGTACGTACGTACGT"
    a = oracle.wfa_align(p, t, oracle.AFFINE2P, 8, 4, 2, 24, 1, ends_free=(0, 0, 0, len(t)))
    assert a.status == 0
    assert a.alignment_span() == ((0, 22), (13, 35))


def test_aligner_span_2(oracle):  # wfaligner.rs:1246-1261
    v = json.load(open(os.path.join(GOLD, "wfa_long_vectors.json")))["span_2"]
    p, t = v["pattern"].encode(), v["text"].encode()
    a = oracle.wfa_align(p, t, oracle.AFFINE2P, 8, 4, 2, 24, 1, ends_free=(0, 0, 0, len(t)))
    assert a.status == 0
    assert a.alignment_span() == ((78, 250), (0, 172))


def test_aligner_ends_free_global(oracle):  # wfaligner.rs:1264-1279
    p = b"AATTTAAGTCTAGGCTACTTTC"
    t = b"CCGACTACTACGAAATTTAAGTATAGGCTACTTTCCGTACGTACGTACGT"
    a = oracle.wfa_align(p, t, oracle.AFFINE, 6, 4, 2, ends_free=(0, 0, 0, len(t)))
    assert a.score == -36
    assert a.cigar_string() == "13I9M1X12M15I"


def test_aligner_ends_free_right_extent(oracle):  # wfaligner.rs:1282-1298
    p = b"AATTTAAGTCTGCTACTTTCACGCAGCT"
    t = b"AATTTCAGTCTGGCTACTTTCACGTACGATGACAGACTCT"
    a = oracle.wfa_align(p, t, oracle.AFFINE, 6, 4, 2, ends_free=(0, len(p), 0, len(t)))
    assert a.score == -24
    assert a.cigar_string() == "5M1X6M1I11M4D1M15I"


def test_aligner_ends_free_left_extent(oracle):  # wfaligner.rs:1301-1316
    p = b"CTTTCACGTACGTGACAGTCTCT"
    t = b"AATTTCAGTCTGGCTACTTTCACGTACGATGACAGACTCT"
    a = oracle.wfa_align(p, t, oracle.AFFINE, 6, 4, 2, ends_free=(0, 0, 0, 0))
    assert a.score == -48
    assert a.cigar_string() == "16I12M1I6M1X4M"


def test_aligner_ends_free_right_overlap(oracle):  # wfaligner.rs:1319-1334
    p = b"CGCGTCTGACTGACTGACTAAACTTTCATGTACCTGACA"
    t = b"AAACTTTCACGTACGTGACATATAGCGATCGATGACT"
    a = oracle.wfa_align(p, t, oracle.AFFINE, 6, 4, 2, ends_free=(0, 0, 0, 0))
    assert a.score == -92
    assert a.cigar_string() == "19D9M1X4M1X5M17I"


def test_clipping_score(oracle):  # wfaligner.rs:1337-1381
    text_lf = b"AAGGAGCTGAGAATTGTTCTTCCAGATACCTTTCCGACCTCTTCTTGGTT"
    text_rf = b"GGAGTGCAGTGGTGCAATCTTGGCTCACTACAACCTCCGCATCCTGGGTT"
    pattern_lf = b"AAGGAGCTGAGAATTGTTCGTCCAGATACCTTTCCGACCTCTTCTTGGTT"
    pattern_rf = b"GGAGTGCAGTGGTGCAATCTTGGCTCACTACAACCTCTGCATCCTGGGTT"
    text = text_lf + b"ATTT" * 10 + text_rf
    pattern = pattern_lf + b"ATTT" * 8 + pattern_rf
    a = oracle.wfa_align(pattern, text, oracle.AFFINE2P, 8, 4, 2, 24, 1)
    assert a.score == -36
    assert a.cigar_string() == "19M1X62M8I37M1X12M"
    assert a.cigar_score() == -36
    assert a.cigar_score_clipped(50) == -20
    assert a.cigar_string(50) == "32M8I"
    b = oracle.wfa_align(pattern, text, oracle.INDEL)
    assert b.score == 12 and b.cigar_score() == 12
    assert b.cigar_score_clipped(19) == 10
    assert b.cigar_score_clipped(0) == 12


def test_memory_modes(oracle):  # wfaligner.rs:1384-1421 (High/Med/Low give one result)
    a = oracle.wfa_align(PATTERN, TEXT, oracle.AFFINE2P, 8, 4, 2, 24, 1)
    assert a.score == -48 and a.cigar_score() == -48 and a.cigar_score_clipped(0) == -48
    assert a.cigar_string() == "1M1X3M1I5M2X8M3I1M1X9M"


def test_invalid_sequence_heuristic_none(oracle):  # wfaligner.rs:1438-1454 (second half)
    v = json.load(open(os.path.join(GOLD, "wfa_long_vectors.json")))["invalid_sequence"]
    a = oracle.wfa_align(v["pattern"].encode(), v["text"].encode(), oracle.AFFINE2P, 8, 4, 2, 24, 1)
    assert a.status == 0 and a.score == -881
    assert a.cigar_score() == -881


def test_get_and_decode_sam_cigar(oracle):  # wfaligner.rs:1590-1676
    a = oracle.wfa_align(b"TCTTTACTCTT", b"TCTTTACTCTT", oracle.AFFINE, 4, 6, 2)
    assert a.sam_cigar(True) == [183]
    assert oracle.decode_sam_cigar([183]) == [(11, "=")]
    assert a.sam_cigar(False) == [176]
    assert oracle.decode_sam_cigar([176]) == [(11, "M")]
    b = oracle.wfa_align(b"TCTTTACTCTT", b"TCTTTACTATT", oracle.AFFINE, 4, 6, 2)
    assert b.sam_cigar(True) == [135, 24, 39]
    assert oracle.decode_sam_cigar([135, 24, 39]) == [(8, "="), (1, "X"), (2, "=")]
    assert b.sam_cigar(False) == [176]


def test_get_alignment_global(oracle):  # wfaligner.rs:1719-1752
    a = oracle.wfa_align(PATTERN, TEXT, oracle.AFFINE, 1, 5, 1)
    assert a.score == -18
    assert a.ops == b"MXMMMIMMMMMXXMMMMMMMMIIIMXMMMMMMMMM"
    assert a.alignment_span() == ((0, 31), (0, 35))


def test_get_alignment_ends_free(oracle):  # wfaligner.rs:1795-1828 (the production flank call shape)
    p = b"AGTGTCAATGGCTAC"
    t = b"GGGGGGGGGGAGTGTCAATGGCTACGGGGGGGGGG"
    a = oracle.wfa_align(p, t, oracle.AFFINE, 1, 5, 1, ends_free=(0, 0, len(t), len(t)))
    assert a.score == 0
    assert a.ops == b"I" * 10 + b"M" * 15 + b"I" * 10
    assert a.alignment_span() == ((0, len(p)), (10, 25))


def test_wfa_score_matches_gotoh(oracle):
    """Not a reference vector: WFA scores must equal a textbook O(nm) affine-gap DP."""
    import random
    rng = random.Random(7)

    def gotoh(p, t, x, o, e):
        INF = 10 ** 9
        n, m = len(p), len(t)
        M = [[INF] * (m + 1) for _ in range(n + 1)]
        I = [[INF] * (m + 1) for _ in range(n + 1)]
        D = [[INF] * (m + 1) for _ in range(n + 1)]
        M[0][0] = 0
        for i in range(n + 1):
            for j in range(m + 1):
                if i > 0:
                    D[i][j] = min(M[i - 1][j] + o + e, D[i - 1][j] + e)
                if j > 0:
                    I[i][j] = min(M[i][j - 1] + o + e, I[i][j - 1] + e)
                if i > 0 and j > 0:
                    M[i][j] = min(M[i][j], M[i - 1][j - 1] + (0 if p[i - 1] == t[j - 1] else x))
                M[i][j] = min(M[i][j], I[i][j], D[i][j])
        return M[n][m]

    for _ in range(60):
        n = rng.randint(0, 40)
        p = bytes(rng.choice(b"ACGT") for _ in range(n))
        t = bytearray(p)
        for _ in range(rng.randint(0, 6)):
            r = rng.random()
            pos = rng.randint(0, len(t))
            if r < 0.4 and len(t):
                t[min(pos, len(t) - 1)] = rng.choice(b"ACGT")
            elif r < 0.7:
                t[pos:pos] = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 4)))
            else:
                del t[pos:pos + rng.randint(1, 4)]
        t = bytes(t)
        for (x, o, e) in [(2, 5, 1), (1, 0, 1), (6, 4, 2)]:
            a = oracle.wfa_align(p, t, oracle.AFFINE, x, o, e)
            assert -a.score == gotoh(p, t, x, o, e), (p, t, x, o, e)
            assert a.cigar_score() == a.score
            # the ops must spell a valid alignment
            pi = ti = 0
            for op in a.ops:
                if op == ord("M"):
                    assert p[pi] == t[ti]
                if op == ord("X"):
                    assert p[pi] != t[ti]
                if op in b"MX":
                    pi += 1; ti += 1
                elif op == ord("I"):
                    ti += 1
                else:
                    pi += 1
            assert (pi, ti) == (len(p), len(t))


# ------------------------------------------------------------------ HMM --

def summarize(spans):
    """builder.rs:191-205"""
    out = []
    for m, s, e in spans:
        if out and out[-1][2] == m:
            out[-1] = (out[-1][0], e, m)
        else:
            out.append((s, e, m))
    return out


def test_annotate_two_perfect_motif_runs(oracle):  # builder.rs:208-216
    hmm = oracle.Hmm([b"CAG", b"A"])
    labels = hmm.label_motifs(hmm.label(b"CAGCAGCAGCAGAAAAA"))
    assert summarize(labels) == [(0, 12, 0), (12, 17, 1)]


def test_annotate_motif_runs_separated_by_insertion(oracle):  # builder.rs:219-240
    motifs = [b"CAG", b"A"]
    hmm = oracle.Hmm(motifs)
    q = b"CAGCAGATCGATCGATCGATCGAAAAA"
    states = hmm.remove_imperfect_motifs(hmm.label(q), q, 6)
    assert summarize(hmm.label_motifs(states)) == [
        (0, 6, 0), (6, 7, 1), (7, 10, 2), (10, 11, 1), (11, 14, 2), (14, 15, 1),
        (15, 18, 2), (18, 19, 1), (19, 22, 2), (22, 27, 1)]


def test_annotate_imperfect_repeat_run(oracle):  # builder.rs:243-250
    hmm = oracle.Hmm([b"CAG", b"A"])
    labels = hmm.label_motifs(hmm.label(b"CAGCAGCTGCAGCAGAAACAG"))
    assert summarize(labels) == [(0, 15, 0), (15, 18, 1), (18, 21, 0)]


def test_parse_aga_repeat(oracle):  # builder.rs:253-273
    hmm = oracle.Hmm([b"AAG", b"CAAC"])
    q = (b"TCTATGCAACCAACTTTCTGTTAGTCATAGTACCCCAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAGAAG"
         b"AAGAAGAATAGAAATGTGTTTAAGAATTCCTCAATAAG")
    states = hmm.remove_imperfect_motifs(hmm.label(q), q, 6)
    assert summarize(hmm.label_motifs(states)) == [
        (0, 6, 2), (6, 14, 1), (14, 36, 2), (36, 93, 0), (93, 108, 2), (108, 111, 0),
        (111, 122, 2), (122, 125, 0)]


@pytest.mark.parametrize("motifs,query,expect", [
    ([b"CAG", b"CCG"], b"CAGCAGCAGCCGCCGCCGCCG", 1.0),                # purity.rs:48-55
    ([b"CAG", b"CCG"], b"CAGCGCAGCCGCCGCCGGG", 17.0 / 20.0),          # purity.rs:58-66
    ([b"CAG", b"CCG"], b"CAGCAGCAGTTTTTTTTCCGCCGCCG", 18.0 / 26.0),   # purity.rs:69-76
    ([b"GCN"], b"GCAGCCGCTGAG", 11.0 / 12.0),                         # purity.rs:79-87
])
def test_calc_purity(oracle, motifs, query, expect):
    hmm = oracle.Hmm(motifs)
    assert hmm.calc_purity(query, hmm.label(query)) == expect


def test_calc_purity_empty(oracle):  # purity.rs:90-96
    hmm = oracle.Hmm([b"CAG", b"CCG"])
    assert hmm.label(b"") == []
    assert math.isnan(hmm.calc_purity(b"", []))


def test_get_base_match(oracle):  # events.rs:124-145
    assert oracle.Hmm([b"A"]).get_base_match(3) == b"A"
    assert oracle.Hmm([b"N"]).get_base_match(3) == b"N"
    assert oracle.Hmm([b"A"]).get_base_match(1) == b" "      # silent state


def test_tutorial_vcf_record(oracle):
    """docs/tutorial.md:42-45: allele CAG x11 (after the padding base) -> MC=11, MS=0(0-33), AP=1."""
    hmm = oracle.Hmm([b"CAG"])
    mc, spans, purity = hmm.annotate(b"CAG" * 11)
    assert mc == [11] and spans == [(0, 0, 33)] and purity == 1.0
    mc, spans, purity = hmm.annotate(b"CAG" * 20)          # the REF allele of the record
    assert mc == [20] and spans == [(0, 0, 60)] and purity == 1.0


def test_replace_invalid_bases(oracle):  # src/hmm/utils.rs:29-42
    assert oracle.replace_invalid_bases(b"ACGTNRYacgt", b"ATCG") == b"ACGTACGTATCG"[:4] + \
        bytes(b"ATCG"[i % 4] for i in range(4, 11))
    assert oracle.replace_invalid_bases(b"GCN", b"ATCGN") == b"GCN"


def test_find_span_exact_and_fallback(oracle):
    """span_locater.rs:7-30 semantics on a hand-made case."""
    import random
    rng = random.Random(3)
    flank = bytes(rng.choice(b"ACGT") for _ in range(250))
    left = bytes(rng.choice(b"ACGT") for _ in range(300))
    right = bytes(rng.choice(b"ACGT") for _ in range(300))
    read = left + flank + right
    span, via, nm = oracle.find_span(flank, read)
    assert span == (300, 550) and via == 1 and nm == 250
    bad = bytearray(flank)
    bad[100] = ord("A") if bad[100] != ord("A") else ord("C")
    read2 = left + bytes(bad) + right
    span, via, nm = oracle.find_span(flank, read2)
    assert span == (300, 550) and via == 2 and nm == 249
    # unrelated read: alignment completes but is rejected by the match threshold
    junk = bytes(rng.choice(b"ACGT") for _ in range(700))
    span, via, nm = oracle.find_span(flank, junk)
    assert span is None and via == 3 and nm < 175


@pytest.mark.parametrize("seq, piece, expect", [  # span_locater.rs:73-119 (the exact search of find_spans, :10-12)
    (b"ABCDEFG", b"CDE", (2, 5)), (b"ABCDEFG", b"XYZ", None), (b"ABCABCABC", b"ABC", (0, 3)),
    (b"ABCDEFG", b"ABC", (0, 3)), (b"ABCDEFG", b"EFG", (4, 7)), (b"ABCDEFG", b"ABCDEFG", (0, 7)),
    (b"ABC", b"ABCDEFG", None), (b"", b"ABC", None), (b"ABCDEFG", b"A", (0, 1)), (b"ABCDEFG", b"G", (6, 7)),
    (b"ABCDEFG", b"D", (3, 4)), (b"AAAAA", b"AA", (0, 2)), (b"ACGTNACGT", b"N", (4, 5)),
])
def test_find_vs_windows_comparison(oracle, seq, piece, expect):
    """The reference's vectors for the exact search: first occurrence, as `windows().position()` finds it.  Where it
    expects None the exact search must miss (what the WFA fallback then makes of such toy inputs is not part of the
    reference's test)."""
    span, via, nm = oracle.find_span(piece, seq)
    if expect is None:
        assert via != 1
    else:
        assert span == expect and via == 1 and nm == len(piece)


# ------------------------------------------------------------- consensus (next row) --

def test_repair_consensus_reference_examples(oracle):
    """src/trgt/genotype/consensus.rs:172-213 -- the reference's own examples for consensus repair.  They sit
    in a commented-out test module written against an older API, so they pin nothing officially
    (parity of this row stays 'unpinned'); the restatement reproduces all three expected strings."""
    assert oracle.repair_consensus(b"CCCCACCCTCCC", [b"CCCCACCCGCCC", b"CCCCCCCCGCCC", b"CCCCCCCCCCCC"]) == b"CCCCCCCCGCCC"
    assert oracle.repair_consensus(b"CCCCCCCCGCCC", [b"CCCCCCGCCC", b"CCCCCCGCCC", b"CCCCCCCCGCCC"]) == b"CCCCCCGCCC"
    assert oracle.repair_consensus(b"CCCCCCCCGCCC", [b"CCCCCAAACCCGCCC", b"CCCCCAAACCCGCCC", b"CCCCCACCCGCAACC"]) == b"CCCCCAAACCCGCCC"


def test_repair_consensus_tie_rules(oracle):
    # column vote: Rust's max_by_key keeps the LAST maximum of [A,T,C,G,-] (consensus.rs:44-53)
    assert oracle.repair_consensus(b"A", [b"A", b"G"]) == b"G"
    assert oracle.repair_consensus(b"AC", [b"AC", b"C"]) in (b"C",)       # A vs '-' tie -> '-' wins
    # insertion: needs more than half of the members and must outnumber members without one (:57, :106-110)
    assert oracle.repair_consensus(b"ACGT", [b"ACTTGT", b"ACTTGT", b"ACGT"]) == b"ACTTGT"
    assert oracle.repair_consensus(b"ACGT", [b"ACTTGT", b"ACGT"]) == b"ACGT"


# ------------------------------------------------------------- read clipping (next row) --

_CLIP_READ = b"CGCTCGTTAAATCACG"
_CLIP_CIGAR = "3=2D2=1X2=5I3="


def _clip(oracle, bases, cigar, ref_pos, region):
    """HiFiRead::clip_to_region on (bases, cigar): None or (bases, clipped ref_pos, clipped CIGAR text)"""
    ops = oracle.encode_bam_cigar(cigar)
    res = oracle.clip_cigar(ops, ref_pos, region)
    if res is None:
        return None
    ref_start, qs, qe, words = res
    packed = oracle.encode_seq4(bases)
    text = "".join(f"{w >> 4}{oracle.BAM_OPS[w & 15]}" for w in words)
    return oracle.decode_seq4(packed, qs, qe - qs), ref_start, text


@pytest.mark.parametrize("bases,cigar,region,expect", [
    (_CLIP_READ, _CLIP_CIGAR, (0, 10), None),                       # clip_region.rs:214-229
    (_CLIP_READ, _CLIP_CIGAR, (23, 33), None),
    (b"AAAAACGCTCGTTAAATCACGAAAAAAAAAA", "5S3=2D2=1X2=5I3=10S", (9, 23),   # :232-239: the original read
     (b"AAAAACGCTCGTTAAATCACGAAAAAAAAAA", 10, "5S3=2D2=1X2=5I3=10S")),
    (_CLIP_READ, _CLIP_CIGAR, (0, 15), (b"CGC", 10, "3=2D")),      # :242-254
    (_CLIP_READ, _CLIP_CIGAR, (12, 17), (b"CTC", 12, "1=2D2=")),   # :257-269
    (_CLIP_READ, _CLIP_CIGAR, (21, 22), (b"C", 21, "1=")),         # :272-284
    (_CLIP_READ, _CLIP_CIGAR, (0, 17), (b"CGCTC", 10, "3=2D2=")),  # :287-299
])
def test_clip_to_region_reference_vectors(oracle, bases, cigar, region, expect):
    assert _clip(oracle, bases, cigar, 10, region) == expect


@pytest.mark.parametrize("bases,cigar,region,meth", [   # the methylation side of the same six tests
    (b"AAAAACGCTCGTTAAATCACGAAAAAAAAAA", "5S3=2D2=1X2=5I3=10S", (9, 23), [10, 20, 30]),   # clip_region.rs:232-239
    (_CLIP_READ, _CLIP_CIGAR, (0, 15), [10]),          # :242-254
    (_CLIP_READ, _CLIP_CIGAR, (12, 17), [20]),         # :257-269
    (_CLIP_READ, _CLIP_CIGAR, (21, 22), [30]),         # :272-284
    (_CLIP_READ, _CLIP_CIGAR, (0, 17), [10, 20]),      # :287-299
])
def test_clip_to_region_methylation_vectors(oracle, bases, cigar, region, meth):
    _, qs, qe, _ = oracle.clip_cigar(oracle.encode_bam_cigar(cigar), 10, region)
    m0, m1 = oracle.meth_range(bases, qs, qe)
    assert [10, 20, 30][m0:m1] == meth


_CB_READ = b"AAAAACGCTCGTTAAATCACGAAAAAAAAAA"
_CB_CIGAR = "5S3=2D2=1X2=5I3=10S"


def _clip_bases(oracle, bases, cigar, left, right):
    """HiFiRead::clip_bases -> None or (bases, methylation of [10, 20, 30], ref_pos, CIGAR text)"""
    res = oracle.clip_bases(bases, oracle.encode_bam_cigar(cigar), 10, left, right)
    if res is None:
        return None
    b0, b1, m0, m1, (ref_pos, words) = res
    return bases[b0:b1], [10, 20, 30][m0:m1], ref_pos, "".join(f"{w >> 4}{oracle.BAM_OPS[w & 15]}" for w in words)


@pytest.mark.parametrize("bases,cigar,left,right,expect", [
    (b"CGCTCGTTAAATCACG", "3=2D2=1X2=5I3=", 16, 0, None),     # clip_bases.rs:147-159 get_none_if_clip_whole_query
    (b"CGCTCGTTAAATCACG", "3=2D2=1X2=5I3=", 0, 16, None),
    (b"CGCTCGTTAAATCACG", "3=2D2=1X2=5I3=", 12, 4, None),
    (_CB_READ, _CB_CIGAR, 3, 0, (b"AACGCTCGTTAAATCACGAAAAAAAAAA", [10, 20, 30], 10, "2S3=2D2=1X2=5I3=10S")),   # :162-177
    (_CB_READ, _CB_CIGAR, 5, 0, (b"CGCTCGTTAAATCACGAAAAAAAAAA", [10, 20, 30], 10, "3=2D2=1X2=5I3=10S")),
    (_CB_READ, _CB_CIGAR, 10, 0, (b"GTTAAATCACGAAAAAAAAAA", [30], 17, "1X2=5I3=10S")),
    (_CB_READ, _CB_CIGAR, 0, 5, (b"AAAAACGCTCGTTAAATCACGAAAAA", [10, 20, 30], 10, "5S3=2D2=1X2=5I3=5S")),     # :180-195
    (_CB_READ, _CB_CIGAR, 0, 10, (b"AAAAACGCTCGTTAAATCACG", [10, 20, 30], 10, "5S3=2D2=1X2=5I3=")),
    (_CB_READ, _CB_CIGAR, 0, 15, (b"AAAAACGCTCGTTAAA", [10, 20], 10, "5S3=2D2=1X2=3I")),
    (_CB_READ, _CB_CIGAR, 5, 5, (b"CGCTCGTTAAATCACGAAAAA", [10, 20, 30], 10, "3=2D2=1X2=5I3=5S")),            # :198-209
    (_CB_READ, _CB_CIGAR, 8, 11, (b"TCGTTAAATCAC", [20, 30], 13, "2D2=1X2=5I2=")),
    (_CB_READ, _CB_CIGAR, 31, 0, None),                                                                       # :212-219
    (_CB_READ, _CB_CIGAR, 0, 31, None),
    (_CB_READ, _CB_CIGAR, 30, 30, None),
    (_CB_READ, _CB_CIGAR, 13, 13, (b"AAATC", [], 20, "5I")),                                                  # :222-230
])
def test_clip_bases_reference_vectors(oracle, bases, cigar, left, right, expect):
    assert _clip_bases(oracle, bases, cigar, left, right) == expect


def test_bamlet_clip_rule(oracle):  # write_bam.rs:80-92: flank_len either side of the span, else the read is skipped
    ops = oracle.encode_bam_cigar(_CB_CIGAR)
    assert oracle.bamlet_clip(_CB_READ, ops, 10, (10, 16), 5) == oracle.clip_bases(_CB_READ, ops, 10, 5, 10)
    assert oracle.bamlet_clip(_CB_READ, ops, 10, (4, 16), 5) is None          # span.0 < flank_len
    assert oracle.bamlet_clip(_CB_READ, ops, 10, (10, 27), 5) is None         # bases.len() < span.1 + flank_len
    assert oracle.bamlet_clip(_CB_READ, None, 0, (10, 16), 5) == (5, 21, 0, 3, None)   # unmapped read: no CIGAR


def test_seq4_alphabet(oracle):  # read.rs:104: htslib's "=ACMGRSVTWYHKDBN", first base in the high nibble
    packed = bytes([0x12, 0x48, 0xF0])
    assert oracle.decode_seq4(packed, 0, 5) == b"ACGTN"
    assert oracle.decode_seq4(packed, 1, 4) == b"CGTN"
    assert oracle.decode_seq4(bytes(range(256)), 0, 512) == bytes(
        oracle.SEQ4_ALPHABET[(i >> 4) if j == 0 else (i & 15)] for i in range(256) for j in (0, 1))
    assert oracle.decode_seq4(b"", 0, 0) == b""


# ------------------------------------------------------------- VCF sample fields (next row) --

def test_vcf_fields_tutorial_record(oracle):
    """docs/tutorial.md:44: AL 33,33  MC 11,11  MS 0(0-33),0(0-33)  AP 1.000000,1.000000, from the oracle's own
    annotation of the two CAG x 11 alleles"""
    h = oracle.Hmm([b"CAG"])
    alleles = []
    for seq in (b"CAG" * 11, b"CAG" * 11):
        mc, spans, pur = h.annotate(seq)
        alleles.append((len(seq), mc, spans, pur))
    assert oracle.vcf_fields(alleles) == (b"33,33", b"11,11", b"0(0-33),0(0-33)", b"1.000000,1.000000")


def test_vcf_fields_shapes(oracle):  # write_vcf.rs:286-343: '_' within an allele, ',' between, '.' for None / NaN
    got = oracle.vcf_fields([(17, [4, 5], [(0, 0, 12), (1, 12, 17)], 1.0), (0, [0, 0], None, float("nan")),
                             (20, [3, 0], [(0, 2, 11)], 17 / 20)])
    assert got == (b"17,0,20", b"4_5,0_0,3_0", b"0(0-12)_1(12-17),.,0(2-11)", b"1.000000,.,0.850000")
    assert oracle.vcf_fields([]) == (b"", b"", b"", b"")
