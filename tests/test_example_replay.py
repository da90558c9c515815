"""BASELINE config 1: `trgt genotype` on the reference's example/ (one CAG locus, 33 real HiFi reads),
replayed from the BAM records through the hot path and compared with the VCF record the reference's own
run prints (docs/tutorial.md:42-45).  Host logic between the leaf calls: oracle/host.py (restatement of
tr.rs / genotype_size.rs / genotype_flank.rs / write_vcf.rs, test infrastructure)."""
import os

import pytest

from oracle import host
from oracle import oracle as orc

EX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example")
# docs/tutorial.md:44
EXPECTED_SAMPLE = "1/1:33,33:30-39,33-33:15,14:11,11:0(0-33),0(0-33):1.000000,1.000000:.,."
EXPECTED_REF = b"C" + b"CAG" * 20
EXPECTED_ALT = b"C" + b"CAG" * 11


def load_example():
    genome = host.read_fasta(os.path.join(EX, "reference.fasta"))
    refs, recs = host.parse_bam(os.path.join(EX, "sample.bam"))
    with open(os.path.join(EX, "repeat.bed")) as f:
        locus = host.Locus.from_bed_line(genome, f.read().strip())
    reads = host.extract_reads(locus, [r[0] for r in refs], recs, host.Params())
    return locus, reads


def check(res):
    assert res is not None
    assert res.sample_field == EXPECTED_SAMPLE
    assert res.vcf_ref_alt == (EXPECTED_REF, [EXPECTED_ALT])
    assert res.n_spanning == 29  # SD 15,14


def test_example_inputs():
    locus, reads = load_example()
    assert (locus.id, locus.contig, locus.start, locus.end, locus.motifs) == ("TR1", "chrA", 10001, 10061, [b"CAG"])
    assert locus.tr == b"CAG" * 20 and len(locus.left_flank) == len(locus.right_flank) == 250
    assert len(reads) == 33
    # real records: soft clips, =/X/I/D CIGARs, both strands, 9.5 - 15.7 kb
    assert {r.rec.flag for r in reads} == {0, 16}
    assert min(r.rec.l_seq for r in reads) == 9513 and max(r.rec.l_seq for r in reads) == 15694


def test_example_replay_oracle():
    locus, reads = load_example()
    res = host.analyze(locus, reads, host.OracleBackend(orc))
    check(res)
    # every clipped read is what clip_to_region leaves: 500 bp either side of the locus where the read reaches that far
    for r in res.reads:
        assert len(r.bases) == r.clip[2] - r.clip[1]


@pytest.mark.gpu
def test_example_replay_engine():
    import trgt_b200
    from tests.engine_backend import EngineBackend
    locus, reads = load_example()
    ref = host.analyze(*load_example(), host.OracleBackend(orc))
    eng = trgt_b200.Engine(device=0)
    try:
        res = host.analyze(locus, reads, EngineBackend(eng))
        check(res)
        # and step by step identical to the oracle: clips, clipped bases, spans, alleles, annotations
        assert [r.clip for r in res.reads] == [r.clip for r in ref.reads]
        assert [r.bases for r in res.reads] == [r.bases for r in ref.reads]
        assert res.spans == ref.spans and res.alleles == ref.alleles and res.classification == ref.classification
        for a, b in zip(res.annotations, ref.annotations):
            assert list(a[0]) == list(b[0]) and (a[1] or None) == (b[1] or None) and a[2] == b[2]
        stats = eng.kernel_stats()
        for k in ("k_clip_cigar", "k_unpack_seq4", "k_flank_exact_t", "k_hmm_lane_viterbi", "k_vcf_fields_write"):
            assert k in stats, (k, sorted(stats))
    finally:
        eng.close()
