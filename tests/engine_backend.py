"""The leaf calls oracle/host.py's worker makes, served by the CUDA engine through the C ABI (GPU tests).
Reads go in as the BAM record holds them: CIGAR words to trgt_clip_reads, 4-bit bases to
trgt_flank_spans_seq4 / trgt_seq4_decode -- no base is decoded or shifted on the host."""
from __future__ import annotations

import numpy as np

import trgt_b200
from trgt_b200.engine import PackedSeq4, PackedSeqs


class EngineBackend:
    def __init__(self, eng: "trgt_b200.Engine"):
        self.eng = eng
        self._p4 = None

    def clip_reads(self, reads, region):
        if not reads:
            return []
        ops = np.concatenate([np.asarray(r.rec.cigar, dtype=np.uint32) for r in reads])
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(r.rec.cigar) for r in reads])
        refs = np.array([r.rec.pos for r in reads], dtype=np.int64)
        clips = self.eng.clip_reads(ops, offs, refs, np.array([region], dtype=np.int64),
                                    np.array([0, len(reads)], dtype=np.uint32))
        out = []
        for r, c in zip(reads, clips):
            if int(c["status"]) != 1:
                assert int(c["status"]) == 0, "unexpected CIGAR operation"
                out.append(None)
                continue
            n, f = int(c["n_ops"]), int(c["first_op"])
            words = [int(c["first_word"])] + [int(w) for w in r.rec.cigar[f + 1:f + n - 1]] + ([int(c["last_word"])] if n > 1 else [])
            out.append((int(c["ref_start"]), int(c["query_start"]), int(c["query_end"]), words))
        return out

    def _pack(self, reads) -> PackedSeq4:
        """the bytes of each record that cover bases[query_start..query_end), back to back (trgt_seq4_t)"""
        chunks, starts, lengths, byte_off = [], [], [], 0
        for r in reads:
            qs, qe = r.clip[1], r.clip[2]
            b = r.rec.seq4[qs >> 1:(qe + 1) >> 1]
            chunks.append(np.frombuffer(b, dtype=np.uint8))
            starts.append(2 * byte_off + (qs & 1))
            lengths.append(qe - qs)
            byte_off += len(b)
        data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
        return PackedSeq4(data, np.array(starts, dtype=np.uint64), np.array(lengths, dtype=np.uint32))

    def clipped_bases(self, reads):
        self._p4 = self._pack(reads)
        dec = self.eng.seq4_decode(self._p4)
        return [dec.get(i) for i in range(len(reads))]

    def find_tr_spans(self, lf, rf, reads, params):
        p4 = self._p4 if self._p4 is not None and len(self._p4) == len(reads) else self._pack(reads)
        spans, _ = self.eng.flank_spans_seq4(PackedSeqs.from_list([lf]), PackedSeqs.from_list([rf]), p4,
                                             np.array([0, len(reads)], dtype=np.uint32), params.scoring,
                                             params.min_flank_id_frac)
        trs = self.eng.flank_trs(copy=True)  # the repeat sequences as the engine cuts them must be the reads' own
        out = []
        for i, r in enumerate(reads):
            if spans[i]["found"]:
                s = (int(spans[i]["start"]), int(spans[i]["end"]))
                assert trs.get(i) == r.bases[s[0]:s[1]]
                out.append(s)
            else:
                assert trs.get(i) == b""
                out.append(None)
        return out

    def align_repair(self, backbone, seqs):
        return self.eng.repair_consensus([(backbone, list(seqs))])[0]

    def label_with_hmm(self, motifs, alleles):
        res = self.eng.label_with_hmm([(list(motifs), list(alleles))])[0]
        return [(a.motif_counts, a.labels, a.purity) for a in res]

    def vcf_fields(self, motifs, seqs, anns):
        self.eng.label_with_hmm([(list(motifs), list(seqs))])  # alleles in genotype order, as trgt_vcf_fields wants them
        return self.eng.vcf_fields()[0]
