"""trgt_b200 -- B200-native per-locus tandem-repeat DP engine (hot path of `trgt genotype`).

The package holds only what that path needs: csrc/ (hand-written sm_100a kernels + the C ABI of
include/trgt_engine.h), engine.py (ctypes binding and the host-side mirror of the reference's
find_tr_spans / align / get_dist_matrix / label_with_hmm) and shard.py (locus shards and the record gather).
Bench and test support (workloads, the phase pipeline, the stand-in host glue) lives in harness/.  There is no
CPU fallback.
"""
from .engine import (Annotation, AnnotationBatch, CigarBatch, BAMLET_CLIP_DTYPE, CLIP_DTYPE, Engine, EXPORTS, PackedSeq4, PackedSeqs,
                     TrgtError,
                     decode_sam_cigar, load_library, HIT_DTYPE, SPAN_DTYPE, VIA_EXACT, VIA_NONE, VIA_WFA,
                     VIA_WFA_REJECTED)

__all__ = ["Annotation", "AnnotationBatch", "CigarBatch", "BAMLET_CLIP_DTYPE", "CLIP_DTYPE", "Engine", "EXPORTS", "PackedSeq4", "PackedSeqs",
           "TrgtError",
           "decode_sam_cigar", "load_library", "HIT_DTYPE", "SPAN_DTYPE", "VIA_EXACT", "VIA_NONE", "VIA_WFA",
           "VIA_WFA_REJECTED"]
