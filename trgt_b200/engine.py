"""ctypes binding of libtrgt_b200.so (include/trgt_engine.h) plus a host-side mirror of the
reference's operator surface for the hot path, so tests read like the reference's own:

    find_tr_spans      src/trgt/genotype/span_locater.rs:32
    align              src/utils/align.rs:14
    get_dist_matrix    src/trgt/genotype/genotype_cluster.rs:250
    label_with_hmm     src/trgt/workflows/tr.rs:454

All compute happens in the CUDA library; there is no CPU fallback.  Importing this module does
not need a GPU, creating an Engine does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _build

VIA_NONE, VIA_EXACT, VIA_WFA, VIA_WFA_REJECTED = 0, 1, 2, 3


class TrgtError(RuntimeError):
    pass


class _Seqs(C.Structure):
    _fields_ = [("data", C.c_void_p), ("offsets", C.c_void_p), ("n", C.c_uint64)]


class _Seq4(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_bytes", C.c_uint64), ("starts", C.c_void_p), ("lengths", C.c_void_p),
                ("n", C.c_uint64)]


class _Scoring(C.Structure):
    _fields_ = [("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32)]


class _Cigars(C.Structure):
    _fields_ = [("n", C.c_uint64), ("offsets", C.POINTER(C.c_uint64)), ("words", C.POINTER(C.c_uint32)),
                ("scores", C.POINTER(C.c_int32)), ("status", C.POINTER(C.c_int32))]


class _SeqsOut(C.Structure):
    _fields_ = [("n", C.c_uint64), ("offsets", C.POINTER(C.c_uint64)), ("data", C.POINTER(C.c_uint8)),
                ("status", C.POINTER(C.c_int32))]


class _Annotations(C.Structure):
    _fields_ = [("n", C.c_uint64), ("motif_count_offsets", C.POINTER(C.c_uint64)),
                ("motif_counts", C.POINTER(C.c_uint32)), ("span_offsets", C.POINTER(C.c_uint64)),
                ("spans", C.POINTER(C.c_uint32)), ("purity", C.POINTER(C.c_double)),
                ("status", C.POINTER(C.c_int32)), ("path_offsets", C.POINTER(C.c_uint64)),
                ("paths", C.POINTER(C.c_uint32))]


CLIP_DTYPE = np.dtype([("ref_start", np.int64), ("query_start", np.uint64), ("query_end", np.uint64),
                       ("first_op", np.uint32), ("n_ops", np.uint32), ("first_word", np.uint32),
                       ("last_word", np.uint32), ("status", np.int32)], align=True)
BAMLET_CLIP_DTYPE = np.dtype([("ref_pos", np.int64), ("base_start", np.uint32), ("base_end", np.uint32),
                              ("meth_start", np.uint32), ("meth_end", np.uint32), ("first_op", np.uint32),
                              ("n_ops", np.uint32), ("first_word", np.uint32), ("last_word", np.uint32),
                              ("status", np.int32), ("pad", np.uint32)], align=True)
SEQ4_ALPHABET = b"=ACMGRSVTWYHKDBN"
SPAN_DTYPE = np.dtype([("found", np.int32), ("start", np.uint32), ("end", np.uint32)])
HIT_DTYPE = np.dtype([("via", np.int32), ("matches", np.int32), ("score", np.int32),
                      ("start", np.uint32), ("end", np.uint32)])

# every symbol include/trgt_engine.h declares
EXPORTS = [
    "trgt_engine_create", "trgt_engine_destroy", "trgt_engine_last_error", "trgt_last_create_error",
    "trgt_engine_stream", "trgt_engine_sm_count", "trgt_engine_sync", "trgt_engine_set_workspace_budget",
    "trgt_engine_set_flank_band_budget", "trgt_engine_set_hmm_lane_path",
    "trgt_host_alloc", "trgt_host_free",
    "trgt_clip_reads", "trgt_seq4_decode", "trgt_flank_spans_seq4", "trgt_flank_upload_seq4",
    "trgt_flank_trs", "trgt_vcf_fields", "trgt_bamlet_clip",
    "trgt_flank_spans", "trgt_align_e2e", "trgt_consensus", "trgt_edit_dist", "trgt_hmm_label",
    "trgt_cluster", "trgt_cluster_trs", "trgt_consensus_trs", "trgt_align_trs",
    "trgt_flank_upload", "trgt_flank_run", "trgt_flank_download", "trgt_flank_free", "trgt_flank_device_views",
    "trgt_flank_fallback_counts",
    "trgt_align_upload", "trgt_align_run", "trgt_align_download", "trgt_align_free",
    "trgt_hmm_upload", "trgt_hmm_run", "trgt_hmm_download", "trgt_hmm_free",
    "trgt_engine_set_profiling", "trgt_engine_reset_stats", "trgt_engine_kernel_count",
    "trgt_engine_kernel_stat", "trgt_engine_launches",
]

_lib = None
_PINNED_OWNERS = {}  # pinned allocations stay alive for the life of the process unless released


def load_library(build: bool = True):
    """dlopen trgt_b200/libtrgt_b200.so (building it with nvcc first if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        try:
            _build.build()
        except RuntimeError:
            if not os.path.exists(_build.LIB_PATH):
                raise
    if not os.path.exists(_build.LIB_PATH):
        raise TrgtError("libtrgt_b200.so is missing: run __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(_build.LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    L.trgt_engine_create.argtypes = [i32, C.POINTER(vp)]
    L.trgt_engine_destroy.argtypes = [vp]
    L.trgt_engine_destroy.restype = None
    L.trgt_engine_last_error.argtypes = [vp]
    L.trgt_engine_last_error.restype = C.c_char_p
    L.trgt_last_create_error.restype = C.c_char_p
    L.trgt_engine_stream.argtypes = [vp]
    L.trgt_engine_stream.restype = vp
    L.trgt_engine_sm_count.argtypes = [vp]
    L.trgt_engine_sync.argtypes = [vp]
    L.trgt_engine_set_workspace_budget.argtypes = [vp, C.c_size_t]
    L.trgt_engine_set_workspace_budget.restype = None
    L.trgt_engine_set_flank_band_budget.argtypes = [vp, i32]
    L.trgt_engine_set_flank_band_budget.restype = None
    L.trgt_engine_set_hmm_lane_path.argtypes = [vp, i32]
    L.trgt_engine_set_hmm_lane_path.restype = None
    L.trgt_host_alloc.argtypes = [C.c_size_t]
    L.trgt_host_alloc.restype = vp
    L.trgt_host_free.argtypes = [vp]
    L.trgt_host_free.restype = None
    sp = C.POINTER(_Seqs)
    L.trgt_flank_spans.argtypes = [vp, sp, sp, sp, vp, u32, _Scoring, C.c_double, vp, vp]
    L.trgt_flank_upload.argtypes = [vp, sp, sp, sp, vp, u32, _Scoring, C.c_double, C.POINTER(vp)]
    s4 = C.POINTER(_Seq4)
    L.trgt_flank_spans_seq4.argtypes = [vp, sp, sp, s4, vp, u32, _Scoring, C.c_double, vp, vp]
    L.trgt_flank_upload_seq4.argtypes = [vp, sp, sp, s4, vp, u32, _Scoring, C.c_double, C.POINTER(vp)]
    L.trgt_seq4_decode.argtypes = [vp, s4, vp, vp]
    L.trgt_clip_reads.argtypes = [vp, vp, vp, vp, u64, vp, vp, u32, vp]
    L.trgt_flank_trs.argtypes = [vp, vp, C.POINTER(_SeqsOut)]
    L.trgt_vcf_fields.argtypes = [vp, vp, C.POINTER(_SeqsOut)]
    L.trgt_flank_run.argtypes = [vp, vp]
    L.trgt_flank_download.argtypes = [vp, vp, vp, vp]
    L.trgt_flank_free.argtypes = [vp, vp]
    L.trgt_flank_free.restype = None
    L.trgt_flank_device_views.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                          C.POINTER(u32), C.POINTER(u32)]
    L.trgt_align_e2e.argtypes = [vp, sp, sp, vp, u32, C.POINTER(_Cigars)]
    L.trgt_align_upload.argtypes = [vp, sp, sp, vp, u32, C.POINTER(vp)]
    L.trgt_align_run.argtypes = [vp, vp]
    L.trgt_align_download.argtypes = [vp, vp, C.POINTER(_Cigars)]
    L.trgt_align_free.argtypes = [vp, vp]
    L.trgt_align_free.restype = None
    L.trgt_consensus.argtypes = [vp, sp, sp, vp, u32, C.POINTER(_SeqsOut)]
    L.trgt_edit_dist.argtypes = [vp, sp, vp, u32, vp]
    L.trgt_cluster.argtypes = [vp, sp, vp, u32, vp, vp, vp]
    L.trgt_cluster_trs.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp]
    L.trgt_bamlet_clip.argtypes = [vp, vp, vp, vp, vp, u32, vp]
    L.trgt_align_trs.argtypes = [vp, vp, vp, vp, vp, u32, vp]
    L.trgt_consensus_trs.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(_SeqsOut)]
    L.trgt_hmm_label.argtypes = [vp, sp, vp, u32, sp, vp, i32, C.POINTER(_Annotations)]
    L.trgt_hmm_upload.argtypes = [vp, sp, vp, u32, sp, vp, i32, C.POINTER(vp)]
    L.trgt_hmm_run.argtypes = [vp, vp]
    L.trgt_hmm_download.argtypes = [vp, vp, C.POINTER(_Annotations)]
    L.trgt_hmm_free.argtypes = [vp, vp]
    L.trgt_hmm_free.restype = None
    L.trgt_engine_set_profiling.argtypes = [vp, i32]
    L.trgt_engine_set_profiling.restype = None
    L.trgt_engine_reset_stats.argtypes = [vp]
    L.trgt_engine_reset_stats.restype = None
    L.trgt_engine_kernel_count.argtypes = [vp]
    L.trgt_engine_kernel_stat.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(u64), C.POINTER(C.c_double)]
    L.trgt_engine_launches.argtypes = [vp]
    L.trgt_engine_launches.restype = u64
    _lib = L
    return L


# ------------------------------------------------------------------ packing helpers ------

class PackedSeqs:
    """CSR set of byte sequences: one contiguous uint8 buffer + uint64 offsets[n+1]."""

    def __init__(self, data: np.ndarray, offsets: np.ndarray):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        assert self.offsets.ndim == 1 and self.offsets.size >= 1
        self._c = _Seqs(self.data.ctypes.data if self.data.size else None, self.offsets.ctypes.data,
                        self.offsets.size - 1)

    @classmethod
    def from_list(cls, seqs: Sequence[bytes]) -> "PackedSeqs":
        offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        if seqs:
            offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        data = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8)
        return cls(data, offs)

    def __len__(self):
        return self.offsets.size - 1

    def get(self, i: int) -> bytes:
        return self.data[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()

    def ref(self):
        return C.byref(self._c)


class PackedSeq4:
    """Reads as BAM stores them (trgt_seq4_t): 4-bit codes of "=ACMGRSVTWYHKDBN", two bases per byte, first
    base in the high nibble; read i = bases [starts[i], starts[i]+lengths[i]) of data."""

    def __init__(self, data: np.ndarray, starts: np.ndarray, lengths: np.ndarray):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.starts = np.ascontiguousarray(starts, dtype=np.uint64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        assert self.starts.shape == self.lengths.shape and self.starts.ndim == 1
        self._c = _Seq4(self.data.ctypes.data if self.data.size else None, self.data.size,
                        self.starts.ctypes.data if self.starts.size else None,
                        self.lengths.ctypes.data if self.lengths.size else None, self.starts.size)

    @classmethod
    def from_ascii(cls, seqs: Sequence[bytes], odd_starts: bool = True) -> "PackedSeq4":
        """Encode ASCII reads the way consecutive BAM records would hold them (test helper; unknown letters
        become N).  With odd_starts every other read begins in the low nibble of a byte, as a clip at an
        odd query position does."""
        lut = np.full(256, 15, dtype=np.uint8)
        for i, ch in enumerate(SEQ4_ALPHABET):
            lut[ch] = i
        nibs, starts, pos = [], [], 0
        for k, s_ in enumerate(seqs):
            if odd_starts and (k & 1) and (pos & 1) == 0:
                nibs.append(np.zeros(1, dtype=np.uint8))
                pos += 1
            elif (not odd_starts or not (k & 1)) and (pos & 1):
                nibs.append(np.zeros(1, dtype=np.uint8))
                pos += 1
            starts.append(pos)
            nibs.append(lut[np.frombuffer(bytes(s_), dtype=np.uint8)])
            pos += len(s_)
        flat = np.concatenate(nibs) if nibs else np.zeros(0, dtype=np.uint8)
        if flat.size & 1:
            flat = np.concatenate([flat, np.zeros(1, dtype=np.uint8)])
        data = ((flat[0::2] << 4) | flat[1::2]).astype(np.uint8)
        return cls(data, np.array(starts, dtype=np.uint64), np.array([len(s_) for s_ in seqs], dtype=np.uint32))

    def __len__(self):
        return self.starts.size

    def ref(self):
        return C.byref(self._c)


def _group_offsets(groups: Sequence[Sequence[bytes]]) -> np.ndarray:
    offs = np.zeros(len(groups) + 1, dtype=np.uint32)
    if groups:
        offs[1:] = np.cumsum([len(g) for g in groups], dtype=np.uint64).astype(np.uint32)
    return offs


def decode_sam_cigar(words: Sequence[int]) -> List[Tuple[int, str]]:
    """WFAligner::decode_sam_cigar (src/wfaligner.rs:961-984)"""
    tab = "MIDNSHP=X"
    return [(int(w) >> 4, tab[int(w) & 0xF]) for w in words]


@dataclass
class Annotation:
    """src/hmm/spans.rs:21-25"""
    labels: Optional[List[Tuple[int, int, int]]]   # (motif_index, start, end) or None
    motif_counts: List[int]
    purity: float


@dataclass
class CigarBatch:
    offsets: np.ndarray   # uint64 [n+1]
    words: np.ndarray     # uint32
    scores: np.ndarray    # int32, -cost
    status: np.ndarray    # int32

    def cigar(self, i: int) -> List[int]:
        return self.words[int(self.offsets[i]):int(self.offsets[i + 1])].tolist()


@dataclass
class AnnotationBatch:
    motif_count_offsets: np.ndarray
    motif_counts: np.ndarray
    span_offsets: np.ndarray
    spans: np.ndarray        # uint32 [n_spans, 3]
    purity: np.ndarray
    status: np.ndarray
    path_offsets: Optional[np.ndarray] = None
    paths: Optional[np.ndarray] = None

    def annotation(self, i: int) -> Annotation:
        mc = self.motif_counts[int(self.motif_count_offsets[i]):int(self.motif_count_offsets[i + 1])].tolist()
        sp = self.spans[int(self.span_offsets[i]):int(self.span_offsets[i + 1])]
        labels = [tuple(int(v) for v in row) for row in sp] or None
        return Annotation(labels, mc, float(self.purity[i]))

    def path(self, i: int) -> List[int]:
        return self.paths[int(self.path_offsets[i]):int(self.path_offsets[i + 1])].tolist()


_memcpy = C.CDLL(None).memcpy
_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
_memcpy.restype = C.c_void_p


def _np_from(ptr, n, dtype, copy=True):
    """numpy view of n elements at a ctypes pointer; copy=False aliases the engine's pinned result
    buffer, which stays valid until the next call of the same kind on that engine."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)
    if not copy:
        return a
    out = np.empty(n, dtype=dtype)
    _memcpy(out.ctypes.data, a.ctypes.data, out.nbytes)  # a foreign call: the copy runs without the GIL
    return out


class Engine:
    """One engine per GPU (trgt_engine_t)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.trgt_engine_create(device, C.byref(h))
        if rc != 0:
            raise TrgtError(f"trgt_engine_create({device}) failed rc={rc}: "
                            f"{self._L.trgt_last_create_error().decode()}")
        self._h = h
        self.device = device

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.trgt_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise TrgtError(f"{what} failed rc={rc}: {self._L.trgt_engine_last_error(self._h).decode()}")

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._L

    @property
    def stream_ptr(self) -> int:
        return int(self._L.trgt_engine_stream(self._h) or 0)

    @property
    def sm_count(self) -> int:
        return self._L.trgt_engine_sm_count(self._h)

    def sync(self):
        self._check(self._L.trgt_engine_sync(self._h), "trgt_engine_sync")

    def set_workspace_budget(self, nbytes: int):
        self._L.trgt_engine_set_workspace_budget(self._h, nbytes)

    def set_hmm_lane_path(self, on: bool):
        """single-motif loci through the per-motif-length lane kernels (default) or the generic ones (diagnostic)"""
        self._L.trgt_engine_set_hmm_lane_path(self._h, int(bool(on)))

    def set_flank_band_budget(self, max_cost: int):
        self._L.trgt_engine_set_flank_band_budget(self._h, max_cost)

    # -- instrumentation ------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._L.trgt_engine_set_profiling(self._h, int(on))

    def reset_stats(self):
        self._L.trgt_engine_reset_stats(self._h)

    def launches(self) -> int:
        return int(self._L.trgt_engine_launches(self._h))

    def kernel_stats(self):
        out = {}
        for i in range(self._L.trgt_engine_kernel_count(self._h)):
            name, n, ms = C.c_char_p(), C.c_uint64(), C.c_double()
            self._L.trgt_engine_kernel_stat(self._h, i, C.byref(name), C.byref(n), C.byref(ms))
            out[name.value.decode()] = (int(n.value), float(ms.value))
        return out

    # -- phase A ---------------------------------------------------------------------------
    def flank_spans_packed(self, left: PackedSeqs, right: PackedSeqs, reads: PackedSeqs,
                           locus_read_offsets: np.ndarray, scoring=(2, 5, 1), min_flank_id_frac: float = 0.7,
                           want_hits: bool = True, spans_out: Optional[np.ndarray] = None,
                           hits_out: Optional[np.ndarray] = None):
        """trgt_flank_spans on packed inputs -> (spans[n_reads], hits[2*n_reads] or None)"""
        lro = np.ascontiguousarray(locus_read_offsets, dtype=np.uint32)
        n_reads = len(reads)
        spans = spans_out if spans_out is not None else np.zeros(n_reads, dtype=SPAN_DTYPE)
        hits = hits_out if hits_out is not None else (np.zeros(2 * n_reads, dtype=HIT_DTYPE) if want_hits else None)
        rc = self._L.trgt_flank_spans(self._h, left.ref(), right.ref(), reads.ref(), lro.ctypes.data, len(left),
                                      _Scoring(*scoring), float(min_flank_id_frac), spans.ctypes.data,
                                      hits.ctypes.data if hits is not None else None)
        self._check(rc, "trgt_flank_spans")
        return spans, hits

    def flank_spans_seq4(self, left: PackedSeqs, right: PackedSeqs, reads: PackedSeq4,
                         locus_read_offsets: np.ndarray, scoring=(2, 5, 1), min_flank_id_frac: float = 0.7,
                         want_hits: bool = True, spans_out: Optional[np.ndarray] = None,
                         hits_out: Optional[np.ndarray] = None):
        """trgt_flank_spans_seq4: the same on reads handed over as BAM 4-bit bases"""
        lro = np.ascontiguousarray(locus_read_offsets, dtype=np.uint32)
        n_reads = len(reads)
        spans = spans_out if spans_out is not None else np.zeros(n_reads, dtype=SPAN_DTYPE)
        hits = hits_out if hits_out is not None else (np.zeros(2 * n_reads, dtype=HIT_DTYPE) if want_hits else None)
        rc = self._L.trgt_flank_spans_seq4(self._h, left.ref(), right.ref(), reads.ref(), lro.ctypes.data, len(left),
                                           _Scoring(*scoring), float(min_flank_id_frac), spans.ctypes.data,
                                           hits.ctypes.data if hits is not None else None)
        self._check(rc, "trgt_flank_spans_seq4")
        return spans, hits

    def seq4_decode(self, reads: PackedSeq4) -> PackedSeqs:
        """rec.seq().as_bytes() (read.rs:104) of the clipped bases, on the device"""
        total = int(reads.lengths.sum(dtype=np.uint64))
        out = np.zeros(max(1, total), dtype=np.uint8)
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        self._check(self._L.trgt_seq4_decode(self._h, reads.ref(), out.ctypes.data, offs.ctypes.data), "trgt_seq4_decode")
        return PackedSeqs(out[:total], offs)

    def clip_reads(self, cigar_ops: np.ndarray, cigar_offsets: np.ndarray, ref_starts: np.ndarray,
                   regions: np.ndarray, locus_read_offsets: np.ndarray) -> np.ndarray:
        """clip_reads (tr.rs:186-196 -> clip_region.rs:19-186) for a chunk of loci -> CLIP_DTYPE[n_reads]"""
        ops = np.ascontiguousarray(cigar_ops, dtype=np.uint32)
        offs = np.ascontiguousarray(cigar_offsets, dtype=np.uint64)
        rs = np.ascontiguousarray(ref_starts, dtype=np.int64)
        reg = np.ascontiguousarray(regions, dtype=np.int64).reshape(-1)
        lro = np.ascontiguousarray(locus_read_offsets, dtype=np.uint32)
        n = rs.size
        out = np.zeros(max(1, n), dtype=CLIP_DTYPE)
        self._check(self._L.trgt_clip_reads(self._h, ops.ctypes.data if ops.size else None, offs.ctypes.data,
                                            rs.ctypes.data, n, reg.ctypes.data, lro.ctypes.data, lro.size - 1,
                                            out.ctypes.data), "trgt_clip_reads")
        return out[:n]

    def find_tr_spans(self, loci: Sequence[Tuple[bytes, bytes, Sequence[bytes]]], search_flank_len: int = 250,
                      min_flank_id_frac: float = 0.7, scoring=(2, 5, 1), return_hits: bool = False):
        """find_tr_spans (span_locater.rs:32-68) for many loci: loci = [(lf, rf, reads)].
        -> per locus a list of Optional[(start, end)]"""
        left = PackedSeqs.from_list([lf[len(lf) - search_flank_len:] for lf, _, _ in loci])
        right = PackedSeqs.from_list([rf[:search_flank_len] for _, rf, _ in loci])
        reads = PackedSeqs.from_list([r for _, _, rs in loci for r in rs])
        lro = _group_offsets([rs for _, _, rs in loci])
        spans, hits = self.flank_spans_packed(left, right, reads, lro, scoring, min_flank_id_frac)
        out, k = [], 0
        for _, _, rs in loci:
            cur = []
            for _ in rs:
                s = spans[k]
                cur.append((int(s["start"]), int(s["end"])) if s["found"] else None)
                k += 1
            out.append(cur)
        return (out, hits) if return_hits else out

    # -- resident batches (upload once, run many times) -------------------------------------
    def flank_upload(self, left: PackedSeqs, right: PackedSeqs, reads: PackedSeqs, locus_read_offsets: np.ndarray,
                     scoring=(2, 5, 1), min_flank_id_frac: float = 0.7):
        lro = np.ascontiguousarray(locus_read_offsets, dtype=np.uint32)
        b = C.c_void_p()
        self._check(self._L.trgt_flank_upload(self._h, left.ref(), right.ref(), reads.ref(), lro.ctypes.data,
                                              len(left), _Scoring(*scoring), float(min_flank_id_frac), C.byref(b)),
                    "trgt_flank_upload")
        self.sync()  # the host arrays may go away after this call
        return b

    def flank_upload_seq4(self, left: PackedSeqs, right: PackedSeqs, reads: PackedSeq4, locus_read_offsets: np.ndarray,
                          scoring=(2, 5, 1), min_flank_id_frac: float = 0.7):
        lro = np.ascontiguousarray(locus_read_offsets, dtype=np.uint32)
        b = C.c_void_p()
        self._check(self._L.trgt_flank_upload_seq4(self._h, left.ref(), right.ref(), reads.ref(), lro.ctypes.data,
                                                   len(left), _Scoring(*scoring), float(min_flank_id_frac), C.byref(b)),
                    "trgt_flank_upload_seq4")
        return b

    def flank_trs(self, b=None, copy: bool = False) -> PackedSeqs:
        """trs of tr.rs:58-62: read[span.start..span.end] per read (empty without a span), cut on the device
        from batch b (None: the last one-shot flank call).  copy=False: views of the engine's pinned buffers."""
        out = _SeqsOut()
        self._check(self._L.trgt_flank_trs(self._h, b, C.byref(out)), "trgt_flank_trs")
        n = int(out.n)
        offs = np.ctypeslib.as_array(out.offsets, shape=(n + 1,))
        total = int(offs[n])
        data = np.ctypeslib.as_array(out.data, shape=(total,)) if total else np.zeros(0, dtype=np.uint8)
        if copy:
            offs, data = offs.copy(), data.copy()
        return PackedSeqs(data, offs)

    def flank_run(self, b):
        self._check(self._L.trgt_flank_run(self._h, b), "trgt_flank_run")

    def flank_download(self, b, n_reads: int, want_hits: bool = True, spans_out=None, hits_out=None):
        """spans_out / hits_out: caller's (pinned) arrays to fill instead of fresh ones"""
        spans = spans_out if spans_out is not None else np.zeros(n_reads, dtype=SPAN_DTYPE)
        hits = (hits_out if hits_out is not None else np.zeros(2 * n_reads, dtype=HIT_DTYPE)) if want_hits else None
        self._check(self._L.trgt_flank_download(self._h, b, spans.ctypes.data,
                                                hits.ctypes.data if hits is not None else None), "trgt_flank_download")
        return spans, hits

    def flank_free(self, b):
        self._L.trgt_flank_free(self._h, b)

    def flank_n_wfa(self, b) -> int:
        n = C.c_uint32()
        self._L.trgt_flank_device_views(b, None, None, None, None, None, C.byref(n))
        return int(n.value)

    def flank_fallback_counts(self, b):
        """(pairs handed to the second cost tier, to the wide-band kernel, to the full-width kernels)"""
        out = (C.c_uint32 * 3)()
        self._L.trgt_flank_fallback_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        self._L.trgt_flank_fallback_counts(b, out)
        return int(out[0]), int(out[1]), int(out[2])

    def align_upload(self, backbones: PackedSeqs, seqs: PackedSeqs, group_seq_offsets: np.ndarray):
        gso = np.ascontiguousarray(group_seq_offsets, dtype=np.uint32)
        b = C.c_void_p()
        self._check(self._L.trgt_align_upload(self._h, backbones.ref(), seqs.ref(), gso.ctypes.data, len(backbones),
                                              C.byref(b)), "trgt_align_upload")
        self.sync()
        return b

    def align_run(self, b):
        self._check(self._L.trgt_align_run(self._h, b), "trgt_align_run")

    def align_download(self, b) -> "CigarBatch":
        out = _Cigars()
        self._check(self._L.trgt_align_download(self._h, b, C.byref(out)), "trgt_align_download")
        return self._cigars(out)

    def align_free(self, b):
        self._L.trgt_align_free(self._h, b)

    def hmm_upload(self, motifs: PackedSeqs, locus_motif_offsets: np.ndarray, alleles: PackedSeqs,
                   allele_locus: np.ndarray, want_paths: bool = False):
        lmo = np.ascontiguousarray(locus_motif_offsets, dtype=np.uint32)
        al = np.ascontiguousarray(allele_locus, dtype=np.uint32)
        b = C.c_void_p()
        self._check(self._L.trgt_hmm_upload(self._h, motifs.ref(), lmo.ctypes.data, lmo.size - 1, alleles.ref(),
                                            al.ctypes.data if al.size else None, int(want_paths), C.byref(b)),
                    "trgt_hmm_upload")
        self.sync()
        return b

    def hmm_run(self, b):
        self._check(self._L.trgt_hmm_run(self._h, b), "trgt_hmm_run")

    def hmm_download(self, b, want_paths: bool = False) -> "AnnotationBatch":
        out = _Annotations()
        self._check(self._L.trgt_hmm_download(self._h, b, C.byref(out)), "trgt_hmm_download")
        return self._annotations(out, want_paths)

    def vcf_fields(self, b=None, raw: bool = False):
        """AL / MC / MS / AP sample fields of every locus of HMM batch b (None: the last hmm_label call), as
        write_vcf.rs:267-343 encodes them -> list of (AL, MC, MS, AP) byte strings per locus"""
        out = _SeqsOut()
        self._check(self._L.trgt_vcf_fields(self._h, b, C.byref(out)), "trgt_vcf_fields")
        n = int(out.n)
        if n == 0:
            return (np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint8)) if raw else []
        offs = np.ctypeslib.as_array(out.offsets, shape=(n + 1,))
        total = int(offs[n])
        if raw:  # views of the engine's pinned buffers: offsets[4 * n_loci + 1], bytes
            return offs, np.ctypeslib.as_array(out.data, shape=(max(1, total),))[:total]
        data = np.ctypeslib.as_array(out.data, shape=(max(1, total),))[:total].tobytes()
        f = [data[int(offs[i]):int(offs[i + 1])] for i in range(n)]
        return [tuple(f[4 * l:4 * l + 4]) for l in range(n // 4)]

    def hmm_free(self, b):
        self._L.trgt_hmm_free(self._h, b)

    @staticmethod
    def _cigars(out, copy: bool = True) -> "CigarBatch":
        n = int(out.n)
        offs = _np_from(out.offsets, n + 1, np.uint64, copy)
        total = int(offs[n]) if n else 0
        return CigarBatch(offs, _np_from(out.words, total, np.uint32, copy), _np_from(out.scores, n, np.int32, copy),
                          _np_from(out.status, n, np.int32, copy))

    @staticmethod
    def _annotations(out, want_paths: bool, copy: bool = True) -> "AnnotationBatch":
        n = int(out.n)
        mco = _np_from(out.motif_count_offsets, n + 1, np.uint64)
        so = _np_from(out.span_offsets, n + 1, np.uint64, copy)
        n_mc = int(mco[n]) if n else 0
        n_sp = int(so[n]) if n else 0
        spans = _np_from(out.spans, 3 * n_sp, np.uint32, copy).reshape(-1, 3)
        res = AnnotationBatch(mco, _np_from(out.motif_counts, n_mc, np.uint32, copy), so, spans,
                              _np_from(out.purity, n, np.float64, copy), _np_from(out.status, n, np.int32, copy))
        if want_paths:
            po = _np_from(out.path_offsets, n + 1, np.uint64, copy)
            res.path_offsets = po
            res.paths = _np_from(out.paths, int(po[n]) if n else 0, np.uint32, copy)
        return res

    def pinned_array(self, nbytes: int) -> np.ndarray:
        """uint8 array in pinned host memory (trgt_host_alloc); freed when the array is collected."""
        ptr = self._L.trgt_host_alloc(max(1, nbytes))
        if not ptr:
            raise MemoryError(f"trgt_host_alloc({nbytes}) failed")
        buf = (C.c_uint8 * max(1, nbytes)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.uint8)
        lib = self._L

        class _Owner:
            def __del__(self_inner):
                lib.trgt_host_free(ptr)
        _PINNED_OWNERS[ptr] = (_Owner(), buf)
        return arr

    # -- phase B ---------------------------------------------------------------------------
    def align_packed(self, backbones: PackedSeqs, seqs: PackedSeqs, group_seq_offsets: np.ndarray,
                     copy: bool = True) -> CigarBatch:
        """copy=False: the result arrays alias the engine's pinned buffers (valid until the next align call)"""
        gso = np.ascontiguousarray(group_seq_offsets, dtype=np.uint32)
        out = _Cigars()
        rc = self._L.trgt_align_e2e(self._h, backbones.ref(), seqs.ref(), gso.ctypes.data, len(backbones), C.byref(out))
        self._check(rc, "trgt_align_e2e")
        return self._cigars(out, copy)

    def align(self, groups: Sequence[Tuple[bytes, Sequence[bytes]]]) -> List[List[List[Tuple[int, str]]]]:
        """utils::align (src/utils/align.rs:14-28) for many (backbone, seqs) groups."""
        bb = PackedSeqs.from_list([b for b, _ in groups])
        sq = PackedSeqs.from_list([s for _, ss in groups for s in ss])
        res = self.align_packed(bb, sq, _group_offsets([ss for _, ss in groups]))
        out, k = [], 0
        for _, ss in groups:
            cur = []
            for _ in ss:
                cur.append(decode_sam_cigar(res.cigar(k)))
                k += 1
            out.append(cur)
        return out

    def consensus_packed(self, backbones: PackedSeqs, seqs: PackedSeqs, group_seq_offsets: np.ndarray):
        """trgt_consensus -> (PackedSeqs of one repaired consensus per group, status int32[n_groups])"""
        gso = np.ascontiguousarray(group_seq_offsets, dtype=np.uint32)
        out = _SeqsOut()
        rc = self._L.trgt_consensus(self._h, backbones.ref(), seqs.ref(), gso.ctypes.data, len(backbones), C.byref(out))
        self._check(rc, "trgt_consensus")
        n = int(out.n)
        offs = _np_from(out.offsets, n + 1, np.uint64)
        total = int(offs[n]) if n else 0
        return PackedSeqs(_np_from(out.data, total, np.uint8), offs), _np_from(out.status, n, np.int32)

    def repair_consensus(self, groups: Sequence[Tuple[bytes, Sequence[bytes]]]) -> List[bytes]:
        """utils::align + repair_consensus (src/trgt/genotype/consensus.rs:5) for many (backbone, seqs) groups."""
        bb = PackedSeqs.from_list([b for b, _ in groups])
        sq = PackedSeqs.from_list([s for _, ss in groups for s in ss])
        cons, status = self.consensus_packed(bb, sq, _group_offsets([ss for _, ss in groups]))
        bad = np.nonzero(status)[0]
        if bad.size:
            raise TrgtError(f"trgt_consensus: group {int(bad[0])} failed with status {int(status[bad[0]])}")
        return [cons.get(i) for i in range(len(groups))]

    def edit_dist_packed(self, seqs: PackedSeqs, locus_seq_offsets: np.ndarray) -> np.ndarray:
        lso = np.ascontiguousarray(locus_seq_offsets, dtype=np.uint32)
        n = np.diff(lso.astype(np.int64))
        total = int((n * (n - 1) // 2).clip(min=0).sum())
        out = np.zeros(total, dtype=np.float64)
        rc = self._L.trgt_edit_dist(self._h, seqs.ref(), lso.ctypes.data, lso.size - 1, out.ctypes.data)
        self._check(rc, "trgt_edit_dist")
        return out

    def get_dist_matrix(self, loci_trs: Sequence[Sequence[bytes]]) -> List[List[float]]:
        """get_dist_matrix (genotype_cluster.rs:250-286) for many loci -> condensed upper triangles."""
        sq = PackedSeqs.from_list([s for trs in loci_trs for s in trs])
        flat = self.edit_dist_packed(sq, _group_offsets(loci_trs))
        out, k = [], 0
        for trs in loci_trs:
            n = len(trs)
            m = n * (n - 1) // 2 if n >= 2 else 0
            out.append(flat[k:k + m].tolist())
            k += m
        return out

    # -- cluster-genotyper glue (next row, rank 3) ------------------------------------------
    def cluster_packed(self, seqs: PackedSeqs, locus_seq_offsets: np.ndarray):
        """trgt_cluster -> (group int32[n_seqs]: 0 group1, 1 group2, 2 neither; central uint32[n_loci, 2];
        n_groups uint32[n_loci])"""
        lso = np.ascontiguousarray(locus_seq_offsets, dtype=np.uint32)
        n_loci = lso.size - 1
        group = np.zeros(max(1, len(seqs)), dtype=np.int32)
        central = np.zeros((max(1, n_loci), 2), dtype=np.uint32)
        ng = np.zeros(max(1, n_loci), dtype=np.uint32)
        rc = self._L.trgt_cluster(self._h, seqs.ref(), lso.ctypes.data, n_loci, group.ctypes.data, central.ctypes.data,
                                  ng.ctypes.data)
        self._check(rc, "trgt_cluster")
        return group[:len(seqs)], central[:n_loci], ng[:n_loci]

    def cluster(self, loci_trs: Sequence[Sequence[bytes]]):
        """cluster() + group1 / group2 + central_read (genotype_cluster.rs:12-39,57-72,154-227) for many loci
        -> per locus (sel list, (central1, central2 or None), n_groups)"""
        sq = PackedSeqs.from_list([s for trs in loci_trs for s in trs])
        lso = _group_offsets(loci_trs)
        group, central, ng = self.cluster_packed(sq, lso)
        out = []
        for l, trs in enumerate(loci_trs):
            a, b = int(lso[l]), int(lso[l + 1])
            out.append((group[a:b].tolist(), tuple(None if c == 0xFFFFFFFF else int(c) for c in central[l]), int(ng[l])))
        return out

    def bamlet_clip(self, b, cigar_ops: np.ndarray, cigar_offsets: np.ndarray, ref_starts: np.ndarray,
                    flank_len: int) -> np.ndarray:
        """clip_bases as BamWriter::write asks for it (write_bam.rs:72-92, clip_bases.rs:9-119) for every read of
        flank batch b (None: the last one-shot call) -> BAMLET_CLIP_DTYPE[n_reads]"""
        ops = np.ascontiguousarray(cigar_ops, dtype=np.uint32)
        offs = np.ascontiguousarray(cigar_offsets, dtype=np.uint64)
        rs = np.ascontiguousarray(ref_starts, dtype=np.int64)
        n = rs.size
        out = np.zeros(max(1, n), dtype=BAMLET_CLIP_DTYPE)
        self._check(self._L.trgt_bamlet_clip(self._h, b, ops.ctypes.data if ops.size else None, offs.ctypes.data,
                                             rs.ctypes.data, int(flank_len), out.ctypes.data), "trgt_bamlet_clip")
        return out[:n]

    def cluster_trs(self, b, reads: np.ndarray, locus_offsets: np.ndarray):
        """trgt_cluster_trs: the same on the repeat sequences of flank batch b (None: the last one-shot call),
        named by read index"""
        rd = np.ascontiguousarray(reads, dtype=np.uint32)
        lo = np.ascontiguousarray(locus_offsets, dtype=np.uint32)
        n_loci = lo.size - 1
        group = np.zeros(max(1, rd.size), dtype=np.int32)
        central = np.zeros((max(1, n_loci), 2), dtype=np.uint32)
        ng = np.zeros(max(1, n_loci), dtype=np.uint32)
        rc = self._L.trgt_cluster_trs(self._h, b, rd.ctypes.data if rd.size else None, lo.ctypes.data, n_loci,
                                      group.ctypes.data, central.ctypes.data, ng.ctypes.data)
        self._check(rc, "trgt_cluster_trs")
        return group[:rd.size], central[:n_loci], ng[:n_loci]

    def align_trs(self, b, backbone_reads: np.ndarray, member_reads: np.ndarray, group_offsets: np.ndarray,
                  copy: bool = True) -> CigarBatch:
        """trgt_align_trs: utils::align with backbones and members named by read index of flank batch b (None: the last
        one-shot call); copy=False: the result arrays alias the engine's pinned buffers"""
        bb = np.ascontiguousarray(backbone_reads, dtype=np.uint32)
        mr = np.ascontiguousarray(member_reads, dtype=np.uint32)
        go = np.ascontiguousarray(group_offsets, dtype=np.uint32)
        out = _Cigars()
        rc = self._L.trgt_align_trs(self._h, b, bb.ctypes.data if bb.size else None, mr.ctypes.data if mr.size else None,
                                    go.ctypes.data, bb.size, C.byref(out))
        self._check(rc, "trgt_align_trs")
        return self._cigars(out, copy)

    def consensus_trs(self, b, backbone_reads: np.ndarray, member_reads: np.ndarray, group_offsets: np.ndarray):
        """trgt_consensus_trs -> (PackedSeqs of one repaired consensus per group, status int32[n_groups])"""
        bb = np.ascontiguousarray(backbone_reads, dtype=np.uint32)
        mr = np.ascontiguousarray(member_reads, dtype=np.uint32)
        go = np.ascontiguousarray(group_offsets, dtype=np.uint32)
        out = _SeqsOut()
        rc = self._L.trgt_consensus_trs(self._h, b, bb.ctypes.data if bb.size else None, mr.ctypes.data if mr.size else None,
                                        go.ctypes.data, bb.size, C.byref(out))
        self._check(rc, "trgt_consensus_trs")
        n = int(out.n)
        offs = _np_from(out.offsets, n + 1, np.uint64)
        total = int(offs[n]) if n else 0
        return PackedSeqs(_np_from(out.data, total, np.uint8), offs), _np_from(out.status, n, np.int32)

    # -- phase C ---------------------------------------------------------------------------
    def hmm_label_packed(self, motifs: PackedSeqs, locus_motif_offsets: np.ndarray, alleles: PackedSeqs,
                         allele_locus: np.ndarray, want_paths: bool = False, copy: bool = True) -> AnnotationBatch:
        lmo = np.ascontiguousarray(locus_motif_offsets, dtype=np.uint32)
        al = np.ascontiguousarray(allele_locus, dtype=np.uint32)
        out = _Annotations()
        rc = self._L.trgt_hmm_label(self._h, motifs.ref(), lmo.ctypes.data, lmo.size - 1, alleles.ref(),
                                    al.ctypes.data if al.size else None, int(want_paths), C.byref(out))
        self._check(rc, "trgt_hmm_label")
        return self._annotations(out, want_paths, copy)

    def label_with_hmm(self, loci: Sequence[Tuple[Sequence[bytes], Sequence[bytes]]]) -> List[List[Annotation]]:
        """label_with_hmm (tr.rs:454-492) for many loci: loci = [(motifs, allele_seqs)]."""
        motifs = PackedSeqs.from_list([m for ms, _ in loci for m in ms])
        lmo = _group_offsets([ms for ms, _ in loci])
        alleles = PackedSeqs.from_list([a for _, als in loci for a in als])
        al = np.array([i for i, (_, als) in enumerate(loci) for _ in als], dtype=np.uint32)
        res = self.hmm_label_packed(motifs, lmo, alleles, al)
        bad = np.nonzero(res.status)[0]
        if bad.size:
            raise TrgtError(f"trgt_hmm_label: allele {int(bad[0])} failed with status {int(res.status[bad[0]])}")
        out, k = [], 0
        for _, als in loci:
            cur = []
            for _ in als:
                cur.append(res.annotation(k))
                k += 1
            out.append(cur)
        return out
