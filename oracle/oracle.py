"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/trgt_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (trgt_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrgt_oracle.so")

INDEL, EDIT, LINEAR, AFFINE, AFFINE2P = 0, 1, 2, 3, 4


def _cpu_stamp() -> str:
    """identifies the host CPU: the library is compiled with -march=native (CPU baseline of bench.py)"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile oracle/*.c into libtrgt_oracle.so (gcc via oracle/Makefile, -O3 -march=native): rebuilt when a source
    is newer or when the library was built on a different CPU (it travels to the GPU box with the snapshot)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stamp_path = os.path.join(_HERE, ".build_cpu")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    try:
        with open(stamp_path) as f:
            built_for = f.read().strip()
    except OSError:
        built_for = ""
    if force or stale or built_for != _cpu_stamp():
        subprocess.run(["make", "-C", _HERE, "-B", "libtrgt_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
        with open(stamp_path, "w") as f:
            f.write(_cpu_stamp())
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("metric", C.c_int), ("x", C.c_int), ("o1", C.c_int), ("e1", C.c_int),
        ("o2", C.c_int), ("e2", C.c_int), ("ends_free", C.c_int),
        ("pattern_begin_free", C.c_int), ("pattern_end_free", C.c_int),
        ("text_begin_free", C.c_int), ("text_end_free", C.c_int),
        ("score_only", C.c_int), ("max_steps", C.c_int),
        ("wfadaptive_min_len", C.c_int), ("wfadaptive_max_dist", C.c_int),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("score", C.c_int), ("ops", C.POINTER(C.c_uint8)),
        ("n_ops", C.c_int64), ("end_k", C.c_int), ("end_offset", C.c_int),
    ]


class _Span(C.Structure):
    _fields_ = [("motif_index", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32)]


class _OptSpan(C.Structure):
    _fields_ = [("found", C.c_int32), ("start", C.c_uint32), ("end", C.c_uint32)]


class _BamletClip(C.Structure):
    _fields_ = [("ref_pos", C.c_int64), ("base_start", C.c_uint64), ("base_end", C.c_uint64),
                ("meth_start", C.c_uint32), ("meth_end", C.c_uint32), ("first_op", C.c_uint32), ("n_ops", C.c_uint32),
                ("first_word", C.c_uint32), ("last_word", C.c_uint32), ("has_cigar", C.c_int32), ("pad", C.c_int32)]


class _Clip(C.Structure):
    _fields_ = [("ref_start", C.c_int64), ("query_start", C.c_uint64), ("query_end", C.c_uint64),
                ("first_op", C.c_uint32), ("n_ops", C.c_uint32), ("first_word", C.c_uint32),
                ("last_word", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.tro_hmm_build.restype = C.c_void_p
        L.tro_hmm_build.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.c_uint32]
        L.tro_hmm_free.argtypes = [C.c_void_p]
        L.tro_hmm_num_states.argtypes = [C.c_void_p]
        L.tro_hmm_em.restype = C.c_double
        L.tro_hmm_em.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.tro_hmm_num_in.argtypes = [C.c_void_p, C.c_int]
        L.tro_hmm_in_state.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.tro_hmm_in_lp.restype = C.c_double
        L.tro_hmm_in_lp.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.tro_hmm_label.restype = C.c_int64
        L.tro_hmm_label.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.c_uint64]
        L.tro_remove_imperfect_motifs.restype = C.c_int64
        L.tro_remove_imperfect_motifs.argtypes = [
            C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64, C.c_char_p, C.c_uint32, C.c_uint32,
            C.POINTER(C.c_uint32), C.c_uint64]
        L.tro_label_motifs.restype = C.c_int64
        L.tro_label_motifs.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64,
                                       C.POINTER(_Span), C.c_uint64]
        L.tro_calc_purity.restype = C.c_double
        L.tro_calc_purity.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64,
                                      C.c_char_p, C.c_uint32]
        L.tro_get_base_match.restype = C.c_uint8
        L.tro_get_base_match.argtypes = [C.c_void_p, C.c_int]
        L.tro_annotate_allele.restype = C.c_int64
        L.tro_annotate_allele.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32,
                                          C.POINTER(C.c_uint32), C.POINTER(_Span), C.c_uint64,
                                          C.POINTER(C.c_double)]
        L.tro_wfa_align.argtypes = [C.POINTER(_Params), C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                    C.POINTER(_Result)]
        L.tro_wfa_result_free.argtypes = [C.POINTER(_Result)]
        L.tro_sam_cigar.restype = C.c_int64
        L.tro_sam_cigar.argtypes = [C.POINTER(C.c_uint8), C.c_int64, C.c_int,
                                    C.POINTER(C.c_uint32), C.c_uint64]
        L.tro_count_matches.argtypes = [C.POINTER(C.c_uint8), C.c_int64]
        L.tro_alignment_span.argtypes = [C.POINTER(C.c_uint8), C.c_int64, C.c_int, C.c_int,
                                         C.c_int] + [C.POINTER(C.c_int)] * 4
        L.tro_cigar_score.argtypes = [C.POINTER(_Params), C.POINTER(C.c_uint8), C.c_int64]
        L.tro_cigar_score_clipped.argtypes = [C.POINTER(_Params), C.POINTER(C.c_uint8), C.c_int64,
                                              C.c_int]
        L.tro_find_span.restype = _OptSpan
        L.tro_find_span.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tro_align_consensus.restype = C.c_int64
        L.tro_align_consensus.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                          C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_int)]
        L.tro_get_dist.restype = C.c_double
        L.tro_get_dist.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        vp = C.c_void_p
        L.tro_flank_batch.argtypes = [vp] * 7 + [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_double, vp, vp, C.c_int]
        L.tro_align_batch.argtypes = [vp] * 5 + [C.c_uint32, vp, C.POINTER(vp), vp, C.c_int]
        L.tro_hmm_batch.argtypes = [vp] * 3 + [C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(vp),
                                    vp, vp, C.c_int]
        L.tro_repair_consensus.restype = C.c_int64
        L.tro_repair_consensus.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, vp, C.c_uint32, vp, vp, vp, C.c_uint64]
        L.tro_clip_cigar.argtypes = [vp, C.c_uint32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_Clip)]
        L.tro_decode_seq4.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
        L.tro_meth_range.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.tro_meth_range.restype = None
        L.tro_clip_bases.argtypes = [vp, C.c_uint32, C.c_int64, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.POINTER(_BamletClip)]
        L.tro_bamlet_clip.argtypes = [vp, C.c_uint32, C.c_int64, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_uint64, C.POINTER(_BamletClip)]
        L.tro_decode_seq4.restype = None
        L.tro_vcf_field.restype = C.c_size_t
        L.tro_vcf_field.argtypes = [C.c_int, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_char_p, C.c_size_t]
        L.tro_free.argtypes = [vp]
        L.tro_free.restype = None
        _lib = L
    return _lib


# ------------------------------------------------------------------ HMM --

def replace_invalid_bases(seq: bytes, allowed: bytes) -> bytes:
    """src/hmm/utils.rs:29-42"""
    return bytes(b if b in allowed else allowed[i % len(allowed)] for i, b in enumerate(seq))


class Hmm:
    """build_hmm (src/hmm/builder.rs:4) + Hmm methods (src/hmm/hmm_model.rs)."""

    def __init__(self, motifs: Sequence[bytes]):
        self.motifs = [bytes(m) for m in motifs]
        data = b"".join(self.motifs)
        offs = [0]
        for m in self.motifs:
            offs.append(offs[-1] + len(m))
        arr = (C.c_uint32 * len(offs))(*offs)
        self._h = lib().tro_hmm_build(data, arr, len(self.motifs))
        if not self._h:
            raise ValueError("invalid motif base")
        self.num_states = lib().tro_hmm_num_states(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tro_hmm_free(self._h)
            self._h = None

    def label(self, query: bytes) -> List[int]:
        cap = (len(query) + 2) * self.num_states + 8
        out = (C.c_uint32 * cap)()
        n = lib().tro_hmm_label(self._h, query, len(query), out, cap)
        if n < 0:
            raise ValueError(f"label failed rc={n}")
        return list(out[:n])

    def remove_imperfect_motifs(self, states: Sequence[int], query: bytes,
                                max_motif_len: int = 6) -> List[int]:
        cap = len(states) + 3 * len(query) + 16
        inp = (C.c_uint32 * max(1, len(states)))(*states)
        out = (C.c_uint32 * cap)()
        n = lib().tro_remove_imperfect_motifs(self._h, inp, len(states), query, len(query),
                                              max_motif_len, out, cap)
        if n < 0:
            raise ValueError(f"remove_imperfect_motifs failed rc={n}")
        return list(out[:n])

    def label_motifs(self, states: Sequence[int]) -> List[Tuple[int, int, int]]:
        """-> [(motif_index, start, end)]"""
        cap = len(states) + 1
        inp = (C.c_uint32 * max(1, len(states)))(*states)
        out = (_Span * cap)()
        n = lib().tro_label_motifs(self._h, inp, len(states), out, cap)
        if n < 0:
            raise ValueError(f"label_motifs failed rc={n}")
        return [(out[i].motif_index, out[i].start, out[i].end) for i in range(n)]

    def calc_purity(self, query: bytes, states: Sequence[int]) -> float:
        inp = (C.c_uint32 * max(1, len(states)))(*states)
        return lib().tro_calc_purity(self._h, inp, len(states), query, len(query))

    def get_base_match(self, state: int) -> bytes:
        return bytes([lib().tro_get_base_match(self._h, state)])

    def annotate(self, allele: bytes):
        """label_with_hmm for one allele (src/trgt/workflows/tr.rs:464-489)
        -> (motif_counts, collapsed spans [(motif,start,end)], purity)"""
        nm = len(self.motifs)
        mc = (C.c_uint32 * max(1, nm))()
        cap = len(allele) + 1
        spans = (_Span * cap)()
        purity = C.c_double()
        n = lib().tro_annotate_allele(self._h, allele, len(allele), mc, spans, cap,
                                      C.byref(purity))
        if n < 0:
            raise ValueError(f"annotate failed rc={n}")
        return (list(mc[:nm]), [(spans[i].motif_index, spans[i].start, spans[i].end)
                                for i in range(n)], purity.value)

    def tables(self):
        """(ems[S][5], in_states[S][..], in_lps[S][..]) for model-table parity tests."""
        L = lib()
        ems, ins, lps = [], [], []
        for s in range(self.num_states):
            ems.append([L.tro_hmm_em(self._h, s, i) for i in range(5)])
            n = L.tro_hmm_num_in(self._h, s)
            ins.append([L.tro_hmm_in_state(self._h, s, i) for i in range(n)])
            lps.append([L.tro_hmm_in_lp(self._h, s, i) for i in range(n)])
        return ems, ins, lps


# ------------------------------------------------------------------ WFA --

@dataclass
class Alignment:
    status: int
    score: int
    ops: bytes            # b"MXID..." forward order
    end_k: int
    end_offset: int
    params: "_Params"
    plen: int
    tlen: int

    def cigar_string(self, flank_len: int = 0) -> str:
        """wfaligner.rs cigar_string(Some(flank_len))"""
        ops = self.ops[flank_len:len(self.ops) - flank_len] if flank_len else self.ops
        out, i = [], 0
        while i < len(ops):
            j = i
            while j < len(ops) and ops[j] == ops[i]:
                j += 1
            out.append(f"{j - i}{chr(ops[i])}")
            i = j
        return "".join(out)

    def _ops_ptr(self):
        buf = (C.c_uint8 * max(1, len(self.ops))).from_buffer_copy(self.ops or b"\0")
        return buf

    def count_matches(self) -> int:
        return lib().tro_count_matches(self._ops_ptr(), len(self.ops))

    def alignment_span(self):
        """get_alignment_span -> ((xstart,xend),(ystart,yend))"""
        xs, xe, ys, ye = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().tro_alignment_span(self._ops_ptr(), len(self.ops), self.params.ends_free,
                                 self.plen, self.tlen, C.byref(xs), C.byref(xe), C.byref(ys),
                                 C.byref(ye))
        return (xs.value, xe.value), (ys.value, ye.value)

    def sam_cigar(self, show_mismatches: bool = True) -> List[int]:
        cap = len(self.ops) + 1
        out = (C.c_uint32 * cap)()
        n = lib().tro_sam_cigar(self._ops_ptr(), len(self.ops), int(show_mismatches), out, cap)
        return list(out[:n])

    def cigar_score(self) -> int:
        return lib().tro_cigar_score(C.byref(self.params), self._ops_ptr(), len(self.ops))

    def cigar_score_clipped(self, flank_len: int) -> int:
        return lib().tro_cigar_score_clipped(C.byref(self.params), self._ops_ptr(),
                                             len(self.ops), flank_len)


def decode_sam_cigar(words: Sequence[int]) -> List[Tuple[int, str]]:
    """WFAligner::decode_sam_cigar (wfaligner.rs:961-984)"""
    tab = "MIDNSHP=X"
    return [(w >> 4, tab[w & 0xF] if (w & 0xF) < len(tab) else "?") for w in words]


def wfa_align(pattern: bytes, text: bytes, metric: int = AFFINE, x: int = 0, o: int = 0,
              e: int = 0, o2: int = 0, e2: int = 0, ends_free: Optional[Tuple[int, int, int, int]] = None,
              score_only: bool = False, max_steps: int = 0, wfadaptive: Optional[Tuple[int, int]] = None) -> Alignment:
    """ends_free = (pattern_begin_free, pattern_end_free, text_begin_free, text_end_free);
    wfadaptive = (min_wavefront_length, max_distance_threshold): WFA2-lib's adaptive heuristic, a measuring device"""
    p = _Params(metric, x, o, e, o2, e2, 0, 0, 0, 0, 0, int(score_only), max_steps, 0, 0)
    if wfadaptive is not None:
        p.wfadaptive_min_len, p.wfadaptive_max_dist = wfadaptive
    if ends_free is not None:
        p.ends_free = 1
        (p.pattern_begin_free, p.pattern_end_free, p.text_begin_free, p.text_end_free) = ends_free
    r = _Result()
    lib().tro_wfa_align(C.byref(p), pattern, len(pattern), text, len(text), C.byref(r))
    ops = bytes(r.ops[:r.n_ops]) if r.ops else b""
    out = Alignment(r.status, r.score, ops, r.end_k, r.end_offset, p, len(pattern), len(text))
    lib().tro_wfa_result_free(C.byref(r))
    return out


def find_span(piece: bytes, seq: bytes, scoring=(2, 5, 1), threshold: Optional[float] = None):
    """find_spans for one read (span_locater.rs:7-30) -> (span or None, via, matches)"""
    if threshold is None:
        threshold = len(piece) * 0.7
    via, nm = C.c_int(), C.c_int()
    r = lib().tro_find_span(piece, len(piece), seq, len(seq), scoring[0], scoring[1], scoring[2],
                            threshold, C.byref(via), C.byref(nm))
    return ((r.start, r.end) if r.found else None), via.value, nm.value


def find_tr_spans(lf: bytes, rf: bytes, reads: Sequence[bytes], search_flank_len: int = 250,
                  min_flank_id_frac: float = 0.7, scoring=(2, 5, 1)):
    """find_tr_spans (span_locater.rs:32-68)"""
    lf_piece = lf[len(lf) - search_flank_len:]
    rf_piece = rf[:search_flank_len]
    thr = search_flank_len * min_flank_id_frac
    out = []
    for r in reads:
        a, _, _ = find_span(lf_piece, r, scoring, thr)
        b, _, _ = find_span(rf_piece, r, scoring, thr)
        if a is not None and b is not None and a[1] <= b[0]:
            out.append((a[1], b[0]))
        else:
            out.append(None)
    return out


def align(backbone: bytes, seqs: Sequence[bytes]) -> List[List[Tuple[int, str]]]:
    """utils::align (src/utils/align.rs:14-28)"""
    out = []
    for s in seqs:
        cap = len(backbone) + len(s) + 2
        buf = (C.c_uint32 * cap)()
        sc = C.c_int()
        n = lib().tro_align_consensus(backbone, len(backbone), s, len(s), buf, cap, C.byref(sc))
        out.append(decode_sam_cigar(buf[:n]))
    return out


def align_words(backbone: bytes, seq: bytes) -> Tuple[List[int], int]:
    cap = len(backbone) + len(seq) + 2
    buf = (C.c_uint32 * cap)()
    sc = C.c_int()
    n = lib().tro_align_consensus(backbone, len(backbone), seq, len(seq), buf, cap, C.byref(sc))
    return list(buf[:n]), sc.value


def get_dist(a: bytes, b: bytes) -> float:
    """genotype_cluster.rs:236-248"""
    return lib().tro_get_dist(a, len(a), b, len(b))


def get_dist_matrix(trs: Sequence[bytes]) -> List[float]:
    """genotype_cluster.rs:250-286 (condensed upper triangle)"""
    n = len(trs)
    return [get_dist(trs[i], trs[j]) for i in range(n) for j in range(i + 1, n)]


# ------------------------------------------------- next row: cluster-genotyper glue --

class _LinkStep(C.Structure):
    _fields_ = [("cluster1", C.c_uint32), ("cluster2", C.c_uint32), ("dissimilarity", C.c_double), ("size", C.c_uint32)]


def ward_linkage(dists: Sequence[float], n: int):
    """kodama::linkage(dists, n, Method::Ward) (genotype_cluster.rs:161)
    -> ([(cluster1, cluster2, dissimilarity, size)], the condensed matrix as the call leaves it)"""
    np = _np()
    d = np.array(list(dists) + [0.0], dtype=np.float64)
    steps = (_LinkStep * max(1, n - 1))()
    L = lib()
    L.tro_ward_linkage.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    if L.tro_ward_linkage(d.ctypes.data, n, steps) != 0:
        raise MemoryError
    return ([(steps[i].cluster1, steps[i].cluster2, steps[i].dissimilarity, steps[i].size) for i in range(max(0, n - 1))],
            d[:-1])


def cluster_locus(dists: Sequence[float], n: int):
    """genotype() of genotype_cluster.rs:57-72 up to the make_consensus calls -> (sel[n]: 0 group1, 1 group2,
    2 other; (central_read of group1, of group2 or None), number of groups)"""
    np = _np()
    d = np.array(list(dists) + [0.0], dtype=np.float64)
    sel = np.zeros(max(1, n), dtype=np.uint8)
    central = (C.c_uint32 * 2)()
    L = lib()
    L.tro_cluster_locus.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    ng = L.tro_cluster_locus(n, d.ctypes.data, sel.ctypes.data, central)
    if ng < 0:
        raise MemoryError
    return sel[:n].tolist(), tuple(None if c == 0xFFFFFFFF else int(c) for c in central), ng


# ------------------------------------------------- batched drivers (numpy) --

def _np():
    import numpy as np
    return np


def flank_batch(left, right, reads, locus_read_off, scoring=(2, 5, 1), min_flank_id_frac=0.7,
                n_threads: int = 1, want_hits: bool = True):
    """tro_flank_batch on CSR sets (objects with .data uint8 / .offsets uint64 numpy arrays).
    -> (spans structured array (found,start,end), hits structured array or None)"""
    np = _np()
    lro = np.ascontiguousarray(locus_read_off, dtype=np.uint32)
    n_reads = reads.offsets.size - 1
    spans = np.zeros(n_reads, dtype=np.dtype([("found", np.int32), ("start", np.uint32), ("end", np.uint32)]))
    hits = np.zeros(2 * n_reads, dtype=np.dtype([("via", np.int32), ("matches", np.int32), ("score", np.int32),
                                                 ("start", np.uint32), ("end", np.uint32)])) if want_hits else None
    lib().tro_flank_batch(left.data.ctypes.data, left.offsets.ctypes.data, right.data.ctypes.data,
                          right.offsets.ctypes.data, reads.data.ctypes.data, reads.offsets.ctypes.data,
                          lro.ctypes.data, lro.size - 1, scoring[0], scoring[1], scoring[2],
                          float(min_flank_id_frac), spans.ctypes.data,
                          hits.ctypes.data if hits is not None else None, n_threads)
    return spans, hits


def align_batch(backbones, seqs, group_seq_off, n_threads: int = 1):
    """tro_align_batch -> (offsets uint64[n+1], words uint32, scores int32)"""
    np = _np()
    gso = np.ascontiguousarray(group_seq_off, dtype=np.uint32)
    n = seqs.offsets.size - 1
    offs = np.zeros(n + 1, dtype=np.uint64)
    scores = np.zeros(max(n, 1), dtype=np.int32)
    wp = C.c_void_p()
    lib().tro_align_batch(backbones.data.ctypes.data, backbones.offsets.ctypes.data, seqs.data.ctypes.data,
                          seqs.offsets.ctypes.data, gso.ctypes.data, gso.size - 1, offs.ctypes.data, C.byref(wp),
                          scores.ctypes.data, n_threads)
    tot = int(offs[n])
    words = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_uint32)), shape=(max(tot, 1),))[:tot].copy()
    lib().tro_free(wp)
    return offs, words, scores[:n]


def hmm_batch(motifs, locus_motif_off, alleles, allele_locus, n_threads: int = 1):
    """tro_hmm_batch -> (mc_off, mc, span_off, spans[n,3], purity, status)"""
    np = _np()
    lmo = np.ascontiguousarray(locus_motif_off, dtype=np.uint32)
    al = np.ascontiguousarray(allele_locus, dtype=np.uint32)
    n = al.size
    nm = np.diff(lmo.astype(np.int64))
    mc_off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        mc_off[1:] = np.cumsum(nm[al])
    mc = np.zeros(max(int(mc_off[n]), 1), dtype=np.uint32)
    span_off = np.zeros(n + 1, dtype=np.uint64)
    purity = np.zeros(max(n, 1), dtype=np.float64)
    status = np.zeros(max(n, 1), dtype=np.int32)
    sp = C.c_void_p()
    lib().tro_hmm_batch(motifs.data.ctypes.data, motifs.offsets.ctypes.data, lmo.ctypes.data, lmo.size - 1,
                        alleles.data.ctypes.data, alleles.offsets.ctypes.data, al.ctypes.data, n,
                        mc_off.ctypes.data, mc.ctypes.data, span_off.ctypes.data, C.byref(sp),
                        purity.ctypes.data, status.ctypes.data, n_threads)
    tot = int(span_off[n])
    spans = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_uint32)), shape=(max(3 * tot, 1),))[:3 * tot].copy()
    lib().tro_free(sp)
    return mc_off, mc[:int(mc_off[n])], span_off, spans.reshape(-1, 3), purity[:n], status[:n]


def repair_consensus(backbone: bytes, seqs: Sequence[bytes], cigars: Optional[Sequence[Sequence[int]]] = None) -> bytes:
    """utils::align + repair_consensus (consensus.rs:5-72), as genotype_cluster.rs:52-53 chains them.
    cigars (optional): run-length SAM words per sequence to use instead of the exact alignments (for measuring
    how sensitive the consensus is to the choice among co-optimal alignments)"""
    np = _np()
    offs = [0]
    for s_ in seqs:
        offs.append(offs[-1] + len(s_))
    seq_off = np.array(offs, dtype=np.uint64)
    words, woff = [], [0]
    for i, s_ in enumerate(seqs):
        w = list(cigars[i]) if cigars is not None else align_words(backbone, s_)[0]
        words += w
        woff.append(len(words))
    warr = np.array(words if words else [0], dtype=np.uint32)
    woffa = np.array(woff, dtype=np.uint64)
    cap = len(backbone) + sum(len(s_) for s_ in seqs) + 8
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().tro_repair_consensus(backbone, len(backbone), b"".join(seqs), seq_off.ctypes.data, len(seqs),
                                   warr.ctypes.data, woffa.ctypes.data, out.ctypes.data, cap)
    if n < 0:
        raise ValueError(f"repair_consensus failed rc={n}")
    return out[:n].tobytes()


# ------------------------------------------------------------------ next row: read clipping --

BAM_OPS = "MIDNSHP=X"
SEQ4_ALPHABET = b"=ACMGRSVTWYHKDBN"


def encode_bam_cigar(text: str) -> List[int]:
    """'3=2D5I' -> BAM words (len<<4)|op"""
    import re
    return [(int(n) << 4) | BAM_OPS.index(op) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", text)]


def clip_cigar(ops: Sequence[int], ref_pos: int, region: Tuple[int, int]):
    """clip_cigar + the query range of clip_to_region (clip_region.rs:19-38,105-186).
    None = no overlap; else (clipped_ref_start, query_start, query_end, clipped ops as BAM words)."""
    np = _np()
    arr = np.array(list(ops) if len(ops) else [0], dtype=np.uint32)
    out = _Clip()
    rc = lib().tro_clip_cigar(arr.ctypes.data, len(ops), ref_pos, region[0], region[1], C.byref(out))
    if rc < 0:
        raise ValueError("Unexpected operation")
    if rc == 0:
        return None
    words = []
    for i in range(out.n_ops):
        if i == 0:
            words.append(out.first_word)
        elif i == out.n_ops - 1:
            words.append(out.last_word)
        else:
            words.append(int(arr[out.first_op + i]))
    return out.ref_start, out.query_start, out.query_end, words


def meth_range(bases: bytes, start: int, end: int) -> Tuple[int, int]:
    """which methylation entries a clip to bases [start, end) keeps (clip_region.rs:40-58, clip_bases.rs:23-44)"""
    m0, m1 = C.c_uint32(), C.c_uint32()
    lib().tro_meth_range(bases, len(bases), start, end, C.byref(m0), C.byref(m1))
    return m0.value, m1.value


def _bamlet_words(out, arr):
    words = []
    for i in range(out.n_ops):
        if i == 0:
            words.append(out.first_word)
        elif i == out.n_ops - 1:
            words.append(out.last_word)
        else:
            words.append(int(arr[out.first_op + i]))
    return words


def clip_bases(bases: bytes, ops: Optional[Sequence[int]], ref_pos: int, left_len: int, right_len: int):
    """HiFiRead::clip_bases (clip_bases.rs:9-119).  None, or (base_start, base_end, meth_start, meth_end,
    cigar) with cigar = None for a read without one, else (ref_pos, BAM words)."""
    np = _np()
    arr = np.array(list(ops) if ops else [0], dtype=np.uint32)
    out = _BamletClip()
    rc = lib().tro_clip_bases(arr.ctypes.data, len(ops) if ops else 0, ref_pos, bases, len(bases), left_len, right_len,
                              C.byref(out))
    if rc < 0:
        raise ValueError("clip_bases: the reference panics")
    if rc == 0:
        return None
    cigar = (out.ref_pos, _bamlet_words(out, arr)) if out.has_cigar else None
    return out.base_start, out.base_end, out.meth_start, out.meth_end, cigar


def bamlet_clip(bases: bytes, ops: Optional[Sequence[int]], ref_pos: int, span: Tuple[int, int], flank_len: int):
    """the clip BamWriter::write asks for (write_bam.rs:80-92); None = skipped"""
    np = _np()
    arr = np.array(list(ops) if ops else [0], dtype=np.uint32)
    out = _BamletClip()
    rc = lib().tro_bamlet_clip(arr.ctypes.data, len(ops) if ops else 0, ref_pos, bases, len(bases), span[0], span[1],
                               flank_len, C.byref(out))
    if rc < 0:
        raise ValueError("clip_bases: the reference panics")
    if rc == 0:
        return None
    cigar = (out.ref_pos, _bamlet_words(out, arr)) if out.has_cigar else None
    return out.base_start, out.base_end, out.meth_start, out.meth_end, cigar


def encode_seq4(bases: bytes) -> bytes:
    """ASCII -> BAM 4-bit (the inverse of rec.seq().as_bytes(); test helper)"""
    codes = [SEQ4_ALPHABET.index(b) if b in SEQ4_ALPHABET else 15 for b in bases]
    if len(codes) & 1:
        codes.append(0)
    return bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))


def decode_seq4(packed: bytes, start: int, length: int) -> bytes:
    """read.rs:104 for bases [start, start+length)"""
    np = _np()
    src = np.frombuffer(packed + b"\0", dtype=np.uint8)
    out = np.zeros(max(1, length), dtype=np.uint8)
    lib().tro_decode_seq4(src.ctypes.data, start, length, out.ctypes.data)
    return out[:length].tobytes()


# ------------------------------------------------------------------ next row: VCF sample fields --

def vcf_fields(alleles):
    """(AL, MC, MS, AP) of one locus (write_vcf.rs:267-343).  alleles = [(allele_len, motif_counts,
    spans or None, purity)] in genotype order, spans = [(motif_index, start, end)]."""
    np = _np()
    n = len(alleles)
    lens = np.array([a[0] for a in alleles] + [0], dtype=np.uint64)
    mc_off = np.zeros(n + 1, dtype=np.uint64)
    sp_off = np.zeros(n + 1, dtype=np.uint64)
    mc, sp = [], []
    for i, (_, counts, spans, _) in enumerate(alleles):
        mc += list(counts)
        sp += [x for s_ in (spans or []) for x in s_]
        mc_off[i + 1] = len(mc)
        sp_off[i + 1] = len(sp) // 3
    mca = np.array(mc + [0], dtype=np.uint32)
    spa = np.array(sp + [0], dtype=np.uint32)
    pur = np.array([a[3] for a in alleles] + [0.0], dtype=np.float64)
    out = []
    for f in range(4):
        buf = C.create_string_buffer(1 << 16)
        ln = lib().tro_vcf_field(f, n, lens.ctypes.data, mc_off.ctypes.data, mca.ctypes.data, sp_off.ctypes.data,
                                 spa.ctypes.data, pur.ctypes.data, buf, len(buf))
        assert ln <= len(buf)
        out.append(buf.raw[:ln])
    return tuple(out)
