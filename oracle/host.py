"""host.py -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).

Python restatement of the reference's HOST-side logic around the hot path, enough to replay BASELINE
config 1 (`trgt genotype` on example/{reference.fasta,repeat.bed,sample.bam}) end to end and compare the
VCF sample field with docs/tutorial.md:44.  The host stays Rust in the reference and is out of the
product's scope; it is restated here only so that the leaf calls of the hot path
(clip, flank spans, consensus alignment, motif HMM, VCF fields) can be driven by REAL BAM records in
the order and with the inputs the reference's worker gives them.  The leaf calls themselves go through
a `backend` (the CPU oracle, or the CUDA engine behind the C ABI in the GPU tests).

Follows (PacificBiosciences/trgt v3.0.0):
  BAM record fields        src/trgt/reads/read.rs:104-150 (from_hts_rec), :166-187 (rq / HP tags)
  mismatch offsets         src/trgt/reads/snp.rs:51-80
  Locus                    src/trgt/locus.rs:26-74, :168-190 (get_tr_and_flanks); src/utils/region.rs:23-39
  analyze                  src/trgt/workflows/tr.rs:24-109
  get_spanning_reads       src/trgt/workflows/tr.rs:111-165, uniform_downsample :167-179
  clip_reads               src/trgt/workflows/tr.rs:186-196
  extract_reads            src/trgt/workflows/tr.rs:268-361 (without the reservoir branch: max_depth*3 reads never reached)
  genotype_size            src/trgt/genotype/genotype_size.rs:6-125
  diploid / haploid        src/trgt/genotype/diploid.rs:5-103, haploid.rs
  get_consensus            src/trgt/genotype/consensus.rs:113-165
  genotype_flank           src/trgt/genotype/genotype_flank.rs:9-298
  median                   src/utils/math.rs:72-98
  VCF sample field         src/trgt/writers/write_vcf.rs:171-397
"""
from __future__ import annotations

import gzip
import math
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

BAM_OPS = "MIDNSHP=X"


# ------------------------------------------------------------------ BAM / FASTA / BED ------------

@dataclass
class BamRecord:
    qname: str
    flag: int
    ref_id: int
    pos: int                 # rec.reference_start(), 0-based
    mapq: int
    cigar: List[int]         # BAM words (len << 4) | op
    seq4: bytes              # bam_get_seq: two bases per byte, first base in the high nibble
    l_seq: int
    qual: bytes
    tags: Dict[str, Tuple[str, object]] = field(default_factory=dict)

    @property
    def reference_end(self) -> int:
        return self.pos + sum(w >> 4 for w in self.cigar if BAM_OPS[w & 15] in "MDN=X")


def _parse_tags(aux: bytes) -> Dict[str, Tuple[str, object]]:
    tags, o = {}, 0
    sizes = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    fmts = {"A": "<c", "c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    while o + 3 <= len(aux):
        tag, ty = aux[o:o + 2].decode(), chr(aux[o + 2])
        o += 3
        if ty in sizes:
            tags[tag] = (ty, struct.unpack_from(fmts[ty], aux, o)[0])
            o += sizes[ty]
        elif ty in "ZH":
            end = aux.index(b"\0", o)
            tags[tag] = (ty, aux[o:end])
            o = end + 1
        elif ty == "B":
            sub = chr(aux[o])
            n, = struct.unpack_from("<i", aux, o + 1)
            o += 5
            tags[tag] = ("B" + sub, struct.unpack_from("<%d%s" % (n, fmts[sub][1]), aux, o))
            o += n * sizes[sub]
        else:
            raise ValueError(f"unknown aux type {ty!r}")
    return tags


def parse_bam(path: str):
    """-> (reference names/lengths, records).  BGZF is a series of gzip members, which gzip reads as one stream."""
    with gzip.open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file")
    l_text, = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, o)
        name = data[o + 4:o + 4 + l_name - 1].decode()
        l_ref, = struct.unpack_from("<i", data, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(data):
        bs, = struct.unpack_from("<i", data, o)
        o += 4
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nref, _npos, _tlen = struct.unpack_from("<iiBBHHHiiii", data, o)
        p = o + 32
        qname = data[p:p + l_rn - 1].decode()
        p += l_rn
        cigar = list(struct.unpack_from("<%dI" % n_cig, data, p))
        p += 4 * n_cig
        seq4 = data[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        qual = data[p:p + l_seq]
        p += l_seq
        recs.append(BamRecord(qname, flag, ref_id, pos, mapq, cigar, seq4, l_seq, qual, _parse_tags(data[p:o + bs])))
        o += bs
    return refs, recs


def read_fasta(path: str) -> Dict[str, bytes]:
    out, name, parts = {}, None, []
    with open(path, "rb") as f:
        for line in f:
            line = line.strip()
            if line.startswith(b">"):
                if name is not None:
                    out[name] = b"".join(parts)
                name, parts = line[1:].split()[0].decode(), []
            elif line:
                parts.append(line)
    if name is not None:
        out[name] = b"".join(parts)
    return out


@dataclass
class Locus:
    id: str
    contig: str
    start: int
    end: int
    left_flank: bytes
    tr: bytes
    right_flank: bytes
    motifs: List[bytes]
    struc: str
    ploidy: int = 2
    genotyper: str = "size"

    @classmethod
    def from_bed_line(cls, genome: Dict[str, bytes], line: str, flank_len: int = 250) -> "Locus":
        chrom, start, end, info = line.split()
        start, end = int(start), int(end)
        fields = dict(kv.split("=", 1) for kv in info.split(";"))
        g = genome[chrom]
        # faidx fetch_seq_string(contig, a, b) is 0-based inclusive: locus.rs:185-187
        fetch = lambda a, b: g[a:b + 1].upper()
        return cls(fields["ID"], chrom, start, end, fetch(start - flank_len, start - 1), fetch(start, end - 1),
                   fetch(end, end + flank_len - 1), [m.encode() for m in fields["MOTIFS"].split(",")], fields["STRUC"])


@dataclass
class Params:
    min_flank_id_frac: float = 0.7
    min_read_qual: float = 0.98
    search_flank_len: int = 250
    max_depth: int = 250
    scoring: Tuple[int, int, int] = (2, 5, 1)


@dataclass
class HiFiRead:
    rec: BamRecord
    mismatch_offsets: List[int]
    start_offset: int
    end_offset: int
    hp_tag: Optional[int]
    bases: bytes = b""               # clipped bases (after clip_reads)
    clip: Optional[tuple] = None     # (clipped_ref_start, query_start, query_end, clipped cigar words)


def extract_snps_offset(cigar: Sequence[int], ref_pos: int, start: int, end: int) -> List[int]:
    """snp.rs:51-80: offsets of the X bases outside [start, end], relative to the region's start / end"""
    out, r = [], ref_pos
    for w in cigar:
        n, op = w >> 4, BAM_OPS[w & 15]
        if op == "X" and not (start <= r <= end):
            diff = r - start if r < start else r - end
            out.extend(diff + i for i in range(n))
            r += n
        elif op in "MX=DN":
            r += n
    return out


def get_rq_tag(rec: BamRecord) -> Optional[float]:
    t = rec.tags.get("rq")
    return float(t[1]) if t and t[0] == "f" else None


def extract_reads(locus: Locus, ref_names: Sequence[str], records: Sequence[BamRecord], params: Params) -> List[HiFiRead]:
    lo, hi = max(0, locus.start - params.search_flank_len), locus.end + params.search_flank_len
    out = []
    for rec in records:
        if rec.ref_id < 0 or ref_names[rec.ref_id] != locus.contig or rec.flag & 4:
            continue
        if not (rec.pos < hi and rec.reference_end > lo):   # bam.fetch: records overlapping the region
            continue
        if rec.flag & (0x800 | 0x100):
            continue
        rq = get_rq_tag(rec)
        if (1.0 if rq is None else rq) < params.min_read_qual:
            continue
        hp = rec.tags.get("HP")
        out.append(HiFiRead(rec, extract_snps_offset(rec.cigar, rec.pos, locus.start, locus.end),
                            rec.pos - locus.start, rec.reference_end - locus.end,
                            int(hp[1]) if hp and hp[0] == "C" else None))
        if len(out) >= params.max_depth * 3:
            break
    return out


# ------------------------------------------------------------------ genotypers --------------------

def _diploid_penalty(gt, sizes, counts) -> float:
    short, long_ = gt
    max_frac = 0.25 if abs(short - long_) <= 100 else 0.05
    pen = 0.0
    for size, count in zip(sizes, counts):
        st = 10 + 2 * abs(short - size) if size != short else 0
        lt = 10 + 2 * abs(long_ - size) if size != long_ else 0
        pen += (float(min(st, lt)) + max_frac * float(max(st, lt))) * float(count)
    return pen


def diploid_genotype(sizes: List[int], counts: List[int]):
    """diploid.rs:5-103 -> [(size, (ci_lo, ci_hi)), (size, ci)]"""
    cands = []
    for i in range(len(sizes)):
        for j in range(i, len(sizes)):
            gt = (sizes[i], sizes[j])
            cands.append((gt, _diploid_penalty(gt, sizes, counts)))
    cands.sort(key=lambda c: c[1])  # stable, as sort_by
    top = cands[0][0]
    short, long_ = min(top), max(top)
    if short != long_ and len(sizes) >= 2:
        coverage = sum(counts)
        hist = sorted(zip(sizes, counts), key=lambda sc: -sc[1])  # stable: sorted_by(|a, b| b.1.cmp(a.1))
        top_frac = hist[0][1] / coverage
        if top_frac > 0.60 and max(sizes) - min(sizes) <= 6:
            short = long_ = hist[0][0]
    s_ci, l_ci = (short, short), (long_, long_)
    for size in sizes:
        if abs(size - short) <= abs(size - long_):
            s_ci = (min(s_ci[0], size), max(s_ci[1], size))
        else:
            l_ci = (min(l_ci[0], size), max(l_ci[1], size))
    return [(short, s_ci), (long_, l_ci)]


def haploid_genotype(sizes: List[int], counts: List[int]):
    """haploid.rs:3-30: the length with the smallest penalty (stable sort: the first one on ties)"""
    def pen(allele):
        return sum((10.0 + 2.0 * abs(allele - s) if s != allele else 0.0) * float(c) for s, c in zip(sizes, counts))
    best = min(range(len(sizes)), key=lambda i: (pen(sizes[i]), i))
    return [(sizes[best], (min(sizes), max(sizes)))]


def _hist(sorted_items):
    uniq, cnt = [], []
    for x in sorted_items:
        if uniq and uniq[-1] == x:
            cnt[-1] += 1
        else:
            uniq.append(x)
            cnt.append(1)
    return uniq, cnt


def get_consensus(sizes: List[int], seqs: List[bytes], counts: List[int]) -> List[bytes]:
    """consensus.rs:113-165; max_by_key returns the LAST maximum"""
    def closest(allele):
        best = None
        for s in seqs:
            if best is None or abs(best - allele) > abs(len(s) - allele):
                best = len(s)
        return best

    def most_frequent(length):
        best = None
        for s, c in zip(seqs, counts):
            if len(s) == length and (best is None or c >= best[1]):
                best = (s, c)
        return best[0]

    out = [most_frequent(closest(sizes[0]))]
    if len(sizes) != 1 and sizes[0] != sizes[1]:
        out.append(most_frequent(closest(sizes[1])))
    return out


def genotype_size(ploidy: int, trs: List[bytes], backend):
    """genotype_size.rs:6-75 -> (gt, alleles, classifications)"""
    sizes, counts = _hist(sorted(len(s) for s in trs))
    gt = haploid_genotype(sizes, counts) if ploidy == 1 else diploid_genotype(sizes, counts)
    allele_lens = [a[0] for a in gt]
    useqs, ucounts = _hist(sorted(trs))
    alleles = get_consensus(allele_lens, useqs, ucounts)
    if len(allele_lens) == 1:
        split = [(useqs, ucounts)]
    else:
        al1, al2 = allele_lens
        s1 = [s for s in useqs if abs(len(s) - al1) <= abs(len(s) - al2)]
        s2 = [s for s in useqs if abs(len(s) - al2) < abs(len(s) - al1)]
        split = [(s1, [c for s, c in zip(useqs, ucounts) if s in s1]), (s2, [c for s, c in zip(useqs, ucounts) if s in s2])]
    fixed = []
    for index, allele in enumerate(alleles):
        seqs, cnts = split[index]
        coverage = sum(cnts)
        ref_count = next((c for s, c in zip(seqs, cnts) if s == allele), 0)
        fixed.append(allele if 2 * ref_count >= coverage else backend.align_repair(allele, seqs))
    alleles = fixed
    if ploidy == 2 and len(alleles) == 1:
        alleles.append(alleles[0])
    cls, tie = [0] * len(trs), 1
    for i, s in enumerate(trs):
        if len(alleles) == 2:
            d1, d2 = abs(len(s) - len(alleles[0])), abs(len(s) - len(alleles[1]))
            if d1 < d2:
                cls[i] = 0
            elif d1 > d2:
                cls[i] = 1
            else:
                tie = (tie + 1) % 2
                cls[i] = tie
    return gt, alleles, cls


def median(data: List[int]) -> Optional[float]:
    """math.rs:72-98 (value-wise: the selection algorithm does not change the result)"""
    if not data:
        return None
    d = sorted(data)
    n = len(d)
    return float(d[n // 2]) if n % 2 else (d[n // 2 - 1] + d[n // 2]) / 2.0


def _ln_sum_exp(a: float, b: float) -> float:
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def genotype_flank(reads: List[HiFiRead], trs: List[bytes], backend):
    """genotype_flank.rs:9-44 -> None or (gt, alleles, assignment)"""
    n = len(reads)
    got = None
    # get_trs_with_hp :46-83
    assign, by, tie, unassigned = [], [[], []], 1, 0
    for read, tr in zip(reads, trs):
        if read.hp_tag == 1:
            assign.append(0); by[0].append(tr)
        elif read.hp_tag == 2:
            assign.append(1); by[1].append(tr)
        else:
            tie = (tie + 1) % 2
            assign.append(tie); by[tie].append(tr)
            unassigned += 1
    if by[0] and by[1] and (n - unassigned) / n >= 0.7:
        got = (by, assign)
    if got is None:  # get_trs_with_clustering :85-137
        if not trs:
            return None
        skip = int(math.floor(n * (1.0 - 0.85) + 0.5))  # get_analysis_region :217-238 (f64::round: half away from zero)
        region = (sorted(r.start_offset for r in reads)[n - 1 - skip], sorted(r.end_offset for r in reads)[skip])
        cnt: Dict[int, int] = {}
        for r in reads:
            for off in r.mismatch_offsets:
                if region[0] <= off <= region[1]:
                    cnt[off] = cnt.get(off, 0) + 1
        snvs = sorted(off for off, c in cnt.items() if c / float(n) >= 0.20)   # call_snvs :283-298
        profiles = []
        for r in reads:                                                     # get_profiles :262-281
            mm = set(r.mismatch_offsets)
            profiles.append(tuple(None if (s < r.start_offset or s > r.end_offset) else (s in mm) for s in snvs))
        full = sorted(p for p in profiles if all(v is not None for v in p))  # get_candidate_gts :240-260
        if len(full) / len(profiles) < 0.40:
            return None
        haps = []
        for p in full:
            if not haps or haps[-1] != p:
                haps.append(p)
        cands = [(h1, h2) for i, h1 in enumerate(haps) for h2 in haps[i:]]
        if len(cands) <= 1:
            return None

        def eval_hap(profile, hap):
            return sum(math.log(0.9) if p == h else math.log(1.0 - 0.9) for p, h in zip(profile, hap) if p is not None)

        def loglik(gt):
            return sum(_ln_sum_exp(eval_hap(p, gt[0]), eval_hap(p, gt[1])) - math.log(2.0) for p in profiles)

        best, best_ll = None, None
        for gt in cands:                  # max_by: the LAST maximum
            ll = loglik(gt)
            if best_ll is None or ll >= best_ll:
                best, best_ll = gt, ll
        if best[0] == best[1]:
            return None
        assign, by, tie = [], [[], []], 1
        for p, tr in zip(profiles, trs):
            d1 = sum(1 for a, b in zip(p, best[0]) if a is not None and a == b)
            d2 = sum(1 for a, b in zip(p, best[1]) if a is not None and a == b)
            if d1 < d2:
                assign.append(0); by[0].append(tr)
            elif d1 > d2:
                assign.append(1); by[1].append(tr)
            else:
                tie = (tie + 1) % 2
                assign.append(tie); by[0].append(tr); by[1].append(tr)
        got = (by, assign)
    by, assign = got
    gt, alleles = [], []
    for group in by:
        med = median([len(s) for s in group])
        if med is None:
            return None
        med = int(med)  # `as usize` truncates
        counts: Dict[bytes, int] = {}
        for s in group:
            counts[s] = counts.get(s, 0) + 1
        top = max(counts.values())
        best = None
        for s in sorted(counts):          # BTreeMap order; min_by_key returns the FIRST minimum
            if counts[s] == top and (best is None or abs(len(s) - med) < best[1]):
                best = (s, abs(len(s) - med))
        backbone, freq = best[0], top / float(len(group))
        allele = backend.align_repair(backbone, group) if freq < 0.5 else backbone
        gt.append((len(allele), (min(len(s) for s in group), max(len(s) for s in group))))
        alleles.append(allele)
    if len(alleles[0]) > len(alleles[1]):
        gt.reverse(); alleles.reverse()
        assign = [(a + 1) % 2 for a in assign]
    return gt, alleles, assign


def filter_impure_trs(purities: Sequence[float], read_quals: Sequence[Optional[float]], rq_cutoff: float = 0.9,
                      purity_cutoff: float = 0.9) -> List[int]:
    """filter_impure_trs (tr.rs:400-452), the part after the HMM: purities[i] is calc_purity of read i's repeat
    sequence (only looked at for reads with rq < rq_cutoff or without rq; the others count as 1.0).  Returns the
    indices of the reads that are kept, in the order the reference returns them (sorted by purity, total_cmp,
    stable); at most max(1, round(0.1 n)) reads below the cutoff are dropped, the least pure first."""
    n = len(purities)
    max_filter = max(1, int(math.floor(0.1 * n + 0.5)))
    eff = [1.0 if (rq is not None and rq >= rq_cutoff) else p for p, rq in zip(purities, read_quals)]

    def total_key(x: float):  # f64::total_cmp: -NaN < -inf < ... < +inf < +NaN
        bits = struct.unpack("<q", struct.pack("<d", x))[0]
        return bits ^ (((bits >> 63) & 0xFFFFFFFFFFFFFFFF) >> 1) if bits < 0 else bits

    order = sorted(range(n), key=lambda i: total_key(eff[i]))
    kept, filtered = [], 0
    for i in order:
        if eff[i] >= purity_cutoff or filtered >= max_filter:
            kept.append(i)
        else:
            filtered += 1
    return kept


# ------------------------------------------------------------------ the per-locus worker ----------

@dataclass
class LocusResult:
    alleles: List[bytes]
    gt: list
    classification: List[int]
    annotations: list            # per allele (motif_counts, labels or None, purity)
    n_spanning: int
    sample_field: str            # GT:AL:ALLR:SD:MC:MS:AP:AM
    vcf_ref_alt: Tuple[bytes, List[bytes]]
    reads: List[HiFiRead] = field(default_factory=list)   # the spanning reads, in genotyping order
    spans: List[Tuple[int, int]] = field(default_factory=list)


def analyze(locus: Locus, reads: List[HiFiRead], backend, params: Params = Params()) -> Optional[LocusResult]:
    """tr.rs:24-109 for ploidy 2 / 1 and the size genotyper, from extracted reads to the VCF sample field"""
    radius = 2 * params.search_flank_len
    region = (locus.start - radius, locus.end + radius)
    clips = backend.clip_reads(reads, region)                                   # tr.rs:186-196
    kept = []
    for r, c in zip(reads, clips):
        if c is not None:
            r.clip = c
            kept.append(r)
    reads = kept
    bases = backend.clipped_bases(reads)
    for r, b in zip(reads, bases):
        r.bases = b
    P = params.search_flank_len
    spans = backend.find_tr_spans(locus.left_flank[len(locus.left_flank) - P:], locus.right_flank[:P], reads, params)
    rs = [(r, s) for r, s in zip(reads, spans) if s is not None]
    rs = [(r, s) for r, s in rs if s[0] >= P and len(r.bases) - s[1] >= P]     # tr.rs:138-144
    if not rs:
        return None
    rs.sort(key=lambda x: x[1][1] - x[1][0])                                     # stable
    if len(rs) > params.max_depth:                                               # uniform_downsample
        fast, step = 0.0, len(rs) / float(params.max_depth)
        for i in range(params.max_depth):
            ind = int(math.floor(fast))
            if ind != i:
                rs[i], rs[ind] = rs[ind], rs[i]
            fast += step
        del rs[params.max_depth:]
    reads, spans = [r for r, _ in rs], [s for _, s in rs]
    trs = [r.bases[s[0]:s[1]] for r, s in rs]
    gt, alleles, cls = genotype_size(locus.ploidy, trs, backend)
    if len(gt) == 2 and abs(gt[0][0] - gt[1][0]) <= 10:
        snp = genotype_flank(reads, trs, backend)
        if snp is not None:
            gt, alleles, cls = snp
    ann = backend.label_with_hmm(locus.motifs, alleles)
    by_hap = [sum(1 for c in cls if c == 0), sum(1 for c in cls if c == 1)]
    geno = [dict(seq=alleles[i], ann=ann[i], ci=gt[i][1], n=by_hap[i]) for i in range(len(gt))]
    if len(geno) != 1 and geno[0]["seq"] != locus.tr and geno[1]["seq"] == locus.tr:
        geno.reverse()
        cls = [1 - c for c in cls]
    # write_vcf.rs:219-259 (set_gt) and :267-397
    seqs, idx = [locus.tr], []
    for a in geno:
        if a["seq"] == locus.tr:
            idx.append(0)
        elif len(seqs) == 1:
            idx.append(1); seqs.append(a["seq"])
        elif geno[0]["seq"] == geno[1]["seq"]:
            idx.append(1)
        else:
            idx.append(2); seqs.append(a["seq"])
    pad = locus.left_flank[-1:]
    al, mc, ms, ap = backend.vcf_fields(locus.motifs, [a["seq"] for a in geno], [a["ann"] for a in geno])
    sample = ":".join(["/".join(str(i) for i in idx), al.decode(), ",".join(f"{a['ci'][0]}-{a['ci'][1]}" for a in geno),
                       ",".join(str(a["n"]) for a in geno), mc.decode(), ms.decode(), ap.decode(),
                       ",".join("." for _ in geno)])  # AM: no MM/ML tags in the replayed reads
    return LocusResult([a["seq"] for a in geno], gt, cls, [a["ann"] for a in geno], len(reads), sample,
                       (pad + seqs[0], [pad + s for s in seqs[1:]]), reads, spans)


class OracleBackend:
    """The leaf calls of the hot path on the CPU oracle."""

    def __init__(self, orc):
        self.orc = orc

    def clip_reads(self, reads, region):
        return [self.orc.clip_cigar(r.rec.cigar, r.rec.pos, region) for r in reads]

    def clipped_bases(self, reads):
        return [self.orc.decode_seq4(r.rec.seq4, r.clip[1], r.clip[2] - r.clip[1]) for r in reads]

    def find_tr_spans(self, lf, rf, reads, params):
        return self.orc.find_tr_spans(lf, rf, [r.bases for r in reads], search_flank_len=len(lf),
                                      min_flank_id_frac=params.min_flank_id_frac, scoring=params.scoring)

    def align_repair(self, backbone, seqs):
        return self.orc.repair_consensus(backbone, list(seqs))

    def label_with_hmm(self, motifs, alleles):
        h = self.orc.Hmm(motifs)
        out = []
        for a in alleles:
            mc, sp, pur = h.annotate(a)
            out.append((mc, sp or None, pur))
        return out

    def vcf_fields(self, motifs, seqs, anns):
        return self.orc.vcf_fields([(len(s), a[0], a[1], a[2]) for s, a in zip(seqs, anns)])
