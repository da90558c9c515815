// synth.cpp -- synthetic HiFi workload generator (host, multi-threaded) -> libtrgt_synth.so.
//
// Produces the packed inputs of the hot path for a catalog of tandem-repeat loci, in the shape the
// reference's per-locus worker hands to it (src/trgt/workflows/tr.rs:32-37): per locus the two flank
// pieces (span_locater.rs:38-39) and D clipped reads = 500-bp context + allele + 500-bp context
// (clip radius 2*flank_len, tr.rs:33) with a HiFi-like error model.  Every random decision is drawn
// from a counter-based generator keyed by (seed, locus, read), so any shard of the catalog can be
// regenerated independently and identically.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64 stream keyed by (seed, a, b)
  uint64_t s;
  Rng(uint64_t seed, uint64_t a, uint64_t b) {
    s = seed ^ (a * 0x9E3779B97F4A7C15ull) ^ (b * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull);
    next(); next();
  }
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  uint8_t base() { return "ACGT"[next() >> 62]; }
  double gauss() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
};

}  // namespace

extern "C" {

typedef struct {
  uint64_t seed;
  uint32_t locus_begin;     // global index of the first locus of this shard
  uint32_t n_loci;
  uint32_t depth;           // reads per locus
  uint32_t context;         // bases of flanking context on each side of the allele in a read (500)
  uint32_t piece;           // search flank length (250)
  double sub_rate, ins_rate, del_rate;  // per base
  double unit_indel_rate;   // per motif copy inside the repeat: gain/loss of one whole copy
  double het_frac;
  double tr_len_median;     // repeat tract length distribution: lognormal(median, sigma), in bp
  double tr_len_sigma;
  uint32_t tr_len_min, tr_len_max;
  // optional fixed motif sets (NULL: draw one 2-6 bp motif per locus with the Adotto length mix)
  const uint8_t *motifs; const uint64_t *motif_off; const uint32_t *locus_motif_off;
  uint32_t n_catalog_loci;  // number of loci the fixed motif sets describe (cycled if n_loci is larger)
  uint32_t threads;
  uint32_t motif_mix;       // 0: motif length mix of the genome-wide catalog (57/8/25/7/2 % for 2..6 bp); 1: uniform 2..6 bp
  uint32_t tr_len_dist;     // 0: lognormal(median, sigma) clipped to [min, max]; 1: log-uniform on [min, max]
  uint32_t het_independent; // 1: the second haplotype of a heterozygous locus draws its own tract length (expansions)
} synth_params;

typedef struct {
  // sizes (filled by synth_plan)
  uint64_t n_reads, read_bytes, n_motifs, motif_bytes, allele_bytes, piece_bytes;
  // buffers (caller-allocated after synth_plan, filled by synth_fill)
  uint8_t *reads; uint64_t *read_off;            // [n_reads+1]
  uint32_t *locus_read_off;                       // [n_loci+1]
  uint8_t *read_hap;                              // [n_reads] 0/1
  uint8_t *left; uint64_t *left_off;              // [n_loci+1]
  uint8_t *right; uint64_t *right_off;
  uint8_t *motifs; uint64_t *motif_off;           // [n_motifs+1]
  uint32_t *locus_motif_off;                      // [n_loci+1]
  uint8_t *alleles; uint64_t *allele_off;         // [2*n_loci+1] true haplotype sequences (2 per locus)
} synth_out;

}  // extern "C"

namespace {

struct LocusDesc {
  std::vector<std::vector<uint8_t>> motifs;
  std::vector<uint8_t> allele[2];
  std::vector<uint8_t> left, right;  // context sequences (length = context)
};

void make_locus(const synth_params &p, uint32_t gl, LocusDesc &d) {
  Rng r(p.seed, gl, 0xFFFFFFFFull);
  d.motifs.clear();
  if (p.motifs) {
    const uint32_t cl = gl % p.n_catalog_loci;
    for (uint32_t m = p.locus_motif_off[cl]; m < p.locus_motif_off[cl + 1]; m++)
      d.motifs.emplace_back(p.motifs + p.motif_off[m], p.motifs + p.motif_off[m + 1]);
  } else {
    const double u = r.uni();  // motif length mix of repeats/repeat_catalog.hg38.bed
    const int n = p.motif_mix == 1 ? 2 + (int)(u * 5.0)
                                   : (u < 0.57 ? 2 : (u < 0.65 ? 3 : (u < 0.90 ? 4 : (u < 0.97 ? 5 : 6))));
    std::vector<uint8_t> m(n);
    for (;;) {
      for (int i = 0; i < n; i++) m[i] = r.base();
      bool homo = true;
      for (int i = 1; i < n; i++) homo = homo && m[i] == m[0];
      if (!homo) break;
    }
    d.motifs.push_back(m);
  }
  // repeat tract: copies of the motifs (N in a catalog motif becomes a concrete base)
  auto draw_len = [&]() {
    double len = p.tr_len_dist == 1
                     ? exp(log((double)p.tr_len_min) + r.uni() * (log((double)p.tr_len_max) - log((double)p.tr_len_min)))
                     : p.tr_len_median * exp(p.tr_len_sigma * r.gauss());
    if (len < p.tr_len_min) len = p.tr_len_min;
    if (len > p.tr_len_max) len = p.tr_len_max;
    return len;
  };
  const size_t nm = d.motifs.size();
  auto copies_of = [&](double len) {
    std::vector<uint32_t> copies(nm);
    for (size_t m = 0; m < nm; m++) {
      const double share = len / (double)nm;
      uint32_t c = (uint32_t)(share / (double)d.motifs[m].size() + 0.5);
      copies[m] = c < 2 ? 2 : c;
    }
    return copies;
  };
  const std::vector<uint32_t> copies = copies_of(draw_len());
  const bool het = r.uni() < p.het_frac;
  for (int h = 0; h < 2; h++) {
    std::vector<uint32_t> c = copies;
    if (h == 1 && het && p.het_independent) {
      c = copies_of(draw_len());
    } else if (h == 1 && het) {
      const size_t m = r.below((uint32_t)nm);
      const int k = 1 + (int)r.below(3);
      if (r.uni() < 0.5 && c[m] > (uint32_t)k + 1) c[m] -= k; else c[m] += k;
    }
    d.allele[h].clear();
    Rng rb(p.seed, gl, 0xFFFFFFF0ull);  // same N-resolution on both haplotypes
    for (size_t m = 0; m < nm; m++)
      for (uint32_t k = 0; k < c[m]; k++)
        for (uint8_t b : d.motifs[m]) d.allele[h].push_back(b == 'N' ? rb.base() : b);
  }
  d.left.resize(p.context);
  d.right.resize(p.context);
  for (auto &b : d.left) b = r.base();
  for (auto &b : d.right) b = r.base();
}

// one read; out == nullptr only counts.  Returns length.
size_t make_read(const synth_params &p, uint32_t gl, uint32_t ri, const LocusDesc &d, int hap, uint8_t *out) {
  Rng r(p.seed, gl, ri);
  size_t n = 0;
  auto emit = [&](uint8_t b) { if (out) out[n] = b; n++; };
  auto noisy = [&](const uint8_t *s, size_t len) {
    for (size_t i = 0; i < len; i++) {
      const double u = r.uni();
      if (u < p.sub_rate) {
        uint8_t b;
        do { b = r.base(); } while (b == s[i]);
        emit(b);
      } else if (u < p.sub_rate + p.ins_rate) {
        emit(s[i]);
        emit(r.base());
      } else if (u < p.sub_rate + p.ins_rate + p.del_rate) {
        // deleted
      } else {
        emit(s[i]);
      }
    }
  };
  noisy(d.left.data(), d.left.size());
  // repeat tract with whole-copy gains/losses (stutter), then per-base noise
  {
    const std::vector<uint8_t> &al = d.allele[hap];
    const size_t unit = d.motifs[0].size();
    std::vector<uint8_t> tract;
    tract.reserve(al.size() + 4 * unit);
    for (size_t i = 0; i < al.size(); i += unit) {
      const size_t l = std::min(unit, al.size() - i);
      const double u = r.uni();
      if (u < p.unit_indel_rate * 0.5) continue;                                      // copy lost
      tract.insert(tract.end(), al.begin() + i, al.begin() + i + l);
      if (u > 1.0 - p.unit_indel_rate * 0.5) tract.insert(tract.end(), al.begin() + i, al.begin() + i + l);  // copy gained
    }
    noisy(tract.data(), tract.size());
  }
  noisy(d.right.data(), d.right.size());
  return n;
}

template <class F>
void parallel_loci(const synth_params &p, F f) {
  uint32_t nt = p.threads ? p.threads : std::max(1u, std::thread::hardware_concurrency());
  if (nt > p.n_loci) nt = p.n_loci ? p.n_loci : 1;
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      const uint64_t lo = (uint64_t)p.n_loci * t / nt, hi = (uint64_t)p.n_loci * (t + 1) / nt;
      f((uint32_t)lo, (uint32_t)hi);
    });
  for (auto &x : th) x.join();
}

}  // namespace

extern "C" {

// Pass 1: sizes.  read_len (caller-allocated, n_loci*depth uint32) receives every read's length.
int synth_plan(const synth_params *pp, synth_out *o, uint32_t *read_len, uint32_t *allele_len, uint32_t *n_motifs_per_locus,
               uint32_t *motif_bytes_per_locus) {
  const synth_params &p = *pp;
  if (p.piece > p.context) return -1;
  parallel_loci(p, [&](uint32_t lo, uint32_t hi) {
    LocusDesc d;
    for (uint32_t l = lo; l < hi; l++) {
      make_locus(p, p.locus_begin + l, d);
      const bool het = d.allele[0] != d.allele[1];
      for (uint32_t ri = 0; ri < p.depth; ri++) {
        const int hap = het ? (int)(ri & 1u) : 0;
        read_len[(size_t)l * p.depth + ri] = (uint32_t)make_read(p, p.locus_begin + l, ri, d, hap, nullptr);
      }
      allele_len[2 * (size_t)l] = (uint32_t)d.allele[0].size();
      allele_len[2 * (size_t)l + 1] = (uint32_t)d.allele[1].size();
      n_motifs_per_locus[l] = (uint32_t)d.motifs.size();
      uint32_t mb = 0;
      for (auto &m : d.motifs) mb += (uint32_t)m.size();
      motif_bytes_per_locus[l] = mb;
    }
  });
  o->n_reads = (uint64_t)p.n_loci * p.depth;
  o->read_bytes = 0;
  for (uint64_t i = 0; i < o->n_reads; i++) o->read_bytes += read_len[i];
  o->n_motifs = 0; o->motif_bytes = 0; o->allele_bytes = 0;
  for (uint32_t l = 0; l < p.n_loci; l++) {
    o->n_motifs += n_motifs_per_locus[l];
    o->motif_bytes += motif_bytes_per_locus[l];
    o->allele_bytes += allele_len[2 * (size_t)l] + allele_len[2 * (size_t)l + 1];
  }
  o->piece_bytes = (uint64_t)p.n_loci * p.piece;
  return 0;
}

// Pass 2: fill.  All offset arrays must already hold their prefix sums (the caller builds them from
// the pass-1 lengths); this writes the bytes.
int synth_fill(const synth_params *pp, synth_out *o) {
  const synth_params &p = *pp;
  parallel_loci(p, [&](uint32_t lo, uint32_t hi) {
    LocusDesc d;
    for (uint32_t l = lo; l < hi; l++) {
      make_locus(p, p.locus_begin + l, d);
      const bool het = d.allele[0] != d.allele[1];
      for (uint32_t ri = 0; ri < p.depth; ri++) {
        const uint64_t r = (uint64_t)l * p.depth + ri;
        const int hap = het ? (int)(ri & 1u) : 0;
        make_read(p, p.locus_begin + l, ri, d, hap, o->reads + o->read_off[r]);
        o->read_hap[r] = (uint8_t)hap;
      }
      memcpy(o->left + o->left_off[l], d.left.data() + (p.context - p.piece), p.piece);
      memcpy(o->right + o->right_off[l], d.right.data(), p.piece);
      uint32_t m0 = o->locus_motif_off[l];
      for (size_t m = 0; m < d.motifs.size(); m++)
        memcpy(o->motifs + o->motif_off[m0 + m], d.motifs[m].data(), d.motifs[m].size());
      for (int h = 0; h < 2; h++)
        memcpy(o->alleles + o->allele_off[2 * (size_t)l + h], d.allele[h].data(), d.allele[h].size());
    }
  });
  return 0;
}

}  // extern "C"
