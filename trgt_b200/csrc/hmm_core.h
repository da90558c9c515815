// hmm_core.h -- motif HMM of `trgt genotype` restructured for one cooperating group per allele.
//
// Replaces (reference, PacificBiosciences/trgt v3.0.0):
//   build_hmm / define_motif_block      src/hmm/builder.rs:4-184
//   Hmm::label (Viterbi + traceback)    src/hmm/hmm_model.rs:54-156, order_states :206-240
//   calc_purity / get_events            src/hmm/purity.rs:6-41, src/hmm/events.rs:17-117
//   remove_imperfect_motifs             src/hmm/operations.rs:6-80
//   Hmm::label_motifs                   src/hmm/hmm_model.rs:158-200
//   skip filter, count_motifs, collapse src/trgt/workflows/tr.rs:471-476, src/hmm/utils.rs:3-27
//   replace_invalid_bases               src/hmm/utils.rs:29-42
//
// Design (not the reference's): the model is never materialised as edge lists.  A state's role
// (ms / match_i / ins_i / del_i / me / skip ...) is decoded from its index and the motif block
// table, and its in-edges are enumerated by arithmetic in the reference's list order, which is
// the tie-break order of the strict '>' argmax.  Only ln() constants computed by the host libm
// are shipped (HmmConsts + the jump-in table), so every score is the same left-to-right f64 sum
// `prev + ln(trans) + ln(emit)` the reference forms.  A column is evaluated in three barriers:
// (1) all emitting states from the previous column, (2) per motif block the silent chain
// del_0..del_{n-2}, me, (3) re, rs, ms.  That is a topological order of the silent sub-graph,
// so it yields the same values as the reference's Kahn layering.  Only two score columns live
// on chip; the per-column back-pointer (in-edge index, one byte per state) goes to HBM, S
// contiguous bytes per column.  Post-processing (purity, imperfect-copy removal, span
// labelling, skip filtering, counting, collapsing) is a single reverse walk over the
// back-pointers that never materialises the state path.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "coop.h"

namespace trgt {

struct HmmConsts {
  double lp_match;       // ln(0.90)            match->match, ms->match_0, match->me
  double lp_ins_exit;    // ln(1.0 - 0.25)      ins->match, ins->me
  double lp_half;        // ln(0.50)            del->match, del->del, me->ms, me->re, skip->skip, skip->me
  double lp_ins_loop;    // ln(0.25)            ins->ins
  double lp_indel_open;  // ln((1.0 - 0.90)/2)  match->ins, match->del
  double lp_end;         // ln(0.10)            re->end
  double lp_one;         // ln(1.00)            start->rs, re->rs, rs->ms, ms->skip, del->me
  double em_hi;          // ln(0.90)
  double em_lo;          // ln(0.03)
  double em_quarter;     // ln(0.25)
  double em_one;         // ln(1.00)
};

struct HmmSpan {
  uint32_t motif_index, start, end;
};

// replace_invalid_bases(seq, ATCG) for one base, utils.rs:29-42
TRGT_HD uint8_t hmm_clean_base(uint8_t b, uint32_t index) {
  if (b == 'A' || b == 'T' || b == 'C' || b == 'G') return b;
  const uint32_t r = index & 3u;
  return r == 0 ? 'A' : (r == 1 ? 'T' : (r == 2 ? 'C' : 'G'));
}
// replace_invalid_bases(motif, ATCGN)
TRGT_HD uint8_t hmm_clean_motif_base(uint8_t b, uint32_t index) {
  if (b == 'A' || b == 'T' || b == 'C' || b == 'G' || b == 'N') return b;
  const uint32_t r = index % 5u;
  return r == 0 ? 'A' : (r == 1 ? 'T' : (r == 2 ? 'C' : (r == 3 ? 'G' : 'N')));
}
// encode_base, hmm_model.rs:243-252 ('#' = 0 is produced by the column loop itself)
TRGT_HD int hmm_symbol(uint8_t b) { return b == 'A' ? 1 : (b == 'T' ? 2 : (b == 'C' ? 3 : 4)); }

// Block table of one locus' model.  Block b < nb-1 is motif b, block nb-1 is the skip block.
struct HmmModel {
  int S;                      // 7 + sum(3 n + 1)                      builder.rs:5-6
  int nb;                     // motifs + 1
  const uint8_t *motif_bytes; // sanitised motif bytes, concatenated
  const uint32_t *blk_moff;   // [nb] offset of the motif's bytes
  const uint32_t *blk_mmoff;  // [nb] offset of the motif's jump-in ln table (indexed by match index)
  const uint16_t *blk_n;      // [nb] motif length (skip block: 0)
  const uint16_t *blk_ms;     // [nb] state index of the block's ms
  const uint16_t *st_blk;     // [S] block of each state (0xFFFF: start, rs, re, end)
  // accessors shared with HmmModelScan (hmm_annotate is written against these)
  TRGT_HD int block_of(int st) const { return st_blk[st]; }
  TRGT_HD int block_ms(int b) const { return blk_ms[b]; }
  TRGT_HD int block_n(int b) const { return blk_n[b]; }
  TRGT_HD uint8_t motif_byte(int b, int i) const { return motif_bytes[blk_moff[b] + i]; }
};

// The same model addressed straight from the packed motif set, without any table: block geometry is
// recomputed by scanning the locus' motif offsets.  For one thread per allele (back-pointer walks),
// where per-thread tables would not fit on chip and loci have one or two motifs anyway.
struct HmmModelScan {
  int S, nb;
  const uint8_t *motifs;    // raw motif bytes of the whole batch
  const uint64_t *moff;     // offsets of this locus' motifs: moff[0..nb-1]
  TRGT_HD int block_n(int b) const { return b == nb - 1 ? 0 : (int)(moff[b + 1] - moff[b]); }
  TRGT_HD int block_ms(int b) const {
    int ms = 2;
    for (int i = 0; i < b; i++) ms += 3 * (int)(moff[i + 1] - moff[i]) + 1;
    return ms;
  }
  TRGT_HD int block_of(int st) const {
    int ms = 2;
    for (int b = 0; b < nb - 1; b++) {
      const int sz = 3 * (int)(moff[b + 1] - moff[b]) + 1;
      if (st < ms + sz) return b;
      ms += sz;
    }
    return nb - 1;
  }
  TRGT_HD uint8_t motif_byte(int b, int i) const { return hmm_clean_motif_base(motifs[moff[b] + i], (uint32_t)i); }
  TRGT_HD uint32_t jump_off(int b, const uint32_t *mm_off) const { return mm_off[block_n(b)]; }
};

TRGT_HD HmmModelScan hmm_model_scan(const uint8_t *motifs, const uint64_t *moff, int nm) {
  HmmModelScan m;
  m.motifs = motifs; m.moff = moff; m.nb = nm + 1;
  int S = 7;
  for (int b = 0; b < nm; b++) S += 3 * (int)(moff[b + 1] - moff[b]) + 1;
  m.S = S;
  return m;
}

#define HMM_THREAD_S 32  // models up to this many states run one allele per lane (k_hmm_viterbi_thread)

// The same again with the whole model in two registers, for the models one lane handles on its own
// (S <= HMM_THREAD_S = 32, i.e. at most 8 motif bases in at most 6 blocks): block lengths as nibbles, the
// sanitised motif bytes concatenated.  Nothing of the model is re-read from memory inside the DP loops.
struct HmmModelPacked {
  int S, nb;
  uint64_t lens;   // nibble b = length of motif block b
  uint64_t offs;   // nibble b = offset of block b's first byte in `bytes`
  uint64_t bytes;  // sanitised motif bytes
  uint64_t joffs;  // byte b = where block b's jump-in ln table starts (mm_off[length])
  TRGT_HD uint32_t jump_off(int b, const uint32_t *) const { return (uint32_t)(joffs >> (8 * b)) & 255u; }
  TRGT_HD int block_n(int b) const { return b == nb - 1 ? 0 : (int)((lens >> (4 * b)) & 15u); }
  TRGT_HD int block_ms(int b) const {
    int ms = 2;
    for (int i = 0; i < b; i++) ms += 3 * (int)((lens >> (4 * i)) & 15u) + 1;
    return ms;
  }
  TRGT_HD int block_of(int st) const {
    int ms = 2;
    for (int b = 0; b < nb - 1; b++) {
      const int sz = 3 * (int)((lens >> (4 * b)) & 15u) + 1;
      if (st < ms + sz) return b;
      ms += sz;
    }
    return nb - 1;
  }
  TRGT_HD uint8_t motif_byte(int b, int i) const {
    return (uint8_t)(bytes >> (8 * ((int)((offs >> (4 * b)) & 15u) + i)));
  }
};

// only for models with S <= 32 (checked by the caller)
TRGT_HD HmmModelPacked hmm_model_pack(const HmmModelScan &m, const uint32_t *mm_off = nullptr) {
  HmmModelPacked p;
  p.S = m.S; p.nb = m.nb; p.lens = 0; p.offs = 0; p.bytes = 0; p.joffs = 0;
  int o = 0;
  for (int b = 0; b < m.nb - 1; b++) {
    const int n = m.block_n(b);
    p.lens |= (uint64_t)n << (4 * b);
    p.offs |= (uint64_t)o << (4 * b);
    if (mm_off) p.joffs |= (uint64_t)(mm_off[n] & 255u) << (8 * b);  // n <= 8: the offset is below 37
    for (int i = 0; i < n; i++) p.bytes |= (uint64_t)m.motif_byte(b, i) << (8 * (o + i));
    o += n;
  }
  return p;
}

enum HmmRoleKind {
  HR_START = 0, HR_RS, HR_RE, HR_END, HR_MS, HR_MATCH, HR_INS, HR_DEL, HR_ME, HR_SKIP_MS, HR_SKIP, HR_SKIP_ME
};

struct HmmRole {
  int kind, b, i, n, ms;
};

template <class M>
TRGT_HD HmmRole hmm_role(const M &m, int st) {
  HmmRole r;
  r.b = -1; r.i = 0; r.n = 0; r.ms = 0;
  if (st == 0) { r.kind = HR_START; return r; }
  if (st == 1) { r.kind = HR_RS; return r; }
  if (st == m.S - 2) { r.kind = HR_RE; return r; }
  if (st == m.S - 1) { r.kind = HR_END; return r; }
  const int b = m.block_of(st);
  r.b = b;
  r.ms = m.block_ms(b);
  const int off = st - r.ms;
  if (b == m.nb - 1) {
    r.kind = off == 0 ? HR_SKIP_MS : (off == 1 ? HR_SKIP : HR_SKIP_ME);
    return r;
  }
  const int n = m.block_n(b);
  r.n = n;
  if (off == 0) { r.kind = HR_MS; }
  else if (off <= n) { r.kind = HR_MATCH; r.i = off - 1; }
  else if (off <= 2 * n) { r.kind = HR_INS; r.i = off - n - 1; }
  else if (off < 3 * n) { r.kind = HR_DEL; r.i = off - 2 * n - 1; }
  else { r.kind = HR_ME; }
  return r;
}

// Fill the block table for one locus.  `motifs`/`moff` are the raw motif bytes and offsets of
// this locus' nm motifs (moff[0..nm], absolute offsets into `motifs`).  mm_off[n] locates the
// jump-in table of motif length n.  Out arrays are caller-provided (shared memory on the device).
// Returns S, or -1 if a motif is empty.
template <class G>
TRGT_HD int hmm_model_build(const G &g, const uint8_t *motifs, const uint64_t *moff, int nm,
                            const uint32_t *mm_off, uint8_t *o_bytes, uint32_t *o_moff,
                            uint32_t *o_mmoff, uint16_t *o_n, uint16_t *o_ms, uint16_t *o_stblk,
                            HmmModel *model) {
  const int nb = nm + 1;
  // block offsets: a short serial prefix sum (nm is 1 for most loci, <= 10 in shipped catalogs)
  int S = 0;
  int bad = 0;
  if (g.lane() == 0) {
    int ms = 2;
    uint32_t bytes = 0;
    for (int b = 0; b < nm; b++) {
      const int n = (int)(moff[b + 1] - moff[b]);
      if (n <= 0) bad = 1;
      o_ms[b] = (uint16_t)ms;
      o_n[b] = (uint16_t)n;
      o_moff[b] = bytes;
      o_mmoff[b] = n > 0 ? mm_off[n] : 0;
      bytes += (uint32_t)(n > 0 ? n : 0);
      ms += 3 * n + 1;
    }
    o_ms[nm] = (uint16_t)ms;
    o_n[nm] = 0;
    o_moff[nm] = bytes;
    o_mmoff[nm] = 0;
    S = ms + 3 + 2;  // skip block (3) + re + end
  }
  S = g.bcast0(S);
  bad = g.bcast0(bad);
  g.sync();
  if (bad) return -1;
  for (int b = g.lane(); b < nb; b += g.size()) {
    const int n = o_n[b];
    const int ms = o_ms[b];
    const int cnt = (b == nm) ? 3 : 3 * n + 1;
    for (int j = 0; j < cnt; j++) o_stblk[ms + j] = (uint16_t)b;
    if (b < nm) {
      const uint8_t *src = motifs + moff[b];
      for (int j = 0; j < n; j++) o_bytes[o_moff[b] + j] = hmm_clean_motif_base(src[j], (uint32_t)j);
    }
  }
  if (g.lane() == 0) {
    o_stblk[0] = 0xFFFF; o_stblk[1] = 0xFFFF; o_stblk[S - 2] = 0xFFFF; o_stblk[S - 1] = 0xFFFF;
  }
  g.sync();
  model->S = S;
  model->nb = nb;
  model->motif_bytes = o_bytes;
  model->blk_moff = o_moff;
  model->blk_mmoff = o_mmoff;
  model->blk_n = o_n;
  model->blk_ms = o_ms;
  model->st_blk = o_stblk;
  return S;
}

#define TRGT_HMM_NONE 255

// One candidate of the argmax, reference order, strict '>' (hmm_model.rs:79-88).
#define TRGT_CAND(from_score, lp)                   \
  do {                                              \
    const double sc_ = ((from_score) + (lp)) + em;  \
    if (sc_ > best) { best = sc_; arg = idx; }      \
    idx++;                                          \
  } while (0)

template <class G>
TRGT_HD void hmm_viterbi_small(const G &g, const HmmModel &m, const HmmConsts &c, const double *mm_lp,
                               const uint8_t *allele, int L, double *sc0, double *sc1, uint8_t *bp);

// Viterbi over '#' + allele + '#'.  sc0/sc1: two columns of S doubles (on-chip).  bp: (L+2)*S bytes.
template <class G>
TRGT_HD void hmm_viterbi(const G &g, const HmmModel &m, const HmmConsts &c, const double *mm_lp,
                         const uint8_t *allele, int L, double *sc0, double *sc1, uint8_t *bp) {
  const int S = m.S, nb = m.nb;
  if (S <= g.size()) {  // one state per lane: the register-resident variant below
    hmm_viterbi_small(g, m, c, mm_lp, allele, L, sc0, sc1, bp);
    return;
  }
  double *prev = sc0, *cur = sc1;
  const double NEG = -INFINITY;
  for (int col = 0; col <= L + 1; col++) {
    const int sym = (col == 0 || col == L + 1)
                        ? 0 : hmm_symbol(hmm_clean_base(allele[col - 1], (uint32_t)(col - 1)));
    uint8_t *bpc = bp + (size_t)col * (size_t)S;
    // (1) emitting states, look back one column (hmm_model.rs:62-69)
    for (int st = g.lane(); st < S; st += g.size()) {
      const HmmRole r = hmm_role(m, st);
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      double em;
      switch (r.kind) {
        case HR_START:
          if (col == 0) { best = c.em_one; arg = 0; }  // hmm_model.rs:91-94
          break;
        case HR_END:
          if (col > 0 && sym == 0) { em = c.em_one; TRGT_CAND(prev[S - 2], c.lp_end); }
          break;
        case HR_MATCH:
          if (col > 0 && sym != 0) {
            const uint8_t mb = m.motif_bytes[m.blk_moff[r.b] + r.i];
            em = (mb == 'N') ? c.em_quarter : (hmm_symbol(mb) == sym ? c.em_hi : c.em_lo);
            if (r.i == 0) {
              TRGT_CAND(prev[r.ms], c.lp_match);
            } else {
              TRGT_CAND(prev[st - 1], c.lp_match);
              TRGT_CAND(prev[r.ms], mm_lp[m.blk_mmoff[r.b] + r.i]);
              TRGT_CAND(prev[r.ms + r.n + r.i], c.lp_ins_exit);                     // ins_{i-1}
              if (r.i >= 2) TRGT_CAND(prev[r.ms + 2 * r.n + r.i - 1], c.lp_half);   // del_{i-2}
            }
          }
          break;
        case HR_INS:
          if (col > 0 && sym != 0) {
            em = c.em_quarter;
            TRGT_CAND(prev[st], c.lp_ins_loop);
            TRGT_CAND(prev[r.ms + 1 + r.i], c.lp_indel_open);
          }
          break;
        case HR_SKIP:
          if (col > 0 && sym != 0) {
            em = c.em_quarter;
            TRGT_CAND(prev[r.ms], c.lp_one);
            TRGT_CAND(prev[st], c.lp_half);
          }
          break;
        default:
          continue;  // silent: steps (2) and (3)
      }
      cur[st] = best;
      bpc[st] = (uint8_t)arg;
    }
    g.sync();
    // (2) per block: del chain then me, same column (silent states, hmm_model.rs:62-69)
    for (int b = g.lane(); b < nb; b += g.size()) {
      const int ms = m.blk_ms[b];
      const double em = 0.0;
      if (b == nb - 1) {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[ms + 1], c.lp_half);
        cur[ms + 2] = best;
        bpc[ms + 2] = (uint8_t)arg;
        continue;
      }
      const int n = m.blk_n[b];
      const int m0 = ms + 1, i0 = ms + 1 + n, d0 = ms + 1 + 2 * n, me = ms + 3 * n;
      for (int i = 0; i + 1 < n; i++) {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[m0 + i], c.lp_indel_open);
        if (i > 0) TRGT_CAND(cur[d0 + i - 1], c.lp_half);
        cur[d0 + i] = best;
        bpc[d0 + i] = (uint8_t)arg;
      }
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[m0 + n - 1], c.lp_match);
        TRGT_CAND(cur[i0 + n - 1], c.lp_ins_exit);
        if (n > 1) TRGT_CAND(cur[d0 + n - 2], c.lp_one);
        cur[me] = best;
        bpc[me] = (uint8_t)arg;
      }
    }
    g.sync();
    // (3) re, rs, then every ms
    for (int b = g.lane(); b < nb; b += g.size()) {
      const double em = 0.0;
      double re_sc, rs_sc;
      int re_arg, rs_arg;
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        for (int bb = 0; bb < nb; bb++) {
          const int me = m.blk_ms[bb] + (bb == nb - 1 ? 2 : 3 * (int)m.blk_n[bb]);
          TRGT_CAND(cur[me], c.lp_half);
        }
        re_sc = best; re_arg = arg;
      }
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[0], c.lp_one);
        TRGT_CAND(re_sc, c.lp_one);
        rs_sc = best; rs_arg = arg;
      }
      {
        const int ms = m.blk_ms[b];
        const int me = ms + (b == nb - 1 ? 2 : 3 * (int)m.blk_n[b]);
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(rs_sc, c.lp_one);
        TRGT_CAND(cur[me], c.lp_half);
        cur[ms] = best;
        bpc[ms] = (uint8_t)arg;
      }
      if (b == 0) {
        cur[S - 2] = re_sc; bpc[S - 2] = (uint8_t)re_arg;
        cur[1] = rs_sc;     bpc[1] = (uint8_t)rs_arg;
      }
    }
    g.sync();
    double *t = prev; prev = cur; cur = t;
  }
}

// The same recurrence for models with at most one state per lane (S <= group size; every locus of
// the genome-wide catalog): each lane decodes its state's role, in-edges and ln terms ONCE, keeps
// them in registers, and the column loop is just loads, adds and compares.  Candidate order, the
// strict '>' and the left-to-right sums are those of hmm_viterbi, so results are bit-identical.
template <class G>
TRGT_HD void hmm_viterbi_small(const G &g, const HmmModel &m, const HmmConsts &c, const double *mm_lp,
                               const uint8_t *allele, int L, double *sc0, double *sc1, uint8_t *bp) {
  const int S = m.S, nb = m.nb;
  const int st = g.lane();
  const double NEG = -INFINITY;
  // ---- per-lane constants: emitting state `st` ----
  int kind = -1, nsrc = 0, base_sym = 0;  // base_sym: 1..4 match base, 5 = N / uniform, 0 = '#' only
  int src0 = 0, src1 = 0, src2 = 0, src3 = 0;
  double lp0 = 0, lp1 = 0, lp2 = 0, lp3 = 0;
  if (st < S) {
    const HmmRole r = hmm_role(m, st);
    kind = r.kind;
    switch (r.kind) {
      case HR_END: nsrc = 1; src0 = S - 2; lp0 = c.lp_end; base_sym = 0; break;
      case HR_MATCH: {
        const uint8_t mb = m.motif_bytes[m.blk_moff[r.b] + r.i];
        base_sym = mb == 'N' ? 5 : hmm_symbol(mb);
        if (r.i == 0) { nsrc = 1; src0 = r.ms; lp0 = c.lp_match; }
        else {
          nsrc = r.i >= 2 ? 4 : 3;
          src0 = st - 1; lp0 = c.lp_match;
          src1 = r.ms; lp1 = mm_lp[m.blk_mmoff[r.b] + r.i];
          src2 = r.ms + r.n + r.i; lp2 = c.lp_ins_exit;
          src3 = r.ms + 2 * r.n + r.i - 1; lp3 = c.lp_half;
        }
        break;
      }
      case HR_INS: nsrc = 2; src0 = st; lp0 = c.lp_ins_loop; src1 = r.ms + 1 + r.i; lp1 = c.lp_indel_open; base_sym = 5; break;
      case HR_SKIP: nsrc = 2; src0 = r.ms; lp0 = c.lp_one; src1 = st; lp1 = c.lp_half; base_sym = 5; break;
      default: break;  // start, silent states
    }
  }
  const bool emitting = kind == HR_START || kind == HR_END || kind == HR_MATCH || kind == HR_INS || kind == HR_SKIP;
  // ---- per-lane constants: block `st` (lanes < nb also run the silent chain of one block) ----
  const bool blk = st < nb;
  int b_ms = 0, b_n = 0, b_me = 0;
  if (blk) {
    b_ms = m.blk_ms[st];
    b_n = st == nb - 1 ? 0 : (int)m.blk_n[st];
    b_me = b_ms + (st == nb - 1 ? 2 : 3 * b_n);
  }
  double *prev = sc0, *cur = sc1;
  for (int col = 0; col <= L + 1; col++) {
    const int sym = (col == 0 || col == L + 1)
                        ? 0 : hmm_symbol(hmm_clean_base(allele[col - 1], (uint32_t)(col - 1)));
    uint8_t *bpc = bp + (size_t)col * (size_t)S;
    // (1) emitting states from the previous column
    if (emitting) {
      double best = NEG;
      int arg = TRGT_HMM_NONE;
      if (kind == HR_START) {
        if (col == 0) { best = c.em_one; arg = 0; }
      } else if (col > 0) {
        double em;
        if (base_sym == 0) em = sym == 0 ? c.em_one : NEG;
        else if (sym == 0) em = NEG;
        else if (base_sym == 5) em = c.em_quarter;
        else em = base_sym == sym ? c.em_hi : c.em_lo;
        if (em != NEG) {
          int idx = 0;
          TRGT_CAND(prev[src0], lp0);
          if (nsrc > 1) TRGT_CAND(prev[src1], lp1);
          if (nsrc > 2) TRGT_CAND(prev[src2], lp2);
          if (nsrc > 3) TRGT_CAND(prev[src3], lp3);
        }
      }
      cur[st] = best;
      bpc[st] = (uint8_t)arg;
    }
    g.sync();
    // (2) per block: del chain then me
    if (blk) {
      const double em = 0.0;
      if (st == nb - 1) {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[b_ms + 1], c.lp_half);
        cur[b_ms + 2] = best;
        bpc[b_ms + 2] = (uint8_t)arg;
      } else {
        const int m0 = b_ms + 1, i0 = b_ms + 1 + b_n, d0 = b_ms + 1 + 2 * b_n;
        double dprev = NEG;
        for (int i = 0; i + 1 < b_n; i++) {
          double best = NEG;
          int arg = TRGT_HMM_NONE, idx = 0;
          TRGT_CAND(cur[m0 + i], c.lp_indel_open);
          if (i > 0) TRGT_CAND(dprev, c.lp_half);
          cur[d0 + i] = best;
          bpc[d0 + i] = (uint8_t)arg;
          dprev = best;
        }
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[m0 + b_n - 1], c.lp_match);
        TRGT_CAND(cur[i0 + b_n - 1], c.lp_ins_exit);
        if (b_n > 1) TRGT_CAND(dprev, c.lp_one);
        cur[b_me] = best;
        bpc[b_me] = (uint8_t)arg;
      }
    }
    g.sync();
    // (3) re, rs, then every ms
    if (blk) {
      const double em = 0.0;
      double re_sc, rs_sc;
      int re_arg, rs_arg;
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        for (int bb = 0; bb < nb; bb++) {
          const int me = m.blk_ms[bb] + (bb == nb - 1 ? 2 : 3 * (int)m.blk_n[bb]);
          TRGT_CAND(cur[me], c.lp_half);
        }
        re_sc = best; re_arg = arg;
      }
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(cur[0], c.lp_one);
        TRGT_CAND(re_sc, c.lp_one);
        rs_sc = best; rs_arg = arg;
      }
      {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(rs_sc, c.lp_one);
        TRGT_CAND(cur[b_me], c.lp_half);
        cur[b_ms] = best;
        bpc[b_ms] = (uint8_t)arg;
      }
      if (st == 0) {
        cur[S - 2] = re_sc; bpc[S - 2] = (uint8_t)re_arg;
        cur[1] = rs_sc;     bpc[1] = (uint8_t)rs_arg;
      }
    }
    g.sync();
    double *t = prev; prev = cur; cur = t;
  }
}

// One lane does a whole allele: for the small models of a genome-wide catalog (S = 14..26) a warp per
// allele leaves most lanes idle in the silent-state steps; with a thread per allele 32 alleles advance
// per instruction.  Block geometry comes from HmmModelScan (no tables); score columns are strided
// arrays (sc[state * stride]) so that neighbouring threads hit different shared-memory banks.
// Candidate order, strict '>' and left-to-right sums as in hmm_viterbi: bit-identical results.
template <class M>
TRGT_HD void hmm_viterbi_thread(const M &m, const HmmConsts &c, const uint32_t *mm_off, const double *mm_lp,
                                const uint8_t *allele, int L, double *prev, double *cur, int stride, uint8_t *bp) {
  const int S = m.S, nb = m.nb;
  const double NEG = -INFINITY;
#define SC(a, st) (a)[(size_t)(st) * (size_t)stride]
  uint8_t base_next = L > 0 ? allele[0] : 0;  // the base of column col + 1, loaded one column ahead
  for (int col = 0; col <= L + 1; col++) {
    const uint8_t base_cur = base_next;
    if (col + 1 <= L) base_next = allele[col];
    const int sym = (col == 0 || col == L + 1)
                        ? 0 : hmm_symbol(hmm_clean_base(base_cur, (uint32_t)(col - 1)));
    uint8_t *bpc = bp + (size_t)col * (size_t)S;
    // ---- emitting states, from the previous column ----
    SC(cur, 0) = col == 0 ? c.em_one : NEG;   // start (hmm_model.rs:91-94)
    bpc[0] = col == 0 ? 0 : TRGT_HMM_NONE;
    int ms = 2;
    for (int b = 0; b < nb - 1; b++) {
      const int n = m.block_n(b);
      const double *jump = mm_lp + m.jump_off(b, mm_off);
      for (int i = 0; i < n; i++) {
        {  // match_i
          const int st = ms + 1 + i;
          double best = NEG;
          int arg = TRGT_HMM_NONE, idx = 0;
          if (col > 0 && sym != 0) {
            const uint8_t mb = m.motif_byte(b, i);
            const double em = (mb == 'N') ? c.em_quarter : (hmm_symbol(mb) == sym ? c.em_hi : c.em_lo);
            if (i == 0) {
              TRGT_CAND(SC(prev, ms), c.lp_match);
            } else {
              TRGT_CAND(SC(prev, st - 1), c.lp_match);
              TRGT_CAND(SC(prev, ms), jump[i]);
              TRGT_CAND(SC(prev, ms + n + i), c.lp_ins_exit);
              if (i >= 2) TRGT_CAND(SC(prev, ms + 2 * n + i - 1), c.lp_half);
            }
          }
          SC(cur, st) = best;
          bpc[st] = (uint8_t)arg;
        }
        {  // ins_i
          const int st = ms + 1 + n + i;
          double best = NEG;
          int arg = TRGT_HMM_NONE, idx = 0;
          if (col > 0 && sym != 0) {
            const double em = c.em_quarter;
            TRGT_CAND(SC(prev, st), c.lp_ins_loop);
            TRGT_CAND(SC(prev, ms + 1 + i), c.lp_indel_open);
          }
          SC(cur, st) = best;
          bpc[st] = (uint8_t)arg;
        }
      }
      ms += 3 * n + 1;
    }
    const int sk = ms;  // skip block: ms_skip, skip, me_skip
    {
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      if (col > 0 && sym != 0) {
        const double em = c.em_quarter;
        TRGT_CAND(SC(prev, sk), c.lp_one);
        TRGT_CAND(SC(prev, sk + 1), c.lp_half);
      }
      SC(cur, sk + 1) = best;
      bpc[sk + 1] = (uint8_t)arg;
    }
    {  // end
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      if (col > 0 && sym == 0) {
        const double em = c.em_one;
        TRGT_CAND(SC(prev, S - 2), c.lp_end);
      }
      SC(cur, S - 1) = best;
      bpc[S - 1] = (uint8_t)arg;
    }
    // ---- silent states, same column: del chain and me per block, me_skip, re, rs, every ms ----
    const double em = 0.0;
    double re_sc = NEG;
    int re_arg = TRGT_HMM_NONE, re_idx = 0;
    ms = 2;
    for (int b = 0; b < nb - 1; b++) {
      const int n = m.block_n(b);
      const int m0 = ms + 1, i0 = ms + 1 + n, d0 = ms + 1 + 2 * n, me = ms + 3 * n;
      double dprev = NEG;
      for (int i = 0; i + 1 < n; i++) {
        double best = NEG;
        int arg = TRGT_HMM_NONE, idx = 0;
        TRGT_CAND(SC(cur, m0 + i), c.lp_indel_open);
        if (i > 0) TRGT_CAND(dprev, c.lp_half);
        SC(cur, d0 + i) = best;
        bpc[d0 + i] = (uint8_t)arg;
        dprev = best;
      }
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      TRGT_CAND(SC(cur, m0 + n - 1), c.lp_match);
      TRGT_CAND(SC(cur, i0 + n - 1), c.lp_ins_exit);
      if (n > 1) TRGT_CAND(dprev, c.lp_one);
      SC(cur, me) = best;
      bpc[me] = (uint8_t)arg;
      {  // this me as candidate `b` of re
        const double sc_ = (best + c.lp_half) + em;
        if (sc_ > re_sc) { re_sc = sc_; re_arg = re_idx; }
        re_idx++;
      }
      ms += 3 * n + 1;
    }
    {  // me_skip, then its turn as the last candidate of re
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      TRGT_CAND(SC(cur, sk + 1), c.lp_half);
      SC(cur, sk + 2) = best;
      bpc[sk + 2] = (uint8_t)arg;
      const double sc_ = (best + c.lp_half) + em;
      if (sc_ > re_sc) { re_sc = sc_; re_arg = re_idx; }
    }
    SC(cur, S - 2) = re_sc;
    bpc[S - 2] = (uint8_t)re_arg;
    double rs_sc;
    {
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      TRGT_CAND(SC(cur, 0), c.lp_one);
      TRGT_CAND(re_sc, c.lp_one);
      rs_sc = best;
      SC(cur, 1) = best;
      bpc[1] = (uint8_t)arg;
    }
    ms = 2;
    for (int b = 0; b < nb; b++) {
      const int n = m.block_n(b);
      const int me = ms + (b == nb - 1 ? 2 : 3 * n);
      double best = NEG;
      int arg = TRGT_HMM_NONE, idx = 0;
      TRGT_CAND(rs_sc, c.lp_one);
      TRGT_CAND(SC(cur, me), c.lp_half);
      SC(cur, ms) = best;
      bpc[ms] = (uint8_t)arg;
      ms += 3 * n + 1;
    }
    double *t = prev; prev = cur; cur = t;
  }
#undef SC
}

// predecessor of `st` through in-edge `e` (inverse of the enumeration order above)
template <class M>
TRGT_HD int hmm_pred(const M &m, const HmmRole &r, int st, int e) {
  switch (r.kind) {
    case HR_END: return m.S - 2;
    case HR_RE: return m.block_ms(e) + (e == m.nb - 1 ? 2 : 3 * m.block_n(e));
    case HR_RS: return e == 0 ? 0 : m.S - 2;
    case HR_MS: return e == 0 ? 1 : r.ms + 3 * r.n;
    case HR_SKIP_MS: return e == 0 ? 1 : r.ms + 2;
    case HR_MATCH:
      if (r.i == 0) return r.ms;
      return e == 0 ? st - 1 : (e == 1 ? r.ms : (e == 2 ? r.ms + r.n + r.i : r.ms + 2 * r.n + r.i - 1));
    case HR_INS: return e == 0 ? st : r.ms + 1 + r.i;
    case HR_DEL: return e == 0 ? r.ms + 1 + r.i : st - 1;
    case HR_ME: return e == 0 ? r.ms + r.n : (e == 1 ? r.ms + 2 * r.n : r.ms + 3 * r.n - 1);
    case HR_SKIP: return e == 0 ? r.ms : st;
    case HR_SKIP_ME: return r.ms + 1;
    default: return 0;
  }
}

struct HmmAnnot {
  double purity;
  uint32_t n_spans;
  int32_t status;  // 0 ok, <0 broken back-pointer chain
};

// One reverse walk from (end, L+1) to start.  mc[nb-1] must be zeroed by the caller when non-null.
// When spans_out is non-null the collapsed spans are written in forward order into the LAST n_spans of
// its n_total slots: n_total is either the count from a previous counting walk or any upper bound (a
// collapsed span covers at least one base, so L is one).  max_motif_len: tr.rs:468 passes 6.
// path_out (optional): receives the state path (Hmm::label).  With path_total == 0 it is written
// in REVERSE order, up to path_cap entries; with path_total = the length found by a previous walk
// it is written in forward order.  *path_len gets the full length.
// how the walk reads a back-pointer: S bytes per column (hmm_viterbi, hmm_viterbi_thread) ...
struct HmmBpBytes {
  const uint8_t *bp;
  int S;
  TRGT_HD int get(int col, int st, const HmmRole &) const { return bp[(size_t)col * (size_t)S + (size_t)st]; }
};

// where the walk leaves collapsed span `idx`: a plain array ...
struct HmmSpanArray {
  HmmSpan *p;
  TRGT_HD bool on() const { return p != nullptr; }
  TRGT_HD void put(uint32_t idx, const HmmSpan &s) const { p[idx] = s; }
};
// ... or slots `stride` apart (the lanes of a warp share a block of scratch, slot j of lane l at j * 32 + l)
struct HmmSpanStrided {
  HmmSpan *p;
  uint32_t stride;
  TRGT_HD bool on() const { return p != nullptr; }
  TRGT_HD void put(uint32_t idx, const HmmSpan &s) const { p[(size_t)idx * stride] = s; }
};

template <class M, class BP, class SP>
TRGT_HD HmmAnnot hmm_annotate_bp(const M &m, const uint8_t *allele, int L, BP &bpr,
                                 int max_motif_len, uint32_t *mc, const SP &spans_out, uint32_t n_total,
                                 uint32_t *path_out, uint64_t path_cap, uint64_t path_total,
                                 uint64_t *path_len);

template <class M>
TRGT_HD HmmAnnot hmm_annotate(const M &m, const uint8_t *allele, int L, const uint8_t *bp,
                              int max_motif_len, uint32_t *mc, HmmSpan *spans_out, uint32_t n_total,
                              uint32_t *path_out, uint64_t path_cap, uint64_t path_total,
                              uint64_t *path_len) {
  HmmBpBytes bpr{bp, m.S};
  return hmm_annotate_bp(m, allele, L, bpr, max_motif_len, mc, HmmSpanArray{spans_out}, n_total, path_out, path_cap,
                         path_total, path_len);
}

template <class M, class BP, class SP>
TRGT_HD HmmAnnot hmm_annotate_bp(const M &m, const uint8_t *allele, int L, BP &bpr,
                                 int max_motif_len, uint32_t *mc, const SP &spans_out, uint32_t n_total,
                                 uint32_t *path_out, uint64_t path_cap, uint64_t path_total,
                                 uint64_t *path_len) {
  HmmAnnot out;
  out.purity = 0.0; out.n_spans = 0; out.status = 0;
  const int S = m.S;
  uint64_t n_match = 0, n_mis = 0, n_ins = 0, n_del = 0, n_skip = 0;
  int st = S - 1, col = L + 1, last = -1;
  int copy_end = 0;
  bool have_p = false;
  uint32_t p_motif = 0, p_start = 0, p_end = 0, emitted = 0;
  uint64_t plen = 0;
  // every column visits each state at most once going backwards; anything longer is a cycle
  const uint64_t max_steps = ((uint64_t)L + 2) * (uint64_t)S + 2;
  uint64_t steps = 0;
  while (st != 0) {
    if (++steps > max_steps) { out.status = -1; break; }
    if (path_out && plen < path_cap) path_out[path_total ? path_total - 1 - plen : plen] = (uint32_t)st;
    plen++;
    const HmmRole r = hmm_role(m, st);
    bool emits = false;
    switch (r.kind) {
      case HR_END: emits = true; break;
      case HR_ME: case HR_SKIP_ME: copy_end = col; break;
      case HR_MS: case HR_SKIP_MS: {
        const int copy_start = col;
        n_del += (uint64_t)(last - st - 1);  // jump-in to match_i counts i deletions, events.rs:44-46
        if (r.kind == HR_MS) {
          bool keep = true;  // operations.rs:46-62
          if (r.n <= max_motif_len) {
            if (copy_end - copy_start < r.n) {
              keep = false;
            } else {
              for (int i = 0; i < r.n; i++) {
                const uint8_t expected = m.motif_byte(r.b, i);
                const uint8_t observed = hmm_clean_base(allele[copy_start + i], (uint32_t)(copy_start + i));
                if (expected != 'N' && observed != expected) keep = false;
              }
            }
          }
          if (keep) {  // a real motif copy: count it and fold it into the run being collapsed
            if (mc) mc[r.b]++;
            if (have_p && p_motif == (uint32_t)r.b && p_start == (uint32_t)copy_end) {
              p_start = (uint32_t)copy_start;
            } else {
              if (have_p) {
                if (spans_out.on() && emitted < n_total) {
                  HmmSpan s; s.motif_index = p_motif; s.start = p_start; s.end = p_end;
                  spans_out.put(n_total - 1 - emitted, s);
                }
                emitted++;
              }
              have_p = true;
              p_motif = (uint32_t)r.b; p_start = (uint32_t)copy_start; p_end = (uint32_t)copy_end;
            }
          }
        }
        break;
      }
      case HR_MATCH: {
        emits = true;
        const uint8_t expected = m.motif_byte(r.b, r.i);
        const uint8_t base = hmm_clean_base(allele[col - 1], (uint32_t)(col - 1));
        if (base == expected || expected == 'N') n_match++; else n_mis++;
        break;
      }
      case HR_INS: emits = true; n_ins++; break;
      case HR_DEL: n_del++; break;
      case HR_SKIP: emits = true; n_skip++; break;
      default: break;  // rs, re
    }
    const int e = bpr.get(col, st, r);
    if (e == TRGT_HMM_NONE) { out.status = -1; break; }
    const int p = hmm_pred(m, r, st, e);
    if (emits) col -= 1;
    last = st;
    st = p;
  }
  if (path_out && plen < path_cap) path_out[path_total ? path_total - 1 - plen : plen] = 0;
  plen++;
  if (path_len) *path_len = plen;
  if (have_p) {
    if (spans_out.on() && emitted < n_total) {
      HmmSpan s; s.motif_index = p_motif; s.start = p_start; s.end = p_end;
      spans_out.put(n_total - 1 - emitted, s);
    }
    emitted++;
  }
  out.n_spans = emitted;
  // purity.rs:11-40
  const double edit = (double)(n_del + n_ins + n_mis + n_skip);
  const uint64_t ref_len = n_match + n_mis + n_del + n_skip;
  const double max_dist = (double)(ref_len > (uint64_t)L ? ref_len : (uint64_t)L);
  out.purity = (max_dist - edit) / max_dist;
  return out;
}

// ---------------------------------------------------------------- single-motif loci, one lane per allele ---
//
// Every locus of a genome-wide catalog has ONE motif of 2..6 bases (S = 3n + 8 = 14..26).  For those
// the recurrence is written out per motif length N (template), with the live score column -- ms, match[N],
// ins[N], del[N-1], skip, re -- in the lane's REGISTERS and every loop unrolled, and the back-pointers of one
// column packed into ONE 32-bit word per allele (4N bits) instead of S bytes:
//   * a state with a single in-edge needs no bits (match_0, del_0, skip_me, end; start);
//   * rs: in-edge 1 (re) in every column but column 0, where it is in-edge 0 (start);
//   * ms and skip_ms always take in-edge 0 (rs): their other candidate me + ln(1/2) is one of the candidates
//     re = rs was maximised over, so it can never be strictly greater (hmm_model.rs:84) -- this also makes
//     ms = skip_ms = rs = re one carried value;
//   * bits 30-31 of the word carry the column's base, so that the walk never reads the allele;
//   * columns 0 and L+1 need no word at all: in column 0 only start, rs, ms, skip_ms are finite (in-edges 0),
//     in column L+1 only `end` (one in-edge).
// Adding ln(1.0) = +0.0 (lp_one, em_one, the emission of a silent state) leaves every score bit for bit as it
// is -- no score is ever -0.0 -- so those additions are not issued; all other sums are the reference's
// left-to-right `(prev + ln(trans)) + ln(emit)` (hmm_model.rs:79-88), candidates in the reference's in-edge
// order with strict '>'.
template <int N>
struct HmmLaneBits {
  static constexpr int I0 = 0;                               // ins_i: bit i
  static constexpr int M0 = N;                               // match_i, i >= 1: two bits at M0 + 2 (i - 1)
  static constexpr int D0 = 3 * N - 2;                       // del_i, i >= 1: bit D0 + (i - 1)
  static constexpr int SKIP = D0 + (N >= 2 ? N - 2 : 0);
  static constexpr int ME = SKIP + 1;                        // two bits
  static constexpr int RE = ME + 2;
  static constexpr int BITS = RE + 1;                        // 4 N for N >= 2
};
#define HMM_LANE_BASE 30  // two-bit code of the (sanitised) base the column emits, for the walk: it never reads the allele
TRGT_HD uint32_t hmm_base_code(uint8_t b) { return ((uint32_t)b >> 1) & 3u; }  // A 0, C 1, T 2, G 3
#define HMM_LANE_NMAX 7  // 4 N <= 28: bits 30-31 of a column's word carry the column's base (HMM_LANE_BASE)

// first maximum, strict '>': `arg` of the reference's argmax whenever one candidate is finite
#define TRGT_LANE_MAX2(a0, a1, best, arg)        \
  do {                                           \
    const bool g_ = (a1) > (a0);                 \
    best = g_ ? (a1) : (a0);                     \
    arg = g_ ? 1u : 0u;                          \
  } while (0)

// The table-free model of a single-motif locus, for the walk (same accessors as HmmModelScan)
template <int N>
struct HmmModelSingle {
  static constexpr int S = 3 * N + 8;
  static constexpr int nb = 2;
  uint64_t bytes;  // sanitised motif bytes
  TRGT_HD int block_n(int b) const { return b == 0 ? N : 0; }
  TRGT_HD int block_ms(int b) const { return b == 0 ? 2 : 3 + 3 * N; }
  TRGT_HD int block_of(int st) const { return st <= 2 + 3 * N ? 0 : 1; }
  TRGT_HD uint8_t motif_byte(int, int i) const { return (uint8_t)(bytes >> (8 * i)); }
};

TRGT_HD uint64_t hmm_pack_motif(const uint8_t *motif, int n) {
  uint64_t v = 0;
  for (int i = 0; i < n; i++) v |= (uint64_t)hmm_clean_motif_base(motif[i], (uint32_t)i) << (8 * i);
  return v;
}

// Viterbi of one allele by one lane.  jump[i] = ln(seed (N - i)), i in 1..N-1 (builder.rs:93-111).
// bp: one word per column 1..L at bp[(col - 1) * stride] (stride = 32: the lanes of a warp write one line).
template <int N>
TRGT_HD void hmm_viterbi_lane(const HmmConsts &c, const double *jump, uint64_t mbytes, const uint8_t *allele, int L,
                              uint32_t *bp, int stride) {
  typedef HmmLaneBits<N> B;
  const double NEG = -INFINITY;
  double M[N], I[N], D[N > 1 ? N - 1 : 1];
#pragma unroll
  for (int i = 0; i < N; i++) { M[i] = NEG; I[i] = NEG; }
#pragma unroll
  for (int i = 0; i + 1 < N; i++) D[i] = NEG;
  double skip = NEG;
  double r = c.em_one;  // column 0: start = ln(1); rs, ms, skip_ms follow through ln(1) edges; re is -inf but never read
  uint8_t base_next = L > 0 ? allele[0] : 0;  // loaded a column ahead
  for (int col = 1; col <= L; col++) {
    const uint8_t base = hmm_clean_base(base_next, (uint32_t)(col - 1));
    if (col < L) base_next = allele[col];
    uint32_t w = hmm_base_code(base) << HMM_LANE_BASE;
    const double em_q = c.em_quarter;
    double nM[N], nI[N];
    // ---- emitting states, from the previous column ----
#pragma unroll
    for (int i = 0; i < N; i++) {  // ins_i: [ins_i, match_i]
      const double a0 = (I[i] + c.lp_ins_loop) + em_q;
      const double a1 = (M[i] + c.lp_indel_open) + em_q;
      unsigned arg;
      TRGT_LANE_MAX2(a0, a1, nI[i], arg);
      w |= arg << (B::I0 + i);
    }
#pragma unroll
    for (int i = 0; i < N; i++) {  // match_i: [match_{i-1}, ms, ins_{i-1}, del_{i-2}]
      const uint8_t mb = (uint8_t)(mbytes >> (8 * i));
      const double em = (mb == 'N') ? c.em_quarter : (mb == base ? c.em_hi : c.em_lo);
      if (i == 0) {
        nM[0] = (r + c.lp_match) + em;
      } else {
        const double a0 = (M[i - 1] + c.lp_match) + em;
        const double a1 = (r + jump[i]) + em;
        const double a2 = (I[i - 1] + c.lp_ins_exit) + em;
        double best;
        unsigned arg;
        TRGT_LANE_MAX2(a0, a1, best, arg);
        if (a2 > best) { best = a2; arg = 2u; }
        if (i >= 2) {
          const double a3 = (D[i >= 2 ? i - 2 : 0] + c.lp_half) + em;
          if (a3 > best) { best = a3; arg = 3u; }
        }
        nM[i] = best;
        w |= arg << (B::M0 + 2 * (i - 1));
      }
    }
    {  // skip: [skip_ms, skip]
      const double a0 = r + em_q;  // (skip_ms + ln 1) + em
      const double a1 = (skip + c.lp_half) + em_q;
      unsigned arg;
      TRGT_LANE_MAX2(a0, a1, skip, arg);
      w |= arg << B::SKIP;
    }
    // ---- silent states of this column ----
#pragma unroll
    for (int i = 0; i < N; i++) { M[i] = nM[i]; I[i] = nI[i]; }
#pragma unroll
    for (int i = 0; i + 1 < N; i++) {  // del_i: [match_i, del_{i-1}]
      const double a0 = M[i] + c.lp_indel_open;
      if (i == 0) {
        D[0] = a0;
      } else {
        const double a1 = D[i - 1] + c.lp_half;
        unsigned arg;
        TRGT_LANE_MAX2(a0, a1, D[i], arg);
        w |= arg << (B::D0 + (i - 1));
      }
    }
    double me;
    {  // me: [match_{N-1}, ins_{N-1}, del_{N-2}]
      const double a0 = M[N - 1] + c.lp_match;
      const double a1 = I[N - 1] + c.lp_ins_exit;
      unsigned arg;
      TRGT_LANE_MAX2(a0, a1, me, arg);
      if (N > 1) {
        const double a2 = D[N > 1 ? N - 2 : 0];  // + ln 1
        if (a2 > me) { me = a2; arg = 2u; }
      }
      w |= arg << B::ME;
    }
    {  // re: [me, skip_me]; rs = re (its other in-edge, start, is -inf here); ms = skip_ms = rs
      const double a0 = me + c.lp_half;
      const double a1 = (skip + c.lp_half) + c.lp_half;  // skip_me = skip + ln(1/2)
      unsigned arg;
      TRGT_LANE_MAX2(a0, a1, r, arg);
      w |= arg << B::RE;
    }
    bp[(size_t)(col - 1) * (size_t)stride] = w;
  }
}

// ... or one packed word per column (hmm_viterbi_lane).  The walk visits the columns in decreasing order, several
// states per column: the word of the current column is kept, and the line a few columns further down is asked for
// ahead of time (the 32 lanes of a warp share every line).
#define HMM_LANE_PREFETCH 8
template <int N>
struct HmmBpWords {
  const uint32_t *w;
  int stride, L;
  int cur_col;
  uint32_t cur_w;
  TRGT_HD HmmBpWords(const uint32_t *w_, int stride_, int L_) : w(w_), stride(stride_), L(L_), cur_col(-1), cur_w(0) {}
  TRGT_HD int get(int col, int st, const HmmRole &r) {
    typedef HmmLaneBits<N> B;
    (void)st;
    if (col <= 0) return 0;          // column 0: rs <- start, ms / skip_ms <- rs
    if (col > L) return 0;           // column L + 1: end <- re
    if (col != cur_col) {
      cur_col = col;
      cur_w = w[(size_t)(col - 1) * (size_t)stride];
#if defined(__CUDA_ARCH__)
      if (col > HMM_LANE_PREFETCH)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(w + (size_t)(col - 1 - HMM_LANE_PREFETCH) * (size_t)stride));
#endif
    }
    const uint32_t v = cur_w;
    switch (r.kind) {
      case HR_INS: return (int)((v >> (B::I0 + r.i)) & 1u);
      case HR_MATCH: return r.i == 0 ? 0 : (int)((v >> (B::M0 + 2 * (r.i - 1))) & 3u);
      case HR_DEL: return r.i == 0 ? 0 : (int)((v >> (B::D0 + (r.i - 1))) & 1u);
      case HR_SKIP: return (int)((v >> B::SKIP) & 1u);
      case HR_ME: return (int)((v >> B::ME) & 3u);
      case HR_RE: return (int)((v >> B::RE) & 1u);
      case HR_RS: return 1;
      default: return 0;             // ms, skip_ms, skip_me, end
    }
  }
};

// ---- the walk of a single-motif allele, table driven ---------------------------------------------------------
// hmm_annotate_bp decodes a state's role and its predecessor arithmetically at every step; for the lane kernels
// that is most of the walk.  Here everything a step needs about state `st` of the model of motif length n is one
// 8-byte table entry, built once per CTA from that same arithmetic (hmm_role / hmm_pred / HmmLaneBits):
//   x = the four predecessors (one byte each, indexed by in-edge);
//   y = shift | mask << 8 | kind << 16 | i << 24   (where the in-edge sits in the column's packed word).
struct HmmModelSingleRt {  // HmmModelSingle with the motif length at run time (only for building tables)
  int n, S, nb;
  TRGT_HD explicit HmmModelSingleRt(int n_) : n(n_), S(3 * n_ + 8), nb(2) {}
  TRGT_HD int block_n(int b) const { return b == 0 ? n : 0; }
  TRGT_HD int block_ms(int b) const { return b == 0 ? 2 : 3 + 3 * n; }
  TRGT_HD int block_of(int st) const { return st <= 2 + 3 * n ? 0 : 1; }
};

struct HmmLaneEntry {
  uint32_t x, y;
};

TRGT_HD HmmLaneEntry hmm_lane_table_entry(int n, int st) {
  const HmmModelSingleRt m(n);
  HmmLaneEntry t;
  t.x = 0; t.y = 0;
  if (st >= m.S) return t;
  const HmmRole r = hmm_role(m, st);
  for (int e = 0; e < 4; e++) t.x |= ((uint32_t)hmm_pred(m, r, st, e) & 255u) << (8 * e);
  const int M0 = n, D0 = 3 * n - 2, SKIP = D0 + (n >= 2 ? n - 2 : 0), ME = SKIP + 1, RE = ME + 2;  // HmmLaneBits<n>
  uint32_t shift = 0, mask = 0;
  switch (r.kind) {
    case HR_INS: shift = (uint32_t)r.i; mask = 1; break;
    case HR_MATCH: if (r.i > 0) { shift = (uint32_t)(M0 + 2 * (r.i - 1)); mask = 3; } break;
    case HR_DEL: if (r.i > 0) { shift = (uint32_t)(D0 + r.i - 1); mask = 1; } break;
    case HR_SKIP: shift = (uint32_t)SKIP; mask = 1; break;
    case HR_ME: shift = (uint32_t)ME; mask = 3; break;
    case HR_RE: shift = (uint32_t)RE; mask = 1; break;
    default: break;  // one in-edge, or rs / ms / skip_ms (see HmmLaneBits)
  }
  t.y = shift | (mask << 8) | ((uint32_t)r.kind << 16) | ((uint32_t)r.i << 24);
  return t;
}

// hmm_annotate_bp for one allele of a single-motif locus over its packed words, driven by `tab` (the 3 n + 8
// entries of motif length n).  Same walk, same results: purity, MC, collapsed spans (into the last n_spans of the
// n_total slots of spans_out).  words: column c at words[(c - 1) * stride].
template <class SP>
TRGT_HD HmmAnnot hmm_walk_table(const HmmLaneEntry *tab, int n, uint64_t mbytes, int L, const uint32_t *words,
                                int stride, int max_motif_len, uint32_t *mc, const SP &spans_out, uint32_t n_total,
                                uint64_t *path_len) {
  HmmAnnot out;
  out.purity = 0.0; out.n_spans = 0; out.status = 0;
  const int S = 3 * n + 8;
  // the motif as two-bit codes (bases i at bits 2 i), 'N' positions as a mask of the same shape
  uint32_t mcodes = 0, nmask = 0;
  for (int k = 0; k < n; k++) {
    const uint8_t mb = (uint8_t)(mbytes >> (8 * k));
    mcodes |= hmm_base_code(mb) << (2 * k);
    if (mb == 'N') nmask |= 3u << (2 * k);
  }
  const uint32_t copy_mask = n < 16 ? (1u << (2 * n)) - 1u : ~0u;
  uint32_t n_match = 0, n_mis = 0, n_ins = 0, n_del = 0, n_skip = 0, n_copies = 0;
  int st = S - 1, col = L + 1, last = -1;
  int copy_end = 0;
  bool have_p = false;
  uint32_t p_start = 0, p_end = 0, emitted = 0;
  uint32_t recent = 0;  // the bases of the columns just left, newest at bits 0-1: at `ms` the copy's first bases
  uint64_t plen = 0;
  const uint64_t max_steps = ((uint64_t)L + 2) * (uint64_t)S + 2;
  // Words of column col and of the one below it.  A column is often left after a single step, so a load issued
  // one column ahead would still be waited for at full latency: the lines further down (a warp's lanes share
  // them) are pulled into L1 HMM_LANE_PREFETCH columns ahead of the walk.
#if defined(__CUDA_ARCH__)
  for (int c = L; c >= 1 && c > L - HMM_LANE_PREFETCH; c--)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(words + (size_t)(c - 1) * (size_t)stride));
#endif
  uint32_t w_cur = 0, w_nxt = L >= 1 ? words[(size_t)(L - 1) * (size_t)stride] : 0u;
  while (st != 0) {
    if (++plen > max_steps) { out.status = -1; break; }
    const HmmLaneEntry t = tab[st];
    const int kind = (int)((t.y >> 16) & 255u), i = (int)(t.y >> 24);
    const bool emits = kind == HR_END || kind == HR_MATCH || kind == HR_INS || kind == HR_SKIP;
    if (kind == HR_MATCH) {  // base == expected || expected == 'N'  (events.rs:88-117)
      const uint32_t d = (((w_cur >> HMM_LANE_BASE) ^ (mcodes >> (2 * i))) & ~(nmask >> (2 * i))) & 3u;
      if (d == 0) n_match++; else n_mis++;
    }
    n_ins += kind == HR_INS ? 1u : 0u;
    n_skip += kind == HR_SKIP ? 1u : 0u;
    n_del += kind == HR_DEL ? 1u : 0u;
    if (kind == HR_ME || kind == HR_SKIP_ME) copy_end = col;
    if (kind == HR_MS || kind == HR_SKIP_MS) {
      const int copy_start = col;
      n_del += (uint32_t)(last - st - 1);  // jump-in to match_i counts i deletions, events.rs:44-46
      if (kind == HR_MS) {
        // operations.rs:46-62: a copy of a short motif is kept only if its first n observed bases spell the motif
        bool keep = true;
        if (n <= max_motif_len) keep = copy_end - copy_start >= n && ((recent ^ mcodes) & ~nmask & copy_mask) == 0;
        if (keep) {
          n_copies++;
          if (have_p && p_start == (uint32_t)copy_end) {
            p_start = (uint32_t)copy_start;
          } else {
            if (have_p) {
              if (spans_out.on() && emitted < n_total) {
                HmmSpan sp; sp.motif_index = 0; sp.start = p_start; sp.end = p_end;
                spans_out.put(n_total - 1 - emitted, sp);
              }
              emitted++;
            }
            have_p = true;
            p_start = (uint32_t)copy_start; p_end = (uint32_t)copy_end;
          }
        }
      }
    }
    // in-edge taken into this state: from the column's word, except where it is implied (HmmLaneBits)
    uint32_t arg = (col >= 1 && col <= L) ? ((w_cur >> (t.y & 255u)) & ((t.y >> 8) & 255u)) : 0u;
    if (kind == HR_RS) arg = col > 0 ? 1u : 0u;
    const int p = (int)((t.x >> (8 * arg)) & 255u);
    if (emits) {
      if (col <= L) recent = (recent << 2) | (w_cur >> HMM_LANE_BASE);
      col -= 1;
      w_cur = w_nxt;
      if (col >= 2) w_nxt = words[(size_t)(col - 2) * (size_t)stride];
#if defined(__CUDA_ARCH__)
      if (col > HMM_LANE_PREFETCH)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(words + (size_t)(col - 1 - HMM_LANE_PREFETCH) * (size_t)stride));
#endif
    }
    last = st;
    st = p;
  }
  plen++;  // the start state
  if (path_len) *path_len = plen;
  if (mc) mc[0] = n_copies;  // count_motifs (hmm/utils.rs:3-9): one motif
  if (have_p) {
    if (spans_out.on() && emitted < n_total) {
      HmmSpan sp; sp.motif_index = 0; sp.start = p_start; sp.end = p_end;
      spans_out.put(n_total - 1 - emitted, sp);
    }
    emitted++;
  }
  out.n_spans = emitted;
  // purity.rs:11-40
  const double edit = (double)((uint64_t)n_del + n_ins + n_mis + n_skip);
  const uint64_t ref_len = (uint64_t)n_match + n_mis + n_del + n_skip;
  const double max_dist = (double)(ref_len > (uint64_t)L ? ref_len : (uint64_t)L);
  out.purity = (max_dist - edit) / max_dist;
  return out;
}

// bytes of on-chip storage one group needs for a model with S states, nb blocks, `mbytes` motif bytes
TRGT_HD size_t hmm_onchip_bytes(int S, int nb, int mbytes) {
  size_t b = 0;
  b += 2 * (size_t)S * sizeof(double);
  b += 2 * (size_t)nb * sizeof(uint32_t);
  b += 2 * (size_t)nb * sizeof(uint16_t);
  b += (size_t)S * sizeof(uint16_t);
  b += (size_t)mbytes;
  return (b + 15) & ~(size_t)15;
}

}  // namespace trgt
