// kernels.cuh -- sm_100a kernels of the tandem-repeat DP engine (included by engine.cu only).
//
//   producer k_clip_cigar      clip_cigar + clip_to_region's query range    clip_region.rs:19-38,105-186
//            k_unpack_seq4     BAM 4-bit bases -> the ASCII reads of phase A   read.rs:104
//   phase A  k_flank_exact     exact flank search by index probes          span_locater.rs:10-12
//            k_flank_band      banded on-chip WFA fallback for the misses  span_locater.rs:14-25
//            k_wfa_score       WFA pass 1 (ring, no history) wfaligner.rs:503-528 (flank), :489 (e2e)
//            k_wfa_trace       WFA pass 2 (cone + back-trace) -> count_matches / span / SAM CIGAR
//            k_flank_combine   find_tr_spans combine rule   span_locater.rs:53-67
//   phase B  k_wfa_score / k_wfa_trace in end-to-end mode, k_cigar_gather   utils/align.rs:14-28
//            k_edit_dist       get_dist_matrix              genotype_cluster.rs:236-286
//            k_consensus_vote  repair_consensus behind the alignments (next row)  consensus.rs:5-111
//   phase C  k_hmm_viterbi, k_hmm_walk, k_hmm_emit          src/hmm/*, tr.rs:454-492
//            k_vcf_fields      AL / MC / MS / AP sample fields (next row)  write_vcf.rs:267-343
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trgt_engine.h"
#include "clip_core.h"
#include "cluster_core.h"
#include "consensus_core.h"
#include "coop.h"
#include "hmm_core.h"
#include "vcf_core.h"
#include "wfa_core.h"

namespace trgt {

// ------------------------------------------------------------------ shared structs ------

struct Counters {
  unsigned int n_work;                  // flank: (read, side) pairs that missed the exact scan
  unsigned int n_trace;                 // e2e: pairs with non-zero cost
  unsigned long long max_trace_ints;    // largest pass-2 workspace any item needs
  unsigned long long words_bound;       // upper bound on CIGAR words pass 2 will emit
  unsigned long long pool_used;         // CIGAR words actually emitted by pass 2
  unsigned int n_failed;                // items that ended with a non-OK status
  unsigned int n_banded;                // flank: pairs settled by k_flank_band_wide
  unsigned int n_tier2;                 // flank: pairs the first cost tier handed to k_flank_band2
  unsigned int n_resid;                 // e2e: pairs k_e2e_lane handed to the warp kernel
  unsigned int n_diff;                  // e2e: members that differ from their backbone (k_e2e_identity)
  unsigned int n_list1;                // flank: pairs the seed pass listed for the band pass (k_flank_seed -> k_flank_band1)
  unsigned int wide_next;               // flank: next position of the work list k_flank_band_wide hands out
};

enum { WFA_MODE_FLANK = 0, WFA_MODE_E2E = 1 };

// where the (pattern, text) of work item `id` lives
struct WfaSrc {
  int mode;
  int x, oe, e;
  // flank mode: id = 2 * read + side
  const uint8_t *reads; const uint64_t *read_off; const uint32_t *read_locus;
  const uint8_t *lp; const uint64_t *lp_off;
  const uint8_t *rp; const uint64_t *rp_off;
  // e2e mode: id = sequence index
  const uint8_t *bb; const uint64_t *bb_off;
  const uint8_t *seqs; const uint64_t *seq_off; const uint32_t *seq_group;
};

__device__ __forceinline__ WfaProb wfa_prob_of(const WfaSrc &s, uint32_t id) {
  WfaProb pr;
  pr.x = s.x; pr.oe = s.oe; pr.e = s.e;
  if (s.mode == WFA_MODE_FLANK) {
    const uint32_t r = id >> 1, side = id & 1u;
    const uint32_t l = s.read_locus[r];
    const uint8_t *pb = side ? s.rp : s.lp;
    const uint64_t *po = side ? s.rp_off : s.lp_off;
    pr.p = pb + po[l];
    pr.P = (int)(po[l + 1] - po[l]);
    pr.t = s.reads + s.read_off[r];
    pr.T = (int)(s.read_off[r + 1] - s.read_off[r]);
    pr.pbf = 0; pr.pef = 0; pr.tbf = pr.T; pr.tef = pr.T;  // span_locater.rs:17
  } else {
    const uint32_t gidx = s.seq_group[id];
    pr.p = s.bb + s.bb_off[gidx];
    pr.P = (int)(s.bb_off[gidx + 1] - s.bb_off[gidx]);
    pr.t = s.seqs + s.seq_off[id];
    pr.T = (int)(s.seq_off[id + 1] - s.seq_off[id]);
    pr.pbf = pr.pef = pr.tbf = pr.tef = 0;
  }
  wfa_unband(pr);
  return pr;
}

// ------------------------------------------------------------------ helpers --------------

// out[i] = group g for every i in [off[g], off[g+1])
__global__ void k_expand_offsets(const uint32_t *__restrict__ off, uint32_t n_groups, uint32_t *__restrict__ out) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gsz)
    for (uint32_t i = off[g]; i < off[g + 1]; i++) out[i] = g;
}

// ------------------------------------------------------------------ producer: clip + decode ---

// clip_reads for a chunk of loci (tr.rs:186-196): one thread per read walks its CIGAR twice (reference
// end, then the clip), a few dozen words.  Bytes per read: 4 * n_ops in, 48 out.
__global__ void __launch_bounds__(256)
k_clip_cigar(const uint32_t *__restrict__ ops, const uint64_t *__restrict__ op_off,
             const long long *__restrict__ ref_starts, const uint32_t *__restrict__ read_locus,
             const long long *__restrict__ regions, uint32_t n_reads, trgt_clip_t *__restrict__ out) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gsz) {
    const uint32_t l = read_locus[r];
    out[r] = clip_cigar_one(ops + op_off[r], (uint32_t)(op_off[r + 1] - op_off[r]), ref_starts[r], regions[2 * l],
                            regions[2 * l + 1]);
  }
}

// Reads r0..r1-1 from BAM 4-bit bases to ASCII, a warp per read: every lane turns 8-9 packed bytes
// into one aligned 16-byte store (seq4_unpack_read).  HBM-bound: 0.5 B read + 1 B written per base.
__global__ void __launch_bounds__(256)
k_unpack_seq4(const uint8_t *__restrict__ data, const uint64_t *__restrict__ starts,
              const uint32_t *__restrict__ lengths, const unsigned long long *__restrict__ out_off, uint32_t r0,
              uint32_t r1, uint8_t *__restrict__ out) {
  const WarpGroup g;
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  for (uint32_t r = r0 + blockIdx.x * wpb + warp; r < r1; r += gridDim.x * wpb)
    seq4_unpack_read(g, data, starts[r], lengths[r], out, (uint64_t)out_off[r]);
}

// clip_bases for the BAMlet (write_bam.rs:72-92, clip_bases.rs:9-119) on the resident reads and spans of a phase-A
// batch: a warp per read counts the read's CG dinucleotides, lane 0 walks its CIGAR.  Bytes per read: the read once,
// 4 * n_ops + 8 + 12 in, 48 out.
__global__ void __launch_bounds__(256)
k_bamlet_clip(const uint8_t *__restrict__ reads, const uint64_t *__restrict__ read_off,
              const trgt_span_t *__restrict__ spans, const uint32_t *__restrict__ ops,
              const uint64_t *__restrict__ op_off, const long long *__restrict__ ref_starts, uint32_t flank_len,
              uint32_t n_reads, trgt_bamlet_clip_t *__restrict__ out) {
  const WarpGroup g;
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  for (uint32_t r = blockIdx.x * wpb + warp; r < n_reads; r += gridDim.x * wpb) {
    const trgt_span_t sp = spans[r];
    const uint64_t o = read_off[r];
    const trgt_bamlet_clip_t c = bamlet_clip_read(g, reads + o, (uint32_t)(read_off[r + 1] - o), sp.found != 0, sp.start,
                                                  sp.end, flank_len, ops + op_off[r], (uint32_t)(op_off[r + 1] - op_off[r]),
                                                  ref_starts[r]);
    if (g.lane() == 0) out[r] = c;
  }
}

// ------------------------------------------------------------------ phase A: flank location ---

#define FL_TXT 1664        // bytes per staged read buffer (reads up to ~1.6 KB take the on-chip path)
#define FL_PIECE 432       // bytes of staged flank piece (pieces up to TRGT_KIDX_MAX_P)
#define FL_WS_INTS 1280    // scratch of the second cost tier: 6*17 header + 3 * 23 diagonals * 17 scores

// copy `bytes` (+16 of slack) starting at global `src` into the 16-byte aligned staging buffer with
// 16-byte cp.async issued by `nthreads` threads (this thread is `tid`); returns the staged address of
// src[0], or nullptr if it does not fit.  The caller commits / waits.
__device__ __forceinline__ const uint8_t *stage_bytes(const uint8_t *src, int bytes, uint8_t *dst, int cap,
                                                      int tid, int nthreads) {
  const uintptr_t a = (uintptr_t)src;
  const int shift = (int)(a & 15u);
  const int chunks = (shift + bytes + 15 + 16) >> 4;
  if (chunks * 16 > cap) return nullptr;
  const uint4 *g = (const uint4 *)(a - (uintptr_t)shift);
  for (int c = tid; c < chunks; c += nthreads) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + 16 * c);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(g + c) : "memory");
  }
  return dst + shift;
}

#define TRGT_VIA_PENDING 4   // internal: missed the exact search, waiting for the banded kernels
#define TRGT_VIA_PENDING2 5  // internal: the one-thread first tier could not settle it

// smaller per-warp footprint of the exact-search kernel (no WFA scratch)
struct __align__(16) FlankExactSmem {
  uint16_t slot[2][TRGT_KIDX_SLOTS];
  uint8_t piece[2][FL_PIECE];
  uint8_t txt[2][FL_TXT];
  int cand[TRGT_CAND_CAP + 4];
};

// Phase A, step 1.  A warp takes one locus at a time: it stages the two flank pieces and builds
// their 8-mer indexes once, then walks the locus' reads with the next read's cp.async in flight and
// runs the exact search (span_locater.rs:10-12) by index probes for both flanks.  Misses are marked
// TRGT_VIA_PENDING for k_flank_band (or, with the banded path switched off, appended to `work`).
// Kept separate from the fallback so that this loop -- all reads go through it -- stays a few
// hundred instructions long and inside the instruction cache.
__global__ void __launch_bounds__(32)
k_flank_exact(WfaSrc src, const uint32_t *__restrict__ locus_read_off, uint32_t l_begin, uint32_t l_end,
              int band_budget, trgt_flank_hit_t *__restrict__ hits, uint32_t *__restrict__ work, Counters *ctr) {
  __shared__ FlankExactSmem sm;
  const WarpGroup g;
  const int lane = g.lane();
  for (uint32_t l = l_begin + blockIdx.x; l < l_end; l += gridDim.x) {
    const uint32_t r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    if (r1 <= r0) continue;
    __syncwarp();
    const uint8_t *pg[2];
    int PL[2];
    pg[0] = src.lp + src.lp_off[l]; PL[0] = (int)(src.lp_off[l + 1] - src.lp_off[l]);
    pg[1] = src.rp + src.rp_off[l]; PL[1] = (int)(src.rp_off[l + 1] - src.rp_off[l]);
    const uint8_t *ps[2];
    ps[0] = stage_bytes(pg[0], PL[0], sm.piece[0], FL_PIECE, lane, 32);
    ps[1] = stage_bytes(pg[1], PL[1], sm.piece[1], FL_PIECE, lane, 32);
    const uint8_t *tg = src.reads + src.read_off[r0];
    int T = (int)(src.read_off[r0 + 1] - src.read_off[r0]);
    const uint8_t *t_s = stage_bytes(tg, T, sm.txt[0], FL_TXT, lane, 32);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    bool indexed[2];
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
      indexed[side] = ps[side] != nullptr && PL[side] >= 16 && PL[side] <= TRGT_KIDX_MAX_P;
      if (indexed[side]) kidx_build(g, KmerIndex{sm.slot[side]}, ps[side], PL[side]);
    }
#pragma unroll 1
    for (uint32_t r = r0; r < r1; r++) {
      const int cur = (int)((r - r0) & 1u);
      const uint8_t *tg_next = nullptr, *t_next = nullptr;
      int T_next = 0;
      if (r + 1 < r1) {
        tg_next = src.reads + src.read_off[r + 1];
        T_next = (int)(src.read_off[r + 2] - src.read_off[r + 1]);
        t_next = stage_bytes(tg_next, T_next, sm.txt[cur ^ 1], FL_TXT, lane, 32);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
      __syncwarp();
#pragma unroll 1
      for (int side = 0; side < 2; side++) {
        const uint8_t *p = ps[side] ? ps[side] : pg[side];
        const uint8_t *t = t_s ? t_s : tg;  // very long reads: straight from global memory
        const int P = PL[side];
        trgt_flank_hit_t h;
        h.via = TRGT_VIA_NONE; h.matches = 0; h.score = 0; h.start = 0; h.end = 0;
        if (P > 0) {
          int pos = indexed[side] ? flank_scan_indexed(g, KmerIndex{sm.slot[side]}, p, P, t, T, sm.cand) : -2;
          if (pos == -2) pos = flank_scan(g, p, P, t, T);
          if (pos >= 0) {
            h.via = TRGT_VIA_EXACT; h.matches = P; h.start = (uint32_t)pos; h.end = (uint32_t)(pos + P);
          } else if (band_budget > 0) {
            h.via = TRGT_VIA_PENDING;
          } else if (lane == 0) {
            const unsigned int slot = atomicAdd(&ctr->n_work, 1u);
            work[slot] = 2 * r + (uint32_t)side;
          }
        }
        if (lane == 0) hits[2 * r + side] = h;
        __syncwarp();
      }
      tg = tg_next; T = T_next; t_s = t_next;
    }
  }
}

// Phase A, step 1, for the usual piece lengths (16 <= P <= FXT_PMAX on every locus): still a warp per
// locus for the per-locus set-up (8-mer index + byte-shifted piece copies, 6.5 KB per warp), but then
// ONE LANE PER (read, flank) PAIR: each lane walks its pair's probes and verifies candidates on its own
// with aligned 16-byte loads of the read where it lies in HBM (flank_exact_thread).  No staging of the
// reads, no warp collectives in the loop; only the sectors a probe or a verification touches are read.
#define FXT_WARPS 4

struct __align__(16) FlankExactTSmem {
  uint16_t slot[2][TRGT_KIDX_SLOTS];
  uint8_t copies[2][FXT_COPIES * FXT_STRIDE];
};

__global__ void __launch_bounds__(32 * FXT_WARPS)
k_flank_exact_t(WfaSrc src, const uint32_t *__restrict__ locus_read_off, uint32_t l_begin, uint32_t l_end,
                int band_budget, trgt_flank_hit_t *__restrict__ hits, uint32_t *__restrict__ work, Counters *ctr,
                uint16_t *__restrict__ kidx_out) {
  __shared__ FlankExactTSmem sm_all[FXT_WARPS];
  FlankExactTSmem &sm = sm_all[threadIdx.x >> 5];
  const WarpGroup g;
  const int lane = g.lane();
  for (uint32_t l = l_begin + blockIdx.x * FXT_WARPS + (threadIdx.x >> 5); l < l_end; l += gridDim.x * FXT_WARPS) {
    const uint32_t r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    if (r1 <= r0) continue;
    __syncwarp();
    const int P0 = (int)(src.lp_off[l + 1] - src.lp_off[l]), P1 = (int)(src.rp_off[l + 1] - src.rp_off[l]);
    fxt_build_copies(g, src.lp + src.lp_off[l], P0, sm.copies[0]);
    fxt_build_copies(g, src.rp + src.rp_off[l], P1, sm.copies[1]);
    kidx_build(g, KmerIndex{sm.slot[0]}, sm.copies[0] + 8, P0);  // copy 0 holds the piece itself at byte 8
    kidx_build(g, KmerIndex{sm.slot[1]}, sm.copies[1] + 8, P1);
    if (kidx_out) {  // both tables (2 KB) for the fallback kernels, which then need not rebuild them
      uint4 *dst = (uint4 *)(kidx_out + (size_t)l * 2 * TRGT_KIDX_SLOTS);
      const uint4 *s4 = (const uint4 *)&sm.slot[0][0];
      for (int i = lane; i < (int)(2 * TRGT_KIDX_SLOTS * sizeof(uint16_t) / 16); i += 32) dst[i] = s4[i];
    }
    const uint32_t n_pairs = 2u * (r1 - r0);
    for (uint32_t pb = 0; pb < n_pairs; pb += 32u) {
      const uint32_t p = pb + (uint32_t)lane;
      const unsigned lanes = __ballot_sync(0xffffffffu, p < n_pairs);  // the lanes that search side by side
      if (p >= n_pairs) continue;
      const uint32_t r = r0 + (p >> 1), side = p & 1u;
      const uint64_t o = src.read_off[r];
      const int T = (int)(src.read_off[r + 1] - o);
      const int P = side ? P1 : P0;
      const int pos = flank_exact_thread(KmerIndex{sm.slot[side]}, sm.copies[side], P, src.reads + o, T, lanes);
      trgt_flank_hit_t h;
      h.via = TRGT_VIA_NONE; h.matches = 0; h.score = 0; h.start = 0; h.end = 0;
      if (pos >= 0) {
        h.via = TRGT_VIA_EXACT; h.matches = P; h.start = (uint32_t)pos; h.end = (uint32_t)(pos + P);
      } else if (band_budget > 0) {
        h.via = TRGT_VIA_PENDING;
      } else {
        const unsigned int slot = atomicAdd(&ctr->n_work, 1u);
        work[slot] = 2 * r + side;
      }
      hits[2 * r + side] = h;
    }
  }
}

#ifndef FB_LT
#define FB_LT 16          // lanes per locus in the first cost tier
#endif
#define FB_LOCI (128 / FB_LT)  // loci in flight per CTA of 128 threads
#define FB_LIST 64        // pending pairs gathered per pass over a locus' reads

struct __align__(16) FlankBandSmem {
  uint16_t slot[2][TRGT_KIDX_SLOTS];
  uint8_t piece[2][FL_PIECE];
  uint16_t list[FB_LIST];  // pending pairs of the pass: (read - first read of the pass) << 1 | side
};

// Phase A, step 2.  Only the (read, flank) pairs left pending, and only the first cost tier of the WFA
// fallback (span_locater.rs:14-25): one mismatch or one 1-bp gap, ~89 % of HiFi misses.  Such an
// alignment lives on 3-4 diagonals -- no work for a warp -- so ONE LANE TAKES ONE PAIR
// (flank_locate_tier1_thread: index seed filter + narrow-band wavefront + back-trace, no collectives,
// history in the lane's local memory, the read taken where it lies in HBM).  A locus has ~13 pending
// pairs at 30x, so a half-warp (FB_LT lanes) takes a locus: it stages and indexes the two pieces once
// (3 KB of shared memory) and then its lanes each take a pending pair.
// What a lane cannot settle goes to `work2` (2*read+side) for k_flank_band2.
__global__ void __launch_bounds__(128, 8)
k_flank_band(WfaSrc src, const uint32_t *__restrict__ locus_read_off, uint32_t l_begin, uint32_t l_end,
             int band_budget, double min_flank_id_frac, trgt_flank_hit_t *__restrict__ hits,
             uint32_t *__restrict__ work2, Counters *ctr, const uint16_t *__restrict__ kidx_in, int prefetch) {
  __shared__ FlankBandSmem sm_all[FB_LOCI];
  const TileGroup<FB_LT> g;
  FlankBandSmem &sm = sm_all[threadIdx.x / FB_LT];
  const int lane = g.lane();
  for (uint32_t l = l_begin + blockIdx.x * FB_LOCI + threadIdx.x / FB_LT; l < l_end; l += gridDim.x * FB_LOCI) {
    const uint32_t r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    bool have_index = false;
    bool indexed[2] = {false, false};
    const uint8_t *ps[2] = {nullptr, nullptr};
    int PL[2] = {0, 0};
#pragma unroll 1
    for (uint32_t rb = r0; rb < r1;) {
      g.sync();
      int n_list = 0;
      uint32_t base = rb;
      for (; base < r1 && n_list + 2 * FB_LT <= FB_LIST && base - rb < 16384u; base += FB_LT) {  // pending pairs, in read order
        const uint32_t r = base + (uint32_t)lane;
        unsigned m = 0;
        if (r < r1) m = (hits[2 * r].via == TRGT_VIA_PENDING ? 1u : 0u) | (hits[2 * r + 1].via == TRGT_VIA_PENDING ? 2u : 0u);
        const unsigned b0 = g.ballot(m & 1u), b1 = g.ballot(m & 2u);
        const unsigned lt = (1u << lane) - 1u;
        int pos = n_list + __popc(b0 & lt) + __popc(b1 & lt);
        if (m & 1u) sm.list[pos++] = (uint16_t)(((r - rb) << 1) | 0u);
        if (m & 2u) sm.list[pos] = (uint16_t)(((r - rb) << 1) | 1u);
        n_list += __popc(b0) + __popc(b1);
      }
      const uint32_t rb_pass = rb;
      rb = base;
      g.sync();
      if (n_list == 0) continue;
      for (int i = lane; prefetch && i < n_list; i += FB_LT) {  // pull the pending reads towards the SM while the pieces are set up
        const uint32_t r = rb_pass + (uint32_t)(sm.list[i] >> 1);
        const uint8_t *t = src.reads + src.read_off[r];
        const int T = (int)(src.read_off[r + 1] - src.read_off[r]);
        for (int o = 0; o < T && o < 4096; o += 128) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(t + o));
      }
      if (!have_index) {  // first pending pair of the locus: stage its pieces, fetch or build their indexes
        have_index = true;
        const uint8_t *pgl = src.lp + src.lp_off[l], *pgr = src.rp + src.rp_off[l];
        PL[0] = (int)(src.lp_off[l + 1] - src.lp_off[l]);
        PL[1] = (int)(src.rp_off[l + 1] - src.rp_off[l]);
        ps[0] = stage_bytes(pgl, PL[0], sm.piece[0], FL_PIECE, lane, FB_LT);
        ps[1] = stage_bytes(pgr, PL[1], sm.piece[1], FL_PIECE, lane, FB_LT);
        if (kidx_in) {  // k_flank_exact_t left both tables in HBM
          const uint4 *s4 = (const uint4 *)(kidx_in + (size_t)l * 2 * TRGT_KIDX_SLOTS);
          for (int i = lane; i < (int)(2 * TRGT_KIDX_SLOTS * sizeof(uint16_t) / 16); i += FB_LT) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared((uint4 *)&sm.slot[0][0] + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(s4 + i) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        g.sync();
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
          indexed[side] = ps[side] != nullptr && PL[side] >= 16 && PL[side] <= TRGT_KIDX_MAX_P;
          if (indexed[side] && !kidx_in) kidx_build(g, KmerIndex{sm.slot[side]}, ps[side], PL[side]);
        }
      }
#pragma unroll 1
      for (int i = lane; i < n_list; i += FB_LT) {  // one pending pair per lane
        const int side = sm.list[i] & 1;
        const uint32_t r = rb_pass + (uint32_t)(sm.list[i] >> 1);
        int deferred = 1;
        FlankHit fh;
        fh.via = 0; fh.matches = 0; fh.score = 0; fh.start = 0; fh.end = 0;
        if (indexed[side]) {
          int ws[FT1_WS_INTS];
          WfaProb pr;
          pr.x = src.x; pr.oe = src.oe; pr.e = src.e;
          pr.p = ps[side]; pr.P = PL[side];
          pr.t = src.reads + src.read_off[r];
          pr.T = (int)(src.read_off[r + 1] - src.read_off[r]);
          pr.pbf = 0; pr.pef = 0; pr.tbf = pr.T; pr.tef = pr.T;  // span_locater.rs:17
          wfa_unband(pr);
          deferred = flank_locate_tier1_thread(pr, band_budget, min_flank_id_frac, ws, &fh, KmerIndex{sm.slot[side]});
        }
        trgt_flank_hit_t h;
        h.via = TRGT_VIA_NONE; h.matches = 0; h.score = 0; h.start = 0; h.end = 0;
        if (!deferred) {
          h.via = fh.via; h.matches = fh.matches; h.score = fh.score;
          h.start = (uint32_t)fh.start; h.end = (uint32_t)fh.end;
        } else {
          const unsigned int slot = atomicAdd(&ctr->n_tier2, 1u);
          work2[slot] = 2 * r + (uint32_t)side;
        }
        hits[2 * r + side] = h;
      }
      g.sync();
    }
  }
}

// ---- bulk copies (TMA, 1-D) and their completion barriers ----
// cp.async.bulk moves a 16-byte aligned run of global memory into shared memory with ONE instruction of ONE
// thread and reports the bytes to an mbarrier; nobody holds registers for the data in flight.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// this thread's arrival, announcing `bytes` of bulk copies that complete on the barrier
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

struct __align__(16) FlankSeedSmem {
  uint16_t slot[2][TRGT_KIDX_SLOTS];  // both tables of the locus as k_flank_exact_t left them: one bulk copy
  uint16_t list[FB_LIST];             // pending pairs of the pass: (read - first read of the pass) << 1 | side
  unsigned long long bar;             // completion of the table copy
};

// Phase A, step 2a (seed pass of the first cost tier, see flank_tier1_seed_thread).  A half-warp per locus scans the
// hit records of its reads; for every pending pair ONE LANE loads all probes of the read up front and takes the hull
// of the diagonals of all index hits: `list1` (2 * read + side, klo << 3 | width - 1) for k_flank_band1, or `work2`
// for k_flank_band2 if the pair is not for this tier.  Two global round trips per pair: hit records, probes.
__global__ void __launch_bounds__(128, 8)
k_flank_seed(WfaSrc src, const uint32_t *__restrict__ locus_read_off, uint32_t l_begin, uint32_t l_end,
             int band_budget, trgt_flank_hit_t *__restrict__ hits, uint2 *__restrict__ list1,
             uint32_t *__restrict__ work2, Counters *ctr, const uint16_t *__restrict__ kidx_in) {
  __shared__ FlankSeedSmem sm_all[FB_LOCI];
  const TileGroup<FB_LT> g;
  FlankSeedSmem &sm = sm_all[threadIdx.x / FB_LT];
  const int lane = g.lane();
  if (lane == 0) mbar_init(&sm.bar, 1);
  mbar_init_fence();
  __syncthreads();
  unsigned parity = 0;
  for (uint32_t l = l_begin + blockIdx.x * FB_LOCI + threadIdx.x / FB_LT; l < l_end; l += gridDim.x * FB_LOCI) {
    const uint32_t r0 = locus_read_off[l], r1 = locus_read_off[l + 1];
    bool have_index = false;
    const int P0 = (int)(src.lp_off[l + 1] - src.lp_off[l]), P1 = (int)(src.rp_off[l + 1] - src.rp_off[l]);
    const uint8_t *pg0 = src.lp + src.lp_off[l], *pg1 = src.rp + src.rp_off[l];
#pragma unroll 1
    for (uint32_t rb = r0; rb < r1;) {
      g.sync();
      int n_list = 0;
      uint32_t base = rb;
      for (; base < r1 && n_list + 2 * FB_LT <= FB_LIST && base - rb < 16384u; base += FB_LT) {  // pending pairs, in read order
        const uint32_t r = base + (uint32_t)lane;
        unsigned m = 0;
        if (r < r1) m = (hits[2 * r].via == TRGT_VIA_PENDING ? 1u : 0u) | (hits[2 * r + 1].via == TRGT_VIA_PENDING ? 2u : 0u);
        const unsigned b0 = g.ballot(m & 1u), b1 = g.ballot(m & 2u);
        const unsigned lt = (1u << lane) - 1u;
        int pos = n_list + __popc(b0 & lt) + __popc(b1 & lt);
        if (m & 1u) sm.list[pos++] = (uint16_t)(((r - rb) << 1) | 0u);
        if (m & 2u) sm.list[pos] = (uint16_t)(((r - rb) << 1) | 1u);
        n_list += __popc(b0) + __popc(b1);
      }
      const uint32_t rb_pass = rb;
      rb = base;
      g.sync();
      if (n_list == 0) continue;
      if (!have_index) {  // first pending pair of the locus: fetch both tables (2.6 KB, contiguous) in one bulk copy
        have_index = true;
        if (lane == 0) {
          mbar_arrive_expect(&sm.bar, (unsigned)sizeof(sm.slot));
          bulk_g2s(&sm.slot[0][0], kidx_in + (size_t)l * 2 * TRGT_KIDX_SLOTS, (unsigned)sizeof(sm.slot), &sm.bar);
        }
        mbar_wait(&sm.bar, parity);
        parity ^= 1u;
      }
#pragma unroll 1
      for (int ib = 0; ib < n_list; ib += FB_LT) {  // one pending pair per lane
        const int i = ib + lane;
        int listed = 0, deferred = 0;
        uint32_t id = 0;
        int klo = 0, khi = 0;
        if (i < n_list) {
          const int side = sm.list[i] & 1;
          const uint32_t r = rb_pass + (uint32_t)(sm.list[i] >> 1);
          id = 2 * r + (uint32_t)side;
          WfaProb pr;
          pr.x = src.x; pr.oe = src.oe; pr.e = src.e;
          pr.p = side ? pg1 : pg0; pr.P = side ? P1 : P0;
          pr.t = src.reads + src.read_off[r];
          pr.T = (int)(src.read_off[r + 1] - src.read_off[r]);
          pr.pbf = 0; pr.pef = 0; pr.tbf = pr.T; pr.tef = pr.T;  // span_locater.rs:17
          wfa_unband(pr);
          listed = pr.P >= 16 && pr.P <= FT1_PMAX &&
                   flank_tier1_seed_thread(KmerIndex{sm.slot[side]}, pr, band_budget, &klo, &khi) == 1;
          deferred = !listed;
        }
        const unsigned bl = g.ballot(listed), bd = g.ballot(deferred);
        unsigned base1 = 0, base2 = 0;
        if (lane == 0) {
          if (bl) base1 = atomicAdd(&ctr->n_list1, (unsigned)__popc(bl));
          if (bd) base2 = atomicAdd(&ctr->n_tier2, (unsigned)__popc(bd));
        }
        base1 = (unsigned)g.bcast0((int)base1);
        base2 = (unsigned)g.bcast0((int)base2);
        const unsigned lt = (1u << lane) - 1u;
        if (listed) list1[base1 + __popc(bl & lt)] = make_uint2(id, (uint32_t)((klo * 8) | (khi - klo)));
        if (deferred) {
          work2[base2 + __popc(bd & lt)] = id;
          trgt_flank_hit_t h;
          h.via = TRGT_VIA_NONE; h.matches = 0; h.score = 0; h.start = 0; h.end = 0;
          hits[id] = h;
        }
      }
      g.sync();
    }
  }
}

#define FB1_THREADS 64

// dynamic shared memory of k_flank_band1 for `rows` history rows (scores that can have a wavefront under the
// scoring, ft1_live_scores): per lane a text window filled by its own bulk copy, its copy-complete barrier, and
// rows * 3 * FT1_WMAX 16-bit history cells (the lanes of a warp interleaved)
__host__ __device__ inline size_t fb1_smem_bytes(int rows) {
  return (size_t)FB1_THREADS * (FT1_WIN_STRIDE + 8) + (size_t)FB1_THREADS * rows * 3 * FT1_WMAX * sizeof(int16_t);
}

// Phase A, step 2b (band pass of the first cost tier, see flank_tier1_band_thread).  ONE LANE PER LISTED PAIR, dense
// warps: the lane fetches the <= 288 bytes of the read its band can touch with one bulk copy into its own window
// and then runs the narrow-band wavefronts, their history and the back-trace entirely on chip; only the piece is
// read through L1 (the ~13 pending pairs of a locus sit next to each other in the list).  Writes the hit, or hands
// the pair on to `work2` (cost above this tier's cap).
__global__ void __launch_bounds__(FB1_THREADS)
k_flank_band1(WfaSrc src, const uint2 *__restrict__ list1, const unsigned int *n_list1_ptr, int band_budget,
              double min_flank_id_frac, const uint8_t *reads_end, int rows, trgt_flank_hit_t *__restrict__ hits,
              uint32_t *__restrict__ work2, Counters *ctr) {
  extern __shared__ __align__(16) unsigned char fb1_raw[];
  const int tid = threadIdx.x;
  uint8_t *win = fb1_raw + (size_t)tid * FT1_WIN_STRIDE;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(fb1_raw + (size_t)FB1_THREADS * FT1_WIN_STRIDE) + tid;
  int16_t *hist = reinterpret_cast<int16_t *>(fb1_raw + (size_t)FB1_THREADS * (FT1_WIN_STRIDE + 8)) +
                  (size_t)(tid >> 5) * rows * 3 * FT1_WMAX * 32 + (tid & 31);
  mbar_init(bar, 1);
  mbar_init_fence();
  __syncthreads();
  unsigned parity = 0;
  const uint32_t n = *n_list1_ptr;
  // which scores can have a wavefront at all: a property of the scoring, the same for every pair
  const int cap1 = wfa_imin(band_budget, wfa_imax(src.x, src.oe));
  unsigned live_g = 0;
  const unsigned live_m = cap1 <= FT1_SMAX ? ft1_live_scores(src.x, src.oe, src.e, cap1, &live_g) : 1u;
  // whole warps walk the list (a lane past its end keeps company): the wavefront loop runs warp-uniform
  for (uint32_t base = blockIdx.x * FB1_THREADS + (uint32_t)(tid & ~31); base < n; base += gridDim.x * FB1_THREADS) {
    const uint32_t i = base + (uint32_t)(tid & 31);
    const bool have = i < n;
    uint32_t id = 0;
    int klo = 0, khi = 0, a0 = 0;
    WfaProb pr;
    pr.p = nullptr; pr.t = nullptr; pr.P = 0; pr.T = 0; pr.x = src.x; pr.oe = src.oe; pr.e = src.e;
    pr.pbf = pr.pef = pr.tbf = pr.tef = 0; pr.blo = 0; pr.bhi = 0;
    if (have) {
      const uint2 ent = list1[i];
      id = ent.x;
      klo = ((int)ent.y) >> 3;
      khi = klo + (int)(ent.y & 7u);
      pr = wfa_prob_of(src, id);
      // the text the band can touch: offsets [first, last) of the read, from the 16-byte boundary below `first`
      const int first = klo > 0 ? klo : 0;
      const int last = wfa_imin(pr.T, khi + pr.P) + 12;
      const uint8_t *a = pr.t + first;
      const int slack = (int)((uintptr_t)a & 15u);
      const uint8_t *a_al = a - slack;
      a0 = first - slack;
      unsigned bytes = (unsigned)(slack + (last - first) + 15) & ~15u;
      if (bytes > FT1_WIN_BYTES) bytes = FT1_WIN_BYTES;
      const size_t room = (size_t)(reads_end - a_al) & ~(size_t)15;  // never past the padded end of the read buffer
      if ((size_t)bytes > room) bytes = (unsigned)room;
      mbar_arrive_expect(bar, bytes);
      bulk_g2s(win, a_al, bytes, bar);
      mbar_wait(bar, parity);
      parity ^= 1u;
    }
    __syncwarp();
    FlankHit fh;
    fh.via = 0; fh.matches = 0; fh.score = 0; fh.start = 0; fh.end = 0;
    const int deferred = flank_tier1_band_thread<32>(pr, klo, khi, band_budget, min_flank_id_frac, win, a0, hist, &fh,
                                                     0xffffffffu, !have, live_m, live_g);
    __syncwarp();
    if (have) {
      trgt_flank_hit_t h;
      h.via = TRGT_VIA_NONE; h.matches = 0; h.score = 0; h.start = 0; h.end = 0;
      if (!deferred) {
        h.via = fh.via; h.matches = fh.matches; h.score = fh.score;
        h.start = (uint32_t)fh.start; h.end = (uint32_t)fh.end;
      } else {
        const unsigned int slot = atomicAdd(&ctr->n_tier2, 1u);
        work2[slot] = id;
      }
      hits[id] = h;
    }
  }
}

struct __align__(16) FlankBand2Smem {
  uint16_t slot[TRGT_KIDX_SLOTS];
  uint8_t piece[FL_PIECE];
  uint8_t txt[FL_TXT];
  int cand[TRGT_CAND_CAP + 4];
  int ws[FL_WS_INTS];
};

// Phase A, step 2b.  The pairs the first tier handed on (~11 %): one warp per pair, second cost tier
// (budget) with the larger scratch; the piece is staged and indexed per pair.  Failures are
// appended to `work` for k_flank_band_wide and the full-width kernels.
__global__ void __launch_bounds__(32)
k_flank_band2(WfaSrc src, const uint32_t *__restrict__ work2, const unsigned int *n_work2_ptr, int band_budget,
              double min_flank_id_frac, trgt_flank_hit_t *__restrict__ hits, uint32_t *__restrict__ work,
              Counters *ctr, const uint16_t *__restrict__ kidx_in) {
  __shared__ FlankBand2Smem sm;
  const WarpGroup g;
  const int lane = g.lane();
  const uint32_t n = *n_work2_ptr;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t id = work2[i];
    const WfaProb gp = wfa_prob_of(src, id);
    __syncwarp();
    const uint8_t *p_s = stage_bytes(gp.p, gp.P, sm.piece, FL_PIECE, lane, 32);
    const uint8_t *t_s = stage_bytes(gp.t, gp.T, sm.txt, FL_TXT, lane, 32);
    if (kidx_in) {  // the piece's table as k_flank_exact_t left it in HBM (1 KB)
      const uint4 *s4 = (const uint4 *)(kidx_in + ((size_t)src.read_locus[id >> 1] * 2 + (id & 1u)) * TRGT_KIDX_SLOTS);
      for (int i = lane; i < (int)(TRGT_KIDX_SLOTS * sizeof(uint16_t) / 16); i += 32) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared((uint4 *)&sm.slot[0] + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(s4 + i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    int deferred = 1;
    FlankHit fh;
    fh.via = 0; fh.matches = 0; fh.score = 0; fh.start = 0; fh.end = 0;
    if (p_s != nullptr && t_s != nullptr && gp.P >= 16 && gp.P <= TRGT_KIDX_MAX_P) {
      const KmerIndex idx{sm.slot};
      if (!kidx_in) kidx_build(g, idx, p_s, gp.P);
      WfaProb pr = gp;
      pr.p = p_s; pr.t = t_s;
      deferred = flank_locate_banded_lean(g, pr, band_budget, min_flank_id_frac, sm.ws, FL_WS_INTS, &fh, idx, sm.cand, 1, 1);
    }
    if (lane == 0) {
      if (!deferred) {
        trgt_flank_hit_t h;
        h.via = fh.via; h.matches = fh.matches; h.score = fh.score;
        h.start = (uint32_t)fh.start; h.end = (uint32_t)fh.end;
        hits[id] = h;
      } else {
        const unsigned int slot = atomicAdd(&ctr->n_work, 1u);
        work[slot] = id;
      }
    }
    __syncwarp();
  }
}

// Phase A, step 3.  Second chance for the pairs k_flank_band deferred (band wider than a warp, cost
// above its budget, scratch too small): one warp per pair with 32 KB of scratch and the general
// banded routine (index or linear seed scan, any band width, history or ring + cone), budgets 18 then 36; the warps
// draw their pairs from a counter (the pairs' costs differ widely).
// What it settles is struck from the work list (0xFFFFFFFF); only pairs without any usable seed are
// left for the full-width kernels.
#ifndef FLW_S0
#define FLW_S0 18  // budgets of the wide pass: FLW_S0, FLW_S1 (if larger), 36  (24, 36: 0.44 ms; 20: 0.37; 18: 0.34; 18, 24, 36: 0.38)
#endif
#ifndef FLW_S1
#define FLW_S1 0
#endif
#ifndef TRGT_WIDE_DYNAMIC
#define TRGT_WIDE_DYNAMIC 1
#endif
#ifndef FLW_WS_INTS
#define FLW_WS_INTS 5120  // (8192: 0.50 ms, 5120: 0.46 ms, 4096: 0.39 ms but 0.2 ms more in the full-width kernels behind it)
#endif

__global__ void __launch_bounds__(32)
k_flank_band_wide(WfaSrc src, uint32_t *__restrict__ work, const unsigned int *n_work_ptr, double min_flank_id_frac,
                  trgt_flank_hit_t *__restrict__ hits, Counters *ctr, const uint16_t *__restrict__ kidx_in) {
  __shared__ __align__(16) int ws[FLW_WS_INTS];
  __shared__ uint64_t keys[32];
  __shared__ __align__(16) uint16_t slot[TRGT_KIDX_SLOTS];
  __shared__ int cand[TRGT_CAND_CAP + 4];
  const WarpGroup g;
  const int lane = g.lane();
  const uint32_t n = *n_work_ptr;
#if TRGT_WIDE_DYNAMIC
  // pairs differ widely in cost (budgets 24 and 36, bands of any width): the warps draw them from a counter
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&ctr->wide_next, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) break;
#else
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
#endif
    const uint32_t id = work[i];
    const WfaProb pr = wfa_prob_of(src, id);
    __syncwarp();
    // the piece's 8-mer table as k_flank_exact_t left it (seed filter by probes instead of a scan of every
    // text position); pieces it does not cover are scanned
    const bool indexed = kidx_in != nullptr && pr.P >= 16 && pr.P <= TRGT_KIDX_MAX_P;
    if (indexed) {
      const uint4 *s4 = (const uint4 *)(kidx_in + ((size_t)src.read_locus[id >> 1] * 2 + (id & 1u)) * TRGT_KIDX_SLOTS);
      for (int c = lane; c < (int)(TRGT_KIDX_SLOTS * sizeof(uint16_t) / 16); c += 32) ((uint4 *)slot)[c] = s4[c];
      __syncwarp();
    }
    const KmerIndex idx{slot};
    int settled = 0;
    FlankHit fh;
    fh.via = 0; fh.matches = 0; fh.score = 0; fh.start = 0; fh.end = 0;
#pragma unroll 1
    for (int b = 0; b < 3 && !settled; b++) {
      const int S = b == 0 ? FLW_S0 : b == 1 ? FLW_S1 : 36;
      if (b == 1 && FLW_S1 <= FLW_S0) continue;
      settled = flank_locate_banded(g, pr, S, min_flank_id_frac, keys, ws, FLW_WS_INTS, &fh, indexed ? &idx : nullptr,
                                    cand) == 0;
    }
    if (lane == 0 && settled) {
      trgt_flank_hit_t h;
      h.via = fh.via; h.matches = fh.matches; h.score = fh.score;
      h.start = (uint32_t)fh.start; h.end = (uint32_t)fh.end;
      hits[id] = h;
      work[i] = 0xFFFFFFFFu;
      atomicAdd(&ctr->n_banded, 1u);
    }
    __syncwarp();
  }
}

// The repeat sequences read[span.start .. span.end) of the spanning reads (tr.rs:58-62), packed back to
// back so that the host genotyper streams ~30 bytes per read instead of touching every ~1 KB read.
__global__ void k_tr_len(const trgt_span_t *__restrict__ spans, uint32_t n_reads, uint32_t *__restrict__ len) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_reads; r += gsz)
    len[r] = (r < n_reads && spans[r].found) ? spans[r].end - spans[r].start : 0u;
}

__global__ void __launch_bounds__(256)
k_tr_gather(const uint8_t *__restrict__ reads, const uint64_t *__restrict__ read_off,
            const trgt_span_t *__restrict__ spans, const unsigned long long *__restrict__ off, uint32_t n_reads,
            uint8_t *__restrict__ out) {
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t r = blockIdx.x * wpb + warp; r < n_reads; r += gridDim.x * wpb) {
    const unsigned long long o = off[r];
    const uint32_t n = (uint32_t)(off[r + 1] - o);
    if (n == 0) continue;
    const uint8_t *src = reads + read_off[r] + spans[r].start;
    for (uint32_t i = lane; i < n; i += 32u) out[o + i] = src[i];
  }
}

// ------------------------------------------------------------------ WFA pass 1 -------------

// Persistent groups (a CTA when BLOCK, else a warp) pull items; ring on chip when it fits.
//   work == nullptr: items are 0..n_direct-1, else work[0..*n_work)
//   ends[i]: result of the i-th item (position in the work list)
#define E2E_NARROW_COST 16    // cost cap of the one-pass path for short consensus pairs
#define E2E_NARROW_INTS 1280  // its history: 6*17 header + 3 * 23 diagonals * 17 scores
#define E2E_NARROW_WORDS 48   // its CIGAR: at most 2 * cost + a few words
#define E2E_MID_COST 256      // cost cap of the banded on-chip ring for long similar pairs (halved until the ring fits)

// Phase B, first steps for short repeat sequences: ONE LANE PER (backbone, member) PAIR.  A member equal
// to its backbone (92 % on HiFi) is a single '=' run (k_e2e_identity, which compacts the others into a
// list); a lane of k_e2e_lane then runs the end-to-end alignment of one listed member itself with cost
// cap E2T_COST and writes the CIGAR straight from its back-trace.  Costlier pairs are appended to `resid`
// for the warp kernel.
#define E2T_COST 8
#define E2T_WORDS 32

// step 1: identity test, one lane per member; the ones that differ are compacted into `diff`
__global__ void __launch_bounds__(256)
k_e2e_identity(WfaSrc src, uint32_t n, WfaEnd *__restrict__ ends, uint32_t *__restrict__ cig_n,
               uint32_t *__restrict__ diff, Counters *ctr) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gsz) {
    const uint32_t id = base + lane;
    bool differs = false;
    if (id < n) {
      const WfaProb pr = wfa_prob_of(src, id);
      if (pr.P == pr.T && (pr.P == 0 || wfa_match_len(pr.p, pr.t, pr.P) == pr.P)) {
        WfaEnd end;
        end.status = TRGT_WFA_OK; end.s = 0; end.k = 0; end.off = pr.T;
        ends[id] = end;
        cig_n[id] = pr.P > 0 ? 1u : 0u;  // a single '=' run
      } else {
        differs = true;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, differs);
    if (bal) {
      unsigned int slot = 0;
      if (lane == 0) slot = atomicAdd(&ctr->n_diff, (unsigned int)__popc(bal));
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (differs) diff[slot + __popc(bal & ((1u << lane) - 1u))] = id;
    }
  }
}

// step 2: the members that differ, ONE LANE PER MEMBER on the first flank tier's machinery (e2e_narrow_lane: cost
// cap E2T_COST, so nothing outside |k| <= (E2T_COST - o) / e is reachable; 16-bit history rows for the scores the
// scoring allows, in shared memory with the lanes of a warp interleaved; all lanes walk the same cells and long
// extensions are parked until the warp has only those left).  Whole warps walk the list; what a lane cannot settle
// (cost above the cap, band wider than E2L_WMAX, CIGAR pool full) goes to `resid` for the warp kernel.
#ifndef E2L_THREADS
#define E2L_THREADS 128
#endif
__global__ void __launch_bounds__(E2L_THREADS)
k_e2e_lane(WfaSrc src, const uint32_t *__restrict__ diff, const unsigned int *n_diff_ptr, WfaEnd *__restrict__ ends,
           uint32_t *__restrict__ cig_n, unsigned long long *__restrict__ cig_off, uint32_t *__restrict__ pool,
           unsigned long long pool_cap, uint32_t *__restrict__ resid, Counters *ctr, int rows) {
  extern __shared__ __align__(16) unsigned char e2l_raw[];
  const int tid = threadIdx.x;
  int16_t *hist = reinterpret_cast<int16_t *>(e2l_raw) + (size_t)(tid >> 5) * rows * 3 * E2L_WMAX * 32 + (tid & 31);
  unsigned live_g = 0;
  const unsigned live_m = ft1_live_scores(src.x, src.oe, src.e, E2T_COST, &live_g);
  const uint32_t n = *n_diff_ptr;
  for (uint32_t base = blockIdx.x * blockDim.x + (uint32_t)(tid & ~31); base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + (uint32_t)(tid & 31);
    const bool have = i < n;
    uint32_t id = 0;
    WfaProb pr;
    pr.p = nullptr; pr.t = nullptr; pr.P = 0; pr.T = 0; pr.x = src.x; pr.oe = src.oe; pr.e = src.e;
    pr.pbf = pr.pef = pr.tbf = pr.tef = 0; pr.blo = 0; pr.bhi = 0;
    if (have) {
      id = diff[i];
      pr = wfa_prob_of(src, id);
    }
    uint32_t wbuf[E2T_WORDS];
    WfaCigarSink sink(wbuf, E2T_WORDS);
    WfaEnd end;
    e2e_narrow_lane<32>(pr, E2T_COST, hist, 0xffffffffu, !have, live_m, live_g, &end, sink);
    __syncwarp();
    if (!have) continue;
    bool done = false;
    if (end.status == TRGT_WFA_OK) {
      const uint32_t nw = sink.finish();
      if (!sink.overflow) {
        unsigned long long off = 0;
        bool ok = true;
        if (nw) {
          off = atomicAdd(&ctr->pool_used, (unsigned long long)nw);
          if (off + nw > pool_cap) ok = false;  // pool full: the warp kernel's generic path takes the pair
        }
        if (ok) {
          for (uint32_t w = 0; w < nw; w++) pool[off + w] = wbuf[w];
          cig_off[id] = off;
          cig_n[id] = nw;
          ends[id] = end;
          done = true;
        }
      }
    }
    if (!done) resid[atomicAdd(&ctr->n_resid, 1u)] = id;
  }
}

template <bool BLOCK>
__global__ void k_wfa_score(WfaSrc src, const uint32_t *__restrict__ work, const unsigned int *n_work_ptr,
                            uint32_t n_direct, WfaEnd *__restrict__ ends, int *gring, size_t gring_stride,
                            int smem_ring_ints, uint32_t *__restrict__ trace_work, uint32_t *__restrict__ cig_n,
                            Counters *ctr, uint32_t *__restrict__ pool = nullptr, unsigned long long pool_cap = 0,
                            unsigned long long *__restrict__ cig_off = nullptr) {
  extern __shared__ int smem_i[];
  const uint32_t n = work ? *n_work_ptr : n_direct;
  uint32_t slot, n_slots;
  int *my_smem;
  int lane0;
  if (BLOCK) {
    slot = blockIdx.x; n_slots = gridDim.x;
    my_smem = smem_i + 40;  // first 40 ints: BlockGroup scratch
    lane0 = threadIdx.x == 0;
  } else {
    const uint32_t wib = threadIdx.x >> 5;
    slot = blockIdx.x * (blockDim.x >> 5) + wib; n_slots = gridDim.x * (blockDim.x >> 5);
    my_smem = smem_i + (size_t)wib * smem_ring_ints;
    lane0 = (threadIdx.x & 31u) == 0;
  }
  // warp variant, end-to-end mode: lanes first test 32 pairs for identity (most repeat sequences of a
  // haplotype equal their backbone), then the warp aligns the remaining ones one at a time
  if (!BLOCK && src.mode == WFA_MODE_E2E) {
    const uint32_t lane = threadIdx.x & 31u;
    // direct mode: 32 members per warp and round (identity test by lane, then the warp aligns the ones that
    // differ one at a time); list mode (members already known to differ): one member per warp and round, so
    // that a short list spreads over all warps instead of queueing 32 deep behind a few
    const uint32_t G = work ? 1u : 32u;
    for (uint32_t base = slot * G; base < n; base += n_slots * G) {
      const uint32_t i = base + lane;
      bool pending = false;
      if (lane < G && i < n) {
        const uint32_t id = work ? work[i] : i;
        const WfaProb pr = wfa_prob_of(src, id);
        if (pr.P == pr.T && (pr.P == 0 || wfa_match_len(pr.p, pr.t, pr.P) == pr.P)) {
          WfaEnd end;
          end.status = TRGT_WFA_OK; end.s = 0; end.k = 0; end.off = pr.T;
          ends[id] = end;  // e2e mode: results are indexed by sequence
          cig_n[id] = pr.P > 0 ? 1u : 0u;  // a single '=' run
        } else {
          pending = true;
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, pending);
      while (todo) {
        const uint32_t src_lane = (uint32_t)__ffs((int)todo) - 1u;
        todo &= todo - 1u;
        const uint32_t ii = base + src_lane;
        const uint32_t id = work ? work[ii] : ii;
        const WfaProb pr = wfa_prob_of(src, id);
        const WarpGroup g;
        WfaEnd end;
        end.status = TRGT_WFA_OOM; end.s = 0; end.k = 0; end.off = 0;
        // short, similar pair: one narrow-band pass with history, CIGAR straight from its back-trace
        bool done = false;
        if (smem_ring_ints >= E2E_NARROW_INTS + E2E_NARROW_WORDS && pool != nullptr) {
          end = wfa_e2e_narrow(g, pr, E2E_NARROW_COST, my_smem, E2E_NARROW_INTS);
          __syncwarp();
          if (end.status == TRGT_WFA_OK) {
            if (lane0) {
              uint32_t *wbuf = (uint32_t *)(my_smem + E2E_NARROW_INTS);
              WfaCigarSink sink(wbuf, E2E_NARROW_WORDS);
              wfa_backtrace(pr, end.s, end.k, end.off, my_smem, sink);
              const uint32_t nw = sink.finish();
              unsigned long long off = 0;
              bool ok = !sink.overflow;
              if (ok && nw) {
                off = atomicAdd(&ctr->pool_used, (unsigned long long)nw);
                if (off + nw > pool_cap) ok = false;  // pool full: the generic path takes this pair
              }
              if (ok) {
                for (uint32_t w = 0; w < nw; w++) pool[off + w] = wbuf[w];
                cig_off[id] = off;
                cig_n[id] = nw;
                ends[id] = end;
              }
              my_smem[0] = ok ? 1 : 0;
            }
            __syncwarp();
            done = my_smem[0] != 0;
            __syncwarp();
          }
        }
        if (done) continue;
        const size_t need = wfa_ring_ints(pr);
        int *ring = (need <= (size_t)smem_ring_ints) ? my_smem : (gring ? gring + (size_t)slot * gring_stride : nullptr);
        // long, similar pair (alleles of kilobases a few hundred edits apart): with cost cap S nothing is reachable
        // outside |k| <= (S - o) / e (wfa_e2e_narrow's argument), so a ring over that band -- on chip -- is the full
        // computation; only a costlier pair needs the ring over all P + T + 1 diagonals in HBM
        bool scored = false;
        if (ring != my_smem) {
          const int o = pr.oe - pr.e;
          int S = E2E_MID_COST;
          WfaProb bp = pr;
          for (;;) {
            const int R = S > o ? (S - o) / pr.e : 0;
            bp.blo = wfa_imax(-pr.P, -R);
            bp.bhi = wfa_imin(pr.T, R);
            if (wfa_ring_ints(bp) <= (size_t)smem_ring_ints || S <= E2E_NARROW_COST) break;
            S >>= 1;
          }
          if (S > E2E_NARROW_COST && wfa_ring_ints(bp) <= (size_t)smem_ring_ints) {
            end = wfa_score_ring(g, bp, my_smem, S);
            __syncwarp();
            scored = end.status == TRGT_WFA_OK;
          }
        }
        if (scored) {
        } else if (ring == nullptr || (ring != my_smem && need > gring_stride)) {
          end.status = TRGT_WFA_OOM; end.s = 0; end.k = 0; end.off = 0;
        } else {
          end = wfa_score_ring(g, pr, ring, wfa_score_cap(pr));
          __syncwarp();
        }
        if (lane0) {
          ends[id] = end;
          if (end.status != TRGT_WFA_OK) {
            atomicAdd(&ctr->n_failed, 1u);
            cig_n[id] = 0;
          } else if (end.s == 0) {
            cig_n[id] = pr.P > 0 ? 1u : 0u;
          } else {
            const unsigned int t = atomicAdd(&ctr->n_trace, 1u);
            trace_work[t] = id;
            const unsigned long long wb = 2ull * (unsigned long long)end.s + 8ull;
            atomicAdd(&ctr->words_bound, wb);
            atomicMax(&ctr->max_trace_ints, (unsigned long long)wfa_trace_ints(pr, end.s) + wb);
          }
        }
      }
    }
    return;
  }
  for (uint32_t i = slot; i < n; i += n_slots) {
    const uint32_t id = work ? work[i] : i;
    if (id == 0xFFFFFFFFu) {  // settled by k_flank_band_wide
      if (lane0) { WfaEnd done; done.status = 1; done.s = 0; done.k = 0; done.off = 0; ends[i] = done; }
      continue;
    }
    const WfaProb pr = wfa_prob_of(src, id);
    const size_t need = wfa_ring_ints(pr);
    int *ring = (need <= (size_t)smem_ring_ints) ? my_smem : (gring ? gring + (size_t)slot * gring_stride : nullptr);
    WfaEnd end;
    if (ring == nullptr || (ring != my_smem && need > gring_stride)) {
      end.status = TRGT_WFA_OOM; end.s = 0; end.k = 0; end.off = 0;
    } else if (BLOCK) {
      const BlockGroup g(smem_i);
      end = wfa_score_ring(g, pr, ring, wfa_score_cap(pr));
      __syncthreads();
    } else {
      const WarpGroup g;
      end = wfa_score_ring(g, pr, ring, wfa_score_cap(pr));
      __syncwarp();
    }
    if (lane0) {
      ends[i] = end;
      if (end.status != TRGT_WFA_OK) {
        atomicAdd(&ctr->n_failed, 1u);
        if (cig_n) cig_n[id] = 0;
      } else if (src.mode == WFA_MODE_FLANK) {
        atomicMax(&ctr->max_trace_ints, (unsigned long long)wfa_trace_ints(pr, end.s));
      } else {
        if (end.s == 0) {
          cig_n[id] = pr.P > 0 ? 1u : 0u;  // a single '=' run
        } else {
          const unsigned int t = atomicAdd(&ctr->n_trace, 1u);
          trace_work[t] = id;
          const unsigned long long wb = 2ull * (unsigned long long)end.s + 8ull;
          atomicAdd(&ctr->words_bound, wb);
          atomicMax(&ctr->max_trace_ints, (unsigned long long)wfa_trace_ints(pr, end.s) + wb);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ WFA pass 2 -------------

// One warp per item.  Workspace: shared memory when the cone fits, else the warp's global slot.
//   flank mode: work[i] = 2*read+side, ends[i]           -> hits[work[i]]
//   e2e mode:   work[i] = sequence id, ends[work[i]]     -> CIGAR words in `pool`, cig_off/cig_n[id]
__global__ void __launch_bounds__(128)
k_wfa_trace(WfaSrc src, const uint32_t *__restrict__ work, const unsigned int *n_work_ptr,
            const WfaEnd *__restrict__ ends, int *gws, size_t gws_stride, int smem_ws_ints,
            double min_flank_id_frac, trgt_flank_hit_t *__restrict__ hits, uint32_t *__restrict__ pool,
            unsigned long long pool_cap, unsigned long long *__restrict__ cig_off, uint32_t *__restrict__ cig_n,
            int32_t *__restrict__ status, Counters *ctr) {
  extern __shared__ int smem_i[];
  const WarpGroup g;
  const uint32_t n = *n_work_ptr;
  const uint32_t wib = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + wib, n_slots = gridDim.x * (blockDim.x >> 5);
  int *my_smem = smem_i + (size_t)wib * smem_ws_ints;
  for (uint32_t i = slot; i < n; i += n_slots) {
    const uint32_t id = work[i];
    if (id == 0xFFFFFFFFu) continue;  // settled by k_flank_band_wide
    const WfaProb pr = wfa_prob_of(src, id);
    const WfaEnd end = (src.mode == WFA_MODE_FLANK) ? ends[i] : ends[id];
    if (end.status != TRGT_WFA_OK) {
      if (g.lane() == 0) {
        if (src.mode == WFA_MODE_FLANK) {
          trgt_flank_hit_t h;
          h.via = TRGT_VIA_NONE; h.matches = 0; h.score = INT_MIN; h.start = 0; h.end = 0;
          hits[id] = h;
        } else {
          status[id] = end.status;
        }
      }
      continue;
    }
    const size_t hist = wfa_trace_ints(pr, end.s);
    const size_t words_cap = (src.mode == WFA_MODE_E2E) ? 2 * (size_t)end.s + 8 : 0;
    const size_t need = hist + words_cap;
    int *ws = (need <= (size_t)smem_ws_ints) ? my_smem : ((gws && need <= gws_stride) ? gws + (size_t)slot * gws_stride : nullptr);
    int rc = ws ? wfa_trace_forward(g, pr, end.s, end.k, ws, hist) : TRGT_WFA_OOM;
    __syncwarp();
    if (g.lane() == 0) {
      if (src.mode == WFA_MODE_FLANK) {
        trgt_flank_hit_t h;
        h.via = TRGT_VIA_NONE; h.matches = 0; h.score = INT_MIN; h.start = 0; h.end = 0;
        if (rc == 0) {
          WfaFlankSink sink(pr.T);
          wfa_backtrace(pr, end.s, end.k, end.off, ws, sink);
          h.matches = sink.matches;
          h.score = -end.s;
          // span_locater.rs:19-25, threshold :46
          if ((double)sink.matches >= (double)pr.P * min_flank_id_frac) {
            h.via = TRGT_VIA_WFA; h.start = (uint32_t)sink.ystart(); h.end = (uint32_t)sink.yend();
          } else {
            h.via = TRGT_VIA_WFA_REJECTED;
          }
        } else {
          atomicAdd(&ctr->n_failed, 1u);
        }
        hits[id] = h;
      } else {
        uint32_t nw = 0;
        if (rc == 0) {
          WfaCigarSink sink((uint32_t *)(ws + hist), (uint32_t)words_cap);
          wfa_backtrace(pr, end.s, end.k, end.off, ws, sink);
          nw = sink.finish();
          if (sink.overflow) { rc = TRGT_WFA_OOM; nw = 0; }
        }
        unsigned long long off = 0;
        if (nw) {
          off = atomicAdd(&ctr->pool_used, (unsigned long long)nw);
          if (off + nw > pool_cap) { rc = TRGT_WFA_OOM; nw = 0; }
        }
        const uint32_t *wsrc = (const uint32_t *)(ws + hist);
        for (uint32_t w = 0; w < nw; w++) pool[off + w] = wsrc[w];
        cig_off[id] = off;
        cig_n[id] = nw;
        status[id] = rc;
        if (rc != 0) atomicAdd(&ctr->n_failed, 1u);
      }
    }
    __syncwarp();
  }
}

// find_tr_spans combine rule, span_locater.rs:53-67
__global__ void k_flank_combine(const trgt_flank_hit_t *__restrict__ hits, uint32_t n_reads,
                                trgt_span_t *__restrict__ spans) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gsz) {
    const trgt_flank_hit_t a = hits[2 * r], b = hits[2 * r + 1];
    const bool fa = a.via == TRGT_VIA_EXACT || a.via == TRGT_VIA_WFA;
    const bool fb = b.via == TRGT_VIA_EXACT || b.via == TRGT_VIA_WFA;
    trgt_span_t s;
    s.found = 0; s.start = 0; s.end = 0;
    if (fa && fb && a.end <= b.start) { s.found = 1; s.start = a.end; s.end = b.start; }
    spans[r] = s;
  }
}

// CSR order gather of the CIGAR words: offsets = exclusive scan of cig_n
__global__ void k_cigar_gather(WfaSrc src, uint32_t n_seqs, const WfaEnd *__restrict__ ends,
                               const uint32_t *__restrict__ pool, const unsigned long long *__restrict__ cig_off,
                               const uint32_t *__restrict__ cig_n, const unsigned long long *__restrict__ out_off,
                               uint32_t *__restrict__ out_words, int32_t *__restrict__ scores,
                               int32_t *__restrict__ status) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n_seqs; id += gsz) {
    const WfaEnd end = ends[id];
    const uint32_t n = cig_n[id];
    const unsigned long long o = out_off[id];
    if (end.status != TRGT_WFA_OK) {
      scores[id] = INT_MIN;  // wfaligner.rs:1448: failed alignments report i32::MIN
      status[id] = end.status;
      continue;
    }
    scores[id] = -end.s;
    if (end.s == 0) {
      status[id] = 0;
      if (n) {
        const uint32_t P = (uint32_t)(src.bb_off[src.seq_group[id] + 1] - src.bb_off[src.seq_group[id]]);
        out_words[o] = (P << 4) | 7u;
      }
    } else {
      const unsigned long long po = cig_off[id];
      for (uint32_t w = 0; w < n; w++) out_words[o + w] = pool[po + w];
    }
  }
}

// ------------------------------------------------------------------ consensus vote ----------

// repair_consensus for many groups, one warp per group (consensus_core.h).  WRITE = false: counting
// pass (lens[g] = consensus length, status[g]); WRITE = true: bytes at data + off[g].
//   counts: per-warp slot of 6 * B_max ints; recs: one record slot per CIGAR word (indexed by the
//   group's first word); align_status: per-sequence status of the alignment pass.
template <bool WRITE, class G>
__device__ __forceinline__ void consensus_vote_groups(const G &g, uint32_t slot, uint32_t n_slots, int *sh, const WfaSrc &src,
                                                      const uint32_t *__restrict__ group_off, uint32_t n_groups,
                                                      const uint32_t *__restrict__ words,
                                                      const unsigned long long *__restrict__ word_off,
                                                      const int32_t *__restrict__ align_status, int *counts,
                                                      size_t counts_stride, ConsRec *recs, uint32_t *__restrict__ lens,
                                                      int32_t *__restrict__ status,
                                                      const unsigned long long *__restrict__ off, uint8_t *__restrict__ data) {
  for (uint32_t gi = slot; gi < n_groups; gi += n_slots) {
    ConsGroup gr;
    gr.B = (int)(src.bb_off[gi + 1] - src.bb_off[gi]);
    gr.s0 = group_off[gi];
    gr.n = group_off[gi + 1] - gr.s0;
    gr.seqs = src.seqs; gr.seq_off = src.seq_off; gr.words = words; gr.word_off = word_off;
    int st = 0;
    for (uint32_t m = (uint32_t)g.lane(); m < gr.n; m += (uint32_t)g.size())
      if (align_status[gr.s0 + m] != 0) st = align_status[gr.s0 + m];
    st = g.min_i(st);  // statuses are <= 0
    if (WRITE && status[gi] != 0) continue;
    long long len = 0;
    if (st == 0) {
      const uint32_t rec_cap = (uint32_t)(word_off[gr.s0 + gr.n] - word_off[gr.s0]) + 1u;
      len = consensus_vote(g, gr, counts + (size_t)slot * counts_stride, recs + word_off[gr.s0] + gi, rec_cap, sh,
                           WRITE ? data + off[gi] : nullptr);
    }
    if (!WRITE && g.lane() == 0) {
      status[gi] = st != 0 ? st : (len == -1 ? TRGT_ITEM_INVALID_BASE : (len < 0 ? TRGT_ITEM_OOM : 0));
      lens[gi] = (st == 0 && len > 0) ? (uint32_t)len : 0u;
    }
    g.sync();
  }
}

// BLOCK = false: a warp per group (groups of a few dozen short repeat sequences: a genome-wide catalog);
// BLOCK = true: a CTA per group (alleles of kilobases: the members' CIGAR walks and the columns spread over four
// warps, and 2 048 groups fill the device instead of a fifth of it).
template <bool WRITE, bool BLOCK>
__global__ void __launch_bounds__(128)
k_consensus_vote(WfaSrc src, const uint32_t *__restrict__ group_off, uint32_t n_groups,
                 const uint32_t *__restrict__ words, const unsigned long long *__restrict__ word_off,
                 const int32_t *__restrict__ align_status, int *counts, size_t counts_stride, ConsRec *recs,
                 uint32_t *__restrict__ lens, int32_t *__restrict__ status, const unsigned long long *__restrict__ off,
                 uint8_t *__restrict__ data) {
  __shared__ int shared[4][2];
  __shared__ int scratch[40];
  if (BLOCK) {
    const BlockGroup g(scratch);
    consensus_vote_groups<WRITE>(g, blockIdx.x, gridDim.x, shared[0], src, group_off, n_groups, words, word_off,
                                 align_status, counts, counts_stride, recs, lens, status, off, data);
  } else {
    const WarpGroup g;
    const uint32_t wib = threadIdx.x >> 5;
    consensus_vote_groups<WRITE>(g, blockIdx.x * (blockDim.x >> 5) + wib, gridDim.x * (blockDim.x >> 5), shared[wib], src,
                                 group_off, n_groups, words, word_off, align_status, counts, counts_stride, recs, lens,
                                 status, off, data);
  }
}

// ------------------------------------------------------------------ edit distance ---------

// get_dist_matrix: one CTA per locus, threads stride the condensed pair index.
// Sequence s of the batch is seqs[seq_off[s] .. seq_off[s+1]) -- or, with `index` and `spans` (the repeat
// sequences of a resident flank batch, read in place), bytes [span.start, span.end) of read index[s].
__global__ void __launch_bounds__(128)
k_edit_dist(const uint8_t *__restrict__ seqs, const uint64_t *__restrict__ seq_off,
            const uint32_t *__restrict__ locus_seq_off, const unsigned long long *__restrict__ pair_off,
            uint32_t n_loci, double *__restrict__ dists, const uint32_t *__restrict__ index,
            const trgt_span_t *__restrict__ spans) {
  for (uint32_t l = blockIdx.x; l < n_loci; l += gridDim.x) {
    const uint32_t s0 = locus_seq_off[l];
    const uint32_t n = locus_seq_off[l + 1] - s0;
    if (n < 2) continue;
    const unsigned long long base = pair_off[l];
    const unsigned long long np = (unsigned long long)n * (n - 1) / 2;
    for (unsigned long long q = threadIdx.x; q < np; q += blockDim.x) {
      // invert q = i*n - i(i+1)/2 + (j-i-1)
      const double nn = 2.0 * (double)n - 1.0;
      long long i = (long long)floor((nn - sqrt(nn * nn - 8.0 * (double)q)) / 2.0);
      if (i < 0) i = 0;
      if (i > (long long)n - 2) i = (long long)n - 2;
      while (i > 0 && (unsigned long long)i * n - (unsigned long long)i * (i + 1) / 2 > q) i--;
      while ((unsigned long long)(i + 1) * n - (unsigned long long)(i + 1) * (i + 2) / 2 <= q) i++;
      const unsigned long long row0 = (unsigned long long)i * n - (unsigned long long)i * (i + 1) / 2;
      const uint32_t j = (uint32_t)(i + 1 + (long long)(q - row0));
      const uint8_t *a, *b;
      int la, lb;
      if (index) {
        const uint32_t ra = index[s0 + i], rb = index[s0 + j];
        a = seqs + seq_off[ra] + spans[ra].start; la = (int)(spans[ra].end - spans[ra].start);
        b = seqs + seq_off[rb] + spans[rb].start; lb = (int)(spans[rb].end - spans[rb].start);
      } else {
        a = seqs + seq_off[s0 + i]; la = (int)(seq_off[s0 + i + 1] - seq_off[s0 + i]);
        b = seqs + seq_off[s0 + j]; lb = (int)(seq_off[s0 + j + 1] - seq_off[s0 + j]);
      }
      int d;
      if ((unsigned long long)la * (unsigned long long)lb > 10000ull) {  // MAX_OPS, genotype_cluster.rs:237-243
        d = la > lb ? la - lb : lb - la;
      } else {
        d = edit_distance_128(a, la, b, lb);
      }
      dists[base + q] = sqrt((double)d);
    }
  }
}

// cluster() + the choice of the two largest groups + central_read (genotype_cluster.rs:12-39,57-72,154-227) on the
// distance matrices k_edit_dist left in HBM: one warp per locus (cluster_core.h), its scratch in a global slot.
__global__ void __launch_bounds__(128)
k_cluster_ward(const uint32_t *__restrict__ locus_seq_off, const unsigned long long *__restrict__ pair_off,
               uint32_t n_loci, double *__restrict__ dists, unsigned char *__restrict__ ws, size_t ws_stride,
               int32_t *__restrict__ group_out, uint32_t *__restrict__ central_out, uint32_t *__restrict__ n_groups_out) {
  const WarpGroup g;
  const uint32_t wib = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + wib, n_slots = gridDim.x * (blockDim.x >> 5);
  for (uint32_t l = slot; l < n_loci; l += n_slots) {
    const uint32_t s0 = locus_seq_off[l];
    const uint32_t n = locus_seq_off[l + 1] - s0;
    const ClusterWs w = cl_carve(ws + (size_t)slot * ws_stride, n);
    const int ng = cl_cluster_locus(g, dists + pair_off[l], n, w, group_out + s0, central_out + 2 * (size_t)l);
    if (g.lane() == 0 && n_groups_out) n_groups_out[l] = (uint32_t)ng;
    __syncwarp();
  }
}

// The repeat sequences of selected reads of a flank batch as one CSR set (lengths first, then the bytes): what
// the align / consensus kernels take, built without a trip through the host.
__global__ void k_trs_len(const trgt_span_t *__restrict__ spans, const uint32_t *__restrict__ index, uint32_t n,
                          uint32_t *__restrict__ len) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gsz) {
    uint32_t v = 0;
    if (i < n) { const trgt_span_t s = spans[index[i]]; v = s.found ? s.end - s.start : 0u; }
    len[i] = v;
  }
}

// lengths of the repeat sequences of the reads index[0..n) (0 without a span), len[n] = 0, and their maximum
__global__ void __launch_bounds__(256)
k_trs_len(const trgt_span_t *__restrict__ spans, const uint32_t *__restrict__ index, uint32_t n,
          uint32_t *__restrict__ len, unsigned int *__restrict__ max_out) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  unsigned int mx = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gsz) {
    uint32_t l = 0;
    if (i < n) {
      const trgt_span_t sp = spans[index[i]];
      l = sp.found ? sp.end - sp.start : 0u;
    }
    len[i] = l;
    mx = l > mx ? l : mx;
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31u) == 0 && mx) atomicMax(max_out, mx);
}

__global__ void __launch_bounds__(256)
k_trs_gather(const uint8_t *__restrict__ reads, const uint64_t *__restrict__ read_off,
             const trgt_span_t *__restrict__ spans, const uint32_t *__restrict__ index,
             const unsigned long long *__restrict__ off, uint32_t n, uint8_t *__restrict__ out) {
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
    const unsigned long long o = off[i];
    const uint32_t len = (uint32_t)(off[i + 1] - o);
    if (len == 0) continue;
    const uint32_t r = index[i];
    const uint8_t *src = reads + read_off[r] + spans[r].start;
    for (uint32_t k = lane; k < len; k += 32u) out[o + k] = src[k];
  }
}

// ------------------------------------------------------------------ phase C: HMM ------------

struct HmmBatch {
  const uint8_t *motifs; const uint64_t *motif_off;   // all motifs
  const uint32_t *locus_motif_off;                     // [n_loci+1]
  const uint8_t *alleles; const uint64_t *allele_off; // [n_alleles+1]
  const uint32_t *allele_locus;
  const uint32_t *list;                                // alleles these kernels take (nullptr: every allele); the
                                                       // kernels' ranges and bp_off / scr_off index this list
  const unsigned long long *bp_off;                    // [n_list+1] into bp (relative to wave base)
  const unsigned long long *scr_off;                   // [n_list+1] into the span scratch (one slot per base)
  const unsigned long long *mc_off;                    // [n_alleles+1] into mc
  const uint32_t *mm_off; const double *mm_lp;         // jump-in ln table
  HmmConsts c;
  int S_max, nb_max, mbytes_max;
  size_t warp_bytes;                                   // hmm_onchip_bytes(S_max, nb_max, mbytes_max)
};

struct HmmWarpMem {
  double *sc0, *sc1;
  uint32_t *moff, *mmoff;
  uint16_t *n, *ms, *stblk;
  uint8_t *bytes;
};

__device__ __forceinline__ HmmWarpMem hmm_carve(unsigned char *base, int S_max, int nb_max) {
  HmmWarpMem w;
  w.sc0 = (double *)base;
  w.sc1 = w.sc0 + S_max;
  w.moff = (uint32_t *)(w.sc1 + S_max);
  w.mmoff = w.moff + nb_max;
  w.n = (uint16_t *)(w.mmoff + nb_max);
  w.ms = w.n + nb_max;
  w.stblk = w.ms + nb_max;
  w.bytes = (uint8_t *)(w.stblk + S_max);
  return w;
}


// One THREAD per allele for small models (every locus of a genome-wide catalog: S = 14..26): 32
// alleles advance per instruction.  Two score columns of HMM_THREAD_S doubles per thread, strided
// over the CTA so that threads of a warp hit different banks.  Larger models are left to
// k_hmm_viterbi.
__global__ void __launch_bounds__(128)
k_hmm_viterbi_thread(HmmBatch hb, uint32_t a0, uint32_t a1, unsigned long long bp_base, uint8_t *__restrict__ bp,
                     int32_t *__restrict__ status, int s_cap) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  double *sc = reinterpret_cast<double *>(smem_b);  // [2][s_cap][128], s_cap = min(HMM_THREAD_S, largest model of the batch)
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t ai = a0 + blockIdx.x * blockDim.x + threadIdx.x; ai < a1; ai += gsz) {
    const uint32_t a = hb.list ? hb.list[ai] : ai;
    const uint32_t l = hb.allele_locus[a];
    const uint32_t m0 = hb.locus_motif_off[l];
    const int nm = (int)(hb.locus_motif_off[l + 1] - m0);
    const uint64_t *moff = hb.motif_off + m0;
    bool bad = false;
    for (int b = 0; b < nm; b++) bad = bad || moff[b + 1] <= moff[b];
    const HmmModelScan model = hmm_model_scan(hb.motifs, moff, nm);
    if (bad) { status[a] = TRGT_ITEM_INVALID_BASE; continue; }
    if (model.S > HMM_THREAD_S) continue;  // k_hmm_viterbi's
    status[a] = 0;
    const int L = (int)(hb.allele_off[a + 1] - hb.allele_off[a]);
    if (L == 0) continue;
    hmm_viterbi_thread(hmm_model_pack(model, hb.mm_off), hb.c, hb.mm_off, hb.mm_lp, hb.alleles + hb.allele_off[a], L, sc + threadIdx.x,
                       sc + (size_t)s_cap * 128 + threadIdx.x, 128, bp + (hb.bp_off[ai] - bp_base));
  }
}

// One warp per allele (models larger than HMM_THREAD_S, or all of them when skip_small is 0): model
// build in shared memory, Viterbi, back-pointers to HBM.
__global__ void __launch_bounds__(128)
k_hmm_viterbi(HmmBatch hb, uint32_t a0, uint32_t a1, unsigned long long bp_base, uint8_t *__restrict__ bp,
              int32_t *__restrict__ status, int skip_small) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  const WarpGroup g;
  const uint32_t wib = threadIdx.x >> 5;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const HmmWarpMem wm = hmm_carve(smem_b + (size_t)wib * hb.warp_bytes, hb.S_max, hb.nb_max);
  for (uint32_t ai = a0 + blockIdx.x * (blockDim.x >> 5) + wib; ai < a1; ai += warps) {
    const uint32_t a = hb.list ? hb.list[ai] : ai;
    const uint32_t l = hb.allele_locus[a];
    const uint32_t m0 = hb.locus_motif_off[l];
    const int nm = (int)(hb.locus_motif_off[l + 1] - m0);
    const uint8_t *allele = hb.alleles + hb.allele_off[a];
    const int L = (int)(hb.allele_off[a + 1] - hb.allele_off[a]);
    if (skip_small) {  // small models were done by k_hmm_viterbi_thread
      int Sq = 7;
      bool bad = false;
      for (int b = 0; b < nm; b++) {
        const long long n = (long long)(hb.motif_off[m0 + b + 1] - hb.motif_off[m0 + b]);
        bad = bad || n <= 0;
        Sq += 3 * (int)n + 1;
      }
      if (bad || Sq <= HMM_THREAD_S) continue;
    }
    HmmModel model;
    const int S = hmm_model_build(g, hb.motifs, hb.motif_off + m0, nm, hb.mm_off, wm.bytes, wm.moff, wm.mmoff,
                                  wm.n, wm.ms, wm.stblk, &model);
    if (g.lane() == 0) status[a] = S < 0 ? TRGT_ITEM_INVALID_BASE : 0;
    if (S >= 0 && L > 0)
      hmm_viterbi(g, model, hb.c, hb.mm_lp, allele, L, wm.sc0, wm.sc1, bp + (hb.bp_off[ai] - bp_base));
    __syncwarp();
  }
}

// One thread per allele: the counting walk over the back-pointers (purity, MC, number of collapsed
// spans, optional path length).  Uses the table-free model (HmmModelScan).
__global__ void __launch_bounds__(128)
k_hmm_walk(HmmBatch hb, uint32_t a0, uint32_t a1, unsigned long long bp_base, const uint8_t *__restrict__ bp,
           uint32_t *__restrict__ mc, double *__restrict__ purity, uint32_t *__restrict__ n_spans,
           unsigned long long *__restrict__ path_len, int32_t *__restrict__ status,
           trgt_motif_span_t *__restrict__ span_scratch) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t ai = a0 + blockIdx.x * blockDim.x + threadIdx.x; ai < a1; ai += gsz) {
    const uint32_t a = hb.list ? hb.list[ai] : ai;
    const uint32_t l = hb.allele_locus[a];
    const uint32_t m0 = hb.locus_motif_off[l];
    const int nm = (int)(hb.locus_motif_off[l + 1] - m0);
    const int L = (int)(hb.allele_off[a + 1] - hb.allele_off[a]);
    uint32_t *my_mc = mc + hb.mc_off[a];
    for (int b = 0; b < nm; b++) my_mc[b] = 0;
    if (status[a] != 0 || L == 0) {
      purity[a] = nan("");  // purity.rs:7-9
      n_spans[a] = 0;
      if (path_len) path_len[a] = 0;
      continue;
    }
    const HmmModelScan model = hmm_model_scan(hb.motifs, hb.motif_off + m0, nm);
    uint64_t plen = 0;
    // span_scratch: one slot per base of the batch's alleles (a collapsed span covers at least one base); the
    // walk fills the allele's slots from the back, so its spans end up in the last n_spans of them, in order
    HmmSpan *sp = span_scratch ? (HmmSpan *)(span_scratch + hb.scr_off[ai]) : nullptr;
    const HmmAnnot an = model.S <= HMM_THREAD_S
                            ? hmm_annotate(hmm_model_pack(model), hb.alleles + hb.allele_off[a], L,
                                           bp + (hb.bp_off[ai] - bp_base), 6, my_mc, sp, (uint32_t)L, nullptr, 0, 0, &plen)
                            : hmm_annotate(model, hb.alleles + hb.allele_off[a], L, bp + (hb.bp_off[ai] - bp_base), 6, my_mc,
                                           sp, (uint32_t)L, nullptr, 0, 0, &plen);
    purity[a] = an.purity;
    n_spans[a] = an.n_spans;
    if (path_len) path_len[a] = plen;
    if (an.status < 0) status[a] = TRGT_ERR_INTERNAL;
  }
}

// spans of the wave's alleles from the walk's scratch slots to their CSR offsets, one thread per allele
__global__ void __launch_bounds__(128)
k_hmm_spans(HmmBatch hb, uint32_t a0, uint32_t a1, const trgt_motif_span_t *__restrict__ span_scratch,
            const uint32_t *__restrict__ n_spans, const unsigned long long *__restrict__ span_off,
            trgt_motif_span_t *__restrict__ spans) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t ai = a0 + blockIdx.x * blockDim.x + threadIdx.x; ai < a1; ai += gsz) {
    const uint32_t a = hb.list ? hb.list[ai] : ai;
    const uint32_t ns = n_spans[a];
    if (ns == 0) continue;
    const uint64_t L = hb.allele_off[a + 1] - hb.allele_off[a];
    const trgt_motif_span_t *src = span_scratch + hb.scr_off[ai] + (L - ns);
    trgt_motif_span_t *dst = spans + span_off[a];
    for (uint32_t i = 0; i < ns; i++) dst[i] = src[i];
  }
}

// One thread per allele: the second walk writes the collapsed spans (and optionally the state path)
// at their CSR offsets.
__global__ void __launch_bounds__(128)
k_hmm_emit(HmmBatch hb, uint32_t a0, uint32_t a1, unsigned long long bp_base, const uint8_t *__restrict__ bp,
           const uint32_t *__restrict__ n_spans, const unsigned long long *__restrict__ span_off,
           trgt_motif_span_t *__restrict__ spans, const unsigned long long *__restrict__ path_off,
           uint32_t *__restrict__ paths, const int32_t *__restrict__ status) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t ai = a0 + blockIdx.x * blockDim.x + threadIdx.x; ai < a1; ai += gsz) {
    const uint32_t a = hb.list ? hb.list[ai] : ai;
    const int L = (int)(hb.allele_off[a + 1] - hb.allele_off[a]);
    const uint32_t ns = n_spans[a];
    const unsigned long long plen = path_off ? path_off[a + 1] - path_off[a] : 0;
    if (L == 0 || status[a] != 0 || (ns == 0 && plen == 0)) continue;
    const uint32_t l = hb.allele_locus[a];
    const uint32_t m0 = hb.locus_motif_off[l];
    const int nm = (int)(hb.locus_motif_off[l + 1] - m0);
    const HmmModelScan model = hmm_model_scan(hb.motifs, hb.motif_off + m0, nm);
    hmm_annotate(model, hb.alleles + hb.allele_off[a], L, bp + (hb.bp_off[ai] - bp_base), 6, nullptr,
                 (HmmSpan *)(spans + span_off[a]), ns, plen ? paths + path_off[a] : nullptr, plen, plen, nullptr);
  }
}

// ---- single-motif loci (every locus of a genome-wide catalog): one lane per allele, score column in registers,
// ---- one packed back-pointer word per column (hmm_core.h, hmm_viterbi_lane) -------------------------------

// The engine sorts these alleles by (motif length, allele length) into `slots`; 32 consecutive slots form a
// group = one warp: same motif length (groups never straddle two lengths: the tail of a length is padded with
// empty slots), similar allele lengths, so the lanes of a warp run the same unrolled code for about the same
// number of columns.  Groups are ordered by their longest allele, longest first, so that the long alleles --
// a lane walks its columns one after the other -- start first and the short ones fill in behind them.
// Column c of lane l of group g is word group_off[g] + (c - 1) * 32 + l: a warp writes one 128-byte line per column.
struct HmmLaneBatch {
  const uint8_t *motifs; const uint64_t *motif_off; const uint32_t *locus_motif_off;
  const uint8_t *alleles; const uint64_t *allele_off; const uint32_t *allele_locus;
  const unsigned long long *mc_off;
  const uint32_t *slots;                  // allele of each slot, 0xFFFFFFFF = padding
  const uint8_t *group_n;                 // motif length of each group
  const unsigned long long *group_off;    // [n_groups+1] first word of each group (absolute; waves subtract their base)
  HmmConsts c;
  double jump[HMM_LANE_NMAX + 1][HMM_LANE_NMAX];  // jump[n][i] = ln(seed (n - i)), builder.rs:93-111
};

#define HMM_LANE_SWITCH(n, F) \
  switch (n) {                \
    case 1: F(1); break;      \
    case 2: F(2); break;      \
    case 3: F(3); break;      \
    case 4: F(4); break;      \
    case 5: F(5); break;      \
    case 6: F(6); break;      \
    case 7: F(7); break;      \
    default: break;           \
  }

__global__ void __launch_bounds__(128)
k_hmm_lane_viterbi(HmmLaneBatch lb, uint32_t g0, uint32_t g1, unsigned long long word_base, uint32_t *__restrict__ bp,
                   int32_t *__restrict__ status) {
  const uint32_t g = g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= g1) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t a = lb.slots[(size_t)g * 32 + lane];
  if (a == 0xFFFFFFFFu) return;
  const uint32_t l = lb.allele_locus[a];
  const uint8_t *motif = lb.motifs + lb.motif_off[lb.locus_motif_off[l]];
  const int L = (int)(lb.allele_off[a + 1] - lb.allele_off[a]);
  status[a] = 0;
  uint32_t *my_bp = bp + (lb.group_off[g] - word_base) + lane;
  const uint8_t *allele = lb.alleles + lb.allele_off[a];
#define F(N) hmm_viterbi_lane<N>(lb.c, lb.jump[N], hmm_pack_motif(motif, N), allele, L, my_bp, 32)
  HMM_LANE_SWITCH(lb.group_n[g], F)
#undef F
}

// The reverse walk of the same alleles over the packed words: purity, MC, collapsed spans (left in the group's
// scratch slots, slot j of lane l at group_off[g] + j * 32 + l, filled from the back) and, optionally, the
// length of the state path.  Table driven (hmm_walk_table): the CTA first builds the 32-entry state table of
// every motif length in shared memory (2 KB), from the same arithmetic the generic walk evaluates per step.
__global__ void __launch_bounds__(128)
k_hmm_lane_walk(HmmLaneBatch lb, uint32_t g0, uint32_t g1, unsigned long long word_base, const uint32_t *__restrict__ bp,
                uint32_t *__restrict__ mc, double *__restrict__ purity, uint32_t *__restrict__ n_spans,
                unsigned long long *__restrict__ path_len, int32_t *__restrict__ status,
                trgt_motif_span_t *__restrict__ span_scratch) {
  __shared__ HmmLaneEntry tab[HMM_LANE_NMAX + 1][32];
  for (int q = threadIdx.x; q < (HMM_LANE_NMAX + 1) * 32; q += blockDim.x)
    tab[q >> 5][q & 31] = (q >> 5) >= 1 ? hmm_lane_table_entry(q >> 5, q & 31) : HmmLaneEntry{0u, 0u};
  __syncthreads();
  const uint32_t g = g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= g1) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t a = lb.slots[(size_t)g * 32 + lane];
  if (a == 0xFFFFFFFFu) return;
  const uint32_t l = lb.allele_locus[a];
  const uint8_t *motif = lb.motifs + lb.motif_off[lb.locus_motif_off[l]];
  const int L = (int)(lb.allele_off[a + 1] - lb.allele_off[a]);
  const int n = lb.group_n[g];
  uint32_t *my_mc = mc + lb.mc_off[a];
  my_mc[0] = 0;
  const HmmSpanStrided sp{span_scratch ? (HmmSpan *)(span_scratch + lb.group_off[g] + lane) : nullptr, 32u};
  uint64_t plen = 0;
  const HmmAnnot an = hmm_walk_table(tab[n], n, hmm_pack_motif(motif, n), L,
                                     bp + (lb.group_off[g] - word_base) + lane, 32, 6, my_mc, sp, (uint32_t)L, &plen);
  purity[a] = an.purity;
  n_spans[a] = an.n_spans;
  if (path_len) path_len[a] = plen;
  if (an.status < 0) status[a] = TRGT_ERR_INTERNAL;
}

// spans of the lane alleles from their groups' scratch slots to their CSR offsets
__global__ void __launch_bounds__(128)
k_hmm_lane_spans(HmmLaneBatch lb, uint32_t g0, uint32_t g1, const trgt_motif_span_t *__restrict__ span_scratch,
                 const uint32_t *__restrict__ n_spans, const unsigned long long *__restrict__ span_off,
                 trgt_motif_span_t *__restrict__ spans) {
  const uint32_t g = g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= g1) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t a = lb.slots[(size_t)g * 32 + lane];
  if (a == 0xFFFFFFFFu) return;
  const uint32_t ns = n_spans[a];
  if (ns == 0) return;
  const uint64_t L = lb.allele_off[a + 1] - lb.allele_off[a];
  const trgt_motif_span_t *src = span_scratch + lb.group_off[g] + lane;
  trgt_motif_span_t *dst = spans + span_off[a];
  for (uint32_t i = 0; i < ns; i++) dst[i] = src[(size_t)(L - ns + i) * 32];
}

// second walk, only when state paths (Hmm::label) are asked for: the path at its CSR offset, forward order
__global__ void __launch_bounds__(128)
k_hmm_lane_emit(HmmLaneBatch lb, uint32_t g0, uint32_t g1, unsigned long long word_base, const uint32_t *__restrict__ bp,
                const unsigned long long *__restrict__ path_off, uint32_t *__restrict__ paths,
                const int32_t *__restrict__ status) {
  const uint32_t g = g0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= g1) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t a = lb.slots[(size_t)g * 32 + lane];
  if (a == 0xFFFFFFFFu || status[a] != 0) return;
  const unsigned long long plen = path_off[a + 1] - path_off[a];
  if (plen == 0) return;
  const uint32_t l = lb.allele_locus[a];
  const uint8_t *motif = lb.motifs + lb.motif_off[lb.locus_motif_off[l]];
  const int L = (int)(lb.allele_off[a + 1] - lb.allele_off[a]);
  const uint32_t *my_bp = bp + (lb.group_off[g] - word_base) + lane;
  const uint8_t *allele = lb.alleles + lb.allele_off[a];
#define F(N)                                                                                              \
  {                                                                                                       \
    const HmmModelSingle<N> model{hmm_pack_motif(motif, N)};                                              \
    HmmBpWords<N> words(my_bp, 32, L);                                                                    \
    hmm_annotate_bp(model, allele, L, words, 6, (uint32_t *)nullptr, HmmSpanArray{nullptr}, 0, paths + path_off[a], \
                    plen, plen, (uint64_t *)nullptr);                                                     \
  }
  HMM_LANE_SWITCH(lb.group_n[g], F)
#undef F
}

// ------------------------------------------------------------------ VCF sample fields ---------

// AL, MC, MS, AP of every locus (write_vcf.rs:267-343): one lane per locus.  WRITE = false fills
// len[4*l + field]; WRITE = true stores the bytes at off[4*l + field].
template <bool WRITE>
__global__ void __launch_bounds__(128)
k_vcf_fields(uint32_t n_loci, const uint32_t *__restrict__ locus_allele_off, const uint64_t *__restrict__ allele_off,
             const unsigned long long *__restrict__ mc_off, const uint32_t *__restrict__ mc,
             const unsigned long long *__restrict__ span_off, const trgt_motif_span_t *__restrict__ spans,
             const double *__restrict__ purity, const int32_t *__restrict__ status, uint32_t *__restrict__ len,
             const unsigned long long *__restrict__ off, uint8_t *__restrict__ out) {
  const uint32_t gsz = gridDim.x * blockDim.x;
  for (uint32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < n_loci; l += gsz) {
    VcfLocus L;
    L.a0 = locus_allele_off[l]; L.a1 = locus_allele_off[l + 1];
    L.allele_off = allele_off; L.mc_off = mc_off; L.mc = mc; L.span_off = span_off; L.spans = spans;
    L.purity = purity; L.status = status;
    for (int f = 0; f < 4; f++) {
      VcfWriter w;
      w.out = WRITE ? out + off[4 * (size_t)l + f] : nullptr;
      w.n = 0;
      vcf_encode_field(L, f, w);
      if (!WRITE) len[4 * (size_t)l + f] = (uint32_t)w.n;
    }
  }
}

}  // namespace trgt
