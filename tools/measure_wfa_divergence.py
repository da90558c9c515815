"""How far can the reference's consensus aligner (BiWFA = MemoryUltraLow with WFA2-lib's default heuristic
wfadaptive(10, 50, 1), commands/genotype.rs:82-86) stray from the exact unidirectional WFA this repository (engine
and oracle) implements?  WFA2-lib is not vendored in the reference tree, so neither BiWFA's choice among
co-optimal alignments nor the heuristic can be pinned; this script BOUNDS their effect on the path's outputs with
two restated devices (CPU only, oracle/):
  1. wfadaptive(10, 50) on the unidirectional aligner (oracle/wfa_oracle.c: wfadaptive_cutoff): how many
     (backbone, member) pairs change score or CIGAR, and how many repaired consensuses change;
  2. the opposite tie-break extreme: align the REVERSED sequences and reverse the CIGAR -- an optimal alignment that
     places every ambiguous gap at the other end of its run.  BiWFA returns some optimal alignment between the two
     extremes, so consensuses that agree under both are insensitive to its choice.
Usage: python tools/measure_wfa_divergence.py [--loci N] [--long-loci M]   -> one JSON line"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from harness import workload
from harness.cluster_pass import consensus_groups, spanning_order
from harness.workload import genotype_glue
from oracle import oracle as orc


def words_of(a):
    return a.sam_cigar(True)


def reverse_tiebreak_words(bb, seq):
    a = orc.wfa_align(bb[::-1], seq[::-1], orc.AFFINE, 2, 5, 1)
    w = words_of(a)
    return list(reversed(w)), a.score


def measure(groups, label):
    n_pairs = n_diff_pairs = 0
    heur_score = heur_cigar = heur_fail = 0
    rev_cigar = 0
    n_groups = cons_heur = cons_rev = 0
    for bb, members in groups:
        exact_w, heur_w, rev_w = [], [], []
        for m in members:
            a = orc.wfa_align(bb, m, orc.AFFINE, 2, 5, 1)
            h = orc.wfa_align(bb, m, orc.AFFINE, 2, 5, 1, wfadaptive=(10, 50))
            rw, rs = reverse_tiebreak_words(bb, m)
            assert rs == a.score
            ew = words_of(a)
            hw = words_of(h) if h.status == 0 else ew
            n_pairs += 1
            n_diff_pairs += m != bb
            heur_fail += h.status != 0
            heur_score += h.status == 0 and h.score != a.score
            heur_cigar += hw != ew
            rev_cigar += rw != ew
            exact_w.append(ew); heur_w.append(hw); rev_w.append(rw)
        c0 = orc.repair_consensus(bb, members, exact_w)
        n_groups += 1
        cons_heur += orc.repair_consensus(bb, members, heur_w) != c0
        cons_rev += orc.repair_consensus(bb, members, rev_w) != c0
    return {"workload": label, "pairs": n_pairs, "pairs_member_differs_from_backbone": int(n_diff_pairs),
            "wfadaptive_changes_score": int(heur_score), "wfadaptive_changes_cigar": int(heur_cigar),
            "wfadaptive_fails": int(heur_fail), "reverse_tiebreak_changes_cigar": int(rev_cigar),
            "groups": n_groups, "consensus_changed_by_wfadaptive": int(cons_heur),
            "consensus_changed_by_reverse_tiebreak": int(cons_rev)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=3000)
    ap.add_argument("--long-loci", type=int, default=6)
    args = ap.parse_args()
    out = []
    # config 4 shape: groups as the bench's glue forms them
    w = workload.generate(args.loci, 30)
    spans, _ = orc.flank_batch(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                               n_threads=os.cpu_count() or 1, want_hits=False)
    g = genotype_glue(w, spans)
    groups = []
    for gi in range(len(g.backbones)):
        ms = [g.seqs.get(i) for i in range(int(g.group_seq_off[gi]), int(g.group_seq_off[gi + 1]))]
        groups.append((g.backbones.get(gi), ms))
    out.append(measure(groups, f"config 4 shape: {args.loci} loci x 30 reads, repeat sequences of ~27 bp"))
    # config 5 shape: the two make_consensus groups of the cluster genotyper
    w = workload.generate(args.long_loci, 40, seed=505, tr_len_dist="loguniform", tr_len_min=5000, tr_len_max=50000,
                          het_independent=True)
    spans, _ = orc.flank_batch(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac,
                               n_threads=os.cpu_count() or 1, want_hits=False)
    sel, off = spanning_order(w, spans)

    def tr(r):
        return w.reads.get(int(r))[int(spans[int(r)]["start"]):int(spans[int(r)]["end"])]

    group = np.zeros(sel.size, dtype=np.int32)
    central = np.full((w.n_loci, 2), 0xFFFFFFFF, dtype=np.uint32)
    for l in range(w.n_loci):
        a, b = int(off[l]), int(off[l + 1])
        trs = [tr(r) for r in sel[a:b]]
        s, c, _ = orc.cluster_locus(orc.get_dist_matrix(trs) if b - a >= 2 else [], b - a)
        group[a:b] = s
        central[l] = [0xFFFFFFFF if x is None else x for x in c]
    bb, members, goff, _ = consensus_groups(sel, off, group, central)
    groups = [(tr(bb[k]), [tr(r) for r in members[int(goff[k]):int(goff[k + 1])]]) for k in range(len(bb))]
    out.append(measure(groups, f"config 5 shape: {args.long_loci} loci x 40 reads, alleles of 5-50 kb"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
