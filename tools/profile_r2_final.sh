#!/bin/bash
# Under gpurun (1 GPU): the round's final evidence.  Launch list of one full-size resident pass (config 4), one
# `--set full` capture per kernel of the step (second launch of each: the first is the warm-up), the config-5 kernels,
# and the two clip kernels from their parity tests.  Each report is summarised on the box (text + key metrics) and
# deleted: gpurun only brings 64 MiB back.  Outputs: gpurun_out/r2z_*.
tag=r2z
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_bench.log 2>&1
sum() {  # sum <name> <mangled substring>
  python tools/ncu_report.py gpurun_out/${tag}_$1.ncu-rep "$2" > gpurun_out/${tag}_ncu_full_$1.txt 2>/dev/null
  python tools/ncu_metrics.py gpurun_out/${tag}_$1.ncu-rep > gpurun_out/${tag}_metrics_$1.json 2>/dev/null
  rm -f gpurun_out/${tag}_$1.ncu-rep
}
cap() {  # cap <name> <kernel regex> <skip> <mangled substring> <bench args...>
  local n=$1 k=$2 skip=$3 mg=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${tag}_$n \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only "$@" > gpurun_out/${tag}_ncu_$n.log 2>&1
  sum $n $mg
}
cap k_flank_exact_t 'k_flank_exact_t$' 1 15k_flank_exact_tE
cap k_flank_seed 'k_flank_seed$' 1 12k_flank_seedE
cap k_flank_band1 'k_flank_band1$' 1 13k_flank_band1E
cap k_flank_band2 'k_flank_band2$' 1 13k_flank_band2E
cap k_flank_band_wide 'k_flank_band_wide$' 1 17k_flank_band_wideE
cap k_e2e_identity 'k_e2e_identity$' 1 14k_e2e_identityE
cap k_e2e_lane 'k_e2e_lane$' 1 10k_e2e_laneE
cap k_hmm_lane_viterbi 'k_hmm_lane_viterbi$' 1 18k_hmm_lane_viterbiE
cap k_hmm_lane_walk 'k_hmm_lane_walk$' 1 15k_hmm_lane_walkE
cap k_wfa_score_warp 'k_wfa_score' 3 11k_wfa_scoreILb0E
cap k_unpack_seq4 'k_unpack_seq4$' 0 13k_unpack_seq4E
cap c5_k_edit_dist 'k_edit_dist$' 0 11k_edit_distE --config 5
cap c5_k_cluster_ward 'k_cluster_ward$' 0 14k_cluster_wardE --config 5
cap c5_k_consensus_vote 'k_consensus_vote' 0 16k_consensus_voteILb0E --config 5
cap c5_k_wfa_score 'k_wfa_score' 1 11k_wfa_scoreILb0E --config 5
cap c5_k_wfa_trace 'k_wfa_trace$' 1 11k_wfa_traceE --config 5
ncu --set full --clock-control none --import-source on -k regex:'k_clip_cigar$' -c 1 -f -o gpurun_out/${tag}_k_clip_cigar \
    python -m pytest tests/test_engine_gpu.py -q -k clip_reads_random_parity > gpurun_out/${tag}_ncu_k_clip_cigar.log 2>&1
sum k_clip_cigar 12k_clip_cigarE
ncu --set full --clock-control none --import-source on -k regex:'k_bamlet_clip$' -c 1 -f -o gpurun_out/${tag}_k_bamlet_clip \
    python -m pytest tests/test_engine_gpu.py -q -k bamlet_clip_parity > gpurun_out/${tag}_ncu_k_bamlet_clip.log 2>&1
sum k_bamlet_clip 13k_bamlet_clipE
rm -f gpurun_out/${tag}_ncu_*.log.tmp
du -sh gpurun_out; ls gpurun_out | wc -l
