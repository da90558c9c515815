#!/bin/bash
# Under gpurun (1 GPU): launch list of one full-size resident pass, then one `--set full` capture of
# each main kernel at full size (whole-shard launches), for the roofline `traffic` figures.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_full.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/ncu_bench_full.log 2>&1
for k in 'k_flank_exact$' 'k_flank_band$' 'k_flank_band2$' 'k_wfa_score' 'k_hmm_viterbi_thread$'; do
  n=$(echo $k | tr -d '$')
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/full_$n \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/ncu_full_$n.log 2>&1
done
ls -la gpurun_out | tail -12
