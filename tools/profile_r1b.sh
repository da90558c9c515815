#!/bin/bash
# Under gpurun (1 GPU): launch list of one full-size resident pass, then one `--set full` capture of the
# kernels named on the command line (default: the main kernels at the end of round 1), whole-shard launches.
# k_unpack_seq4 launches once per run (outside the resident step), the others once per warm-up and per step.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_full.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/ncu_bench_full.log 2>&1
KERNELS=${@:-'k_flank_exact_t$ k_flank_band$ k_flank_band2$ k_e2e_thread$ k_hmm_viterbi_thread$ k_unpack_seq4$'}
for k in $KERNELS; do
  n=$(echo $k | tr -d '$')
  skip=1; [ "$n" = "k_unpack_seq4" ] && skip=0
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/full_$n \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/ncu_full_$n.log 2>&1
done
ls -la gpurun_out | tail -12
