#!/bin/bash
# Under gpurun (1 GPU): launch list of one full-size resident pass, then one `--set full` capture of the
# kernels named on the command line, whole-shard launches (second launch of each: the first is the warm-up).
# Usage: tools/profile_r2.sh <tag> 'k_a$' 'k_b$' ...   ->  gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_<kernel>.ncu-rep
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in "$@"; do
  n=$(echo $k | tr -d '$')
  skip=1; [ "$n" = "k_unpack_seq4" ] && skip=0
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${tag}_$n \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_$n.log 2>&1
done
ls -la gpurun_out | tail -12
