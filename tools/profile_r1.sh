#!/bin/bash
# Run under gpurun: launch list of one bench step + one full ncu capture of the dominant kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${1:-k_flank_locate} -s 2 -c 1 -f -o gpurun_out/prof_${1:-k_flank_locate} \
    python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
