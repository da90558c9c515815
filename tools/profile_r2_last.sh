#!/bin/bash
# Under gpurun (1 GPU): the launch list and `--set full` summaries of the two kernels changed after the r2z/r2y
# captures (k_flank_exact_t: funnel-shift copies; k_flank_band_wide: counter-drawn pairs, budgets 18/36).
# Outputs: gpurun_out/r2x_*.
tag=r2x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_bench.log 2>&1
cap() {  # cap <name> <kernel regex> <mangled substring>
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/${tag}_$1 \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_$1.log 2>&1
  python tools/ncu_report.py gpurun_out/${tag}_$1.ncu-rep "$3" > gpurun_out/${tag}_ncu_full_$1.txt 2>/dev/null
  python tools/ncu_metrics.py gpurun_out/${tag}_$1.ncu-rep > gpurun_out/${tag}_metrics_$1.json 2>/dev/null
  rm -f gpurun_out/${tag}_$1.ncu-rep
}
cap k_flank_exact_t 'k_flank_exact_t$' 15k_flank_exact_tE
cap k_flank_band_wide 'k_flank_band_wide$' 17k_flank_band_wideE
head -22 gpurun_out/${tag}_ncu_full_k_flank_exact_t.txt; head -22 gpurun_out/${tag}_ncu_full_k_flank_band_wide.txt
