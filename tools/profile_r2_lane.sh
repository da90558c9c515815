#!/bin/bash
# Under gpurun (1 GPU): evidence for the phase-B lane kernel that replaced k_e2e_thread after the r2z capture --
# a fresh launch list of the resident pass, one `--set full` capture of k_e2e_lane (summarised on the box), and
# memcheck over the end-to-end alignment parity tests.  Outputs: gpurun_out/r2y_*.
tag=r2y
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_e2e_lane$' -s 1 -c 1 -f -o gpurun_out/${tag}_k_e2e_lane \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --resident-only > gpurun_out/${tag}_ncu_k_e2e_lane.log 2>&1
python tools/ncu_report.py gpurun_out/${tag}_k_e2e_lane.ncu-rep 10k_e2e_laneE > gpurun_out/${tag}_ncu_full_k_e2e_lane.txt 2>/dev/null
python tools/ncu_metrics.py gpurun_out/${tag}_k_e2e_lane.ncu-rep > gpurun_out/${tag}_metrics_k_e2e_lane.json 2>/dev/null
rm -f gpurun_out/${tag}_k_e2e_lane.ncu-rep
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -x \
    -k "align or consensus" > gpurun_out/${tag}_sanitizer_memcheck_align.txt 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_sanitizer_memcheck_align.txt
cat gpurun_out/${tag}_ncu_full_k_e2e_lane.txt | head -50
