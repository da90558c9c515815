for cl in 7813 3907 1954; do for ht in 8 12 16; do
python bench.py --steps 5 --warmup 2 --no-cpu-baseline --chunk-loci $cl --host-threads $ht > gpurun_out/e2e_sweep_${cl}_${ht}.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_sweep_${cl}_${ht}.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("chunk", $cl, "threads", $ht, "e2e ms", round(e["ms_per_step"],2), "loci/s", round(e["value"]), e.get("phase_ms_summed_over_host_threads"))
PY
done; done
