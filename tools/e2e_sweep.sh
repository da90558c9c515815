# end-to-end pass for a few driver settings; one line each
run() { python bench.py --steps 5 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/e2e_sweep_tmp.json 2>gpurun_out/e2e_sweep_tmp.err
python - "$@" <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/e2e_sweep_tmp.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print(" ".join(sys.argv[1:]), "| chunks", e.get("chunks"), "e2e ms", round(e["ms_per_step"],2), "loci/s", round(e["value"]))
except Exception as ex:
    print(" ".join(sys.argv[1:]), "FAILED", ex, open("gpurun_out/e2e_sweep_tmp.err").read()[-400:])
PY
}
run --upload-slots 0 --host-threads 8
run --upload-slots 0 --host-threads 8 --guided 1
run --upload-slots 0 --host-threads 8 --guided 1 --min-chunk-loci 800
run --upload-slots 0 --host-threads 8 --guided 1 --min-chunk-loci 3000
run --upload-slots 0 --host-threads 8 --guided 1 --chunk-loci 11000
run --upload-slots 0 --host-threads 12 --guided 1
run --upload-slots 0 --host-threads 6 --guided 1
