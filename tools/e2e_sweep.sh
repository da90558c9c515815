# end-to-end pass for a few driver settings; one line each
run() { python bench.py --steps 5 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/e2e_sweep_tmp.json 2>gpurun_out/e2e_sweep_tmp.err
python - "$@" <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/e2e_sweep_tmp.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print(" ".join(sys.argv[1:]), "| e2e ms", round(e["ms_per_step"],2), "loci/s", round(e["value"]), "parity", str((d.get("parity") or {}).get("result","-"))[:9])
except Exception as ex:
    print(" ".join(sys.argv[1:]), "FAILED", ex, open("gpurun_out/e2e_sweep_tmp.err").read()[-400:])
PY
}
run --upload-slots 0 --host-threads 8
run --uploaders 1 --host-threads 8 --chunk-loci 7813
run --uploaders 2 --host-threads 8 --chunk-loci 7813
run --uploaders 2 --host-threads 8 --chunk-loci 3907
run --uploaders 2 --host-threads 12 --chunk-loci 3907 --max-inflight 6
run --uploaders 3 --host-threads 12 --chunk-loci 3907 --max-inflight 6
run --uploaders 2 --host-threads 8 --chunk-loci 1954 --max-inflight 8
