"""Print the headline numbers and the per-kernel table of bench.py JSON lines (files given as arguments)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    e2e = d.get("e2e") or {}
    print(f, "value %.3g ms %.2f | e2e %.3g ms %.1f h2d %.2f GB" % (
        d["value"], d["ms_per_step"], e2e.get("value", 0), e2e.get("ms_per_step", 0), e2e.get("h2d_bytes_per_step", 0) / 1e9))
    print(" phases", e2e.get("phase_ms_summed_over_host_threads"))
    print(" cpu", d.get("cpu_baseline"), d.get("parity"))
    print(" roofline", d.get("roofline"))
    for k, v in (d.get("kernels") or {}).items():
        if "ms_per_launch" not in v:
            print("  %-24s %s" % (k, v))
            continue
        print("  %-24s %.3f ms  share %s  GB/s %s" % (k, v["ms_per_launch"], v["share"] and round(v["share"], 3),
                                                      v["achieved_gbs"] and round(v["achieved_gbs"])))
