"""Markdown table of the step's kernels: CUDA-event time and share from a bench line, algorithmic bytes and achieved
GB/s from the same line, DRAM traffic / issue-active / warps-active / lanes-per-instruction from the ncu captures
gpurun_out/<tag>_<kernel>.ncu-rep.  Usage: r2_table.py <bench.json> <tag> [--traffic-json out.json]"""
import csv, io, json, os, subprocess, sys

bench, tag = sys.argv[1], sys.argv[2]
d = json.loads(open(bench).read().strip().splitlines()[-1])
peak = d["roofline"]["peak"]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def ncu(kernel):
    rep = os.path.join(root, "gpurun_out", f"{tag}_{kernel}.ncu-rep")
    if not os.path.exists(rep):
        return {}
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    h, u, v = rows[0], rows[1], rows[2]
    m = {}
    for a, b, c in zip(h, u, v):
        m[a] = (b, c)
    def val(k, scale=None):
        if k not in m:
            return None
        unit, x = m[k]
        try:
            x = float(x.replace(",", ""))
        except ValueError:
            return None
        if scale == "bytes":
            x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return x
    return {"traffic": (val("dram__bytes_read.sum", "bytes") or 0) + (val("dram__bytes_write.sum", "bytes") or 0),
            "issue": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "lanes": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "inst": val("smsp__inst_executed.sum"), "regs": val("launch__registers_per_thread")}

traffic = {}
print("| Kernel | ms / launch | share | algorithmic bytes | achieved GB/s | frac of %d | DRAM traffic (ncu) | warp-instr | issue-active | warps active | lanes / instr | regs |" % peak)
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, v in d["kernels"].items():
    if v["ms_per_launch"] * max(1.0, v["launches_per_step"]) < 0.03 and k != "k_unpack_seq4":
        continue
    n = ncu("k_wfa_score_warp" if k == "k_wfa_score_warp" else k)
    ab, gbs = v.get("algorithmic_bytes"), v.get("achieved_gbs")
    if n.get("traffic"):
        traffic[k] = int(n["traffic"])
    f = lambda x, fmt: (fmt % x) if x is not None else "—"
    print("| `%s` | %.3f | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        k, v["ms_per_launch"], f(v["share"] and 100 * v["share"], "%.1f %%"), f(ab and ab / 1e9, "%.3f GB"), f(gbs, "%.0f"),
        f(gbs and gbs / peak, "%.3f"), f(n.get("traffic") and n["traffic"] / 1e9, "%.2f GB") + (f(n.get("traffic") and ab and n["traffic"] / ab, " (%.2f×)") if ab else ""),
        f(n.get("inst") and n["inst"] / 1e6, "%.0f M"), f(n.get("issue"), "%.0f %%"), f(n.get("warps"), "%.0f %%"), f(n.get("lanes"), "%.1f"), f(n.get("regs"), "%d")))
if "--traffic-json" in sys.argv:
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    old = json.load(open(out)) if os.path.exists(out) else {}
    old.update(traffic)
    old["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per whole-shard launch (125 000 loci, 30x), from the ncu "
                       "--set full captures summarised in r2_ncu_full_*.txt (kernels that no longer exist keep their round-1 entry)")
    json.dump(old, open(out, "w"), indent=1)
