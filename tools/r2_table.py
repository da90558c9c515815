"""Markdown table of the step's kernels: CUDA-event time and share from a bench line, algorithmic bytes and achieved
GB/s from the same line, DRAM traffic / issue-active / warps-active / lanes-per-instruction from the ncu captures'
key metrics <tag>_metrics_<kernel>.json (gpurun_out/ or profiles/).  Usage: r2_table.py <bench.json> <tag> [--traffic-json out.json]"""
import json, os, sys

bench, tag = sys.argv[1], sys.argv[2]
d = json.loads(open(bench).read().strip().splitlines()[-1])
peak = d["roofline"]["peak"]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def ncu(kernel):
    """key metrics of the kernel's capture: gpurun_out/<tag>_metrics_<kernel>.json (tools/ncu_metrics.py)"""
    for d_ in ("gpurun_out", "profiles"):
        f = os.path.join(root, d_, f"{tag}_metrics_{kernel}.json")
        if os.path.exists(f):
            m = json.load(open(f))
            return {"traffic": (m.get("dram_read") or 0) + (m.get("dram_write") or 0), "issue": m.get("issue_active_pct"),
                    "warps": m.get("warps_active_pct"), "lanes": m.get("lanes_per_inst"), "inst": m.get("warp_inst"),
                    "regs": m.get("regs")}
    return {}

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
