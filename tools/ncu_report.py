"""One text summary per ncu report for profiles/: raw metrics (ncu_summary.py), instruction share by function
(ncu_by_func.py) and stall reasons.  Usage: ncu_report.py <report.ncu-rep> <mangled-kernel-substring> [lib.so] > out.txt"""
import csv, io, os, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
lib = [sys.argv[3]] if len(sys.argv) > 3 else []
here = os.path.dirname(os.path.abspath(__file__))
print(subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), rep], capture_output=True, text=True).stdout.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
vals = dict(zip(rows[0], rows[2]))
for k in ("l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
    if k in vals:
        print(f"{k} = {vals[k]}")
print("\n--- instruction share by function")
out = subprocess.run([sys.executable, os.path.join(here, "ncu_by_func.py"), rep, kern] + lib, capture_output=True, text=True).stdout
print("\n".join(out.splitlines()[:16]))
print("--- stall reasons (warps stalled per issue)")
for k, v in sorted(vals.items()):
    if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
        try:
            if float(v) >= 0.05:
                print(k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v)
        except ValueError:
            pass
