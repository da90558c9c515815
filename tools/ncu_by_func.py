"""Usage: ncu_by_func.py <report> <mangled-kernel-substring> [lib.so].
Aggregate the ncu per-line instruction shares of one kernel by enclosing function (line ranges
taken from the source files by a crude scan for TRGT_HD / __global__ / template heads)."""
import os, re, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
lib = [sys.argv[3]] if len(sys.argv) > 3 else []   # the .so the report was captured with (default: the current build)
here = os.path.dirname(os.path.abspath(__file__))
out = subprocess.run([sys.executable, os.path.join(here, "ncu_by_line.py"), rep, kern] + lib + ["--all"], capture_output=True, text=True).stdout
csrc = os.path.join(os.path.dirname(here), "trgt_b200", "csrc")
bounds = {}
for f in os.listdir(csrc):
    heads = []
    for n, line in enumerate(open(os.path.join(csrc, f)), 1):
        m = re.match(r"(?:TRGT_HD|TRGT_D|__device__ __forceinline__|__global__)?\s*[\w:<>\*& ]*?\b(\w+)\(", line) if re.match(r"^(TRGT_HD|TRGT_D|__device__|__global__|k_\w+\()", line) else None
        if m:
            heads.append((n, m.group(1)))
    bounds[f] = heads
agg = collections.Counter(); samp = collections.Counter()
for l in out.splitlines():
    m = re.match(r"(\S+):(\d+)\s+inst\s+([\d.]+)%\s+samples\s+([\d.]+)%", l)
    if not m:
        continue
    f, ln, p, sp = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))
    name = f
    for n, fn in bounds.get(f, []):
        if n <= ln:
            name = f"{f}:{fn}"
    agg[name] += p; samp[name] += sp
for k, v in agg.most_common(30):
    print(f"{k:45s} inst {v:5.1f}%  samples {samp[k]:5.1f}%")
