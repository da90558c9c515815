"""Join an ncu SASS source page (per-instruction counters) with nvdisasm line info, and print the
hottest source lines of one kernel.  Usage: ncu_by_line.py <report.ncu-rep> <kernel-substring> [so]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
so = os.path.abspath(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trgt_b200", "libtrgt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate kernel section
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and kern in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//---")), len(dis))
line_of = {}   # instruction index -> (file, line)
cur = ("?", 0)
idx = 0
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        line_of[idx] = cur
        idx += 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = hdr.index("Instructions Executed")
cs = hdr.index("# Samples") if "# Samples" in hdr else None
agg = collections.Counter(); samp = collections.Counter()
n = 0
for r in rows[h + 1:]:
    if len(r) <= ci:
        continue
    k = line_of.get(n, ("?", 0))
    try:
        agg[k] += int(float(r[ci] or 0))
        if cs is not None:
            samp[k] += int(float(r[cs] or 0))
    except ValueError:
        pass
    n += 1
tot = sum(agg.values()) or 1
tots = sum(samp.values()) or 1
print(f"{n} SASS instructions, {len(line_of)} with line info, {tot} warp-instructions executed")
for k, v in agg.most_common(100000 if '--all' in sys.argv else 45):
    print(f"{k[0]}:{k[1]:<5d} inst {100*v/tot:5.1f}%  samples {100*samp[k]/tots:5.1f}%")
