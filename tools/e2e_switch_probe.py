import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trgt_b200
from harness import workload
from harness.pipeline import ChunkedHotPath
n_loci, threads = 125000, 8
engines = [trgt_b200.Engine(0) for _ in range(threads)]
w = workload.generate(n_loci, 30, alloc_reads=engines[0].pinned_array)
w.pack_seq4(alloc=engines[0].pinned_array)
chp = ChunkedHotPath(engines, w, chunk_loci=-(-n_loci // (2 * threads)), glue_threads=2, use_seq4=True, upload_slots=0)
def timeit(label, n=6):
    for _ in range(2):
        chp.run_e2e()
    t0 = time.perf_counter()
    for _ in range(n):
        chp.run_e2e()
    print(f"{label}: {(time.perf_counter() - t0) / n * 1e3:.2f} ms per step", flush=True)
for si in (0.005, 0.001, 0.0002, 0.00005):
    sys.setswitchinterval(si)
    timeit(f"switch interval {si*1e3:.2f} ms")
