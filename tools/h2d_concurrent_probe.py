"""What host->device rate does ONE GPU reach when the e2e driver's copy pattern is reproduced without any kernels:
T streams each copying a chunk of pinned memory at the same time, with and without device->host traffic in the
other direction?  (The end-to-end pass moves 2.19 GB up and 0.26 GB down per step.)  One JSON line per case."""
import json, sys, time
import torch

dev = torch.device("cuda", 0)
def run(n_streams, chunk_mb, total_mb, d2h_frac):
    n_chunks = max(n_streams, int(total_mb // chunk_mb))
    src = [torch.empty(chunk_mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
    dst = [torch.empty(chunk_mb << 20, dtype=torch.uint8, device=dev) for _ in range(n_streams)]
    back_n = int((chunk_mb << 20) * d2h_frac)
    back = [torch.empty(max(1, back_n), dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    def once():
        for c in range(n_chunks):
            s = streams[c % n_streams]
            with torch.cuda.stream(s):
                dst[c % n_streams].copy_(src[c % n_streams], non_blocking=True)
                if back_n:
                    back[c % n_streams].copy_(dst[c % n_streams][:back_n], non_blocking=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    up = n_chunks * (chunk_mb << 20)
    print(json.dumps({"streams": n_streams, "chunk_mb": chunk_mb, "chunks": n_chunks, "d2h_frac": d2h_frac,
                      "h2d_gbs": up / dt / 1e9, "ms_for_2.19GB": 2.19e9 / (up / dt) * 1e3}))

for ns, mb in [(1, 1024), (1, 137), (2, 137), (8, 137), (8, 34), (16, 137)]:
    for frac in (0.0, 0.12):
        run(ns, mb, 2190, frac)
