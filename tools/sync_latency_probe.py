"""GPU aid: what does one small synchronised round trip (4 KB up, a trivial kernel, 4 KB down, stream synchronise) cost
on an idle GPU, and while other streams keep the host->device link saturated with 96 MB copies -- the situation of a
phase-B / C call in the end-to-end driver?  One line per case."""
import threading, time
import torch

dev = torch.device("cuda", 0)
small_h = torch.empty(4096, dtype=torch.uint8).pin_memory()
small_d = torch.empty(4096, dtype=torch.uint8, device=dev)
back_h = torch.empty(4096, dtype=torch.uint8).pin_memory()
s_small = torch.cuda.Stream(dev)

def round_trips(n=300):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        with torch.cuda.stream(s_small):
            small_d.copy_(small_h, non_blocking=True)
            small_d.add_(1)
            back_h.copy_(small_d, non_blocking=True)
        s_small.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[int(len(ts) * 0.95)] * 1e6

print("idle: median %.0f us, p95 %.0f us" % round_trips())
for n_up, mb in ((1, 96), (4, 96), (7, 96), (7, 8)):
    stop = False
    big_h = [torch.empty(mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_up)]
    big_d = [torch.empty(mb << 20, dtype=torch.uint8, device=dev) for _ in range(n_up)]
    def uploader(i):
        st = torch.cuda.Stream(dev)
        while not stop:
            with torch.cuda.stream(st):
                for _ in range(2):   # two copies queued ahead, as a chunk's pieces are
                    big_d[i].copy_(big_h[i], non_blocking=True)
            st.synchronize()
    th = [threading.Thread(target=uploader, args=(i,)) for i in range(n_up)]
    for t in th:
        t.start()
    time.sleep(0.3)
    med, p95 = round_trips()
    stop = True
    for t in th:
        t.join()
    print("%d streams uploading %d MB copies: median %.0f us, p95 %.0f us" % (n_up, mb, med, p95))
