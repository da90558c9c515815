import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import trgt_b200
eng = trgt_b200.Engine(0)
n = 1 << 30
a = eng.pinned_array(n); a[:] = 1
t = torch.from_numpy(a)
print("is_pinned (torch view of trgt_host_alloc):", t.is_pinned())
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def bw(src, label):
    for _ in range(2): d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(label, "%.1f GB/s" % (n * 5 / (time.perf_counter() - t0) / 1e9))
bw(t, "trgt_host_alloc buffer via torch copy_")
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(1)
bw(h, "torch pinned")
# through cudaMemcpyAsync directly (cuda-python not needed: use torch's stream with cudart via ctypes)
import ctypes as C
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
for src, label in ((a.ctypes.data, "cudaMemcpy from trgt_host_alloc"), (h.data_ptr(), "cudaMemcpy from torch pinned")):
    rt.cudaMemcpy(d.data_ptr(), src, n, 1)
    t0 = time.perf_counter()
    for _ in range(5): rt.cudaMemcpy(d.data_ptr(), src, n, 1)
    print(label, "%.1f GB/s" % (n * 5 / (time.perf_counter() - t0) / 1e9))
