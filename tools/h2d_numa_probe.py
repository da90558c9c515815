"""Pinned H2D bandwidth from each NUMA node of the host (is the 44 GB/s of the end-to-end pass a NUMA effect?)."""
import glob, os, time
import torch

def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", nodes, "affinity", len(os.sched_getaffinity(0)))
dev = torch.device("cuda", 0)
props = torch.cuda.get_device_properties(0)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0" if hasattr(props, "pci_bus_id") else None
try:
    print("gpu", bus, "numa_node", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
except Exception as e:
    print("gpu numa unknown", e)
allowed = os.sched_getaffinity(0)
for node in nodes + [None]:
    if node is not None:
        c = [x for x in cpus_of(node) if x in allowed]
        if not c:
            print("node", node, "no allowed cpus"); continue
        os.sched_setaffinity(0, c)
    else:
        os.sched_setaffinity(0, allowed)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)  # first touch on this node
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("node", node, "H2D GB/s %.1f" % (n / dt / 1e9))
    del h, d
