"""Aggregate pinned host->device bandwidth of one box with 1..N ranks copying at the same time, with the default
placement of threads and pinned buffers and with both bound to the GPU's own NUMA node.  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_scale_probe.py
Prints one JSON line on rank 0: per-rank and aggregate GB/s for each placement (the ceiling of the end-to-end
pass, whose 2.2 GB of reads per rank and step all cross this link)."""
import glob
import json
import os
import time

import torch
import torch.distributed as dist


def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    props = torch.cuda.get_device_properties(local)
    bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    try:
        gpu_node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    except Exception:
        gpu_node = -1
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    allowed = sorted(os.sched_getaffinity(0))
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    res = {}
    for mode in ("default", "gpu_node"):
        if mode == "gpu_node":
            if gpu_node < 0 or gpu_node not in nodes:
                res[mode] = None
                continue
            c = [x for x in cpus_of(gpu_node) if x in allowed]
            if not c:
                res[mode] = None
                continue
            os.sched_setaffinity(0, c)
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        h.fill_(1)   # first touch under the current affinity
        for _ in range(2):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(8):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        res[mode] = n / dt / 1e9
        if world > 1:
            dist.barrier()
        del h
        os.sched_setaffinity(0, allowed)
    mine = {"rank": rank, "gpu": bus, "gpu_numa_node": gpu_node, "allowed_cpus": len(allowed), **res}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        agg = {m: (sum(r[m] for r in allr) if all(r[m] is not None for r in allr) else None) for m in ("default", "gpu_node")}
        print(json.dumps({"world": world, "numa_nodes": nodes, "host_cpus": len(allowed), "aggregate_gbs": agg, "ranks": allr}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
