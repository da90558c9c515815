"""Multi-GPU plumbing of the hot path: loci are independent (src/trgt/workflows/tr.rs:24-109), so the
catalog shards by locus across ranks with no data-path collective.  The only exchange is one
variable-length gather of per-locus records to rank 0 at the end of a pass.
Works with any torch.distributed backend: NCCL on device tensors, gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_loci: int, world: int) -> List[int]:
    """Contiguous locus shards, sizes differing by at most one: bounds[r]..bounds[r+1] is rank r's."""
    base, rem = divmod(n_loci, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


def record_parts(results) -> List[np.ndarray]:
    """Byte views of what a VCF writer needs per locus group: allele sequences + offsets, MC, MS, AP."""
    return [x for r in results for x in (
        np.ascontiguousarray(r.annotations.motif_counts).view(np.uint8),
        np.ascontiguousarray(r.annotations.spans).reshape(-1).view(np.uint8),
        np.ascontiguousarray(r.annotations.purity).view(np.uint8),
        r.glue.backbones.data, np.ascontiguousarray(r.glue.backbones.offsets).view(np.uint8))]


class RecordGather:
    """Sizes all-gather, then one padded gather from a (pinned) staging buffer to rank 0."""

    def __init__(self, device: torch.device):
        self.device = device
        self.cap = 0
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def _drain(self):
        """Wait for the previous call's staging copy and gather: they read self.pin / self.dev asynchronously."""
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _ensure(self, n: int):
        if n <= self.cap:
            return
        self._drain()  # earlier operations may still use the buffers that are about to be replaced
        self.cap = int(n * 1.5) + (1 << 16)
        cuda = self.device.type == "cuda"
        self.pin = torch.empty(self.cap, dtype=torch.uint8, pin_memory=cuda)
        self.pin_np = self.pin.numpy()
        self.dev = torch.empty(self.cap, dtype=torch.uint8, device=self.device) if cuda else self.pin
        self.recv = ([torch.empty(self.cap, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
                     if self.rank == 0 else None)
        self.host = torch.empty(self.cap * self.world, dtype=torch.uint8, pin_memory=cuda) if self.rank == 0 else None

    def __call__(self, parts: Sequence[np.ndarray]) -> Optional[List[np.ndarray]]:
        """-> on rank 0 the payload of every rank (views of an internal buffer), else None"""
        n = int(sum(p.size for p in parts))
        if self.world == 1:
            self._ensure(n)
            if parts:
                np.concatenate(parts, out=self.pin_np[:n])
            return [self.pin_np[:n]]
        size = torch.tensor([n], dtype=torch.int64, device=self.device)
        sizes_t = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(sizes_t, size)
        sizes = [int(v) for v in sizes_t.cpu().tolist()]
        mx = max(sizes)
        self._ensure(mx)
        self._drain()  # the staging buffer is rewritten below
        if parts:
            np.concatenate(parts, out=self.pin_np[:n])
        if self.dev is not self.pin:
            self.dev[:n].copy_(self.pin[:n], non_blocking=True)
        dist.gather(self.dev[:mx], [t[:mx] for t in self.recv] if self.rank == 0 else None, dst=0)
        if self.rank != 0:
            self._drain()  # the caller may re-enter (or drop its arrays) as soon as this returns
            return None
        out, o = [], 0
        for t, sz in zip(self.recv, sizes):
            self.host[o:o + sz].copy_(t[:sz], non_blocking=True)
            out.append((o, sz))
            o += sz
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        h = self.host.numpy()
        return [h[a:a + sz] for a, sz in out]
