"""Multi-GPU plumbing: hypotheses shard across ranks (one process per GPU); the scene grid and the
models are replicated; the only exchange is the all-gather of the per-rank top-k records
(k x 64 bytes) after K4.  SURVEY.md 8(e).

The reference has no counterpart (single-threaded, S4/algorithms/match4pcsBase.cc:1888-1901); the
contract kept is that the merged result equals the serial scan's: order (score desc, index asc),
independent of the number of ranks."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .engine import HYP_DTYPE, topk_merge


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n hypotheses for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_topk(local: np.ndarray, k: int, group=None) -> np.ndarray:
    """Host-buffer variant (gloo or nccl): all-gather the local top-k record arrays and merge."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return topk_merge([local], k)
    buf = np.zeros(k, HYP_DTYPE)
    buf["index"] = -1
    buf[: len(local)] = local[:k]
    t = torch.from_numpy(buf.view(np.uint8).copy())
    backend = dist.get_backend(group)
    if backend == "nccl":
        t = t.cuda()
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    rec = out.cpu().numpy().view(HYP_DTYPE).reshape(world, k)
    return topk_merge([rec[r] for r in range(world)], k)


class DeviceTopkGather:
    """NCCL path with no host round trip before the collective: K4 writes its k records straight
    into the all-gather send buffer on the device (pgp_topk_dev), the all-gather runs on the same
    stream order, and only the gathered world*k*64 bytes come back for the final merge."""

    def __init__(self, engine, k: int, group=None, slots: int = 2):
        import torch
        import torch.distributed as dist

        self.engine, self.k, self.group = engine, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.send = torch.zeros(k * 64, dtype=torch.uint8, device="cuda")
        self.recv = torch.zeros(self.world * k * 64, dtype=torch.uint8, device="cuda")
        self._free = [torch.zeros(self.world * k * 64, dtype=torch.uint8).pin_memory() for _ in range(max(1, slots))]

    def __call__(self, obj: int, index_base: int) -> np.ndarray:
        return self.collect(self.submit(obj, index_base))

    # Pipelined form: submit() only enqueues (K4 -> all-gather -> async D2H into a pinned slot + an event) and returns a
    # ticket, so the next step's scoring kernel can start without a host round trip; collect() waits for the ticket's event
    # and does the deterministic host merge.  All work of a step stays on the one stream, in order.
    def submit(self, obj: int, index_base: int):
        import torch
        import torch.distributed as dist

        host = torch.empty(self.world * self.k * 64, dtype=torch.uint8).pin_memory() if self._free == [] else self._free.pop()
        self.engine.topk_device(obj, self.k, index_base, self.send)
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
            host.copy_(self.recv, non_blocking=True)
        else:
            host.copy_(self.send, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev

    def collect(self, ticket) -> np.ndarray:
        host, ev = ticket
        ev.synchronize()
        rec = host.numpy().view(HYP_DTYPE).reshape(self.world, self.k).copy()
        self._free.append(host)
        return topk_merge([rec[r] for r in range(self.world)], self.k)
