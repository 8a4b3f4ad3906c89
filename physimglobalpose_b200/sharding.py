"""Multi-GPU plumbing on the host side.  Hypotheses (or bases) shard across ranks, one process per GPU; the scene grid and the
models are replicated; the only exchange on the path is the all-gather of the per-rank selection records after K4 -- and that
lives behind the C ABI (pgp_comm_init / pgp_topk_begin / pgp_topk_end, csrc/pgp_comm.cu: NCCL + a deterministic merge in C++).
SURVEY.md 8(e).

What is left here is the launcher glue: the contiguous shard of a rank, handing the 128-byte NCCL id from rank 0 to the others
over whatever host channel the launcher offers (torch.distributed's store under torchrun), and a host-buffer gather for the gloo
tests of the merge rule.

The reference has no counterpart (single-threaded, S4/algorithms/match4pcsBase.cc:1888-1901); the contract kept is that the merged
result equals the serial scan's: order (score desc, index asc), independent of the number of ranks."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .engine import HYP_DTYPE, PoseEngine, topk_merge


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n hypotheses (or bases) for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def comm_init_from_env(engine: PoseEngine, rank: int, world: int, group=None) -> None:
    """Joins `engine` to a communicator of `world` ranks (pgp_comm_init).  The NCCL unique id is created on rank 0
    (pgp_comm_unique_id) and handed to the other ranks through torch.distributed -- the host channel a torchrun launch already
    has; nothing of torch touches the data path afterwards."""
    if world == 1:
        engine.comm_init(None, 0, 1)
        return
    import torch.distributed as dist

    box = [PoseEngine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(box[0], rank, world)


def gather_topk(local: np.ndarray, k: int, group=None) -> np.ndarray:
    """Host-buffer variant over torch.distributed (gloo in the CPU tests): all-gather the local top-k record arrays and apply
    the library's merge rule (pgp_topk_merge).  Test scaffolding for the merge; the product path is pgp_topk on a communicator."""
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
