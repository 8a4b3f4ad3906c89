"""world_size-2 CPU tests (gloo) of the multi-GPU merge plumbing.  Each rank holds the selection of its hypothesis shard in the
library's wire format (header + k records, what K4 writes into the exchange slot); the all-gather runs over gloo instead of
NCCL, and the merge is the library's own pgp_exchange_merge -- the function pgp_topk / pgp_improving_chain call after the
ncclAllGather.  The merged top-k and the merged improving chain must equal the serial scan over the whole list."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from physimglobalpose_b200 import sharding
from physimglobalpose_b200.engine import HYP_DTYPE, exchange_merge, wire_block


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_all(n):
    rng = np.random.default_rng(42)
    r = np.zeros(n, HYP_DTYPE)
    r["index"] = np.arange(n)
    r["count"] = rng.integers(0, 30, size=n)                 # many ties: the index tie-break matters
    r["score"] = r["count"].astype(np.float32) / np.float32(30)
    r["T"] = rng.normal(size=(n, 12)).astype(np.float32)
    return r


def _serial_chain(rec):
    keep, best = [], 0
    for i in range(len(rec)):
        if rec["count"][i] > best:
            keep.append(i); best = rec["count"][i]
    return rec[keep]


def _worker(rank, world, port, n, k, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allrec = _make_all(n)
    lo, hi = sharding.shard_range(n, rank, world)
    part = allrec[lo:hi]
    local_top = part[np.lexsort((part["index"], -part["score"]))][:k]
    local_chain = _serial_chain(part)

    def gather(blk):
        t = torch.from_numpy(blk.view(np.uint8).copy())
        out = torch.empty(world * t.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(out, t)
        return out.numpy().view(HYP_DTYPE)

    # explicit index bases (the records already carry global indices)
    top = exchange_merge(gather(wire_block(local_top, hi - lo, k)), world, k, kind=0)
    # PGP_INDEX_AUTO: the records carry shard-local indices, the bases come from the exchanged batch sizes
    loc = local_top.copy(); loc["index"] -= lo
    top_auto = exchange_merge(gather(wire_block(loc, hi - lo, k)), world, k, kind=0, auto_base=True)
    chain = exchange_merge(gather(wire_block(local_chain, hi - lo, 255)), world, 255, kind=1, mode="count")
    legacy = sharding.gather_topk(local_top, k)
    np.savez(os.path.join(out_dir, f"merged_{rank}.npz"), top=top, top_auto=top_auto, chain=chain, legacy=legacy)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_exchange_equals_the_serial_scan(tmp_path):
    n, k, world = 5000, 64, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    allrec = _make_all(n)
    want = allrec[np.lexsort((allrec["index"], -allrec["score"]))][:k]
    want_chain = _serial_chain(allrec)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"merged_{r}.npz"))
        for key in ("top", "top_auto", "legacy"):                    # every rank ends with the same, serial-order result
            assert got[key].tobytes() == want.tobytes(), key
        assert got["chain"].tobytes() == want_chain.tobytes()


def test_exchange_merge_is_independent_of_the_sharding():
    allrec = _make_all(6000)
    want = allrec[np.lexsort((allrec["index"], -allrec["score"]))][:64]
    want_chain = _serial_chain(allrec)
    for world in (1, 2, 3, 4, 8):
        blocks, chains = [], []
        for r in range(world):
            lo, hi = sharding.shard_range(len(allrec), r, world)
            part = allrec[lo:hi]
            blocks.append(wire_block(part[np.lexsort((part["index"], -part["score"]))][:64], hi - lo, 64))
            chains.append(wire_block(_serial_chain(part), hi - lo, 255))
        assert exchange_merge(np.concatenate(blocks), world, 64).tobytes() == want.tobytes()
        assert exchange_merge(np.concatenate(chains), world, 255, kind=1).tobytes() == want_chain.tobytes()
    # an empty shard (more ranks than hypotheses) contributes a zero header
    blocks = [wire_block(allrec[:0], 0, 64), wire_block(want, len(allrec), 64)]
    assert exchange_merge(np.concatenate(blocks), 2, 64).tobytes() == want.tobytes()
    # a local chain that overflowed its k slots is an error, not a silently short chain
    bad = wire_block(want_chain, len(allrec), 255)
    bad[0]["score"] = 1.0
    with pytest.raises(Exception):
        exchange_merge(bad, 1, 255, kind=1)
