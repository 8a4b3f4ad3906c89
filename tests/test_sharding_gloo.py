"""world_size-2 CPU test (gloo) of the multi-GPU merge plumbing: each rank holds the top-k of its
hypothesis shard; the all-gather + deterministic merge must equal the serial top-k."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from physimglobalpose_b200 import sharding
from physimglobalpose_b200.engine import HYP_DTYPE


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


def _worker(rank, world, port, n, k, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allrec = _make_all(n)
    lo, hi = sharding.shard_range(n, rank, world)
    part = allrec[lo:hi]
    local = part[np.lexsort((part["index"], -part["score"]))][:k]
    merged = sharding.gather_topk(local, k)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gather_equals_serial_topk(tmp_path):
    n, k, world = 5000, 64, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    allrec = _make_all(n)
    want = allrec[np.lexsort((allrec["index"], -allrec["score"]))][:k]
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"merged_{r}.npy"))
        assert np.array_equal(got["index"], want["index"])       # every rank ends with the same, serial-order result
        assert np.array_equal(got["T"], want["T"])
