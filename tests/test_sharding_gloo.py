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


def _cap_split(counts, max_hyp, rank):
    """pgp_generated_cap_split through ctypes: (keep, index_base, n_total) of one rank."""
    import ctypes as C
    from physimglobalpose_b200 import _lib
    lib = _lib.load()
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    k, b, t = C.c_int64(), C.c_int64(), C.c_int64()
    rc = lib.pgp_generated_cap_split(counts.ctypes.data_as(C.c_void_p), len(counts), int(max_hyp), int(rank), C.byref(k), C.byref(b), C.byref(t))
    assert rc == 0
    return k.value, b.value, t.value


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
    # a GENERATED request: every rank generated `mine` hypotheses from its base range; the counts are exchanged (8 bytes per rank,
    # what pgp_comm_sync_generated all-gathers) and the global cap is applied by the library's own host function
    mine = torch.tensor([1000 + 337 * rank], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, mine)
    counts = np.array([int(c.item()) for c in cnts], dtype=np.int64)
    caps = np.array([_cap_split(counts, cap, rank) for cap in (0, 500, 1000, 1200, 5000)], dtype=np.int64)
    np.savez(os.path.join(out_dir, f"merged_{rank}.npz"), top=top, top_auto=top_auto, chain=chain, legacy=legacy, caps=caps, counts=counts)
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
    # the capped request: the ranks' kept ranges tile exactly the prefix [0, min(cap, total)) of the concatenated list
    got = [np.load(os.path.join(str(tmp_path), f"merged_{r}.npz")) for r in range(world)]
    counts = got[0]["counts"]
    assert counts.tolist() == [1000, 1337] and all(g["counts"].tolist() == counts.tolist() for g in got)
    for ci, cap in enumerate((0, 500, 1000, 1200, 5000)):
        total = int(counts.sum()) if cap <= 0 else min(cap, int(counts.sum()))
        end = 0
        for r in range(world):
            keep, base, n_total = (int(x) for x in got[r]["caps"][ci])
            assert n_total == total and base == end and 0 <= keep <= counts[r]
            end = base + keep
        assert end == total


def test_cap_split_equals_the_single_list_prefix():
    """For any split of a generated list into per-rank counts, the kept pieces are the prefix the single-GPU generator keeps."""
    rng = np.random.default_rng(3)
    for trial in range(300):
        world = int(rng.integers(1, 9))
        counts = rng.integers(0, 2000, size=world)
        if trial % 7 == 0:
            counts[rng.integers(0, world)] = 0
        total = int(counts.sum())
        for cap in (0, -1, 1, total // 2, total, total + 5):
            want_total = total if cap <= 0 else min(cap, total)
            kept = np.zeros(total, dtype=bool)
            starts = np.concatenate([[0], np.cumsum(counts)])
            for r in range(world):
                keep, base, n_total = _cap_split(counts, cap, r)
                assert n_total == want_total
                assert base == min(int(starts[r]), want_total)
                kept[int(starts[r]):int(starts[r]) + keep] = True
                assert keep == 0 or base == int(starts[r])               # a rank that keeps anything starts where its shard starts
            assert kept[:want_total].all() and not kept[want_total:].any()


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
