"""GPU tests of the multi-GPU layer of the C ABI (csrc/pgp_comm.cu) and of the pipelined host-buffer API.

Single-GPU box: the communicator degenerates to world = 1 (no NCCL call), which still exercises the exchange slots, the wire
format, the merge and the base-range generator (a request generated in two halves == generated at once).  With >= 2 devices
(`gpurun --gpus 2`) the group tests run the real ncclAllGather path and assert sharding invariance: the same request on 1 and on
n devices gives byte-identical global top-k / improving chain."""
import numpy as np
import pytest

from physimglobalpose_b200 import synth
from physimglobalpose_b200.engine import PoseEngine, PoseGroup

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _setup(e, prob, obj=0):
    e.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta)
    e.set_model(obj, prob.model_xyz, prob.model_nrm)


def test_topk_begin_end_equals_topk(engine, small_problem):
    prob, T = small_problem
    _setup(engine, prob)
    engine.score_lcp(0, T, "count")
    want = engine.topk(0, 16, 1000)
    tickets = [engine.topk_begin(0, 16, 1000) for _ in range(5)]      # several exchanges in flight
    for t in tickets:
        assert engine.topk_end(t).tobytes() == want.tobytes()
    # PGP_INDEX_AUTO on one rank = base 0
    assert np.array_equal(engine.topk_end(engine.topk_begin(0, 16, -1))["index"], want["index"] - 1000)
    with pytest.raises(Exception):
        [engine.topk_begin(0, 16, 0) for _ in range(9)]               # only 8 slots
    # (drain the slots the failed loop took)
    for t in range(8):
        try:
            engine.topk_end(t)
        except Exception:
            pass


def test_world1_communicator_is_the_local_answer(small_problem):
    prob, T = small_problem
    e = PoseEngine(0)
    _setup(e, prob)
    e.score_lcp(0, T, "weighted")
    want_top, want_chain = e.topk(0, 32), e.improving_chain(0)
    e.comm_init(None, 0, 1)
    assert e.comm_world == 1 and e.comm_rank == 0
    assert e.topk(0, 32).tobytes() == want_top.tobytes()
    assert e.improving_chain(0).tobytes() == want_chain.tobytes()
    e.close()


def test_two_batches_in_flight(engine, port_lib):
    """pgp_score_lcp_begin twice before the first pgp_score_lcp_end: both batches come back complete and correct, each batch's
    top-k is the one queued right behind its own scoring launch; a third begin is refused."""
    import torch
    prob = synth.make_problem(600, 20000, 0.01, seed=41)
    _setup(engine, prob)
    Ts = [synth.make_hypotheses(prob, 40000, seed=s) for s in (42, 43)]
    want = []
    for T in Ts:
        c, s = engine.score_lcp(0, T, "count")
        want.append((c, s, engine.topk(0, 8)))
    host = [(torch.from_numpy(T.reshape(-1, 12).copy()).pin_memory(), torch.zeros(len(T), dtype=torch.int32).pin_memory(),
             torch.zeros(len(T), dtype=torch.float32).pin_memory()) for T in Ts]
    for rep in range(3):
        tickets = []
        for Th, ch, sh in host:
            ch.zero_(); sh.zero_()
            engine.score_lcp_begin(0, Th.data_ptr(), Th.shape[0], ch.data_ptr(), sh.data_ptr(), "count")
            tickets.append(engine.topk_begin(0, 8, 0))
        with pytest.raises(Exception):
            engine.score_lcp_begin(0, host[0][0].data_ptr(), 8, host[0][1].data_ptr(), host[0][2].data_ptr(), "count")
        for i, ((Th, ch, sh), t) in enumerate(zip(host, tickets)):
            redo = engine.score_lcp_end()
            top = engine.topk_end(t)
            if redo:
                top = engine.topk(0, 8)
            assert np.array_equal(ch.numpy().astype(np.uint32), want[i][0])
            assert np.array_equal(sh.numpy(), want[i][1])
            assert top.tobytes() == want[i][2].tobytes()


@pytest.mark.parametrize("mode", [0, 1])
def test_base_range_generation_is_range_invariant(engine, mode):
    """Bases [0, B) generated at once == the concatenation of [0, a), [a, b), [b, B): every draw is keyed by the global base
    index, so the split (the GPU count) cannot change a hypothesis."""
    prob = synth.make_segment_problem(500, 700, 0.005, seed=19)
    _setup(engine, prob)
    if mode == 1:
        engine.build_ppf_map(0)
    B = 45
    n = engine.generate_pcs(0, seed=5, max_hyp=3_000_000, n_bases=B, mode=mode)
    T_all, _, _ = engine.get_generated(0)
    ids_all, inv_all, ok_all = engine.get_bases(0)
    assert n > 0
    parts, ids = [], []
    for lo, hi in ((0, 7), (7, 7), (7, 30), (30, B)):
        m = engine.generate_pcs_range(0, lo, hi, seed=5, max_hyp=3_000_000, n_bases=B, mode=mode)
        if hi > lo:
            parts.append(engine.get_generated(0)[0])
            ids.append(engine.get_bases(0)[0][: hi - lo])
        else:
            assert m == 0
    assert np.array_equal(np.concatenate(ids), ids_all)
    assert np.array_equal(np.concatenate(parts), T_all)


def test_group_on_one_device_equals_the_context(engine, small_problem):
    prob, T = small_problem
    _setup(engine, prob)
    want_c, want_s = engine.score_lcp(0, T, "count")
    want_top, want_chain = engine.topk(0, 20), engine.improving_chain(0)
    g = PoseGroup([0])
    g.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta)
    g.set_model(0, prob.model_xyz, prob.model_nrm)
    c, s = g.score_lcp(0, T, "count")
    assert np.array_equal(c, want_c) and np.array_equal(s, want_s)
    assert g.topk(0, 20).tobytes() == want_top.tobytes()
    assert g.improving_chain(0).tobytes() == want_chain.tobytes()
    g.close()


def test_sector_gather_benchmark_runs(engine):
    l2 = engine.bench_sector_gather(32 << 20, 64)
    assert 100.0 < l2 < 100000.0          # GB/s: sanity only, the number itself is reported by bench.py


@pytest.mark.skipif("_n_devices() < 2")
@pytest.mark.parametrize("mode", ["count", "weighted"])
def test_group_scoring_is_sharding_invariant(mode):
    """The same hypothesis list on 1 device and sharded over all devices of the box: identical counts, scores, global top-64 and
    improving chain (ncclAllGather + merge inside libpgp.so)."""
    nd = _n_devices()
    prob = synth.make_problem(700, 30000, 0.01, seed=51)
    T = synth.make_hypotheses(prob, 50_001, seed=52)
    T = T[np.random.default_rng(1).permutation(len(T))]
    # a long improving chain that spans every shard: the 120 best hypotheses, ascending, spread evenly over the list
    g = PoseGroup([0])
    g.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta)
    g.set_model(0, prob.model_xyz, prob.model_nrm)
    c0, _ = g.score_lcp(0, T, "count")
    g.close()
    # the list in ascending order of the count, so that the chain has one element per distinct count value -- but only for the
    # ~60 largest distinct values (a rank may contribute at most PGP_CHAIN_EXCHANGE_CAP = 255 elements); the rest stays shuffled
    uniq = np.unique(c0)
    thr = uniq[-min(60, len(uniq))]
    hi = np.flatnonzero(c0 >= thr)
    lo = np.flatnonzero(c0 < thr)
    hi = hi[np.argsort(c0[hi], kind="stable")]
    # interleave: the low ones first in every shard's share, the sorted high ones spread evenly so that every shard holds chain elements
    order = np.empty(len(T), np.int64)
    slots = np.unique(np.linspace(0, len(T) - 1, len(hi)).astype(np.int64))
    lo = np.concatenate([lo, hi[len(slots):]])
    hi = hi[: len(slots)]
    order[slots] = hi
    order[np.setdiff1d(np.arange(len(T)), slots)] = lo
    T = T[order]
    results = []
    for devs in ([0], list(range(nd)), list(range(nd - 1, -1, -1))[:2]):
        g = PoseGroup(devs)
        g.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta)
        g.set_model(0, prob.model_xyz, prob.model_nrm)
        c, s = g.score_lcp(0, T, mode)
        results.append((c, s, g.topk(0, 64), g.improving_chain(0)))
        g.close()
    for r in results[1:]:
        assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1])
        assert r[2].tobytes() == results[0][2].tobytes()
        assert r[3].tobytes() == results[0][3].tobytes()
    assert len(results[0][3]) >= (20 if mode == "count" else 8)   # (the list is ordered by the count; the weighted chain follows the score)


@pytest.mark.skipif("_n_devices() < 2")
@pytest.mark.parametrize("pcs_mode", [0, 1])
def test_group_generation_is_sharding_invariant(pcs_mode):
    """configs[2] in small: bases split across the devices, each device generates and scores its own hypotheses; the global
    top-64 and chain equal the single-device request's, including under a global hypothesis cap."""
    nd = _n_devices()
    prob = synth.make_segment_problem(500, 700, 0.005, seed=23)
    out = []
    for devs in ([0], list(range(nd))):
        g = PoseGroup(devs)
        g.set_scene(prob.scene_xyz, prob.scene_nrm, prob.delta)
        g.set_model(0, prob.model_xyz, prob.model_nrm)
        if pcs_mode == 1:
            g.build_ppf_map(0)
        row = []
        for cap in (5_000_000, 1500):
            n = g.generate_pcs(0, seed=9, max_hyp=cap, n_bases=60, mode=pcs_mode)
            g.score_generated(0, "weighted")
            row.append((n, g.topk(0, 64), g.improving_chain(0)))
        out.append(row)
        g.close()
    for a, b in zip(out[0], out[1]):
        assert a[0] == b[0] and a[0] > 0
        assert a[1].tobytes() == b[1].tobytes()
        assert a[2].tobytes() == b[2].tobytes()
    assert out[0][1][0] == min(1500, out[0][0][0])          # the cap binds exactly like the single-device generator's
