#!/usr/bin/env python
"""Mints golden vectors for the PCS -> LCP path FROM THE REFERENCE ENGINE ITSELF
(oracle/_ref/libs4ref.so = the reference sources compiled in place from /root/reference by
oracle/Makefile).  Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

The reference's own test-suite holds no golden vectors for this path, so these
files are what pins the oracle restatement (oracle/lcp_oracle.c) and the CUDA path on machines
that do not have the reference tree.  Inputs are stored next to the outputs, so nothing has to be
regenerated bit-identically elsewhere."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.pyoracle import RefOracle  # noqa: E402
from physimglobalpose_b200 import synth  # noqa: E402


def main():
    # ---- LCP: count + weighted (with and without a prior image), centring, chain
    prob = synth.make_problem(300, 6000, 0.01, seed=101)
    T = synth.make_hypotheses(prob, 96, seed=102)
    rng = np.random.default_rng(103)
    perm = rng.permutation(len(T))
    T = T[perm]                                   # GT not first: the improving chain gets longer
    args = (prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
    ref = RefOracle(*args)
    counts = ref.verify(T)
    frac, best_index = ref.verify_running_best(T)
    ws, wn, reg = ref.weighted_verify(T, reg_of=int(np.argmax(counts)))
    cP, cQ = ref.centroids()
    img = rng.integers(0, 10001, size=(480, 640)).astype(np.uint16)
    K = np.array([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]], np.float32)
    ref_img = RefOracle(*args, K=K, prior_img=img)
    priors = ref_img.priors()
    ws_img, wn_img = ref_img.weighted_verify(T)
    np.savez_compressed(os.path.join(HERE, "lcp_small.npz"), scene_xyz=prob.scene_xyz, scene_nrm=prob.scene_nrm, model_xyz=prob.model_xyz,
                        model_nrm=prob.model_nrm, delta=np.float64(prob.delta), T=T, counts=counts, running_frac=frac,
                        running_best=np.int64(best_index), weighted_score=ws, weighted_nreg=wn, registered_of=np.int64(np.argmax(counts)),
                        registered=reg, cP=cP, cQ=cQ, prior_img=img, K=K, priors=priors, weighted_score_img=ws_img, weighted_nreg_img=wn_img)

    # ---- a second delta (the shipped default 5 mm) on the same clouds
    ref5 = RefOracle(*args[:-1], 0.005)
    np.savez_compressed(os.path.join(HERE, "lcp_small_d5.npz"), counts=ref5.verify(T), weighted_score=ref5.weighted_verify(T)[0])

    # ---- PCS: pair extraction, quad join, rigid transforms for a few bases (operMode 0 pieces) on a
    # segment-sized request (the scene is one object's visible surface)
    seg = synth.make_segment_problem(600, 1500, 0.005, seed=201)
    sargs = (seg.scene_xyz, seg.scene_nrm, seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.delta)
    ref = RefOracle(*sargs)
    pairs_out, quads_out, rigid_out = {}, {}, {}
    bases, invs = [], []
    for seed in range(1, 400):
        ok, b, inv = ref.select_quadrilateral(seed)
        if not ok:
            continue
        cP_xyz, _ = ref.centred(0)
        d1 = np.float32(np.linalg.norm(cP_xyz[b[0]] - cP_xyz[b[1]]))
        d2 = np.float32(np.linalg.norm(cP_xyz[b[2]] - cP_xyz[b[3]]))
        p1 = ref.extract_pairs(d1, np.float32(seg.delta))
        p2 = ref.extract_pairs(d2, np.float32(seg.delta))
        if len(p1) == 0 or len(p2) == 0:
            continue
        q = ref.find_quads(b, inv[0], inv[1], np.float32(seg.delta), p1, p2)
        if len(q) == 0:
            continue
        k = len(bases)
        bases.append(b.copy()); invs.append(inv.copy())
        qs = q[:256]
        Ts, oks, poses = [], [], []
        for quad in qs:
            ok2, T4, P4 = ref.rigid_from_quad(b, quad)
            oks.append(ok2); Ts.append(T4[:3]); poses.append(P4)
        pairs_out[f"b{k}_d1"] = d1; pairs_out[f"b{k}_d2"] = d2
        pairs_out[f"b{k}_p1"] = p1; pairs_out[f"b{k}_p2"] = p2
        quads_out[f"b{k}_quads"] = q
        rigid_out[f"b{k}_T"] = np.array(Ts, np.float32).reshape(-1, 3, 4); rigid_out[f"b{k}_ok"] = np.array(oks, bool)
        rigid_out[f"b{k}_pose"] = np.array(poses, np.float64).reshape(-1, 4, 4)
        if len(bases) == 4:
            break
    scP, scQ = ref.centroids()
    np.savez_compressed(os.path.join(HERE, "pcs_small.npz"), scene_xyz=seg.scene_xyz, scene_nrm=seg.scene_nrm, model_xyz=seg.model_xyz,
                        model_nrm=seg.model_nrm, delta=np.float64(seg.delta), cP=scP, cQ=scQ, bases=np.array(bases, np.int32), invariants=np.array(invs, np.float32),
                        **pairs_out, **quads_out, **rigid_out)
    mint_stocs()
    mint_mode2()
    mint_full_shapes()
    print("golden vectors written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


def mint_stocs():
    """operMode 1 (the shipped generator): computePPF keys, the model's PPF map, SelectQuadrilateralStoCS draws with the
    engine seed pinned (oracle/Makefile, second patch) and ExtractCongruentSet in mode 1 -- all from the reference itself."""
    from oracle import pyoracle
    from physimglobalpose_b200 import _lib
    lib = _lib.load()                                    # only for the host-side seed derivation pgp_stocs_engine_seed
    seg = synth.make_segment_problem(300, 400, 0.005, seed=301)
    n = len(seg.model_xyz)
    # the model's map: the reference's computePPF over all ordered pairs of the model cloud (an oracle whose P is the model)
    mref = RefOracle(seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.delta)
    allp = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2).astype(np.int32)
    keys4, offsets, pairs = pyoracle.group_ppf_keys(mref.compute_ppf(allp), n)
    ref = RefOracle(seg.scene_xyz, seg.scene_nrm, seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.delta)
    ref.set_ppf_map(keys4, offsets, pairs)
    rng = np.random.default_rng(302)
    sp = rng.integers(0, len(seg.scene_xyz), size=(4000, 2)).astype(np.int32)
    scene_keys = ref.compute_ppf(sp)
    user_seed, n_bases = 77, 48
    ok_l, ids_l, inv_l, att_l = [], [], [], []
    for b in range(n_bases):
        ok, ids, inv, used = False, np.zeros(4, np.int32), np.zeros(2, np.float32), 16
        for a in range(16):
            ok, ids, inv = ref.select_stocs(int(lib.pgp_stocs_engine_seed(user_seed, b, a)))
            if ok:
                used = a
                break
        ok_l.append(ok); ids_l.append(ids.copy() if ok else np.zeros(4, np.int32)); inv_l.append(inv.copy() if ok else np.zeros(2, np.float32)); att_l.append(used)
    quads = {}
    kept = 0
    for b in range(n_bases):
        if not ok_l[b] or kept >= 6:
            continue
        q = ref.congruent_set_mode1(ids_l[b], float(inv_l[b][0]), float(inv_l[b][1]))
        if 0 < len(q) <= 20000:
            quads[f"quads_b{b}"] = q
            kept += 1
    np.savez_compressed(os.path.join(HERE, "stocs_small.npz"), scene_xyz=seg.scene_xyz, scene_nrm=seg.scene_nrm, model_xyz=seg.model_xyz,
                        model_nrm=seg.model_nrm, delta=np.float64(seg.delta), map_keys=keys4, map_offsets=offsets, map_pairs=pairs,
                        scene_pairs=sp, scene_keys=scene_keys, user_seed=np.int64(user_seed), base_ok=np.array(ok_l, bool),
                        base_ids=np.array(ids_l, np.int32), base_inv=np.array(inv_l, np.float32), base_attempts=np.array(att_l, np.int32), **quads)


def mint_mode2():
    """operMode 2 (V4PCS): tetrahedron bases drawn by the reference's SelectTetrahedronBase after srand(seed), and the congruent
    quads its ExtractCongruentSet finds for them (six ExtractPairs + FindCongruentQuadrilateralsV4PCS) -- from the reference itself."""
    seg = synth.make_segment_problem(300, 400, 0.005, seed=401)
    ref = RefOracle(seg.scene_xyz, seg.scene_nrm, seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.delta)
    bases, quads, offs = [], [], [0]
    for seed in range(1, 13):
        ok, b = ref.select_tetrahedron(seed)
        assert ok
        q = ref.congruent_set_mode2(b)
        q = np.array(sorted(map(tuple, q.tolist())), np.int32).reshape(-1, 4)      # the reference's order is an unordered_set iteration
        bases.append(b); quads.append(q); offs.append(offs[-1] + len(q))
    out = os.path.join(HERE, "mode2_small.npz")
    np.savez_compressed(out, scene_xyz=seg.scene_xyz, scene_nrm=seg.scene_nrm, model_xyz=seg.model_xyz, model_nrm=seg.model_nrm,
                        delta=np.float64(seg.delta), bases=np.array(bases, np.int32), quads=np.concatenate(quads), quad_offsets=np.array(offs, np.int64))
    print("mode2_small.npz:", len(bases), "bases,", offs[-1], "quads,", os.path.getsize(out), "bytes")


def mint_full_shapes():
    """BASELINE configs[1] and configs[4] shapes at FULL cloud sizes, from the reference itself: the clouds come from the seeded
    generator (synth.make_problem, numpy only: identical on every machine), so only the hypothesis sample and the reference's
    numbers are stored.  Verify / WeightedVerify of Match4PCSBase on a sample of hypotheses spread over the benchmark's batch."""
    out = {}
    for tag, (nm, ns, n_hyp, n_sample, seed_p, seed_t) in {"c2": (2000, 100000, 100000, 400, 1234, 4321), "c5": (30000, 300000, 2000, 60, 51, 52)}.items():
        prob = synth.make_problem(nm, ns, 0.01, seed=seed_p)
        T = synth.make_hypotheses(prob, n_hyp, seed=seed_t)
        idx = np.unique(np.r_[0, 1, 2, np.random.default_rng(7).choice(n_hyp, n_sample - 3, replace=False)]).astype(np.int64)
        ref = RefOracle(prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
        counts = ref.verify(T[idx])
        ws, wn = ref.weighted_verify(T[idx])
        out.update({f"{tag}_shape": np.array([nm, ns, n_hyp, seed_p, seed_t], np.int64), f"{tag}_idx": idx, f"{tag}_counts": counts,
                    f"{tag}_wscore": ws, f"{tag}_wnreg": wn, f"{tag}_T_check": T[idx[:4]]})
        print(tag, "reference counts on", len(idx), "hypotheses; max", counts.max(), "at", idx[counts.argmax()])
    np.savez_compressed(os.path.join(HERE, "lcp_full_shapes.npz"), **out)


if __name__ == "__main__":
    main()
