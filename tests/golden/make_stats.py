#!/usr/bin/env python
"""Mints tests/golden/pcs_stats.npz: the DISTRIBUTION of what the reference's own Perform_N_steps returns over seeds
(S4/algorithms/match4pcsBase.cc:1823-1927; SURVEY.md 7: generation is random on both sides -- rand() / the engine seed there,
counter-based hashes here -- so end-to-end outputs can only be compared as statistics).

Per seed and operMode (0 = Super4PCS pairs + Verify, 1 = StoCS + PPF map + WeightedVerify, the shipped mode): best LCP, length of
the improving chain (hypothesisSet), number of transforms verified, pose error of the best hypothesis against the ground truth.
The request is a test-scene-sized object (500-pt model, 700-pt camera-visible segment, delta = 5 mm); the clouds are regenerated
from the seeded generator, the file holds only the statistics.  Run here (needs /root/reference via oracle/_ref):
    python tests/golden/make_stats.py"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from oracle import pyoracle, stocs_port
from physimglobalpose_b200 import synth

N_SEEDS = 24
PROBLEM = dict(n_model=500, n_segment=700, delta=0.005, seed=23)


def main():
    seg = synth.make_segment_problem(**PROBLEM)
    keys, offs, pairs = stocs_port.build_ppf_map(seg.model_xyz - synth.seq_centroid_f32(seg.model_xyz), seg.model_nrm)
    out = {}
    for mode in (0, 1):
        best, chain, ntr, terr, rerr, secs = [], [], [], [], [], []
        for s in range(1, N_SEEDS + 1):
            ref = pyoracle.RefOracle(seg.scene_xyz, seg.scene_nrm, seg.model_xyz, seg.model_nrm, seg.model_xyz, seg.model_nrm, seg.delta, srand_seed=s)
            if mode == 1:
                ref.set_ppf_map(keys, offs, pairs)
            t0 = time.perf_counter()
            r = ref.perform_n_steps(mode=mode, seed=s)
            secs.append(time.perf_counter() - t0)
            best.append(r["best_lcp"]); chain.append(len(r["chain_score"])); ntr.append(len(r["transforms"]))
            dt, da = synth.pose_error(r["chain_pose"][-1], seg.gt_pose) if len(r["chain_pose"]) else (np.inf, np.inf)
            terr.append(dt); rerr.append(da)
            print(f"mode {mode} seed {s}: best {best[-1]:.4f} chain {chain[-1]} transforms {ntr[-1]} err {dt * 1e3:.1f} mm {np.degrees(da):.1f} deg  {secs[-1]:.2f} s", flush=True)
        out[f"mode{mode}_best"] = np.array(best, np.float32); out[f"mode{mode}_chain"] = np.array(chain, np.int32)
        out[f"mode{mode}_transforms"] = np.array(ntr, np.int32); out[f"mode{mode}_terr"] = np.array(terr); out[f"mode{mode}_rerr"] = np.array(rerr)
        out[f"mode{mode}_seconds"] = np.array(secs)
    out["problem"] = np.array([PROBLEM["n_model"], PROBLEM["n_segment"], PROBLEM["seed"]], np.int64)
    out["delta"] = np.float32(PROBLEM["delta"])
    np.savez_compressed(os.path.join(HERE, "pcs_stats.npz"), **out)


if __name__ == "__main__":
    main()
