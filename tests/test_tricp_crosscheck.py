"""Independent cross-check of the trimmed-ICP restatement (oracle/lcp_oracle.c::lo_tricp).

pcl::recognition::TrimmedICP is not vendored and PCL is not installed, so lo_tricp is "restated, parity unpinned" (SURVEY.md 8c).
What CAN be done without PCL is to make sure a bug of the restatement cannot hide behind that label: the same published algorithm
(Chetverikov et al.; PCL trimmed_icp.h: NN of every transformed source point in the target, keep the n smallest squared distances,
energy = their sum, Umeyama / SVD without scale on the kept pairs using the ORIGINAL source coordinates, loop while
energy / old_energy < ratio) is written a second time with entirely different building blocks -- scipy's cKDTree for the exact
1-NN, numpy's SVD for the rotation, float64 throughout -- and the two must walk the same iterations to the same pose.
The device kernel K5 is compared with lo_tricp in tests/test_gpu_pcs_tricp.py; this test closes the other side of that chain."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from physimglobalpose_b200 import synth


def tricp_numpy(src, tgt, T0, trim=0.5, ratio=0.99, max_iter=100):
    src = np.asarray(src, np.float64); tgt = np.asarray(tgt, np.float64)
    T = np.asarray(T0, np.float64).reshape(3, 4).copy()
    tree = cKDTree(tgt)
    n_keep = min(len(src), int(abs(np.float32(trim) * np.float32(len(src)))))      # abs(numPoints) on a float (UCTState.cpp:181,194)
    energy, it = np.finfo(np.float32).max, 0
    while True:
        q = src @ T[:, :3].T + T[:, 3]
        d, idx = tree.query(q)
        order = np.argsort(d * d, kind="stable")[:n_keep]
        old, energy = energy, float(np.sum(d[order] ** 2))
        s, g = src[order], tgt[idx[order]]
        cs, cg = s.mean(0), g.mean(0)
        H = (s - cs).T @ (g - cg)
        U, _, Vt = np.linalg.svd(H)
        D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
        R = Vt.T @ D @ U.T
        T = np.concatenate([R, (cg - R @ cs)[:, None]], axis=1)
        it += 1
        if not (energy / old < ratio and it < max_iter):
            return T, it, energy


@pytest.mark.parametrize("seed", [3, 4, 5])
@pytest.mark.parametrize("trim", [0.5, 0.9])
def test_lo_tricp_equals_an_independent_implementation(port_lib, seed, trim):
    prob = synth.make_segment_problem(800, 600, 0.005, seed=seed)
    rng = np.random.default_rng(seed)
    # start a few mm / degrees off the ground truth, in the scene -> model direction the call site uses (tform = inverse pose)
    R = synth.rot_axis_angle(rng.normal(size=3), 0.06)
    pose = prob.gt_pose.copy()
    pose[:3, :3] = pose[:3, :3] @ R
    pose[:3, 3] += rng.normal(0, 0.004, size=3)
    inv = np.linalg.inv(pose)[:3, :]
    o = port_lib.PortOracle(prob.scene_xyz, prob.scene_nrm, prob.model_xyz, prob.model_nrm, prob.model_xyz, prob.model_nrm, prob.delta)
    T_c, it_c, e_c = o.tricp(prob.scene_xyz, prob.model_xyz, inv.astype(np.float32), trim=trim, ratio=0.99, max_iter=100)
    T_n, it_n, e_n = tricp_numpy(prob.scene_xyz, prob.model_xyz, inv.astype(np.float32), trim=trim, ratio=0.99, max_iter=100)
    assert it_c == it_n, (it_c, it_n)
    assert abs(e_c - e_n) <= 1e-4 * max(e_n, 1e-9)
    dt = np.linalg.norm(T_c[:, 3] - T_n[:, 3])
    ang = np.linalg.norm(T_c[:, :3].astype(np.float64) - T_n[:, :3]) / np.sqrt(2.0)      # small-angle: |R_a - R_b|_F = sqrt(2) * angle
    assert dt < 1e-5 and ang < 1e-5, (dt, ang)                      # well inside north_star's 1e-4 m / 1e-4 rad (the port returns fp32)
    # and the refinement did its job: the trimmed energy went down from the first iteration's
    _, _, e_first = tricp_numpy(prob.scene_xyz, prob.model_xyz, inv.astype(np.float32), trim=trim, ratio=0.99, max_iter=1)
    assert it_c >= 2 and e_c < e_first
