#!/usr/bin/env python
"""Mints the configs[0] fixture (BASELINE.json: "test-scene frame-000000 (APC shelf) using the provided mask.png in place of
FCN, PCS hypotheses + LCP scoring, CPU reference") FROM THE REFERENCE'S OWN test-scene and engine.  Run where /root/reference
exists:

    make -C oracle ref && python tests/golden/make_c1.py

Inputs taken from the reference tree (read-only, nothing is copied verbatim into the repo):
  test-scene/frame-000000.{depth,mask}.png, gt_info.yml     the RGB-D frame, class mask (ids 2/3/8) and intrinsics
  src/physim_pose_estimation/models_visualization/<obj>.ply   meshes, sampled here into stand-in search/validation clouds
                                                              (the real model_search.ply files are a download)
Segment preparation restates, without PCL (PPE = src/physim_pose_estimation):
  depth decode     utilities::readDepthImage        PPE/src/misc/utilities.cpp:47-61   ((d << 13) | (d >> 3)) as u16, / 10000
  back-projection  utilities::convert3dUnOrganizedRGB  :210-228   fp32 ((v - cx) * depth) / fx, 0.1 < depth < 2.0
  mask             GTSegmentation::compute2dSegment  PPE/src/segmentation/Segmentation.cpp:187-207 (prior image = 10000 in the mask)
  1 cm voxel centroids, normals (local PCA in 2 cm instead of pcl::MovingLeastSquares), radius-outlier removal (3 cm, >= 10),
  normals flipped to the camera and re-normalised      Segmentation.cpp:211-252, PPE/src/hypothesis_generation/ObjectPoseCandidateSet.cpp:28-51
The three PCL filters are un-vendored third-party code, so these steps are stand-ins, not parity claims; what the fixture pins
is the engine: the reference's own Perform_N_steps (operMode 0) generates the hypotheses on these clouds and its own Verify /
WeightedVerify score them -- the CUDA path must reproduce those numbers on the same prepared clouds (SURVEY.md 8d "C1")."""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
from oracle.pyoracle import RefOracle  # noqa: E402
from oracle import segment_port  # noqa: E402

K = np.array([[6.13998108e+02, 0.0, 3.22453583e+02], [0.0, 6.13998169e+02, 2.39678940e+02], [0.0, 0.0, 1.0]], np.float32)   # gt_info.yml:4
OBJECTS = [(8, "kleenex_tissue_box"), (2, "expo_dry_erase_board_eraser"), (3, "folgers_classic_roast_coffee")]             # gt_info.yml:14-19, obj_config.yml


def read_mesh(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        nv = nf = 0
        while True:
            line = f.readline().strip()
            if line.startswith(b"element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith(b"element face"):
                nf = int(line.split()[-1])
            elif line == b"end_header":
                break
        v = np.frombuffer(f.read(nv * 12), "<f4").reshape(nv, 3).astype(np.float64)
        faces = np.frombuffer(f.read(nf * 13), np.dtype([("n", "u1"), ("i", "<i4", (3,))]))
        assert np.all(faces["n"] == 3)
        return v, faces["i"].astype(np.int64)


def sample_mesh(v, f, n, rng):
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    pick = rng.choice(len(f), size=n, p=area / area.sum())
    r1, r2 = np.sqrt(rng.uniform(size=n)), rng.uniform(size=n)
    p = (1 - r1)[:, None] * a[pick] + (r1 * (1 - r2))[:, None] * b[pick] + (r1 * r2)[:, None] * c[pick]
    nrm = cr[pick] / np.linalg.norm(cr[pick], axis=1, keepdims=True)
    # outward orientation (the meshes are closed and roughly star-shaped around their centroid)
    s = np.sign(np.einsum("ij,ij->i", nrm, p - v.mean(axis=0)))
    return p.astype(np.float32), (nrm * np.where(s == 0, 1, s)[:, None]).astype(np.float32)


def _rle(a):
    """run-length encoding (values, run lengths) of a flat array"""
    starts = np.flatnonzero(np.r_[True, a[1:] != a[:-1]])
    return a[starts].astype(np.int32), np.diff(np.r_[starts, len(a)]).astype(np.int32)


def main():
    raw = np.array(Image.open(os.path.join(REF, "test-scene", "frame-000000.depth.png"))).astype(np.uint16)
    dec = ((raw << np.uint16(13)) | (raw >> np.uint16(3))).astype(np.uint16)
    depth_m = segment_port.decode_depth(raw)
    mask = np.array(Image.open(os.path.join(REF, "test-scene", "frame-000000.mask.png"))).astype(np.uint8)
    assert depth_m.shape == mask.shape == (480, 640)
    rng = np.random.default_rng(2024)
    out = dict(K=K, delta=np.float64(0.005), depth_raw_crop=raw[200:216, 300:316], depth_dec_crop=dec[200:216, 300:316],
               depth_raw_rle=np.stack(_rle(raw.ravel())), mask_all_rle=np.stack(_rle(mask.ravel())))
    for cls, name in OBJECTS:
        seg_xyz, seg_nrm, n_raw = segment_port.prepare_segment(depth_m, mask, cls, K)
        v, f = read_mesh(os.path.join(REF, "src/physim_pose_estimation/models_visualization", name + ".ply"))
        mod_xyz, mod_nrm = sample_mesh(v, f, 1500, rng)
        prior_img = np.where(mask == cls, 10000, 0).astype(np.uint16)
        ref = RefOracle(seg_xyz, seg_nrm, mod_xyz, mod_nrm, mod_xyz, mod_nrm, 0.005, K=K, prior_img=prior_img)
        res = ref.perform_n_steps(mode=0, seed=7)                      # the reference's own generator + running-best scan
        T = res["transforms"]
        if len(T) > 1500:
            T = T[np.sort(rng.choice(len(T), 1500, replace=False))]
        counts = ref.verify(T)
        ws, wn = ref.weighted_verify(T)
        print(f"{name}: {n_raw} valid-depth pixels -> {len(seg_xyz)} segment points, model extent {np.ptp(mod_xyz, axis=0)}, "
              f"{len(res['transforms'])} reference hypotheses ({len(T)} kept), best count {counts.max()}/1500, best weighted {ws.max():.4f}, "
              f"reference chain length {len(res['chain_score'])} best LCP {res['best_lcp']:.4f}")
        out.update({f"{name}_seg_xyz": seg_xyz, f"{name}_seg_nrm": seg_nrm, f"{name}_model_xyz": mod_xyz, f"{name}_model_nrm": mod_nrm,
                    f"{name}_mask_rle": np.flatnonzero(np.diff(np.r_[0, (mask == cls).ravel().astype(np.int8), 0])).astype(np.int32),
                    f"{name}_priors": ref.priors(), f"{name}_T": T, f"{name}_counts": counts, f"{name}_wscore": ws, f"{name}_wnreg": wn,
                    f"{name}_n_raw": np.int64(n_raw)})
    np.savez_compressed(os.path.join(HERE, "c1_test_scene.npz"), names=np.array([n for _, n in OBJECTS]), **out)
    print("written", os.path.join(HERE, "c1_test_scene.npz"), os.path.getsize(os.path.join(HERE, "c1_test_scene.npz")), "bytes")


if __name__ == "__main__":
    main()
