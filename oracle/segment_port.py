"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the segment preparation that precedes the PCS -> LCP path
(PPE = /root/reference/src/physim_pose_estimation):
  depth decode     utilities::readDepthImage           PPE/src/misc/utilities.cpp:47-61   ((d << 13) | (d >> 3)) as u16, / 10000
  mask             GTSegmentation::compute2dSegment    PPE/src/segmentation/Segmentation.cpp:187-207
  back-projection  utilities::convert3dUnOrganizedRGB  PPE/src/misc/utilities.cpp:210-228  fp32 ((v - cx) * depth) / fx, 0.1 < depth < 2.0
  voxel centroids  pcl::VoxelGrid, leaf 1 cm           Segmentation.cpp:226-229  (centroid per voxel, output in voxel-index order)
  normals          pcl::MovingLeastSquares, r = 2 cm   Segmentation.cpp:231-238  -- restated as local PCA (smallest eigenvector)
  outlier removal  pcl::RadiusOutlierRemoval 3 cm / 10 PPE/src/hypothesis_generation/ObjectPoseCandidateSet.cpp:28-32
  normal flip      pcl::flipNormalTowardsViewpoint + renormalise  :39-51
PCL is not vendored in the reference tree and not installed: the three PCL filters are RESTATED, PARITY UNPINNED (VoxelGrid's
index convention and MLS's polynomial projection in particular).  The device kernels (k7_segment.cu) are checked against THIS file;
tests/golden/make_c1.py uses it to prepare the configs[0] fixture.  Only tests/ and tests/golden/ import it."""
from __future__ import annotations

import numpy as np


def decode_depth(raw_u16: np.ndarray) -> np.ndarray:
    raw = raw_u16.astype(np.uint16)
    dec = ((raw << np.uint16(13)) | (raw >> np.uint16(3))).astype(np.uint16)
    return (dec.astype(np.float32) / np.float32(10000)).astype(np.float32)


def backproject(depth_m: np.ndarray, mask: np.ndarray, cls: int, K: np.ndarray) -> np.ndarray:
    obj = np.where(mask == cls, depth_m, np.float32(0))
    o64 = obj.astype(np.float64)                                     # `depth > 0.1 && depth < 2.0` promotes the float to double
    u, v = np.nonzero((o64 > 0.1) & (o64 < 2.0))                     # row-major pixel order, like the double loop
    d = obj[u, v].astype(np.float32)
    x = ((v.astype(np.float32) - K[0, 2]) * d / K[0, 0]).astype(np.float32)
    y = ((u.astype(np.float32) - K[1, 2]) * d / K[1, 1]).astype(np.float32)
    return np.stack([x, y, d], axis=1)


def voxel_centroids(pts: np.ndarray, leaf: float = 0.01) -> np.ndarray:
    leaf = np.float32(leaf)
    ijk = np.floor(pts / leaf).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(key, kind="stable")
    key, pts = key[order], pts[order]
    starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    cnt = np.diff(np.r_[starts, len(key)])
    return (np.add.reduceat(pts.astype(np.float64), starts, axis=0) / cnt[:, None]).astype(np.float32)


def pca_normals(cen: np.ndarray, radius: float = 0.02) -> np.ndarray:
    from scipy.spatial import cKDTree
    tree = cKDTree(cen)
    nrm = np.zeros_like(cen)
    for i, nb in enumerate(tree.query_ball_point(cen, radius)):
        q = cen[nb].astype(np.float64)
        if len(nb) >= 3:
            w, vec = np.linalg.eigh(np.cov((q - q.mean(axis=0)).T))
            n = vec[:, 0]
        else:
            n = -cen[i].astype(np.float64)
        if np.dot(n, cen[i]) > 0:
            n = -n
        nrm[i] = (n / np.linalg.norm(n)).astype(np.float32)
    return nrm


def radius_outlier_keep(cen: np.ndarray, radius: float = 0.03, min_neighbors: int = 10) -> np.ndarray:
    from scipy.spatial import cKDTree
    tree = cKDTree(cen)
    return np.array([len(nb) >= min_neighbors for nb in tree.query_ball_point(cen, radius)])       # the point itself counts, as in PCL


def prepare_segment(depth_m, mask, cls, K, leaf=0.01, normal_radius=0.02, outlier_radius=0.03, min_neighbors=10):
    pts = backproject(depth_m, mask, cls, K)
    cen = voxel_centroids(pts, leaf)
    nrm = pca_normals(cen, normal_radius)
    keep = radius_outlier_keep(cen, outlier_radius, min_neighbors)
    return cen[keep], nrm[keep], len(pts)
