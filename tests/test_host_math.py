"""CPU property tests of the conservative bounds the CUDA path relies on (numpy restatements of the device formulas, fp32
where the device uses fp32).  They do not run the kernels -- tests/test_gpu_lcp.py does, against the oracle -- they check that
the MATH of the bounds is sound on random inputs:
  * domination pruning of candidate lists   (physimglobalpose_b200/csrc/k1_fine.cu: dominated(), k1f_classify / k1w_count)
  * spectral-norm bound of the group cull    (physimglobalpose_b200/csrc/k3_lcp.cu: score_hypothesis, `snorm`)
  * separable gap distance field K1d         (physimglobalpose_b200/csrc/k1_fine.cu: k1d_pass)
  * tri-state voxel labels K1b               (physimglobalpose_b200/csrc/k1_fine.cu: k1f_classify, k1_build_fine)"""
import numpy as np

f32 = np.float32


def _d2(q, p):
    """(q - p).squaredNorm() in the reference's association, fp32: dx^2 + (dy^2 + dz^2)  (kdtree.h:423)."""
    d = (q - p).astype(f32)
    return f32(d[..., 0] * d[..., 0]) + (f32(d[..., 1] * d[..., 1]) + f32(d[..., 2] * d[..., 2]))


def _dominated(p, b, hs, c2p, c2b, margin):
    spread = f32(2.0) * hs * (abs(b[0] - p[0]) + abs(b[1] - p[1]) + abs(b[2] - p[2]))
    return (c2p - c2b) - spread > margin + f32(1e-5) * (c2p + c2b)


def test_domination_pruning_keeps_existence_and_the_nearest_point():
    rng = np.random.default_rng(1)
    delta = f32(0.01)
    r2 = f32(delta * delta)
    dhi2 = f32((0.01 * (1 + 1e-5)) ** 2)
    margin = f32(1e-5) * dhi2
    hs = f32(0.5 * 0.00125 + 7e-6)                    # half a 1.25 mm voxel + inflate
    removed_total = kept_total = 0
    for trial in range(300):
        c = rng.uniform(-0.3, 0.3, 3).astype(f32)
        n = int(rng.integers(2, 40))
        # points on a jittered plane patch at 0..12 mm from the voxel: the regime of the AMBIG / IN voxels
        off = rng.uniform(0.0, 0.012)
        pts = (c + np.c_[rng.uniform(-0.015, 0.015, (n, 2)), np.full(n, off) + rng.uniform(-0.001, 0.001, n)]).astype(f32)
        c2 = _d2(pts, c)
        # online rule of k1f_classify: dominator = closest-to-centre candidate seen so far
        kept, best, bc2 = [], None, f32(np.inf)
        for i in range(n):
            a = np.abs(pts[i] - c)
            lo = np.maximum(a - hs, 0).astype(f32)
            if f32(lo @ lo) > dhi2:
                continue
            if best is None or not _dominated(pts[i], pts[best], hs, c2[i], bc2, margin):
                kept.append(i)
            if c2[i] < bc2:
                bc2, best = c2[i], i
        # exact-dominator rule of k1w_count: dominator = the closest-to-centre point of all
        star = int(np.argmin(c2))
        # ... on top of K1c's own bound: lo(p)^2 <= min(min_p' hi(p')^2 (1 + 4e-5), (delta (1 + 1e-5))^2)   (wlist_threshold)
        a_all = np.abs(pts - c).astype(f32)
        lo_all = np.maximum(a_all - hs, 0).astype(f32); hi_all = (a_all + hs).astype(f32)
        lo2 = (lo_all * lo_all).sum(1).astype(f32); hi2 = (hi_all * hi_all).sum(1).astype(f32)
        thr = min(f32(hi2.min() * f32(1.0 + 4e-5)), dhi2)
        kept_w = [i for i in range(n) if lo2[i] <= thr and not _dominated(pts[i], pts[star], hs, c2[i], c2[star], margin)]
        removed_total += n - len(kept_w); kept_total += len(kept_w)
        q = (c + rng.uniform(-float(hs), float(hs), (400, 3))).astype(f32)
        d2 = _d2(q[:, None, :], pts[None, :, :])                               # (400, n)
        exists = (d2 <= r2).any(1)
        if kept:
            assert np.array_equal(exists, (d2[:, kept] <= r2).any(1))          # Verify: existence unchanged
        else:
            assert not exists.any()
        # WeightedVerify: nearest in-range point (ties -> smaller index) is never pruned
        masked = np.where(d2 <= r2, d2, np.inf)
        has = np.isfinite(masked).any(1)
        nearest = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), masked), axis=1)[:, 0]
        assert set(nearest[has].tolist()) <= set(kept_w)
    assert removed_total > kept_total                                           # and the pruning does prune


def test_spectral_norm_bound_of_the_group_cull():
    rng = np.random.default_rng(2)
    for trial in range(2000):
        A = rng.normal(size=(3, 3)) * rng.uniform(0.01, 5.0)
        if trial % 3 == 0:                                                      # rotations: the bound must be tight (~1)
            A, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        A = A.astype(f32)
        G = np.abs((A.T.astype(f32) @ A).astype(f32))
        snorm = f32(np.sqrt(G.sum(1).max())) * f32(1 + 1e-5)
        smax = np.linalg.svd(A.astype(np.float64), compute_uv=False)[0]
        assert snorm >= smax * (1 - 1e-6)
        if trial % 3 == 0:
            assert snorm < 1.001


def test_gap_distance_field_is_a_lower_bound():
    rng = np.random.default_rng(3)
    W, h = 6, 0.005
    for trial in range(20):
        dims = rng.integers(6, 14, 3)
        pts = rng.uniform(0, 1, (int(rng.integers(1, 12)), 3)) * dims * h * 0.999
        cell = np.floor(pts / h).astype(int)
        occ = np.zeros(dims, bool)
        occ[cell[:, 0], cell[:, 1], cell[:, 2]] = True
        S = np.where(occ, 0, W * W)
        for ax in range(3):                                                     # the three k1d_pass sweeps
            out = np.full_like(S, W * W)
            for t in range(-W, W + 1):
                gap = max(abs(t) - 1, 0)
                sh = np.full_like(S, W * W)
                src = [slice(None)] * 3; dst = [slice(None)] * 3
                if t >= 0:
                    src[ax] = slice(t, None); dst[ax] = slice(0, S.shape[ax] - t)
                else:
                    src[ax] = slice(0, t); dst[ax] = slice(-t, None)
                sh[tuple(dst)] = S[tuple(src)]
                out = np.minimum(out, sh + gap * gap)
            S = np.minimum(out, W * W)
        # brute force: min over occupied cells of the squared gap vector, capped
        idx = np.stack(np.meshgrid(*[np.arange(d) for d in dims], indexing="ij"), -1)
        occ_idx = np.argwhere(occ)
        gaps = np.maximum(np.abs(idx[..., None, :] - occ_idx) - 1, 0)
        brute = np.minimum((gaps ** 2).sum(-1).min(-1), W * W)
        assert np.array_equal(S, brute)
        # and it bounds the true distance from any position inside a cell to any scene point from below
        q = rng.uniform(0, 1, (500, 3)) * dims * h * 0.999
        qc = np.floor(q / h).astype(int)
        true = np.sqrt(((q[:, None, :] - pts[None]) ** 2).sum(-1)).min(1)
        lower = h * np.sqrt(S[qc[:, 0], qc[:, 1], qc[:, 2]])
        assert np.all(lower <= true + 1e-12)


def test_tristate_labels_are_conservative():
    """The IN / OUT labels of K1b (k1_fine.cu: k1f_classify, thresholds and `inflate` from k1_build_fine) restated in numpy on a
    small scene: for queries anywhere inside a voxel's INFLATED box -- where K3's fast FMA transform may place a query whose
    reference-rounded position the exact test uses -- an IN voxel always has a scene point with fp32 d2 <= delta^2 and an OUT voxel
    never has one, so only AMBIG voxels need the exact test."""
    rng = np.random.default_rng(5)
    delta = f32(0.01)
    r2 = f32(delta * delta)
    # scene: a jittered plane patch and a small box corner, centred like the engine does
    n = 400
    pts = np.c_[rng.uniform(-0.06, 0.06, (n, 2)), rng.uniform(-0.001, 0.001, n)]
    pts[: n // 3, 2] += 0.03 + rng.uniform(0, 0.02, n // 3)
    pts = (pts - pts.mean(0)).astype(f32)
    h = f32(delta * f32(1.0 + 1.0 / 256.0))
    lo = (pts.min(0) - f32(2.5) * h).astype(f32)
    dim = ((pts.max(0).astype(np.float64) - lo) / h).astype(int) + 4
    F = 8
    hf = f32(h / F)
    maxabs = f32(max(np.abs(lo).max(), np.abs(lo + h * dim.astype(f32)).max()))
    pos_bound = f32(3.0) * maxabs
    eps_pos = f32(32.0 * 5.9604645e-8) * (pos_bound + maxabs)
    inflate = hf * (f32(2e-3) + f32(16.0 * 5.9604645e-8) * f32(dim.max() * F)) + eps_pos
    dlo2 = f32((0.01 * (1 - 1e-5)) ** 2)
    dhi2 = f32((0.01 * (1 + 1e-5)) ** 2)
    hs = f32(0.5) * hf + inflate
    # voxels: a random sample of those within ~2 delta of some scene point (elsewhere everything is trivially OUT)
    seeds = pts[rng.integers(0, n, 6000)] + rng.uniform(-0.02, 0.02, (6000, 3)).astype(f32)
    vidx = np.floor((seeds - lo) / hf).astype(np.int64)
    c = (lo + (vidx.astype(f32) + f32(0.5)) * hf).astype(f32)                           # voxel centres
    a = np.abs(pts[None, :, :] - c[:, None, :]).astype(f32)                             # (voxels, points, 3)
    lo_d = np.maximum(a - hs, 0).astype(f32)
    hi_d = (a + hs).astype(f32)
    mind2 = (lo_d * lo_d).sum(2).astype(f32)
    maxd2 = (hi_d * hi_d).sum(2).astype(f32)
    is_in = (maxd2 <= dlo2).any(1)
    is_out = (mind2 > dhi2).all(1)
    assert is_in.sum() > 500 and is_out.sum() > 500 and (~is_in & ~is_out).sum() > 200    # all three classes are exercised
    # queries anywhere in the inflated box, including its corners
    for rep in range(6):
        e = rng.uniform(-1, 1, c.shape) if rep else np.sign(rng.uniform(-1, 1, c.shape))
        q = (c + e.astype(f32) * hs).astype(f32)
        exists = (_d2(q[:, None, :], pts[None, :, :]) <= r2).any(1)
        assert exists[is_in].all()
        assert not exists[is_out].any()
