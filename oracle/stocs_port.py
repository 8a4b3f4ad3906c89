"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's operMode-1 base sampler.

Follows, statement by statement (S4 = /root/reference/src/3rdparty/super4pcs/src/super4pcs):
  compute_ppf        Match4PCSBase::computePPF            S4/algorithms/match4pcsBase.cc:582-598, approximate_bin :150-160
  select_stocs       Match4PCSBase::SelectQuadrilateralStoCS  :600-792
  MinStd / discrete_draw   std::default_random_engine (minstd_rand0) + std::discrete_distribution<int> as libstdc++ 13
                     evaluates them (bits/random.h, bits/random.tcc: accumulate -> normalize -> partial_sum -> lower_bound
                     of generate_canonical<double, 53>)
  try_quadrilateral  Match4PCSBase::TryQuadrilateral :415-464 with distSegmentToSegment :81-148 (Scalar = double)

Pinned against the compiled reference (oracle/_ref, built with the engine-seed patch of oracle/Makefile) by
tests/test_oracle_golden.py::test_stocs_port_equals_reference_live and the golden vectors tests/golden/stocs_small.npz.
Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def approximate_bin(val: int, disc: int) -> int:
    lower = val - int(math.fmod(val, disc))          # C++ % truncates toward zero
    upper = lower + disc
    return lower if (val - lower) < (upper - val) else upper


def _dot3(a, b) -> f32:
    # Eigen's unrolled 3-element reduction: a0 b0 + (a1 b1 + a2 b2)
    return f32(f32(a[0] * b[0]) + f32(f32(a[1] * b[1]) + f32(a[2] * b[2])))


def _cross(a, b):
    return np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
                     f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], f32)


def _norm(a) -> f32:
    return np.sqrt(_dot3(a, a), dtype=f32)


def _angle_deg_int(y: f32, x: f32) -> int:
    a = f32(math.atan2(float(y), float(x)))          # atan2f, emulated by rounding the double result
    return int(float(f32(a * f32(180.0))) / math.pi)


def compute_ppf(p1, n1, p2, n2):
    u = (p1 - p2).astype(f32)
    k0 = approximate_bin(int(f32(_norm(u) * f32(1000.0))), 5)
    k1 = approximate_bin(_angle_deg_int(_norm(_cross(n1, u)), _dot3(n1, u)), 10)
    k2 = approximate_bin(_angle_deg_int(_norm(_cross(n2, u)), _dot3(n2, u)), 10)
    k3 = approximate_bin(_angle_deg_int(_norm(_cross(n1, n2)), _dot3(n1, n2)), 10)
    return (k0, k1, k2, k3)


def ppf_keys_all_pairs(xyz, nrm) -> np.ndarray:
    """compute_ppf of ALL ordered pairs (i, j) of a cloud, vectorised with the same fp32 operation order: (n*n, 4) int keys, row
    i*n + j.  The generator of PPFMap.txt is not in the reference tree; this is the map the CPU arm of bench.py hands to the
    reference's Perform_N_steps (rows are grouped by pyoracle.group_ppf_keys)."""
    P = np.ascontiguousarray(xyz, f32)
    N = np.ascontiguousarray(nrm, f32)
    n = len(P)
    u = (P[:, None, :] - P[None, :, :]).astype(f32)                      # p1 - p2
    n1 = np.broadcast_to(N[:, None, :], (n, n, 3))
    n2 = np.broadcast_to(N[None, :, :], (n, n, 3))

    def dot(a, b):
        return (a[..., 0] * b[..., 0]).astype(f32) + ((a[..., 1] * b[..., 1]).astype(f32) + (a[..., 2] * b[..., 2]).astype(f32)).astype(f32)

    def cross(a, b):
        return np.stack([(a[..., 1] * b[..., 2]).astype(f32) - (a[..., 2] * b[..., 1]).astype(f32),
                         (a[..., 2] * b[..., 0]).astype(f32) - (a[..., 0] * b[..., 2]).astype(f32),
                         (a[..., 0] * b[..., 1]).astype(f32) - (a[..., 1] * b[..., 0]).astype(f32)], axis=-1).astype(f32)

    def norm(a):
        return np.sqrt(dot(a, a).astype(f32)).astype(f32)

    def angle(y, x):
        a = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(f32)
        return ((a * f32(180.0)).astype(f32).astype(np.float64) / math.pi).astype(np.int64)      # C++ int(): truncation toward zero

    def abin(val, disc):
        lower = val - np.fmod(val, disc)
        upper = lower + disc
        return np.where((val - lower) < (upper - val), lower, upper)

    k0 = abin((norm(u) * f32(1000.0)).astype(f32).astype(np.int64), 5)
    k1 = abin(angle(norm(cross(n1, u)), dot(n1, u)), 10)
    k2 = abin(angle(norm(cross(n2, u)), dot(n2, u)), 10)
    k3 = abin(angle(norm(cross(n1, n2)), dot(n1, n2)), 10)
    return np.stack([k0, k1, k2, k3], axis=-1).reshape(n * n, 4).astype(np.int32)


def build_ppf_map(xyz, nrm):
    """(keys4, offsets, pairs) of the cloud's PPF map, rows in key order (see ppf_keys_all_pairs)."""
    from oracle import pyoracle
    return pyoracle.group_ppf_keys(ppf_keys_all_pairs(xyz, nrm), len(xyz))


class MinStd:
    M = 2147483647

    def __init__(self, seed: int):
        self.x = seed % self.M
        if self.x == 0:
            self.x = 1

    def next(self) -> int:
        self.x = (self.x * 16807) % self.M
        return self.x

    def canonical(self) -> float:
        r = 2147483646.0
        s = float(self.next() - 1)
        s += float(self.next() - 1) * r
        ret = s / (r * r)
        return ret if ret < 1.0 else math.nextafter(1.0, 0.0)


def discrete_draw(w: np.ndarray, gen: MinStd) -> int:
    n = len(w)
    if n < 2:
        return 0
    p = w.astype(np.float64)
    s = 0.0
    for v in p:                                       # std::accumulate, sequential
        s += v
    u = gen.canonical()
    acc = 0.0
    for i in range(n - 1):
        acc += p[i] / s
        if not (acc < u):
            return i
    return n - 1


def _seg_seg(p1, p2, q1, q2):
    """distSegmentToSegment<Vector3f, double>: the vectors and their dot products are fp32 (Eigen order), the case analysis
    runs in double, the closing distance is fp32 again with the invariants narrowed to float."""
    k_small = 0.0001
    u, v, w = (p2 - p1).astype(f32), (q2 - q1).astype(f32), (p1 - q1).astype(f32)
    a, b, c, d, e = (float(_dot3(x, y)) for x, y in ((u, u), (u, v), (v, v), (u, w), (v, w)))
    f = a * c - b * b
    s1, s2, t1, t2 = 0.0, f, 0.0, f
    if f < k_small:
        s1, s2, t1, t2 = 0.0, 1.0, e, c
    else:
        s1, t1 = (b * e - c * d), (a * e - b * d)
        if s1 < 0.0:
            s1, t1, t2 = 0.0, e, c
        elif s1 > s2:
            s1, t1, t2 = s2, e + b, c
    if t1 < 0.0:
        t1 = 0.0
        if -d < 0.0:
            s1 = 0.0
        elif -d > a:
            s1 = s2
        else:
            s1, s2 = -d, a
    elif t1 > t2:
        t1 = t2
        if (-d + b) < 0.0:
            s1 = 0
        elif (-d + b) > a:
            s1 = s2
        else:
            s1, s2 = (-d + b), a
    inv1 = 0.0 if abs(s1) < k_small else s1 / s2
    inv2 = 0.0 if abs(t1) < k_small else t1 / t2
    i1, i2 = f32(inv1), f32(inv2)
    r = ((w + (i1 * u).astype(f32)).astype(f32) - (i2 * v).astype(f32)).astype(f32)
    return float(_norm(r)), inv1, inv2


def try_quadrilateral(P, ids):
    pt = [P[i].astype(f32) for i in ids]
    best, min_d, inv = None, np.finfo(np.float32).max, (0.0, 0.0)
    for i in range(4):
        for j in range(4):
            if i == j:
                continue
            k = 0
            while k in (i, j):
                k += 1
            l = 0
            while l in (i, j, k):
                l += 1
            d, i1, i2 = _seg_seg(pt[i], pt[j], pt[k], pt[l])
            d = f32(d)
            if d < min_d:
                min_d, best, inv = d, (i, j, k, l), (f32(i1), f32(i2))
    if best is None:
        return False, None, None
    return True, np.array([ids[b] for b in best], np.int32), np.array(inv, f32)


def select_stocs(P, N, prior, keyset, engine_seed):
    """P, N: centred scene positions / unit normals (n,3) f32; prior (n,) f32; keyset: set of 4-int key tuples of the model's
    PPF map.  Returns (ok, ids[4], inv[2]) like SelectQuadrilateralStoCS + TryQuadrilateral."""
    n = len(P)
    gen = MinStd(engine_seed)
    orig = prior.astype(f32)
    curr = orig.copy()
    b1 = discrete_draw(curr, gen)

    def edge(b, i):
        return f32(1.0) if compute_ppf(P[b], N[b], P[i], N[i]) in keyset else f32(0.0)

    def normalise_and_draw(cur):
        s = f32(0.0)
        for v in cur:
            s = f32(s + v)
        cur = (cur / s).astype(f32)
        return cur, discrete_draw(cur, gen)

    # point 2
    nxt = np.zeros(n, f32)
    for i in range(n):
        if i == b1 or curr[i] == 0:
            continue
        nxt[i] = f32(f32(orig[i] * orig[b1]) * edge(b1, i))
    if not np.any(nxt != 0):
        return False, None, None
    curr, b2 = normalise_and_draw(nxt)
    # point 3
    v1 = (P[b2] - P[b1]).astype(f32)
    nxt = np.zeros(n, f32)
    for i in range(n):
        v2 = (P[i] - P[b1]).astype(f32)
        d = float(_dot3(v1, v2))
        ang = f32(float(f32(f32(math.acos(d)) * f32(180.0))) / math.pi) if -1.0 <= d <= 1.0 else f32(np.nan)
        other = f32(f32(180.0) - ang)
        ang = other if other < ang else ang
        if i == b1 or i == b2 or curr[i] == 0 or ang < 30:
            continue
        nxt[i] = f32(f32(curr[i] * orig[b2]) * edge(b2, i))
    if not np.any(nxt != 0):
        return False, None, None
    curr, b3 = normalise_and_draw(nxt)
    # point 4
    (x1, y1, z1), (x2, y2, z2), (x3, y3, z3) = (P[b].astype(np.float64) for b in (b1, b2, b3))
    denom = f32(-x3 * y2 * z1 + x2 * y3 * z1 + x3 * y1 * z2 - x1 * y3 * z2 - x2 * y1 * z3 + x1 * y2 * z3)
    nxt = np.zeros(n, f32)
    if denom != 0:
        A = f32((-y2 * z1 + y3 * z1 + y1 * z2 - y3 * z2 - y1 * z3 + y2 * z3) / float(denom))
        B = f32((x2 * z1 - x3 * z1 - x1 * z2 + x3 * z2 + x1 * z3 - x2 * z3) / float(denom))
        Cc = f32((-x2 * y1 + x3 * y1 + x1 * y2 - x3 * y2 - x1 * y3 + x2 * y3) / float(denom))
    for i in range(n):
        if i in (b1, b2, b3) or curr[i] == 0:
            continue
        if denom != 0:
            lin = f32(f32(f32(A * P[i][0]) + f32(B * P[i][1])) + f32(Cc * P[i][2]))
            pd = f32(abs(float(lin) - 1.0))
            if float(pd) > 0.01 or any(float(_norm((P[i] - P[b]).astype(f32))) < 0.01 for b in (b1, b2, b3)):
                continue
        nxt[i] = f32(f32(curr[i] * orig[b3]) * edge(b3, i))
    if not np.any(nxt != 0):
        return False, None, None
    curr, b4 = normalise_and_draw(nxt)
    return try_quadrilateral(P, [b1, b2, b3, b4])
