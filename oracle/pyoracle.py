"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

  RefOracle   oracle/_ref/libs4ref.so       the reference engine compiled in place (oracle/Makefile `ref`)
  PortOracle  oracle/_build/liblcp_oracle.so  the plain-C restatement (oracle/lcp_oracle.c)
  v4pcs_quads_port                          numpy restatement of ExtractCongruentSet in operMode 2 (V4PCS), pinned against
                                            tests/golden/mode2_small.npz and live against RefOracle

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (physimglobalpose_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libs4ref.so")
PORT_SO = os.path.join(HERE, "_build", "liblcp_oracle.so")

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u16p = C.POINTER(C.c_uint16)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def build_port() -> str:
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def build_ref() -> str:
    """Needs /root/reference; on the GPU box the prebuilt .so travels with the snapshot."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return REF_SO


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class _Base:
    prefix = ""

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def _common_init(self, scene_xyz, scene_nrm, search_xyz, search_nrm, val_xyz, val_nrm):
        self.P, self.Pn = _f32(scene_xyz), _f32(scene_nrm)
        self.Q, self.Qn = _f32(search_xyz), _f32(search_nrm)
        self.V, self.Vn = _f32(val_xyz), _f32(val_nrm)
        self.nP, self.nQ, self.nV = len(self.P), len(self.Q), len(self.V)

    def centroids(self):
        cP = np.zeros(3, np.float32)
        cQ = np.zeros(3, np.float32)
        self._fn("get_centroids", None, [C.c_void_p, _f32p, _f32p])(self.h, _p(cP, _f32p), _p(cQ, _f32p))
        return cP, cQ

    def priors(self):
        out = np.zeros(self.nP, np.float32)
        self._fn("get_priors", None, [C.c_void_p, _f32p])(self.h, _p(out, _f32p))
        return out

    def centred(self, which):
        n = [self.nP, self.nQ, self.nV][which]
        xyz = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        self._fn("get_centred", None, [C.c_void_p, C.c_int, _f32p, _f32p])(self.h, which, _p(xyz, _f32p), _p(nrm, _f32p))
        return xyz, nrm

    def verify(self, T):
        T = _f32(T).reshape(-1, 12)
        out = np.zeros(len(T), np.uint32)
        self._fn("verify_batch", None, [C.c_void_p, _f32p, C.c_int64, _u32p])(self.h, _p(T, _f32p), len(T), _p(out, _u32p))
        return out

    def verify_running_best(self, T):
        T = _f32(T).reshape(-1, 12)
        frac = np.zeros(len(T), np.float32)
        best = C.c_int64(-1)
        self._fn("verify_running_best", None, [C.c_void_p, _f32p, C.c_int64, _f32p, _i64p])(
            self.h, _p(T, _f32p), len(T), _p(frac, _f32p), C.byref(best))
        return frac, best.value

    def weighted_verify(self, T, reg_of=-1):
        T = _f32(T).reshape(-1, 12)
        scores = np.zeros(len(T), np.float32)
        nreg = np.zeros(len(T), np.int32)
        reg = np.full(self.nV, -1, np.int32)
        self._fn("weighted_verify_batch", None, [C.c_void_p, _f32p, C.c_int64, _f32p, _i32p, C.c_int64, _i32p])(
            self.h, _p(T, _f32p), len(T), _p(scores, _f32p), _p(nreg, _i32p), reg_of, _p(reg, _i32p))
        if reg_of >= 0:
            return scores, nreg, reg[: nreg[reg_of]].copy()
        return scores, nreg

    def rigid_from_quad(self, base, quad):
        b = np.ascontiguousarray(base, np.int32)
        q = np.ascontiguousarray(quad, np.int32)
        T16 = np.zeros(16, np.float32)
        P16 = np.zeros(16, np.float64)
        ok = self._fn("rigid_from_quad", C.c_int, [C.c_void_p, _i32p, _i32p, _f32p, _f64p])(
            self.h, _p(b, _i32p), _p(q, _i32p), _p(T16, _f32p), _p(P16, _f64p))
        # column-major -> row-major 4x4
        return bool(ok), T16.reshape(4, 4).T.copy(), P16.reshape(4, 4).T.copy()

    def extract_pairs(self, dist, eps, cap=1 << 24):
        buf = np.zeros((cap, 2), np.int32)
        n = self._fn("extract_pairs", C.c_int64, [C.c_void_p, C.c_float, C.c_float, _i32p, C.c_int64])(
            self.h, dist, eps, _p(buf, _i32p), cap)
        assert n <= cap
        return buf[:n].copy()


class PortOracle(_Base):
    prefix = "lo_"

    def __init__(self, scene_xyz, scene_nrm, search_xyz, search_nrm, val_xyz, val_nrm, delta,
                 K=None, prior_img=None):
        if not os.path.exists(PORT_SO):
            build_port()
        self.lib = C.CDLL(PORT_SO)
        self._common_init(scene_xyz, scene_nrm, search_xyz, search_nrm, val_xyz, val_nrm)
        K9 = None if K is None else _f32(K).reshape(9)
        img = None if prior_img is None else np.ascontiguousarray(prior_img, np.uint16)
        rows, cols = (0, 0) if img is None else img.shape
        create = self._fn("create", C.c_void_p, [_f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int,
                                                 C.c_double, _f32p, _u16p, C.c_int, C.c_int])
        self.h = create(_p(self.P, _f32p), _p(self.Pn, _f32p), self.nP, _p(self.Q, _f32p), _p(self.Qn, _f32p), self.nQ,
                        _p(self.V, _f32p), _p(self.Vn, _f32p), self.nV, float(delta), _p(K9, _f32p), _p(img, _u16p), rows, cols)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._fn("destroy", None, [C.c_void_p])(self.h)
                self.h = None
        except Exception:      # interpreter shutdown
            pass

    def verify_mt(self, T, nthreads):
        T = _f32(T).reshape(-1, 12)
        out = np.zeros(len(T), np.uint32)
        s = self._fn("verify_batch_mt", C.c_double, [C.c_void_p, C.c_int, _f32p, C.c_int64, _u32p])(
            self.h, nthreads, _p(T, _f32p), len(T), _p(out, _u32p))
        return out, s

    def nn_ids(self, T):
        T = _f32(T).reshape(12)
        out = np.zeros(self.nV, np.int32)
        self._fn("nn_ids", None, [C.c_void_p, _f32p, _i32p])(self.h, _p(T, _f32p), _p(out, _i32p))
        return out

    def improving_chain(self, scores, cap=4096):
        s = _f32(scores)
        idx = np.zeros(cap, np.int64)
        n = self._fn("improving_chain", C.c_int64, [_f32p, C.c_int64, _i64p, C.c_int64])(_p(s, _f32p), len(s), _p(idx, _i64p), cap)
        return idx[: min(n, cap)].copy()

    def tricp(self, src, tgt, T, trim=0.5, ratio=0.99, max_iter=100):
        src, tgt = _f32(src), _f32(tgt)
        Tio = _f32(T).reshape(12).copy()
        e = C.c_float(0)
        it = self._fn("tricp", C.c_int, [_f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_float, C.c_float, C.c_int, _f32p])(
            _p(src, _f32p), len(src), _p(tgt, _f32p), len(tgt), _p(Tio, _f32p), trim, ratio, max_iter, C.byref(e))
        return Tio.reshape(3, 4), it, e.value


def port_remove_explained(seg_xyz, model_xyz, placed_poses44, threshold=0.008):
    """lo_remove_explained: mask of the segment points explained by the model placed at the given poses."""
    build_port()
    lib = C.CDLL(PORT_SO)
    seg, mod = _f32(seg_xyz).reshape(-1, 3), _f32(model_xyz).reshape(-1, 3)
    T = np.ascontiguousarray(np.asarray(placed_poses44, np.float64).reshape(-1, 4, 4)[:, :3, :], dtype=np.float32)
    flags = np.zeros(len(seg), np.uint8)
    f = lib.lo_remove_explained
    f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_float, C.POINTER(C.c_ubyte)]
    kept = f(_p(seg, _f32p), len(seg), _p(mod, _f32p), len(mod), _p(T, _f32p), len(T), threshold, flags.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return flags.astype(bool), kept


class RefOracle(_Base):
    prefix = "ref_"

    def __init__(self, scene_xyz, scene_nrm, search_xyz, search_nrm, val_xyz, val_nrm, delta,
                 K=None, prior_img=None, srand_seed=1):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(REF_SO)
        self._common_init(scene_xyz, scene_nrm, search_xyz, search_nrm, val_xyz, val_nrm)
        self._args = (delta, K, prior_img, srand_seed)
        self.h = self._create()

    def _create(self):
        delta, K, prior_img, srand_seed = self._args
        K9 = None if K is None else _f32(K).reshape(9)
        img = None if prior_img is None else np.ascontiguousarray(prior_img, np.uint16)
        rows, cols = (0, 0) if img is None else img.shape
        create = self._fn("create", C.c_void_p, [_f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int,
                                                 C.c_double, _f32p, _u16p, C.c_int, C.c_int, C.c_uint])
        return create(_p(self.P, _f32p), _p(self.Pn, _f32p), self.nP, _p(self.Q, _f32p), _p(self.Qn, _f32p), self.nQ,
                      _p(self.V, _f32p), _p(self.Vn, _f32p), self.nV, float(delta), _p(K9, _f32p), _p(img, _u16p), rows, cols,
                      srand_seed)

    def __del__(self):
        for h in [getattr(self, "h", None)] + list(getattr(self, "_extra", [])):
            if h:
                self._fn("destroy", None, [C.c_void_p])(h)
        self.h = None
        self._extra = []

    def diameter(self):
        return self._fn("get_diameter", C.c_float, [C.c_void_p])(self.h)

    def verify_mt(self, T, nthreads):
        """One separately-initialised matcher per thread (kdtree.h:311 member stack)."""
        extra = getattr(self, "_extra", [])
        while len(extra) < nthreads - 1:
            extra.append(self._create())
        self._extra = extra
        handles = (C.c_void_p * nthreads)(self.h, *extra[: nthreads - 1])
        T = _f32(T).reshape(-1, 12)
        out = np.zeros(len(T), np.uint32)
        s = self._fn("verify_batch_mt", C.c_double, [C.POINTER(C.c_void_p), C.c_int, _f32p, C.c_int64, _u32p])(
            handles, nthreads, _p(T, _f32p), len(T), _p(out, _u32p))
        return out, s

    def find_quads(self, base, inv1, inv2, eps, pairs1, pairs2, cap=1 << 22):
        b = np.ascontiguousarray(base, np.int32)
        p1 = np.ascontiguousarray(pairs1, np.int32)
        p2 = np.ascontiguousarray(pairs2, np.int32)
        out = np.zeros((cap, 4), np.int32)
        n = self._fn("find_quads", C.c_int64, [C.c_void_p, _i32p, C.c_float, C.c_float, C.c_float, _i32p, C.c_int64,
                                               _i32p, C.c_int64, _i32p, C.c_int64])(
            self.h, _p(b, _i32p), inv1, inv2, eps, _p(p1, _i32p), len(p1), _p(p2, _i32p), len(p2), _p(out, _i32p), cap)
        return out[: min(n, cap)].copy()

    def select_quadrilateral(self, seed):
        b = np.zeros(4, np.int32)
        inv = np.zeros(2, np.float32)
        ok = self._fn("select_quadrilateral", C.c_int, [C.c_void_p, C.c_uint, _i32p, _f32p])(self.h, seed, _p(b, _i32p), _p(inv, _f32p))
        return bool(ok), b, inv

    def perform_n_steps(self, mode=0, seed=1, cap=4096):
        poses = np.zeros((cap, 16), np.float64)
        scores = np.zeros(cap, np.float32)
        nt = C.c_int64(0)
        best = C.c_float(0)
        bi = C.c_int(-1)
        stage = np.zeros(3, np.float32)
        n = self._fn("perform_n_steps", C.c_int, [C.c_void_p, C.c_int, C.c_uint, _f64p, _f32p, C.c_int, _i64p, _f32p,
                                                  C.POINTER(C.c_int), _f32p])(
            self.h, mode, seed, _p(poses, _f64p), _p(scores, _f32p), cap, C.byref(nt), C.byref(best), C.byref(bi), _p(stage, _f32p))
        T = np.zeros((nt.value, 12), np.float32)
        self._fn("get_transforms", None, [C.c_void_p, _f32p, C.c_int64])(self.h, _p(T, _f32p), nt.value)
        reg = np.zeros(self.nV, np.int32)
        nr = self._fn("get_registered", C.c_int64, [C.c_void_p, _i32p, C.c_int64])(self.h, _p(reg, _i32p), self.nV)
        bases = np.zeros((256, 4), np.int32)
        inv = np.zeros((256, 2), np.float32)
        nb = self._fn("get_bases", C.c_int, [C.c_void_p, _i32p, _f32p, C.c_int])(self.h, _p(bases, _i32p), _p(inv, _f32p), 256)
        return dict(chain_pose=poses[:n].reshape(-1, 4, 4).transpose(0, 2, 1).copy(), chain_score=scores[:n].copy(),
                    transforms=T.reshape(-1, 3, 4), best_lcp=best.value, best_index=bi.value, stage_s=stage,
                    registered=reg[:nr].copy(), bases=bases[:nb].copy(), invariants=inv[:nb].copy())

    # ---- operMode 1 (StoCS + PPF map)
    def set_ppf_map(self, keys4, offsets, pairs):
        k = np.ascontiguousarray(keys4, np.int32).reshape(-1, 4)
        o = np.ascontiguousarray(offsets, np.int64)
        p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self._fn("set_ppf_map", None, [C.c_void_p, _i32p, _i64p, _i32p, C.c_int64])(self.h, _p(k, _i32p), _p(o, _i64p), _p(p, _i32p), len(k))

    def compute_ppf(self, pairs):
        """Match4PCSBase::computePPF of scene index pairs -> (n, 4) int32."""
        p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        out = np.zeros((len(p), 4), np.int32)
        self._fn("compute_ppf", None, [C.c_void_p, _i32p, C.c_int64, _i32p])(self.h, _p(p, _i32p), len(p), _p(out, _i32p))
        return out

    def select_stocs(self, engine_seed):
        b = np.zeros(4, np.int32)
        inv = np.zeros(2, np.float32)
        ok = self._fn("select_stocs", C.c_int, [C.c_void_p, C.c_uint, _i32p, _f32p])(self.h, int(engine_seed), _p(b, _i32p), _p(inv, _f32p))
        return bool(ok), b, inv

    def congruent_set_mode1(self, base, inv1, inv2, cap=1 << 22):
        b = np.ascontiguousarray(base, np.int32)
        out = np.zeros((cap, 4), np.int32)
        n = self._fn("congruent_set_mode1", C.c_int64, [C.c_void_p, _i32p, C.c_float, C.c_float, _i32p, C.c_int64])(
            self.h, _p(b, _i32p), inv1, inv2, _p(out, _i32p), cap)
        return out[: min(n, cap)].copy()


    # ---- operMode 2 (V4PCS: tetrahedron base, six-distance join)
    def congruent_set_mode2(self, base, cap=1 << 22):
        """ExtractCongruentSet in operMode 2 for one base (4 scene ids): (n, 4) model ids, in the reference's own order
        (an unordered_set iteration: compare as sets)."""
        b = np.ascontiguousarray(base, np.int32)
        out = np.zeros((cap, 4), np.int32)
        n = self._fn("congruent_set_mode2", C.c_int64, [C.c_void_p, _i32p, _i32p, C.c_int64])(self.h, _p(b, _i32p), _p(out, _i32p), cap)
        return out[: min(n, cap)].copy()

    def select_tetrahedron(self, seed):
        b = np.zeros(4, np.int32)
        ok = self._fn("select_tetrahedron", C.c_int, [C.c_void_p, C.c_uint, _i32p])(self.h, int(seed), _p(b, _i32p))
        return bool(ok), b


def v4pcs_quads_port(P_centred, Q_centred, base, eps):
    """numpy restatement of ExtractCongruentSet in operMode 2 (match4pcsBase.cc:1929-2039): the six pair extractions
    (pairCreationFunctor.h:167-253: |float norm - d| <= eps compared in double, both orientations, no self pairs) and the
    connectivity join of FindCongruentQuadrilateralsV4PCS (:978-1044).  Returns the quads (v1, v2, v3, v4) sorted."""
    P = np.asarray(P_centred, np.float32); Q = np.asarray(Q_centred, np.float32)
    b = [P[i] for i in base]
    def norm32(v):
        v = v.astype(np.float32)
        return np.sqrt((v[..., 0] * v[..., 0] + (v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2])).astype(np.float32)).astype(np.float32)
    def dist(i, j):                                     # (base_3D_[i].pos() - base_3D_[j].pos()).norm(), Eigen: x^2 + y^2 + z^2 left to right
        v = (b[i] - b[j]).astype(np.float32)
        return np.float32(np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])))
    d = {1: dist(0, 1), 2: dist(0, 2), 3: dist(0, 3), 4: dist(1, 2), 5: dist(1, 3), 6: dist(2, 3)}
    D = norm32(Q[:, None, :] - Q[None, :, :])           # (q_i - q_j).norm() in the accelerator's association dx^2 + (dy^2 + dz^2)
    n = len(Q)
    off = ~np.eye(n, dtype=bool)
    A = {k: (np.abs(D.astype(np.float64) - np.float64(d[k])) <= np.float64(np.float32(eps))) & off for k in d}
    quads = []
    for v1, v2 in np.argwhere(A[1]):
        v3s = np.flatnonzero(A[2][v1] & A[4][:, v2])    # (v1, v3) in pairs2 and (v3, v2) in pairs4
        if len(v3s) == 0:
            continue
        c4 = A[3][v1] & A[5][:, v2]                     # (v1, v4) in pairs3 and (v4, v2) in pairs5
        for v3 in v3s:
            for v4 in np.flatnonzero(c4 & A[6][:, v3]): # (v4, v3) in pairs6
                quads.append((v1, v2, v3, v4))
    return np.array(sorted(quads), np.int32).reshape(-1, 4)


def group_ppf_keys(keys4_all, n):
    """All-ordered-pairs keys (n*n, 4; row t = pair (t // n, t % n)) -> map rows (keys4, offsets, pairs) in key order,
    pairs inside a key in (i, j) order; the diagonal is skipped."""
    k = np.asarray(keys4_all, np.int64).reshape(n * n, 4)
    t = np.arange(n * n)
    keep = (t // n) != (t % n)
    k, t = k[keep], t[keep]
    order = np.lexsort((t, k[:, 3], k[:, 2], k[:, 1], k[:, 0]))
    k, t = k[order], t[order]
    new = np.ones(len(k), bool)
    new[1:] = np.any(k[1:] != k[:-1], axis=1)
    starts = np.flatnonzero(new)
    offsets = np.append(starts, len(k)).astype(np.int64)
    pairs = np.stack([t // n, t % n], axis=1).astype(np.int32)
    return k[starts].astype(np.int32), offsets, pairs
