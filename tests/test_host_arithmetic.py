"""CPU tests of two exact-arithmetic helpers the kernels rely on, through their host-callable copies in the C ABI (no GPU, no compute
on a device: `pgp_host_*` are plain host functions compiled from the same source as the device code).

* chain_sum_equal (csrc/k2_pcs.cu): the StoCS sampler normalises its weights by their SEQUENTIAL float sum (std::accumulate,
  S4/algorithms/match4pcsBase.cc:652-657).  For m equal weights the kernel replaces the m-step chain by a closed form; it must be
  bit-identical to the chain for every value and count, including the ties-to-even cases and a stalled chain.
* max_eigvec4 (csrc/k5_tricp.cu): the rigid fit of trimmed ICP takes the eigenvector of the largest eigenvalue of Horn's 4x4 matrix
  from the characteristic polynomial instead of a Jacobi diagonalisation; it must agree with a symmetric eigensolver, and decline
  (so that the kernel falls back to Jacobi) when the top eigenvalue is nearly double."""
import ctypes as C

import numpy as np
import pytest

from physimglobalpose_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _chain(c: np.float32, m: int) -> np.float32:
    # numpy's cumsum accumulates sequentially in the array's dtype: exactly acc = fl(acc + c), m times from 0
    return np.cumsum(np.full(m, c, dtype=np.float32), dtype=np.float32)[-1]


def _float_from(mant: int, exp: int) -> np.float32:
    return np.array([(exp << 23) | mant], dtype=np.uint32).view(np.float32)[0]


def test_chain_sum_equals_the_sequential_float_chain(lib):
    rng = np.random.default_rng(11)
    cases = []
    for trial in range(4000):
        kind = trial % 5
        mant = int(rng.integers(0, 1 << 23))
        if kind == 1:
            mant &= ~((1 << int(rng.integers(0, 23))) - 1)              # trailing zeros: ties happen
        elif kind == 2:
            mant = (mant & ~0xFFF) | 0x800                               # ...1000 0000 0000: exact ties against a coarser ulp
        elif kind == 3:
            mant = 1 << int(rng.integers(0, 23))
        c = _float_from(mant, int(rng.integers(100, 140)))
        m = int(rng.integers(1, 40)) if kind == 4 else int(rng.integers(1, 20000))
        cases.append((c, m))
    # what the sampler actually meets: fl(1 / k), k points still in play, m <= k of them with an existing PPF
    for k in (3, 7, 100, 257, 1317, 2000, 4093):
        for m in (1, 2, k // 3 + 1, k - 1, k):
            cases.append((np.float32(1.0) / np.float32(k), m))
    bad = 0
    for c, m in cases:
        got = np.float32(lib.pgp_host_chain_sum_equal(C.c_float(float(c)), C.c_longlong(m)))
        want = _chain(c, m)
        bad += int(got.view(np.uint32) != want.view(np.uint32))
    assert bad == 0
    # a chain that stalls: c below half an ulp of the running sum
    c = _float_from(0x123456, 100)
    assert np.float32(lib.pgp_host_chain_sum_equal(C.c_float(float(c)), C.c_longlong(40_000_000))).view(np.uint32) == _chain(c, 40_000_000).view(np.uint32)
    assert lib.pgp_host_chain_sum_equal(C.c_float(1.5), C.c_longlong(0)) == 0.0


def _horn_matrix(H: np.ndarray) -> np.ndarray:
    (Sxx, Sxy, Sxz), (Syx, Syy, Syz), (Szx, Szy, Szz) = H
    return np.array([[Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx],
                     [Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz],
                     [Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy],
                     [Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz]], dtype=np.float64)


def _eig(lib, N: np.ndarray):
    N = np.ascontiguousarray(N, dtype=np.float64)
    q = np.zeros(4, dtype=np.float64)
    ok = lib.pgp_host_max_eigvec4(N.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
    return ok, q


def test_max_eigvec4_matches_a_symmetric_eigensolver(lib):
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(3000):
        scale = 10.0 ** float(rng.integers(-6, 7))
        H = scale * (rng.random((3, 3)) - 0.5)
        if trial % 4 == 1:
            H[2] *= 1e-9                                                 # planar source cloud
        elif trial % 4 == 2:                                             # a clean fit: H = diag . rotation
            th = rng.random() * 6.28
            R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
            H = (np.diag([1.0, 0.6, 0.3]) * scale) @ R.T
        N = _horn_matrix(H)
        ok, q = _eig(lib, N)
        w, V = np.linalg.eigh(N)
        gap = (w[-1] - w[-2]) / max(abs(w[-1]), 1e-300)
        if not ok:
            assert gap < 1e-3                                            # declining is only allowed near a double eigenvalue
            continue
        worst = max(worst, 1.0 - abs(float(q @ V[:, -1])) / float(np.linalg.norm(q)))
    assert worst < 1e-12


def test_max_eigvec4_declines_a_double_eigenvalue(lib):
    for e in (1e-9, 1e-12, 0.0):
        ok, _ = _eig(lib, _horn_matrix(np.diag([1.0, e, 0.5 * e])))      # collinear source points: rotation about x is free
        assert ok == 0
    ok, q = _eig(lib, _horn_matrix(np.diag([1.0, 1e-2, 0.5e-2])))
    assert ok == 1 and abs(abs(q[0]) / np.linalg.norm(q) - 1.0) < 1e-12
    assert _eig(lib, np.zeros((4, 4)))[0] == 0
