"""The per-lane solver math of the kernels (matrix_b200/csrc/ppx_math.cuh), compiled for the host and held against
numpy, the oracle and the reference's recorded outputs -- Jacobi SVD with its look-ahead, Jacobi eigensolver,
Gauss-Jordan inverse, LU solve, and the square pseudo-inverse solve with its LU fast path and SVD fallback -- so that
the numerics of the GPU kernels are exercised by the CPU suite as well.

tests/cpp/math_host.cpp is a test harness around the device header; it is not part of the product (the product has
no CPU path) and is built into a scratch directory here.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from matrix_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 1 << 13


@pytest.fixture(scope="module")
def mh(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("math_host") / "math_host.so")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
           os.path.join(ROOT, "tests", "cpp", "math_host.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cm(a, rows, cols):
    """records [n, rows*cols] column-major -> matrices [n, rows, cols]"""
    return a.reshape(a.shape[0], cols, rows).transpose(0, 2, 1)


@pytest.mark.parametrize("shape", [(6, 6), (4, 3)])
def test_device_jacobi_svd_on_the_host(mh, shape):
    m, k = shape
    A = np.ascontiguousarray(synth.dense(N, m, k, seed=0x51 + m))
    US, w2, V = np.empty((N, m * k)), np.empty((N, k)), np.empty((N, k * k))
    assert mh.mh_svd(m, k, _p(A), _p(US), _p(w2), _p(V), C.c_size_t(N)) == 0
    Am, Um, Vm = _cm(A, m, k), _cm(US, m, k), _cm(V, k, k)
    scale = np.abs(Am).max(axis=(1, 2))
    assert (np.abs(np.einsum("nij,nkj->nik", Um, Vm) - Am).max(axis=(1, 2)) / scale).max() < 1e-14
    assert np.abs(np.einsum("nji,njk->nik", Vm, Vm) - np.eye(k)).max() < 1e-14
    G = np.einsum("nji,njk->nik", Um, Um)
    off = G - np.einsum("nii->ni", G)[:, :, None] * np.eye(k)
    assert (np.abs(off).max(axis=(1, 2)) / np.einsum("nii->ni", G).max(axis=1)).max() < 1e-14  # columns orthogonal
    sv = np.sort(np.sqrt(w2), axis=1)
    ref = np.sort(np.linalg.svd(Am, compute_uv=False), axis=1)
    assert (np.abs(sv - ref).max(axis=1) / ref.max(axis=1)).max() < 1e-14


def test_device_jacobi_svd_rank_deficient_and_clustered(mh, fixtures):
    """the edge matrices of the fixtures (zero, identity, rank 4, all ones, Hilbert ...) and clustered spectra"""
    A = np.ascontiguousarray(fixtures["dense6.A"][192:])
    n = A.shape[0]
    US, w2, V = np.empty((n, 36)), np.empty((n, 6)), np.empty((n, 36))
    assert mh.mh_svd(6, 6, _p(A), _p(US), _p(w2), _p(V), C.c_size_t(n)) == 0
    Am = _cm(A, 6, 6)
    ref = np.sort(np.linalg.svd(Am, compute_uv=False), axis=1)
    sv = np.sort(np.sqrt(w2), axis=1)
    assert np.all(np.abs(sv - ref).max(axis=1) <= 1e-14 * np.maximum(ref.max(axis=1), 1e-300))
    assert np.abs(np.einsum("nij,nkj->nik", _cm(US, 6, 6), _cm(V, 6, 6)) - Am).max() < 1e-13
    rng = np.random.default_rng(3)
    Q1 = np.linalg.qr(rng.standard_normal((256, 6, 6)))[0]
    Q2 = np.linalg.qr(rng.standard_normal((256, 6, 6)))[0]
    sig = np.array([1.0, 1.0 + 1e-9, 0.5, 1e-6, 1e-6 * (1 + 1e-7), 1e-10])
    B = np.ascontiguousarray(np.einsum("nij,j,nkj->nik", Q1, sig, Q2).transpose(0, 2, 1).reshape(256, 36))
    US, w2, V = np.empty((256, 36)), np.empty((256, 6)), np.empty((256, 36))
    assert mh.mh_svd(6, 6, _p(B), _p(US), _p(w2), _p(V), C.c_size_t(256)) == 0
    assert np.abs(np.sort(np.sqrt(w2), axis=1) - np.sort(sig)).max() < 1e-14
    assert np.abs(np.einsum("nji,njk->nik", _cm(V, 6, 6), _cm(V, 6, 6)) - np.eye(6)).max() < 1e-14


@pytest.mark.parametrize("n_", [4, 6])
def test_device_jacobi_eig_on_the_host(mh, n_):
    S = np.ascontiguousarray(synth.symmetric(N, n_, seed=0x61 + n_))
    d, Z = np.empty((N, n_)), np.empty((N, n_ * n_))
    assert mh.mh_eig(n_, _p(S), _p(d), _p(Z), C.c_size_t(N)) == 0
    Sm, Zm = _cm(S, n_, n_), _cm(Z, n_, n_)
    ref = np.linalg.eigvalsh(Sm)
    assert np.all(np.diff(d, axis=1) >= 0) and np.abs(d - ref).max() < 1e-13
    res = np.einsum("nij,njk->nik", Sm, Zm) - Zm * d[:, None, :]
    assert np.abs(res).max() < 1e-13
    assert np.abs(np.einsum("nji,njk->nik", Zm, Zm) - np.eye(n_)).max() < 1e-13


def test_device_lu_and_gauss_jordan_on_the_host(mh, port, fixtures):
    A = np.ascontiguousarray(np.vstack([synth.dense(N, 6, 6, seed=0x71), fixtures["dense6.A"][192:]]))
    b = np.ascontiguousarray(synth.dense(A.shape[0], 6, 1, seed=0x72))
    n = A.shape[0]
    Ai, x = np.empty((n, 36)), np.empty((n, 6))
    s1, s2 = np.empty(n, dtype=np.int8), np.empty(n, dtype=np.int8)
    assert mh.mh_inverse6(_p(A), _p(Ai), _p(s1), C.c_size_t(n)) == 0
    assert mh.mh_lu_solve6(_p(A), _p(b), _p(x), _p(s2), C.c_size_t(n)) == 0
    rAi = port.inverse(6, A)
    rx, rst = port.linsolve_lu(6, A, b)
    # same pivot sequence as the reference's LU: same SINGULAR decisions (zero matrix, rank 4, all ones, ...)
    assert np.array_equal(s1 != 0, rst == 3) and np.array_equal(s2 != 0, rst == 3) and (rst == 3).sum() >= 3
    ok = rst != 3
    kappa = np.linalg.cond(_cm(A[ok], 6, 6))
    err = np.abs(Ai[ok] - rAi[ok]).max(axis=1) / np.abs(rAi[ok]).max(axis=1)
    assert np.all(err <= 1e-13 * np.maximum(1.0, kappa))
    errx = np.abs(x[ok] - rx[ok]).max(axis=1) / np.abs(rx[ok]).max(axis=1)
    assert np.all(errx <= 1e-13 * np.maximum(1.0, kappa))


def _graded_triangular():
    """I + t triu(ones, 1) and its transpose: every partial-pivoting pivot is 1 while cond grows like t^5 (3.5e18 at
    t = 1e3), so pivot ratios say nothing about the conditioning -- the reference truncates the smallest singular value
    there (SVD::solve, linalg.hpp:889-901) and an unguarded LU shortcut would return A^-1 b ~ 1e15 instead."""
    mats = []
    for t in (1.0, 10.0, 1e2, 3e2, 1e3, 1e4):
        M = np.eye(6) + t * np.triu(np.ones((6, 6)), 1)
        for X in (M, M.T, M[::-1, ::-1].copy()):
            mats.append(X.T.reshape(36))  # column-major record
    return np.array(mats)


def test_square_pinv_solve_paths_on_the_host(mh, port, fixtures):
    """LU fast path vs forced Jacobi SVD vs the oracle's SVD::solve: well-conditioned systems agree to cond * eps on
    both paths; graded condition numbers up to 1e14 cross the pivot test; rank-deficient systems (truncation active in
    the reference) are answered by the fallback on either setting."""
    A = synth.dense(N, 6, 6, seed=0x81)
    rng = np.random.default_rng(5)
    Q1 = np.linalg.qr(rng.standard_normal((512, 6, 6)))[0]
    Q2 = np.linalg.qr(rng.standard_normal((512, 6, 6)))[0]
    smin = 10.0 ** -np.linspace(1, 14, 512)
    sig = np.stack([np.ones(512), np.full(512, 0.7), np.full(512, 0.4), np.full(512, 0.2), np.sqrt(smin), smin], axis=1)
    G = np.einsum("nij,nj,nkj->nik", Q1, sig, Q2).transpose(0, 2, 1).reshape(512, 36)
    edge = fixtures["dense6.A"][192:]
    A = np.ascontiguousarray(np.vstack([A, G, edge, _graded_triangular()]))
    n = A.shape[0]
    b = np.ascontiguousarray(synth.dense(n, 6, 1, seed=0x82))
    ref = port.linsolve_svd(6, 6, A, b)
    kappa = np.linalg.cond(_cm(A, 6, 6))
    out = {}
    for allow in (1, 0):
        x = np.empty((n, 6))
        assert mh.mh_square_pinv_solve6(_p(A), _p(b), _p(x), allow, C.c_size_t(n)) == 0
        out[allow] = x
        fin = np.isfinite(ref).all(axis=1) & (kappa < 1e13)
        err = np.abs(x[fin] - ref[fin]).max(axis=1) / np.maximum(np.abs(ref[fin]).max(axis=1), 1e-300)
        assert np.all(err <= 1e-12 * np.maximum(1.0, kappa[fin]))
        # numerically rank-deficient systems: the reference truncates; so must both settings (minimum-norm answer)
        rd = np.isfinite(ref).all(axis=1) & (kappa > 1e15) & (np.abs(ref).max(axis=1) > 0)
        assert rd.sum() >= 2
        errd = np.abs(x[rd] - ref[rd]).max(axis=1) / np.abs(ref[rd]).max(axis=1)
        assert np.all(errd < 1e-6)
    well = kappa < 1e3
    assert np.abs(out[1][well] - out[0][well]).max() < 1e-11


def test_device_lie_maps_and_kinematics_on_the_host(mh, port, fixtures):
    """se3::exp, SE3::log, so3::exp, SO3::log and the UR5 forward kinematics + Jacobian as the kernels compute them
    (closed forms, branch-free reciprocals) against the oracle at the 1e-12 bar, incl. the recorded edge cases
    (theta = 0, 1e-9 ... pi) of the reference itself."""
    xi = np.ascontiguousarray(np.vstack([synth.se3_twists(N), fixtures["se3_exp.xi"]]))
    n = xi.shape[0]
    T, back = np.empty((n, 16)), np.empty((n, 6))
    assert mh.mh_se3_exp(_p(xi), _p(T), C.c_size_t(n)) == 0
    rT = port.se3_exp(xi)
    assert np.abs(T - rT).max() < 1e-14
    assert mh.mh_SE3_log(_p(np.ascontiguousarray(rT)), _p(back), C.c_size_t(n)) == 0
    rback = port.SE3_log(rT)
    # theta = acos(c) carries the rounding of c divided by sin(theta): 1e-12 everywhere except next to theta = pi
    theta = np.linalg.norm(xi[:, :3], axis=1)
    tol = 1e-12 + 4 * np.finfo(float).eps / np.maximum(np.abs(np.sin(theta)), 1e-9)
    assert np.all(np.abs(back - rback).max(axis=1) <= tol * np.maximum(1.0, np.abs(rback).max(axis=1)))
    w = np.ascontiguousarray(xi[:, :3])
    R, wb = np.empty((n, 9)), np.empty((n, 3))
    assert mh.mh_so3_exp(_p(w), _p(R), C.c_size_t(n)) == 0
    assert np.abs(R - port.so3_exp(w)).max() < 1e-14
    Rm = np.ascontiguousarray(fixtures["SO3_log.R"])
    wl = np.empty((Rm.shape[0], 3))
    assert mh.mh_SO3_log(_p(Rm), _p(wl), C.c_size_t(Rm.shape[0])) == 0
    ref = fixtures["SO3_log.w"]
    assert np.all(np.abs(wl - ref).max(axis=1) <= 1e-9)  # near theta = pi the log is ill-conditioned: 1e-16 / sin(theta)
    gen = np.linalg.norm(ref, axis=1) < 3.0
    assert np.abs(wl[gen] - ref[gen]).max() < 1e-12
    q = np.ascontiguousarray(fixtures["ur5.q"])
    S = np.ascontiguousarray(synth.UR5_SCREWS, dtype=np.float64)
    M = np.ascontiguousarray(synth.UR5_HOME, dtype=np.float64)
    Tq, Js = np.empty((q.shape[0], 16)), np.empty((q.shape[0], 36))
    assert mh.mh_fk_jacobian6(_p(S), _p(M), _p(q), _p(Tq), _p(Js), C.c_size_t(q.shape[0])) == 0
    assert np.abs(Tq - fixtures["ur5.T"]).max() < 1e-14 and np.abs(Js - fixtures["ur5.Js"]).max() < 1e-14


MIXED_CHAIN = np.array([
    [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],            # revolute about +z        (aligned)
    [0.3, -0.5, 0.8, 0.2, -0.1, 0.4],          # general screw, |w| != 1
    [0.0, -1.0, 0.0, -0.089, 0.0, 0.425],      # revolute about -y        (aligned, negative sign)
    [0.0, 0.0, 0.0, 0.0, 0.6, 0.8],            # prismatic (w = 0)
    [1.0, 0.0, 0.0, 0.0, 0.3, -0.2],           # revolute about +x        (aligned)
    [0.0, 2.0, 0.0, 0.1, 0.0, 0.3],            # along y but not unit: general path
])
MIXED_HOME = np.array([0.0, 1.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.3, -0.2, 0.5, 1.0])


def test_axis_aligned_and_general_joints_mix_on_the_host(mh, port):
    """poe_joint: joints whose axis is a unit coordinate axis take the structured path, everything else the general
    one -- a chain that mixes both (incl. a negative axis, a prismatic joint and a non-unit axis) against the oracle,
    with angles that hit the |x| < EPS_SP identity branch of se3::exp on aligned and general joints alike"""
    n = 4096
    q = synth.joint_angles(n, seed=0xC4A1)
    q[:64, 0] = 0.0
    q[64:128, 2] = 1e-9
    q[128:192, 1] = 1e-9
    q[192:256, 4] = -5e-8
    q = np.ascontiguousarray(q)
    T, Js = np.empty((n, 16)), np.empty((n, 36))
    assert mh.mh_fk_jacobian6(_p(MIXED_CHAIN), _p(MIXED_HOME), _p(q), _p(T), _p(Js), C.c_size_t(n)) == 0
    rT, rJ = port.poe_fk_jacobian(MIXED_CHAIN, MIXED_HOME, q)
    assert (np.abs(T - rT).max(axis=1) / np.abs(rT).max(axis=1)).max() < 1e-12
    assert (np.abs(Js - rJ).max(axis=1) / np.abs(rJ).max(axis=1)).max() < 1e-12
    assert np.all(T[:, [3, 7, 11]] == 0.0) and np.all(T[:, 15] == 1.0)
