"""The device state machine of inverseSpace (matrix_b200/csrc/ppx_codo.cuh), compiled for the host and replayed
against the reference's fixtures and the oracle -- the GPU kernel's per-lane logic checked without a GPU.

tests/cpp/codo_host.cpp is a test harness around the device header; it is not part of the product (the product has
no CPU path) and is built into a scratch directory here.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from matrix_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS_SP = 1.1920928955078125e-07
CONVERGED, DIVERGED = 1, 2


@pytest.fixture(scope="module")
def codo(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("codo_host") / "codo_host.so")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
           os.path.join(ROOT, "tests", "cpp", "codo_host.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    dp = C.POINTER(C.c_double)

    def run(S, M, lo, hi, Tt, init):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (S, M, lo, hi, Tt, init)]
        n = arrs[5].shape[0]
        q, x = np.empty((n, 6)), np.empty((n, 6))
        fn, st, ev = np.empty(n), np.empty(n, dtype=np.int8), np.empty(n, dtype=np.int32)
        rc = lib.codo_host_run(*[a.ctypes.data_as(dp) for a in arrs], q.ctypes.data_as(dp), x.ctypes.data_as(dp),
                               fn.ctypes.data_as(dp), st.ctypes.data_as(C.POINTER(C.c_byte)),
                               ev.ctypes.data_as(C.POINTER(C.c_int)), C.c_size_t(n))
        assert rc == 0
        return q, x, fn, st, ev

    return run


def test_codo_state_machine_replays_reference_fixtures(codo, fixtures):
    """tests/golden/ref_outputs.npz codo.*: outputs of the reference's own kinematics<6>::inverseSpace"""
    fx = fixtures
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    q, x, fn, st, ev = codo(S, M, fx["codo.lo"], fx["codo.hi"], fx["codo.Ttarget"], fx["codo.init"])
    rq, rx, rfn, rst = fx["codo.q_out"], fx["codo.x_raw"], fx["codo.fnorm"], fx["codo.status"]
    assert np.array_equal(st, rst)
    # last row: illegal start outside the box -> DIVERGED, y = MAX_SP, init returned (optimization.hpp:552-556)
    assert st[-1] == DIVERGED and fn[-1] == rfn[-1] and np.array_equal(q[-1], fx["codo.init"][-1])
    solved, rsolved = fn <= 0.5 * EPS_SP, rfn <= 0.5 * EPS_SP
    assert np.array_equal(solved, rsolved)
    ok = solved & (st == CONVERGED)
    assert ok.mean() > 0.95
    # same root, to the optimiser's tolerance (the iterates differ in the last bits: reciprocal-based divisions, one
    # pass for residual and Jacobian, fused multiply-adds)
    assert np.abs(q[ok] - rq[ok]).max() < 1e-6
    assert np.abs(x[ok] - rx[ok]).max() < 1e-6
    # the work is the reference's: one evaluation per residual evaluation (start + one per trust-region trial)
    assert ev[ok].mean() < 8


def test_codo_state_machine_matches_oracle(codo, port):
    """8192 seeded instances incl. the stalled ones (ITMAX trips): status, solved set and roots vs the C oracle"""
    n = 1 << 13
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    qi = synth.joint_angles(n, seed=synth.SEEDS["ik_newton_step"]) * 0.95
    Tt = port.poe_fk_jacobian(S, M, qi + synth.ik_offsets(n), want_Js=False)[0]
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    q, x, fn, st, ev = codo(S, M, lo, hi, Tt, qi)
    rq, rx, rfn, rst = port.ik_codo(S, M, lo, hi, Tt, qi)
    assert (st == rst).mean() > 0.999
    solved, rsolved = fn <= 0.5 * EPS_SP, rfn <= 0.5 * EPS_SP
    assert (solved == rsolved).mean() > 0.998
    both = solved & rsolved & (st == CONVERGED) & (rst == CONVERGED)
    assert both.mean() > 0.99
    Tq = port.poe_fk_jacobian(S, M, q, want_Js=False)[0]
    assert np.abs(Tq[both] - Tt[both]).max() < 5e-7
    Jf = port.poe_fk_jacobian(S, M, rq, want_T=False)[1]
    sv = np.linalg.svd(Jf.reshape(n, 6, 6), compute_uv=False)
    kf = sv[:, 0] / sv[:, -1]
    ok = both & (kf < 1e2)
    assert np.all(np.abs(q[ok] - rq[ok]).max(axis=1) < 1e-6 * kf[ok])
    # a start on the bound itself is legal; when the gradient points out of the box the reference's scaling matrix
    # divides by zero and the optimiser reports DIVERGED on its first trip (optimization.hpp:787-808, 586-589)
    edge = qi[:64].copy()
    edge[:, 2] = np.pi
    q, x, fn, st, ev = codo(S, M, lo, hi, Tt[:64], edge)
    rq, rx, rfn, rst = port.ik_codo(S, M, lo, hi, Tt[:64], edge)
    assert np.array_equal(st, rst) and (st == DIVERGED).sum() >= 8 and (st == CONVERGED).sum() >= 8
    assert np.array_equal(q[st == DIVERGED], edge[st == DIVERGED]) and np.all(ev[st == DIVERGED] == 1)
    same = (st == CONVERGED) & (fn <= 0.5 * EPS_SP) & (rfn <= 0.5 * EPS_SP)
    assert np.abs(q[same] - rq[same]).max() < 1e-5
