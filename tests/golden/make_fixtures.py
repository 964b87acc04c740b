"""Generate tests/golden/ref_outputs.npz from the reference itself.

Run in the build container (where /root/reference exists):

    cd oracle && make ref && cd .. && python tests/golden/make_fixtures.py

It feeds small seeded inputs plus the edge cases listed below through
oracle/_ref/libppx_ref_strict.so -- the UNMODIFIED reference headers behind
oracle/ref_shim.cpp, compiled with the pinned flags and -ffp-contract=off -- and
stores inputs and outputs.  The committed .npz is what pins the C restatement
(oracle/ppx_oracle.c) and the CUDA path on machines where /root/reference does
not exist (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from oracle_lib import Oracle  # noqa: E402
from matrix_b200 import synth  # noqa: E402

N = 192  # random instances per op


def edge_twists():
    """theta in {0, 1e-9, 1e-7, 1.2e-7, 1e-6, 1e-3, pi-1e-6, pi} about three axes."""
    thetas = [0.0, 1e-9, 1e-7, 1.2e-7, 1e-6, 1e-3, 0.5, np.pi - 1e-3, np.pi - 1e-6, np.pi]
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [-1, 2, 0.5]], dtype=float)
    axes /= np.linalg.norm(axes, axis=1, keepdims=True)
    rows = []
    for t in thetas:
        for a in axes:
            rows.append(np.concatenate([t * a, [0.3, -0.7, 1.1]]))
    rows.append(np.zeros(6))
    return np.array(rows)


def edge_rotations(ref):
    """Rotations that hit every branch of SO3::log (liegroup.hpp:64-84), incl. exact pi turns."""
    mats = [np.eye(3), np.diag([1.0, -1.0, -1.0]), np.diag([-1.0, 1.0, -1.0]), np.diag([-1.0, -1.0, 1.0])]
    out = [m.T.reshape(9) for m in mats]
    w = edge_twists()[:, :3]
    return np.vstack([np.array(out), ref.so3_exp(w)])


def edge_lstsq():
    """Singular / rank-deficient / scaled 4x3 systems (QR SINGULAR path, linalg.hpp:549-552)."""
    A = []
    b = []
    base = np.array([3.5, 1.6, 3.7, 4.3, 2.7, -5.7, -0.8, -9.8, -3.1, -6.0, 1.9, 6.9])
    A.append(base), b.append(np.ones(4))
    A.append(np.zeros(12)), b.append(np.array([1.0, 2.0, 3.0, 4.0]))
    z = base.copy()
    z[4:8] = 0.0  # zero second column
    A.append(z), b.append(np.ones(4))
    d = base.copy()
    d[8:12] = d[0:4]  # duplicated column (rank 2, not flagged by the reference's test on the column norm)
    A.append(d), b.append(np.array([1.0, -1.0, 0.5, 2.0]))
    A.append(base * 1e-9), b.append(np.ones(4))
    A.append(base * 1e9), b.append(np.ones(4))
    t = base.copy()
    t[0] = -t[0]
    A.append(t), b.append(np.array([0.0, 0.0, 0.0, 0.0]))
    return np.array(A), np.array(b)


def edge_dense6():
    e = [np.zeros(36), np.eye(6).reshape(36), np.diag([3.0, -2.0, 1.0, 0.0, 1e-9, 5.0]).reshape(36),
         np.ones(36), np.arange(36, dtype=float), np.triu(np.arange(1, 37, dtype=float).reshape(6, 6)).T.reshape(36)]
    h = np.array([[1.0 / (i + j + 1) for j in range(6)] for i in range(6)])  # Hilbert
    e.append(h.T.reshape(36))
    return np.array(e)


def edge_sym4():
    g = np.array([2, -1, 1, 4, -1, 2, -1, 1, 1, -1, 2, -2, 4, 1, -2, 3], dtype=float)
    e = [g, np.zeros(16), np.eye(4).reshape(16), np.diag([1.0, 1.0, 2.0, 2.0]).reshape(16), np.ones(16),
         np.diag([4.0, -3.0, 2.0, -1.0]).reshape(16)]
    t = np.diag([2.0, 2.0, 2.0, 2.0]) + np.diag([1.0, 1.0, 1.0], 1) + np.diag([1.0, 1.0, 1.0], -1)
    e.append(t.reshape(16))
    return np.array(e)


def main():
    ref = Oracle("ref_strict")
    assert ref.kind == "reference"
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    fx = {}

    xi = np.vstack([synth.se3_twists(N), edge_twists()])
    fx["se3_exp.xi"] = xi
    fx["se3_exp.T"] = ref.se3_exp(xi)
    fx["SE3_log.T"] = fx["se3_exp.T"]
    fx["SE3_log.xi"] = ref.SE3_log(fx["SE3_log.T"])
    fx["se3_exp_log.xi"] = ref.se3_exp_log(xi)

    w = xi[:, :3]
    fx["so3_exp.w"] = w
    fx["so3_exp.R"] = ref.so3_exp(w)
    Rm = edge_rotations(ref)
    fx["SO3_log.R"] = Rm
    fx["SO3_log.w"] = ref.SO3_log(Rm)
    for k, nm in enumerate(["ljac", "ljacinv", "rjac", "rjacinv"]):
        fx[f"so3_{nm}.J"] = ref.so3_jac(k, w)

    T1 = ref.se3_exp(synth.se3_twists(N))
    T2 = ref.se3_exp(synth.se3_twists(N, first=N))
    fx["SE3_mul.A"], fx["SE3_mul.B"] = T1, T2
    fx["SE3_mul.C"] = ref.SE3_mul(T1, T2)
    fx["SE3_inv.Ti"] = ref.SE3_inv(T1)
    fx["SE3_Adt.Ad"] = ref.SE3_Adt(T1)
    fx["se3_adt.ad"] = ref.se3_adt(xi)

    q = np.vstack([synth.joint_angles(N), np.zeros((1, 6)),
                   np.array([[0.0, -0.5 * np.pi, 0.0, 0.0, 0.5 * np.pi, 0.0]]),
                   np.full((1, 6), np.pi), np.full((1, 6), 1e-9)])
    fx["ur5.q"] = q
    fx["ur5.T"], fx["ur5.Js"] = ref.poe_fk_jacobian(S, M, q)
    # a 4-joint and a 7-joint chain with non-unit, non-axis-aligned screws
    for nj in (4, 7):
        sc = synth.uniform_records(0xC4A1 + nj, 0, nj, 6, -1.0, 1.0)
        home = ref.se3_exp(synth.se3_twists(1, first=5))[0]
        qq = synth.joint_angles(64, njoints=nj, seed=0xC4A2 + nj)
        fx[f"chain{nj}.screws"], fx[f"chain{nj}.home"], fx[f"chain{nj}.q"] = sc, home, qq
        fx[f"chain{nj}.T"], fx[f"chain{nj}.Js"] = ref.poe_fk_jacobian(sc, home, qq)

    A, b = synth.lstsq_4x3(N)
    Ae, be = edge_lstsq()
    A, b = np.vstack([A, Ae]), np.vstack([b, be])
    fx["lstsq.A"], fx["lstsq.b"] = A, b
    fx["lstsq.x_qr"], fx["lstsq.status_qr"] = ref.linsolve_qr(4, 3, A, b)
    fx["lstsq.x_svd"] = ref.linsolve_svd(4, 3, A, b)
    fx["lstsq.U"], fx["lstsq.w"], fx["lstsq.V"] = ref.svd(4, 3, A)
    fx["lstsq.pinv"] = ref.pinv(4, 3, A)

    A6 = np.vstack([synth.dense(N, 6, 6), edge_dense6()])
    b6 = synth.dense(A6.shape[0], 6, 1, seed=0xB6)
    fx["dense6.A"], fx["dense6.b"] = A6, b6
    fx["dense6.U"], fx["dense6.w"], fx["dense6.V"] = ref.svd(6, 6, A6)
    fx["dense6.x_svd"] = ref.linsolve_svd(6, 6, A6, b6)
    fx["dense6.x_qr"], fx["dense6.status_qr"] = ref.linsolve_qr(6, 6, A6, b6)
    fx["dense6.x_lu"], fx["dense6.status_lu"] = ref.linsolve_lu(6, A6, b6)
    fx["dense6.inverse"] = ref.inverse(6, A6)
    fx["dense6.det"] = ref.det(6, A6)
    fx["dense6.norm2"] = ref.norm2(6, 6, A6)

    S4 = np.vstack([synth.symmetric(N, 4), edge_sym4()])
    fx["sym4.S"] = S4
    fx["sym4.d_sorted"], fx["sym4.z_sorted"] = ref.eig_sym(4, S4, True)
    fx["sym4.d_unsorted"], fx["sym4.z_unsorted"] = ref.eig_sym(4, S4, False)
    S6 = synth.symmetric(64, 6)
    fx["sym6.S"] = S6
    fx["sym6.d_sorted"], fx["sym6.z_sorted"] = ref.eig_sym(6, S6, True)

    qi = synth.joint_angles(N, seed=synth.SEEDS["ik_newton_step"])
    Tt = ref.poe_fk_jacobian(S, M, qi + synth.ik_offsets(N), want_Js=False)[0]
    fx["ik.q"], fx["ik.Ttarget"] = qi, Tt
    fx["ik.q_out"], fx["ik.Vs"] = ref.ik_newton_step(S, M, qi, Tt)
    fx["ik.q_solved"], fx["ik.iters"] = ref.ik_solve(S, M, qi, Tt, 20)

    # kinematics<6>::inverseSpace (CoDo trust region, demo/robotics.hpp:126-152) with the demo's (-pi, pi) joint ranges
    qc = synth.joint_angles(N, seed=synth.SEEDS["ik_newton_step"]) * 0.95
    Tc = ref.poe_fk_jacobian(S, M, qc + synth.ik_offsets(N), want_Js=False)[0]
    qc = np.vstack([qc, np.array([[4.0, 0, 0, 0, 0, 0]])])  # last row: illegal start, outside the box
    Tc = np.vstack([Tc, Tc[:1]])
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    fx["codo.init"], fx["codo.Ttarget"], fx["codo.lo"], fx["codo.hi"] = qc, Tc, lo, hi
    fx["codo.q_out"], fx["codo.x_raw"], fx["codo.fnorm"], fx["codo.status"] = ref.ik_codo(S, M, lo, hi, Tc, qc)

    out = os.path.join(HERE, "ref_outputs.npz")
    np.savez_compressed(out, **fx)
    print(f"wrote {out}: {len(fx)} arrays, {os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
