"""Parity of the CUDA path with the CPU oracle (run on the B200 box: pytest -m gpu).

Every test calls libppx_cuda.so through its C ABI (matrix_b200.batch / matrix_b200.host are thin
ctypes wrappers) and compares with the C restatement of the reference (oracle/ppx_oracle.c, pinned
bit-for-bit to the reference by tests/test_oracle.py) on identical seeded inputs, with the reference's
golden vectors and with the reference outputs recorded in tests/golden/ref_outputs.npz.

Tolerances (fp64, north_star: 1e-12 relative):
  REL  = 1e-12  per-instance inf-norm relative error  |gpu - ref|_inf / |ref|_inf
  solves (QR / SVD / IK step) are compared condition-aware: error <= REL * max(1, kappa(A)), because
  the reference's own FMA and non-FMA builds already differ by ~kappa * 1e-16;
  the fraction inside the flat 1e-12 bar is asserted separately.
  SVD / eig: sorted singular values / eigenvalues relative to the largest, reconstruction residual,
  orthogonality; vectors only up to sign.
"""
import os

import numpy as np
import pytest

from matrix_b200 import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REL = 1e-12
EPS = 2.220446049250313e-16
LAYOUTS = ["aos", "soa"]


@pytest.fixture(scope="module")
def mb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from matrix_b200 import batch
    return batch


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run(mb, fn, layout, *arrays, **kw):
    """Call a batch function on numpy AoS inputs in the given layout; return numpy AoS outputs."""
    lay = mb.AOS if layout == "aos" else mb.SOA
    n = arrays[0].shape[0]
    ins = [dev(a) if layout == "aos" else mb.to_soa(dev(a)) for a in arrays]
    out = fn(*ins, layout=lay, **kw)
    single = not isinstance(out, tuple)
    outs = (out,) if single else out

    def back(t):
        if t is None:
            return None
        if t.dtype != torch.float64 or t.dim() == 1:  # dense per-instance vectors (status, iterations, fnorm)
            return t.cpu().numpy()
        return (t if layout == "aos" else mb.to_aos(t, n)).cpu().numpy()

    res = tuple(back(t) for t in outs)
    torch.cuda.synchronize()
    return res[0] if single else res


def rel_err(a, b, floor=0.0):
    """per-instance |a-b|_inf / max(|b|_inf, floor)"""
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    den = np.maximum(np.abs(b).max(axis=1), floor)
    den = np.where(den == 0.0, 1.0, den)
    return np.abs(a - b).max(axis=1) / den


N = 1 << 16
RAGGED = 1000 * 33 + 7  # not a multiple of the warp, block or tile size


# --------------------------------------------------------------------------------------- Lie ops
@pytest.mark.parametrize("layout", LAYOUTS)
def test_se3_exp(mb, cpu_ref, layout):
    xi = synth.se3_twists(RAGGED)
    T = run(mb, mb.se3_exp, layout, xi)
    ref = cpu_ref.se3_exp(xi)
    assert rel_err(T, ref).max() < REL
    assert np.all(T[:, [3, 7, 11]] == 0.0) and np.all(T[:, 15] == 1.0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_SE3_log_same_inputs(mb, cpu_ref, layout):
    xi = synth.se3_twists(RAGGED)
    T = cpu_ref.se3_exp(xi)
    got = run(mb, mb.SE3_log, layout, T)
    assert rel_err(got, cpu_ref.SE3_log(T)).max() < REL


@pytest.mark.parametrize("layout", LAYOUTS)
def test_se3_exp_log_fused_round_trip(mb, cpu_ref, layout):
    xi = synth.se3_twists(N)  # theta in [0.05, pi - 0.05]
    got = run(mb, mb.se3_exp_log, layout, xi)
    ref = cpu_ref.se3_exp_log(xi)
    assert rel_err(got, ref).max() < REL
    assert rel_err(got, xi).max() < 1e-11  # identity, conditioned by 1/sin^2(theta)


def test_lie_edge_cases_against_recorded_reference(mb, fixtures):
    """theta in {0, 1e-9, 1e-7, 1.2e-7, 1e-6, 1e-3, .., pi-1e-6, pi}: every branch of exp / log."""
    xi = fixtures["se3_exp.xi"]
    T = run(mb, mb.se3_exp, "aos", xi)
    assert rel_err(T, fixtures["se3_exp.T"]).max() < REL
    # identity branch is exact (liegroup.hpp:258-261)
    small = np.linalg.norm(xi[:, :3], axis=1) < 1.1920928955078125e-07
    assert np.array_equal(T[small], fixtures["se3_exp.T"][small])
    lg = run(mb, mb.SE3_log, "aos", fixtures["SE3_log.T"])
    ref = fixtures["SE3_log.xi"]
    Tin = fixtures["SE3_log.T"]
    ct = (Tin[:, 0] + Tin[:, 5] + Tin[:, 10] - 1) / 2
    # same branch as the reference everywhere, including its (0, p) result at theta = pi
    assert np.array_equal(np.all(lg[:, :3] == 0, axis=1), np.all(ref[:, :3] == 0, axis=1))
    # acos amplifies last-bit differences by 1/sin(theta): condition-aware bound
    s = np.sqrt(np.maximum(1 - ct * ct, 1e-300))
    tol = np.where(np.abs(ct) < 1, REL + 8 * EPS / np.maximum(s, 1e-8) ** 2, REL)
    assert np.all(rel_err(lg, ref, floor=1.0) <= tol)
    w = fixtures["so3_exp.w"]
    assert rel_err(run(mb, mb.so3_exp, "aos", w), fixtures["so3_exp.R"]).max() < REL
    R = fixtures["SO3_log.R"]
    got = run(mb, mb.SO3_log, "aos", R)
    ref = fixtures["SO3_log.w"]
    c = (R[:, 0] + R[:, 4] + R[:, 8] - 1) / 2
    s = np.sqrt(np.maximum(1 - c * c, 1e-300))
    tol = np.where(np.abs(c) < 1, REL + 8 * EPS / np.maximum(s, 1e-8) ** 2, REL)
    assert np.all(rel_err(got, ref, floor=1.0) <= tol)
    # exact pi turns: the three sub-branches of liegroup.hpp:68-84
    assert np.allclose(np.abs(got[1:4]), np.pi * np.eye(3), atol=1e-15)


MATMUL_MENU = [(3, 3, 3), (4, 4, 4), (6, 6, 6), (6, 6, 1), (4, 4, 1), (3, 3, 1)]  # unrolled kernels
MATMUL_GENERIC = [(4, 3, 3), (5, 4, 2), (2, 7, 3)]                                 # runtime-shape kernel


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("shape", MATMUL_MENU + MATMUL_GENERIC)
def test_matmul_is_bit_identical_to_the_reference_operator(mb, cpu_ref, layout, shape):
    """SURVEY 8(a) row a2, MatrixS::operator* (matrixs.hpp:596-640): k-j-i accumulation with separately rounded
    products and sums -- integer-style parity: every bit equal to the reference built without FMA contraction."""
    m, n, l = shape
    cnt = RAGGED if shape in MATMUL_MENU else 4099
    A = synth.dense(cnt, m, n, seed=0xA20 + m)
    B = synth.dense(cnt, n, l, seed=0xA40 + l)
    got = run(mb, lambda a_, b_, layout: mb.matmul(m, n, l, a_, b_, layout=layout), layout, A, B)
    assert np.array_equal(got, cpu_ref.matmul(m, n, l, A, B))


def test_matmul_host_face_and_errors(mb, port):
    from matrix_b200 import host
    from matrix_b200._lib import PpxCudaError
    A, B = synth.dense(300_001, 6, 6, seed=0xA61), synth.dense(300_001, 6, 6, seed=0xA62)
    assert np.array_equal(host.matmul(6, 6, 6, A, B), port.matmul(6, 6, 6, A, B))
    with pytest.raises(PpxCudaError):
        host.matmul(9, 2, 2, np.zeros((4, 18)), np.zeros((4, 4)))
    with pytest.raises(ValueError):
        host.matmul(3, 3, 3, np.zeros((4, 9)), np.zeros((5, 9)))
    # out-of-range explicit n and short second operands are refused before anything is launched (batch._in)
    a = torch.zeros((8, 9), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        mb.matmul(3, 3, 3, a, a[:4].contiguous())
    with pytest.raises(ValueError):
        mb.matmul(3, 3, 3, a, a, count=9)
    with pytest.raises(ValueError):
        mb.SE3_mul(torch.zeros((8, 16), dtype=torch.float64, device="cuda"),
                   torch.zeros((7, 16), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        mb.se3_exp(torch.zeros((8, 6), dtype=torch.float64, device="cuda"),
                   out=torch.zeros((4, 16), dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("layout", LAYOUTS)
def test_so3_exp_log_jac(mb, port, layout):
    w = synth.se3_twists(N)[:, :3].copy()
    R = run(mb, mb.so3_exp, layout, w)
    ref = port.so3_exp(w)
    assert rel_err(R, ref).max() < REL
    assert rel_err(run(mb, mb.SO3_log, layout, ref), port.SO3_log(ref)).max() < REL
    for k, name in enumerate(["ljac", "ljacinv", "rjac", "rjacinv"]):
        lay = mb.AOS if layout == "aos" else mb.SOA
        x = dev(w) if layout == "aos" else mb.to_soa(dev(w))
        J = mb.so3_jac(name, x, layout=lay)
        J = (J if layout == "aos" else mb.to_aos(J, w.shape[0])).cpu().numpy()
        refJ = port.so3_jac(k, w)
        # ljacinv's coefficient 1/x^2 - (1+cos x)/(2 x sin x) cancels near pi: bound by its conditioning
        th = np.linalg.norm(w, axis=1)
        tol = REL * np.maximum(1.0, 1.0 / np.abs(np.sin(th))) if k in (1, 3) else REL
        assert np.all(rel_err(J, refJ, floor=1.0) <= tol), name


@pytest.mark.parametrize("layout", LAYOUTS)
def test_SE3_mul_inv_Adt_adt(mb, port, layout):
    xi = synth.se3_twists(N)
    A = port.se3_exp(xi)
    B = port.se3_exp(synth.se3_twists(N, first=N))
    assert rel_err(run(mb, mb.SE3_mul, layout, A, B), port.SE3_mul(A, B)).max() < REL
    assert rel_err(run(mb, mb.SE3_inv, layout, A), port.SE3_inv(A)).max() < REL
    assert rel_err(run(mb, mb.SE3_Adt, layout, A), port.SE3_Adt(A)).max() < REL
    ad = run(mb, mb.se3_adt, layout, xi)
    assert np.array_equal(ad, port.se3_adt(xi))  # pure data movement: bit-exact


# --------------------------------------------------------------------------------------- kinematics
@pytest.mark.parametrize("layout", LAYOUTS)
def test_ur5_fk_jacobian(mb, cpu_ref, layout):
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q = synth.joint_angles(RAGGED)
    T, Js = run(mb, kin.fk_jacobian, layout, q)
    rT, rJ = cpu_ref.poe_fk_jacobian(synth.UR5_SCREWS, synth.UR5_HOME, q)
    assert rel_err(T, rT).max() < REL
    assert rel_err(Js, rJ).max() < REL
    assert np.array_equal(Js[:, :6], np.broadcast_to(synth.UR5_SCREWS[0], (RAGGED, 6)))  # demo/robotics.hpp:97
    # single-output variants give the same bits
    assert np.array_equal(run(mb, kin.forwardSpace, layout, q), T)
    assert np.array_equal(run(mb, kin.jacobiSpace, layout, q), Js)


def test_fk_jacobian_recorded_reference_and_other_chains(mb, fixtures):
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    T, Js = run(mb, kin.fk_jacobian, "aos", fixtures["ur5.q"])
    assert rel_err(T, fixtures["ur5.T"]).max() < REL and rel_err(Js, fixtures["ur5.Js"], floor=1.0).max() < REL
    for nj in (4, 7):  # non-unit, non-axis-aligned screws
        k = mb.Kinematics(fixtures[f"chain{nj}.screws"], fixtures[f"chain{nj}.home"])
        for layout in LAYOUTS:
            T, Js = run(mb, k.fk_jacobian, layout, fixtures[f"chain{nj}.q"])
            assert rel_err(T, fixtures[f"chain{nj}.T"]).max() < REL
            assert rel_err(Js, fixtures[f"chain{nj}.Js"]).max() < REL


def test_mixed_axis_aligned_and_general_chain(mb, cpu_ref):
    """axis-aligned joints take the structured kernel path (poe_joint<1..3>), all others the general one: a chain
    mixing both, incl. a negative axis, a prismatic joint and a non-unit axis, FK + Jacobian and the IK Newton step"""
    from test_math_host import MIXED_CHAIN, MIXED_HOME
    n = 20011
    q = synth.joint_angles(n, seed=0xC4A2)
    q[:64, 0] = 0.0
    q[64:128, 2] = 1e-9
    q[128:192, 1] = 1e-9
    kin = mb.Kinematics(MIXED_CHAIN, MIXED_HOME)
    rT, rJ = cpu_ref.poe_fk_jacobian(MIXED_CHAIN, MIXED_HOME, q)
    for layout in LAYOUTS:
        T, Js = run(mb, kin.fk_jacobian, layout, q)
        assert rel_err(T, rT).max() < REL and rel_err(Js, rJ).max() < REL
    Tt = cpu_ref.poe_fk_jacobian(MIXED_CHAIN, MIXED_HOME, q + synth.ik_offsets(n), want_Js=False)[0]
    qo, Vs = run(mb, kin.ik_newton_step, "soa", q, Tt, want_Vs=True)
    rq, rV = cpu_ref.ik_newton_step(MIXED_CHAIN, MIXED_HOME, q, Tt)
    assert rel_err(Vs, rV).max() < REL
    sv = np.linalg.svd(rJ.reshape(n, 6, 6), compute_uv=False)
    kappa = sv[:, 0] / sv[:, -1]
    err = np.abs((qo - q) - (rq - q)).max(axis=1) / np.maximum(np.abs(rq - q).max(axis=1), 1e-300)
    ok = kappa < 1e6
    assert ok.mean() > 0.9 and np.all(err[ok] <= REL * np.maximum(1.0, kappa[ok]))


WRIST_CHAIN = np.array([  # a 6R arm with a spherical wrist in its home pose: axes z y y x y x (PUMA / ABB / KUKA style)
    [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],
    [0.0, 1.0, 0.0, -0.40, 0.0, 0.03],
    [0.0, 1.0, 0.0, -0.40, 0.0, 0.48],
    [-1.0, 0.0, 0.0, 0.0, -0.44, 0.0],
    [0.0, 1.0, 0.0, -0.44, 0.0, 0.90],
    [1.0, 0.0, 0.0, 0.0, 0.44, 0.0],
])
WRIST_HOME = np.array([0.0, 0.0, -1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.44, 1.0])


UPRIGHT_CHAIN = np.array([  # the same kind of arm drawn with the forearm up in the home pose: axes z y y z y z
    [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],
    [0.0, 1.0, 0.0, -0.35, 0.0, 0.05],
    [0.0, -1.0, 0.0, 0.80, 0.0, -0.05],
    [0.0, 0.0, 1.0, 0.0, -0.07, 0.0],
    [0.0, 1.0, 0.0, -1.20, 0.0, 0.07],
    [0.0, 0.0, -1.0, 0.0, 0.07, 0.0],
])
UPRIGHT_HOME = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.07, 0.0, 1.30, 1.0])


@pytest.mark.parametrize("chain", ["ur5", "wrist", "upright"])
def test_pattern_kernels_match_the_generic_ones_and_the_oracle(mb, cpu_ref, chain):
    """Chains whose joint axes follow a precompiled pattern (FixedKinds: z y y y z y, z y y x y x) run IK kernels in
    which the per-joint branch is resolved at compile time; PPX_IK_PATTERNS=0 keeps them on the generic kernels.  Same
    arithmetic, so the two must agree bit for bit -- Newton step in both solve modes, and the whole Newton loop -- and
    both with the oracle."""
    from matrix_b200 import _lib
    S, M = {"ur5": (synth.UR5_SCREWS, synth.UR5_HOME), "wrist": (WRIST_CHAIN, WRIST_HOME),
            "upright": (UPRIGHT_CHAIN, UPRIGHT_HOME)}[chain]
    n = 50_003
    q = synth.joint_angles(n, seed=0xC4B1)
    Tt = cpu_ref.poe_fk_jacobian(S, M, q + synth.ik_offsets(n), want_Js=False)[0]
    kin = mb.Kinematics(S, M)
    out = {}
    old = os.environ.get("PPX_IK_PATTERNS")
    try:
        for mode in ("1", "0"):
            os.environ["PPX_IK_PATTERNS"] = mode
            res = []
            for lu in (True, False):
                _lib.set_square_solve_lu(lu)
                for layout in LAYOUTS:
                    res.append(run(mb, kin.ik_newton_step, layout, q, Tt, want_Vs=True))
            _lib.set_square_solve_lu(True)
            qs, it, st = kin.inverseSpaceNewton(dev(Tt), dev(q), 20)
            res.append((qs.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()))
            out[mode] = res
    finally:
        _lib.set_square_solve_lu(True)
        if old is None:
            os.environ.pop("PPX_IK_PATTERNS", None)
        else:
            os.environ["PPX_IK_PATTERNS"] = old
    for a, b in zip(out["1"], out["0"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    rq, rV = cpu_ref.ik_newton_step(S, M, q, Tt)
    qo, Vs = out["1"][0]
    assert rel_err(Vs, rV).max() < REL
    Js = cpu_ref.poe_fk_jacobian(S, M, q, want_T=False)[1]
    sv = np.linalg.svd(Js.reshape(n, 6, 6), compute_uv=False)
    kappa = sv[:, 0] / sv[:, -1]
    err = np.abs((qo - q) - (rq - q)).max(axis=1) / np.maximum(np.abs(rq - q).max(axis=1), 1e-300)
    ok = kappa < 1e6
    assert ok.mean() > 0.9 and np.all(err[ok] <= REL * np.maximum(1.0, kappa[ok]))


def test_golden_ur5(mb, golden):
    g = golden["ur5"]
    kin = mb.Kinematics(np.array(g["screws"], dtype=float), np.array(g["home"]))
    q = (np.array(g["q_times_pi"]) * 3.141592653589793).reshape(1, 6)
    T, Js = run(mb, kin.fk_jacobian, "aos", q)
    assert np.abs(T[0] - np.array(g["fk"])).max() < g["tol"]
    assert np.abs(Js[0].reshape(6, 6)[1:] - np.array(g["jacobian_cols_1_to_5"])).max() < g["tol"]
    # test/lie_test.cpp:211-216: Newton-SVD IK from (0,-1.5,0,0,1.5,0) reaches the target pose
    q0 = dev(np.array(g["ik_start"]).reshape(1, 6))
    Tt = dev(np.array(g["ik_target"]).reshape(1, 16))
    qs, it, st = kin.inverseSpaceNewton(Tt, q0, g["ik_max_iter"])
    Tq = kin.forwardSpace(qs).cpu().numpy()
    assert np.abs(Tq[0] - np.array(g["ik_target"])).max() < g["tol"]
    assert int(st[0]) == mb.CONVERGED and 1 <= int(it[0]) <= 20


def _ik_inputs(port, n, first=0):
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    q = synth.joint_angles(n, first=first, seed=synth.SEEDS["ik_newton_step"])
    Tt = port.poe_fk_jacobian(S, M, q + synth.ik_offsets(n, first=first), want_Js=False)[0]
    Js = port.poe_fk_jacobian(S, M, q, want_T=False)[1]
    sv = np.linalg.svd(Js.reshape(n, 6, 6), compute_uv=False)
    return q, Tt, sv[:, 0] / sv[:, -1]


@pytest.mark.parametrize("layout", LAYOUTS)
def test_ik_newton_step(mb, cpu_ref, layout):
    n = 1 << 15
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q, Tt, kappa = _ik_inputs(cpu_ref, n)
    qo, Vs = run(mb, kin.ik_newton_step, layout, q, Tt, want_Vs=True)
    rq, rV = cpu_ref.ik_newton_step(synth.UR5_SCREWS, synth.UR5_HOME, q, Tt)
    assert rel_err(Vs, rV).max() < REL
    # dq = pinv(Js) Vs: forward error of the step is bounded by kappa(Js)
    err = np.abs((qo - q) - (rq - q)).max(axis=1) / np.maximum(np.abs(rq - q).max(axis=1), 1e-300)
    ok = kappa < 1e3  # away from UR5 singularities
    assert ok.mean() > 0.9
    assert np.all(err[ok] <= REL * np.maximum(1.0, kappa[ok]))
    # the reference's own FMA / non-FMA builds agree inside the flat bar for 99.26 % of instances
    assert (err[ok] < REL).mean() > 0.98


def test_ik_solve_loop(mb, port):
    n = 1 << 13
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q, Tt, kappa = _ik_inputs(port, n, first=5000)
    qs, it, st = kin.inverseSpaceNewton(dev(Tt), dev(q), 20)
    qs, it, st = qs.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    rq, rit = port.ik_solve(synth.UR5_SCREWS, synth.UR5_HOME, q, Tt, 20)
    conv = st == mb.CONVERGED
    assert conv.mean() > 0.99
    # converged instances reproduce the target pose (the property lie_test.cpp:216 checks)
    Tq = kin.forwardSpace(dev(qs)).cpu().numpy()
    assert np.abs(Tq[conv] - Tt[conv]).max() < 1.1920928955078125e-07
    # same number of Newton steps as the reference wherever the problem is well conditioned
    ok = conv & (kappa < 1e2) & (rit < 20)
    assert (it[ok] == rit[ok]).mean() > 0.98
    # and the same solution, down to the floor of the reference's algorithm: once the residual rotation is below
    # sqrt(eps) ~ 1.5e-8, (trace-1)/2 rounds to 1 and SE3::log returns a ZERO rotation part (liegroup.hpp:246-249),
    # so Newton cannot resolve q below ~1.5e-8 * kappa; the reference's own FMA / non-FMA builds differ by that much
    Jf = port.poe_fk_jacobian(synth.UR5_SCREWS, synth.UR5_HOME, rq, want_T=False)[1]
    sv = np.linalg.svd(Jf.reshape(n, 6, 6), compute_uv=False)
    kf = sv[:, 0] / sv[:, -1]
    same = ok & (it == rit) & (kf < 1e2)
    assert same.mean() > 0.5
    assert np.all(np.abs(qs[same] - rq[same]).max(axis=1) < 1e-7 * kf[same])


# --------------------------------------------------------------------------------------- solvers
@pytest.mark.parametrize("layout", LAYOUTS)
def test_linsolve_qr_4x3(mb, cpu_ref, layout):
    A, b = synth.lstsq_4x3(RAGGED)
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.QR, 4, 3, a_, b_, layout=layout), layout, A, b)
    rx, rst = cpu_ref.linsolve_qr(4, 3, A, b)
    assert np.array_equal(st, rst)
    kappa = np.linalg.cond(A.reshape(-1, 3, 4).transpose(0, 2, 1))
    err = rel_err(x, rx)
    assert np.all(err <= REL * np.maximum(1.0, kappa))
    assert (err < REL).mean() > 0.999
    # least-squares optimality: A^T (A x - b) = 0
    Am = A.reshape(-1, 3, 4).transpose(0, 2, 1)
    r = np.einsum("nij,nj->ni", Am, x) - b
    g = np.einsum("nij,ni->nj", Am, r)
    assert np.all(np.abs(g).max(axis=1) <= 1e-13 * kappa * np.maximum(1.0, np.abs(x).max(axis=1)))


def test_linsolve_qr_singular_and_golden(mb, fixtures, golden):
    A, b = fixtures["lstsq.A"], fixtures["lstsq.b"]
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.QR, 4, 3, a_, b_, layout=layout), "aos", A, b)
    assert np.array_equal(st, fixtures["lstsq.status_qr"])
    sing = st == mb.SINGULAR
    assert sing.sum() >= 2 and np.array_equal(x[sing], b[sing][:, :3])  # untouched head of b
    g = golden["linsolve_4x3"]
    xg, sg = run(mb, lambda a_, b_, layout: mb.linsolve(mb.QR, 4, 3, a_, b_, layout=layout), "aos",
                 np.array(g["A"]).reshape(1, 12), np.array(g["b"], dtype=float).reshape(1, 4))
    assert np.abs(xg[0] - np.array(g["x"])).max() < 5e-15 and sg[0] == 0
    xs, _ = run(mb, lambda a_, b_, layout: mb.linsolve(mb.SVD, 4, 3, a_, b_, layout=layout), "aos",
                np.array(g["A"]).reshape(1, 12), np.array(g["b"], dtype=float).reshape(1, 4))
    assert np.abs(xs[0] - np.array(g["x"])).max() < 5e-15


@pytest.mark.parametrize("shape", [(3, 3), (4, 4), (6, 4), (6, 6)])
def test_linsolve_qr_other_shapes(mb, port, shape):
    m, n = shape
    A = synth.dense(4096, m, n, seed=0x51 + m * 8 + n)
    b = synth.dense(4096, m, 1, seed=0x52 + m * 8 + n)
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.QR, m, n, a_, b_, layout=layout), "soa", A, b)
    rx, rst = port.linsolve_qr(m, n, A, b)
    kappa = np.linalg.cond(A.reshape(-1, n, m).transpose(0, 2, 1))
    assert np.array_equal(st, rst)
    assert np.all(rel_err(x, rx) <= REL * np.maximum(1.0, kappa))


def _svd_checks(A, U, w, V, m, n, rw):
    cnt = A.shape[0]
    Am = A.reshape(cnt, n, m).transpose(0, 2, 1)
    Um = U.reshape(cnt, n, m).transpose(0, 2, 1)
    Vm = V.reshape(cnt, n, n).transpose(0, 2, 1)
    wmax = np.maximum(rw.max(axis=1), 1e-300)
    # singular values: sorted, relative to the largest (Golub-Reinsch guarantees absolute accuracy only)
    assert np.all(np.abs(np.sort(w, axis=1) - np.sort(rw, axis=1)).max(axis=1) / wmax < REL)
    assert np.all(w >= 0)
    rec = np.einsum("nij,nj,nkj->nik", Um, w, Vm)
    nrm = np.maximum(np.abs(Am).max(axis=(1, 2)), 1e-300)
    assert np.all(np.abs(rec - Am).max(axis=(1, 2)) / nrm < 1e-13)
    assert np.abs(np.einsum("nji,njk->nik", Vm, Vm) - np.eye(n)).max() < 1e-13
    full = w.min(axis=1) > 1e-8 * wmax
    assert np.abs(np.einsum("nji,njk->nik", Um[full], Um[full]) - np.eye(n)).max() < 1e-12


@pytest.mark.parametrize("layout", LAYOUTS)
def test_svd_6x6(mb, cpu_ref, layout):
    n = 1 << 14
    A = synth.dense(n, 6, 6)
    U, w, V = run(mb, lambda a_, layout: mb.svdcmp(6, 6, a_, layout=layout), layout, A)
    _, rw, _ = cpu_ref.svd(6, 6, A)
    _svd_checks(A, U, w, V, 6, 6, rw)


@pytest.mark.parametrize("shape", [(4, 3), (3, 3), (4, 4), (5, 4), (6, 4)])
def test_svd_other_shapes(mb, port, shape):
    m, n = shape
    A = synth.dense(4096, m, n, seed=0x61 + m * 8 + n)
    U, w, V = run(mb, lambda a_, layout: mb.svdcmp(m, n, a_, layout=layout), "soa", A)
    _, rw, _ = port.svd(m, n, A)
    _svd_checks(A, U, w, V, m, n, rw)
    # values-only call returns the same singular values
    _, w2, _ = run(mb, lambda a_, layout: mb.svdcmp(m, n, a_, layout=layout, want_U=False, want_V=False), "soa", A)
    assert np.abs(w2 - w).max() < 1e-14


def test_svd_rank_deficient_edge_cases(mb, fixtures, golden):
    A = fixtures["dense6.A"][-7:]  # zeros, identity, diag with a zero, ones, arange, triangular, Hilbert
    U, w, V = run(mb, lambda a_, layout: mb.svdcmp(6, 6, a_, layout=layout), "aos", A)
    rw = fixtures["dense6.w"][-7:]
    # the reference's SVD returns NaN for the all-zero matrix (0/0 in its anorm-scaled split test);
    # the Jacobi path returns the exact answer w = 0 -- a deliberate difference
    assert np.isnan(rw[0]).all() and np.all(w[0] == 0)
    good = ~np.isnan(rw).any(axis=1)
    wmax = np.maximum(rw[good].max(axis=1), 1e-300)
    assert np.all(np.abs(np.sort(w[good], axis=1) - np.sort(rw[good], axis=1)).max(axis=1) / wmax < REL)
    rec = np.einsum("nij,nj,nkj->nik", U.reshape(-1, 6, 6).transpose(0, 2, 1), w, V.reshape(-1, 6, 6).transpose(0, 2, 1))
    assert np.abs(rec - A.reshape(-1, 6, 6).transpose(0, 2, 1)).max() < 1e-12 * 36
    # test/matrix_test.cpp:831-835 and test/solver_test.cpp:41-104 (sigma_max to 10 eps) for the compiled shapes
    g = golden["svd_reconstruct_4x3"]
    Ag = np.array(g["A"], dtype=float).reshape(1, 12)
    U, w, V = run(mb, lambda a_, layout: mb.svdcmp(4, 3, a_, layout=layout), "aos", Ag)
    rec = U[0].reshape(3, 4).T @ np.diag(w[0]) @ V[0].reshape(3, 3)
    assert np.abs(rec - Ag.reshape(3, 4).T).max() < g["tol"]
    for case in golden["matrix_norm2"]["cases"]:
        if (case["m"], case["n"]) in [(5, 4), (4, 4)]:
            _, w, _ = run(mb, lambda a_, layout: mb.svdcmp(case["m"], case["n"], a_, layout=layout), "aos",
                          np.array(case["A"]).reshape(1, -1))
            assert abs(w.max() - case["norm2"]) < golden["matrix_norm2"]["tol"]


def test_svd_with_nan_and_inf_instances_terminates_and_spares_the_others(mb, port):
    """NaN / inf matrices mixed into a batch: the sweeps end by warp vote and the single-precision preconditioner runs
    a fixed number of sweeps, so such a lane can neither hang its warp nor disturb the 31 ordinary matrices next to it"""
    n = 4099
    A = synth.dense(n, 6, 6, seed=0xBAD)
    bad = np.arange(5, n, 37)
    A[bad[0::3], 7] = np.nan
    A[bad[1::3], 0] = np.inf
    A[bad[2::3]] = np.nan
    good = np.ones(n, dtype=bool)
    good[bad] = False
    for layout in LAYOUTS:
        U, w, V = run(mb, lambda a_, layout: mb.svdcmp(6, 6, a_, layout=layout), layout, A)
        _, rw, _ = port.svd(6, 6, A[good])
        assert np.abs(np.sort(w[good], axis=1) - np.sort(rw, axis=1)).max() < 1e-12 * rw.max()
        rec = np.einsum("nij,nj,nkj->nik", U[good].reshape(-1, 6, 6).transpose(0, 2, 1), w[good],
                        V[good].reshape(-1, 6, 6).transpose(0, 2, 1))
        assert np.abs(rec - A[good].reshape(-1, 6, 6).transpose(0, 2, 1)).max() < 1e-13
        assert not np.isfinite(w[bad]).all(axis=1).any()  # the poisoned instances do not pretend to have an answer


@pytest.mark.timeout(120)
def test_nan_inputs_end_every_iterative_kernel(mb, port):
    """a few NaN instances in a batch: every data-dependent loop (Jacobi sweeps by warp vote, the persistent Newton
    loop and the trust-region optimiser with their iteration caps) still ends, and the ordinary instances next to them
    get the ordinary answers"""
    n = 2051
    bad = np.arange(3, n, 41)
    good = np.ones(n, dtype=bool)
    good[bad] = False
    S4 = synth.symmetric(n, 4, seed=0xBA1)
    S4[bad, 5] = np.nan
    d, z = mb.eig(4, dev(S4))
    rd, _ = port.eig_sym(4, S4[good], True)
    assert np.abs(d.cpu().numpy()[good] - rd).max() < 1e-12
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    kin = mb.Kinematics(S, M)
    q = synth.joint_angles(n, seed=0xBA2)
    Tt = port.poe_fk_jacobian(S, M, q + synth.ik_offsets(n), want_Js=False)[0]
    qn, Tn = q.copy(), Tt.copy()
    qn[bad[::2], 3] = np.nan
    Tn[bad[1::2], 12] = np.nan
    from matrix_b200 import _lib
    for lu in (True, False):
        _lib.set_square_solve_lu(lu)
        try:
            out = kin.ik_newton_step(dev(qn), dev(Tn)).cpu().numpy()
        finally:
            _lib.set_square_solve_lu(True)
        ref = port.ik_newton_step(S, M, q[good], Tt[good])[0]
        assert np.abs(out[good] - ref).max() < 1e-6 and not np.isfinite(out[bad]).all()
    qs, it, st = kin.inverseSpaceNewton(dev(Tn), dev(qn), 20)
    assert (st.cpu().numpy()[good] == mb.CONVERGED).mean() > 0.98 and int(it.max()) <= 20
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    qc, stc, fn = kin.inverseSpace(dev(Tn), dev(np.where(np.isfinite(qn), qn, 0.0) * 0.95), lo, hi)
    torch.cuda.synchronize()
    assert set(np.unique(stc.cpu().numpy())) <= {mb.CONVERGED, mb.DIVERGED}


@pytest.mark.parametrize("layout", LAYOUTS)
def test_linsolve_svd_6x6(mb, port, layout):
    n = 1 << 14
    A = synth.dense(n, 6, 6)
    b = synth.dense(n, 6, 1, seed=0xB6)
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.SVD, 6, 6, a_, b_, layout=layout), layout, A, b)
    rx = port.linsolve_svd(6, 6, A, b)
    kappa = np.linalg.cond(A.reshape(-1, 6, 6).transpose(0, 2, 1))
    err = rel_err(x, rx)
    assert np.all(err <= REL * np.maximum(1.0, kappa))
    assert (err < REL).mean() > 0.97
    assert np.all(st == 0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_eig_sym_4(mb, cpu_ref, layout):
    n = 1 << 15
    S = synth.symmetric(n, 4)
    d, z = run(mb, lambda s_, layout: mb.eig(4, s_, sorted=True, layout=layout), layout, S)
    rd, rz = cpu_ref.eig_sym(4, S, True)
    scale = np.abs(rd).max(axis=1)
    assert np.all(np.abs(d - rd).max(axis=1) / scale < REL)
    assert np.all(np.diff(d, axis=1) >= 0)  # ascending, like eigsrt (linalg.hpp:1137-1151)
    Sm = S.reshape(n, 4, 4)  # symmetric: no transpose needed
    Zm = z.reshape(n, 4, 4).transpose(0, 2, 1)  # [i, row, col]
    res = np.einsum("nij,njk->nik", Sm, Zm) - Zm * d[:, None, :]
    assert np.abs(res).max() < 1e-13 * 4
    assert np.abs(np.einsum("nji,njk->nik", Zm, Zm) - np.eye(4)).max() < 1e-13
    # eigenvectors agree with the reference's up to sign wherever the eigenvalue is isolated
    Rm = rz.reshape(n, 4, 4).transpose(0, 2, 1)
    gap = np.min(np.diff(rd, axis=1), axis=1) / scale
    ok = gap > 1e-3
    dots = np.abs(np.einsum("nji,nji->ni", Zm[ok], Rm[ok]))
    assert np.abs(dots - 1).max() < 1e-10
    # unsorted / values-only variants carry the same spectrum
    d2, _ = run(mb, lambda s_, layout: mb.eig(4, s_, sorted=False, vectors=False, layout=layout), layout, S)
    assert np.abs(np.sort(d2, axis=1) - d).max() < 1e-13


def test_eig_golden_and_edge_cases(mb, fixtures, golden):
    g = golden["eig_sym_4"]
    d, z = run(mb, lambda s_, layout: mb.eig(4, s_, layout=layout), "aos", np.array(g["S"], dtype=float).reshape(1, 16))
    assert np.abs(d[0] - np.array(g["d"])).max() < 1e-14
    lit = np.array(g["z_literal"])
    expected = np.array([-lit[c + 4 * r] for c in range(4) for r in range(4)]).reshape(4, 4)  # [col, row]
    Z = z[0].reshape(4, 4)
    for c in range(4):  # each eigenvector up to a global sign (the reference's own literal is "* (-1)")
        assert min(np.abs(Z[c] - expected[c]).max(), np.abs(Z[c] + expected[c]).max()) < g["tol"]
    S = fixtures["sym4.S"][-7:]
    d, z = run(mb, lambda s_, layout: mb.eig(4, s_, layout=layout), "aos", S)
    rd = fixtures["sym4.d_sorted"][-7:]
    # zero matrix and all-ones matrix: the reference's tred2/tqli divide 0 by 0 and return NaN; Jacobi returns
    # the exact spectra (0,0,0,0) and (0,0,0,4) -- a deliberate difference
    assert np.isnan(rd[1]).all() and np.all(d[1] == 0)
    assert np.isnan(rd[4]).all() and np.abs(d[4] - np.array([0, 0, 0, 4.0])).max() < 1e-14
    good = ~np.isnan(rd).any(axis=1)
    assert good.sum() == 5 and np.abs(d[good] - rd[good]).max() < 1e-13
    for n in (2, 3, 5, 6):
        Sn = synth.symmetric(2048, n, seed=0x71 + n)
        dn, zn = run(mb, lambda s_, layout: mb.eig(n, s_, layout=layout), "soa", Sn)
        ref = np.linalg.eigvalsh(Sn.reshape(-1, n, n))
        assert np.abs(dn - ref).max() < 1e-13 * n
    d6, _ = run(mb, lambda s_, layout: mb.eig(6, s_, layout=layout), "aos", fixtures["sym6.S"])
    assert np.abs(d6 - fixtures["sym6.d_sorted"]).max() < 1e-13 * 6


# --------------------------------------------------------------------------------------- plumbing
def test_layouts_give_identical_bits(mb):
    xi = synth.se3_twists(4097)
    assert np.array_equal(run(mb, mb.se3_exp_log, "aos", xi), run(mb, mb.se3_exp_log, "soa", xi))
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q = synth.joint_angles(4097)
    a, b = run(mb, kin.fk_jacobian, "aos", q), run(mb, kin.fk_jacobian, "soa", q)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 129, 1025])
def test_empty_and_ragged_batches(mb, port, n):
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q = synth.joint_angles(n)
    for layout in LAYOUTS:
        guard = torch.full((n + 64, 16), 7.0, dtype=torch.float64, device="cuda")
        if layout == "aos":
            T, Js = kin.fk_jacobian(dev(q).view(n, 6), layout=mb.AOS, out_T=guard[:n])
            T = T.cpu().numpy()
            assert torch.all(guard[n:] == 7.0)  # nothing written past the batch
        else:
            T, Js = run(mb, kin.fk_jacobian, layout, q)
        if n:
            assert rel_err(T, port.poe_fk_jacobian(synth.UR5_SCREWS, synth.UR5_HOME, q)[0]).max() < REL


def test_soa_padding_is_untouched(mb):
    n, ld = 1000, 1024
    xi = synth.se3_twists(n)
    x = torch.full((6, ld), float("nan"), dtype=torch.float64, device="cuda")
    x[:, :n] = dev(xi).t()
    out = torch.full((16, ld), 7.0, dtype=torch.float64, device="cuda")
    mb.se3_exp(x, layout=mb.SOA, n=n, out=out)
    assert torch.all(out[:, n:] == 7.0) and torch.isfinite(out[:, :n]).all()


def test_device_generator_matches_numpy(mb):
    t = torch.empty(100003, dtype=torch.float64, device="cuda")
    mb.fill_uniform(t, synth.SEEDS["ur5_fk_jacobian"], 12345, -np.pi, np.pi)
    assert np.array_equal(t.cpu().numpy(), synth.uniform(synth.SEEDS["ur5_fk_jacobian"], 12345, 100003, -np.pi, np.pi))


def test_host_abi_matches_device_abi(mb, port):
    from matrix_b200 import host
    n = 300007  # several pipeline chunks for the wide outputs
    q = host.pinned_empty((n, 6))
    q[:] = synth.joint_angles(n)
    kin = host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    T, Js = kin.fk_jacobian(q)
    dk = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    dT, dJ = dk.fk_jacobian(dev(q))
    assert np.array_equal(T, dT.cpu().numpy()) and np.array_equal(Js, dJ.cpu().numpy())
    xi = synth.se3_twists(50001)
    assert np.array_equal(host.se3_exp_log(xi), mb.se3_exp_log(dev(xi)).cpu().numpy())
    A, b = synth.lstsq_4x3(50001)
    x, st = host.linsolve("QR", 4, 3, A, b)
    rx, rst = port.linsolve_qr(4, 3, A, b)
    assert np.array_equal(st, rst) and np.abs(x - rx).max() < 1e-9
    d, z = host.eig(4, synth.symmetric(1000, 4))
    assert np.all(np.diff(d, axis=1) >= 0)


def test_errors_are_loud(mb):
    from matrix_b200._lib import PpxCudaError
    x = torch.zeros((8, 6), dtype=torch.float64, device="cuda")
    with pytest.raises(PpxCudaError):
        mb.linsolve(mb.QR, 7, 2, torch.zeros((8, 14), dtype=torch.float64, device="cuda"),
                    torch.zeros((8, 7), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        mb.se3_exp(x, layout=5)
    with pytest.raises(TypeError):
        mb.se3_exp(x.cpu())


# --------------------------------------------------------------------------------------- full BASELINE sizes
def test_full_size_properties_16M_fk_jacobian(mb):
    """BASELINE config 2 at full size (2^24), checked through size-independent properties:
    R orthonormal, bottom row exact, Js column 0 = S0, |omega-part of each column| = 1 (unit screws),
    and equality with an independent second launch on a shard (bitwise)."""
    n = 1 << 24
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    q = torch.empty((6, n), dtype=torch.float64, device="cuda")
    mb.fill_uniform(q, synth.SEEDS["ur5_fk_jacobian"], 0, -np.pi, np.pi)
    T, Js = kin.fk_jacobian(q, layout=mb.SOA)
    R = T[[0, 1, 2, 4, 5, 6, 8, 9, 10]].view(3, 3, n)  # [col, row, i]
    G = torch.einsum("cri,dri->cdi", R, R)
    eye = torch.eye(3, dtype=torch.float64, device="cuda")[:, :, None]
    assert (G - eye).abs().max().item() < 1e-14
    assert torch.all(T[[3, 7, 11]] == 0) and torch.all(T[15] == 1)
    S0 = torch.tensor(synth.UR5_SCREWS[0], device="cuda")[:, None]
    assert torch.all(Js[:6] == S0)
    wn = Js.view(6, 6, n)[:, :3].norm(dim=1)
    assert (wn - 1).abs().max().item() < 1e-14
    lo, hi = 5_000_000, 5_065_537
    T2, J2 = kin.fk_jacobian(q[:, lo:hi].contiguous(), layout=mb.SOA)
    assert torch.equal(T2, T[:, lo:hi]) and torch.equal(J2, Js[:, lo:hi])
    del T, Js, G, R
    torch.cuda.empty_cache()


def test_full_size_round_trip_1M(mb):
    """BASELINE config 1 at full size: log(exp(xi)) = xi."""
    n = 1 << 20
    xi = dev(synth.se3_twists(n))
    back = mb.se3_exp_log(xi)
    err = (back - xi).abs().max(dim=1).values / xi.abs().max(dim=1).values
    assert err.max().item() < 1e-11


# --------------------------------------------------------------------------------------- C++ host API
def test_cpp_host_api(mb, tmp_path):
    """include/ppx/ppx.hpp (C++14): the reference's own test bodies, replayed on the GPU (tests/cpp/host_api_test.cpp)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_api_test"
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_api_test.cpp"), "-L",
                           os.path.join(root, "matrix_b200"), "-lppx_cuda",
                           f"-Wl,-rpath,{os.path.join(root, 'matrix_b200')}", "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "all passed" in r.stdout, r.stdout + r.stderr


# --------------------------------------------------------------------------------------- SURVEY 8(f) rows 2-3
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_lu_solve_inverse_det(mb, port, layout, n):
    cnt = 20011
    A = synth.dense(cnt, n, n, seed=0x81 + n)
    b = synth.dense(cnt, n, 1, seed=0x91 + n)
    kappa = np.linalg.cond(A.reshape(cnt, n, n).transpose(0, 2, 1))
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.LU, n, n, a_, b_, layout=layout), layout, A, b)
    rx, rst = port.linsolve_lu(n, A, b)
    assert np.array_equal(st, rst)
    assert np.all(rel_err(x, rx) <= REL * np.maximum(1.0, kappa))
    Ai = run(mb, lambda a_, layout: mb.inverse(n, a_, layout=layout), layout, A)
    assert np.all(rel_err(Ai, port.inverse(n, A)) <= REL * np.maximum(1.0, kappa))
    d = run(mb, lambda a_, layout: mb.det(n, a_, layout=layout), layout, A)
    rd = port.det(n, A)
    assert np.all(np.abs(d[:, 0] - rd) <= 1e-13 * np.maximum(np.abs(rd), np.abs(A).max(axis=1) ** n))


def test_lu_singular_and_golden(mb, fixtures, golden):
    A6, b6 = fixtures["dense6.A"], fixtures["dense6.b"]
    x, st = run(mb, lambda a_, b_, layout: mb.linsolve(mb.LU, 6, 6, a_, b_, layout=layout), "aos", A6, b6)
    assert np.array_equal(st, fixtures["dense6.status_lu"])
    sing = st == mb.SINGULAR
    assert sing.sum() >= 3 and np.array_equal(x[sing], b6[sing])  # b untouched (linalg.hpp:1182-1186)
    Ai = run(mb, lambda a_, layout: mb.inverse(6, a_, layout=layout), "aos", A6)
    assert np.all(Ai[sing] == 0.0)  # MatrixS::I() returns {} (linalg.hpp:938-941)
    d = run(mb, lambda a_, layout: mb.det(6, a_, layout=layout), "aos", A6)
    assert np.all(d[sing] == 0.0)
    g = golden["inverse_2x2"]  # test/matrix_test.cpp:809-817
    A = np.array(g["A"], dtype=float).reshape(1, 4)
    assert np.abs(run(mb, lambda a_, layout: mb.inverse(2, a_, layout=layout), "aos", A)[0] - np.array(g["inverse"])).max() < 1e-15
    assert abs(run(mb, lambda a_, layout: mb.det(2, a_, layout=layout), "aos", A)[0, 0] - g["det"]) < 1e-14


@pytest.mark.parametrize("shape", [(4, 3), (6, 6), (5, 4)])
def test_pinv_and_norm2(mb, port, shape):
    m, n = shape
    cnt = 8192
    A = synth.dense(cnt, m, n, seed=0xA1 + m * 8 + n)
    kappa = np.linalg.cond(A.reshape(cnt, n, m).transpose(0, 2, 1))
    for layout in LAYOUTS:
        P = run(mb, lambda a_, layout: mb.pinv(m, n, a_, layout=layout), layout, A)
        if (m, n) != (6, 4):
            assert np.all(rel_err(P, port.pinv(m, n, A)) <= REL * np.maximum(1.0, kappa))
        s = run(mb, lambda a_, layout: mb.norm2(m, n, a_, layout=layout), layout, A)
        assert np.abs(s[:, 0] / port.norm2(m, n, A) - 1).max() < REL


def test_norm2_golden(mb, golden):
    """test/solver_test.cpp:41-104: sigma_max of a 5x4, a 4x5 (wide) and a 4x4 matrix to 10 eps"""
    g = golden["matrix_norm2"]
    for case in g["cases"]:
        A = np.array(case["A"]).reshape(1, -1)
        s = run(mb, lambda a_, layout: mb.norm2(case["m"], case["n"], a_, layout=layout), "aos", A)
        assert abs(s[0, 0] - case["norm2"]) < g["tol"], case["m"]


@pytest.mark.parametrize("scale", [1e-250, 1e-100, 1e-9, 1e9, 1e100, 1e250])
def test_svd_eig_scale_with_their_input(mb, port, scale):
    """The Jacobi path pre-scales by an exact power of two (integer exponent arithmetic), so singular values and
    eigenvalues scale exactly with the input at any magnitude.  The reference is compared where it is itself valid:
    its absolute EPS_DP tests (linalg.hpp:863, 1105) break it for small inputs and its unscaled f*f + h*h
    (:803, 846) overflows for large ones."""
    A = synth.dense(2048, 6, 6, seed=0xC1)
    _, w1, _ = run(mb, lambda a_, layout: mb.svdcmp(6, 6, a_, layout=layout), "soa", A)
    _, w2, _ = run(mb, lambda a_, layout: mb.svdcmp(6, 6, a_, layout=layout), "soa", A * scale)
    assert np.all(np.isfinite(w2))
    assert np.all(np.abs(np.sort(w2, axis=1) / scale - np.sort(w1, axis=1)).max(axis=1) / w1.max(axis=1) < 1e-13)
    S = synth.symmetric(2048, 4, seed=0xC3)
    d1, _ = run(mb, lambda s_, layout: mb.eig(4, s_, layout=layout), "soa", S)
    d2, _ = run(mb, lambda s_, layout: mb.eig(4, s_, layout=layout), "soa", S * scale)
    assert np.all(np.isfinite(d2)) and np.abs(d2 / scale - d1).max() < 1e-13
    if 1e-3 <= scale <= 1e100:
        _, rw, _ = port.svd(6, 6, A * scale)
        assert np.all(np.abs(np.sort(w2, axis=1) - np.sort(rw, axis=1)).max(axis=1) / rw.max(axis=1) < REL)
        rd, _ = port.eig_sym(4, S * scale, True)
        assert np.all(np.abs(d2 - rd).max(axis=1) / np.abs(rd).max(axis=1) < REL)


# --------------------------------------------------------------------------------------- SURVEY 8(f) row 4
def test_inverse_space_codo(mb, port):
    """kinematics<6>::inverseSpace (bounded trust-region dogleg, demo/robotics.hpp:126-152) vs the oracle."""
    n = 1 << 13
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    q = synth.joint_angles(n, seed=synth.SEEDS["ik_newton_step"]) * 0.95  # inside the (-pi, pi) joint ranges
    Tt = port.poe_fk_jacobian(S, M, q + synth.ik_offsets(n), want_Js=False)[0]
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)  # demo/main.cpp:363-368
    kin = mb.Kinematics(S, M)
    for layout in LAYOUTS:
        qs, st, fn = run(mb, lambda t_, q_, layout: kin.inverseSpace(t_, q_, lo, hi, layout=layout), layout, Tt, q)
        rq, rx, rfn, rst = port.ik_codo(S, M, lo, hi, Tt, q)
        assert (st == rst).mean() > 0.999
        conv = (st == mb.CONVERGED) & (rst == mb.CONVERGED)
        assert conv.mean() > 0.99
        # CoDo reports CONVERGED unless the trust region collapses -- also when it stops on "no improvement" or on a
        # box-constrained minimum (optimization.hpp:590-593, 772-775) -- so "solved" is |f| <= FTOLA = EPS_SP (:501, 567)
        solved, rsolved = fn <= 0.5 * 1.1920928955078125e-07, rfn <= 0.5 * 1.1920928955078125e-07
        assert solved[conv].mean() > 0.99 and abs(solved.mean() - rsolved.mean()) < 5e-3
        assert (solved == rsolved).mean() > 0.995
        conv &= solved & rsolved
        # solved instances reproduce the target pose to that accuracy, like the reference's result does
        Tq = port.poe_fk_jacobian(S, M, qs, want_Js=False)[0]
        assert np.abs(Tq[conv] - Tt[conv]).max() < 5e-7
        # the trajectories are rounding-sensitive (trust-region decisions); the converged joint vectors still agree to
        # the optimiser's tolerance times the conditioning of the solution
        Jf = port.poe_fk_jacobian(S, M, rq, want_T=False)[1]
        sv = np.linalg.svd(Jf.reshape(n, 6, 6), compute_uv=False)
        kf = sv[:, 0] / sv[:, -1]
        ok = conv & (kf < 1e2)
        assert ok.mean() > 0.5
        assert np.all(np.abs(qs[ok] - rq[ok]).max(axis=1) < 1e-6 * kf[ok])
    # illegal start (outside the box) -> DIVERGED and init returned (optimization.hpp:552-556, robotics.hpp:151)
    bad = q[:4].copy()
    bad[:, 0] = 4.0
    qs, st, fn = run(mb, lambda t_, q_, layout: kin.inverseSpace(t_, q_, lo, hi, layout=layout), "aos", Tt[:4], bad)
    assert np.all(st == mb.DIVERGED) and np.array_equal(qs, bad)
    # demo/main.cpp:369-375: the demo's own configuration, start 0.05 rad away on every joint
    qd = np.array([[0.336058, 1.69911, -0.445658, -1.25227, 0.00013744, -2.71554]])
    Td = port.poe_fk_jacobian(S, M, qd, want_Js=False)[0]
    qs, st, fn = run(mb, lambda t_, q_, layout: kin.inverseSpace(t_, q_, lo, hi, layout=layout), "aos", Td, qd - 0.05)
    assert st[0] == mb.CONVERGED and np.abs(qs - qd).max() < 1e-5


# --------------------------------------------------------------------------------------- full BASELINE sizes, configs 3-5
def _soa_workload(name, n=None):
    from matrix_b200.workloads import WORKLOADS
    cls = WORKLOADS[name]
    w = cls(n or cls.n_default, torch.device("cuda", 0))
    w.setup()
    w.step()
    torch.cuda.synchronize()
    return w


def test_full_size_64M_linsolve_qr(mb):
    """BASELINE config 3 at full size (2^26 systems): least-squares optimality A^T (A x - b) = 0 and status, in chunks."""
    w = _soa_workload("linsolve_qr_4x3")
    assert int((w.status != 0).sum()) == 0  # U(-1,1) 4x3 matrices: no squared column norm below eps
    worst = 0.0
    step = 1 << 23
    for lo in range(0, w.n, step):
        A = w.A[:, lo:lo + step].view(3, 4, -1)  # [col, row, i]
        x = w.x[:, lo:lo + step]
        r = torch.einsum("cri,ci->ri", A, x) - w.b[:, lo:lo + step]
        g = torch.einsum("cri,ri->ci", A, r)
        # |A^T r| relative to |A|^2 |x| + |A| |b| (backward-stable least squares)
        scale = (A * A).sum(dim=(0, 1)) * x.abs().max(dim=0).values + 2.0
        worst = max(worst, float((g.abs().max(dim=0).values / scale).max()))
    assert worst < 1e-13


def test_full_size_16M_svd_and_eig(mb):
    """BASELINE config 4 at full size: U diag(w) V^T = A, V^T V = I (SVD 6x6); S z = z d, sorted d (eig 4x4)."""
    w = _soa_workload("svd_6x6")
    step = 1 << 21
    for lo in range(0, w.n, step):
        A = w.A[:, lo:lo + step].view(6, 6, -1)  # [col, row, i]
        U = w.U[:, lo:lo + step].view(6, 6, -1)
        V = w.V[:, lo:lo + step].view(6, 6, -1)
        s = w.w[:, lo:lo + step]
        rec = torch.einsum("jri,ji,jci->cri", U, s, V)  # sum_j U[:,j] s_j V[c,j]
        assert float((rec - A).abs().max()) < 1e-13
        G = torch.einsum("ari,bri->abi", V, V)
        assert float((G - torch.eye(6, dtype=torch.float64, device="cuda")[:, :, None]).abs().max()) < 1e-13
        assert bool((s >= 0).all())
    del w
    torch.cuda.empty_cache()
    e = _soa_workload("eig_sym_4")
    S = e.S.view(4, 4, -1)
    Z = e.z.view(4, 4, -1)  # [col, row, i]
    res = torch.einsum("kri,jki->jri", S, Z) - Z * e.d[:, None, :]
    assert float(res.abs().max()) < 1e-13
    assert bool((e.d[1:] >= e.d[:-1]).all())


def test_full_size_32M_ik_newton_step(mb):
    """BASELINE config 5 at its per-GPU size (2^25 = 256M / 8): one Newton step reduces the pose error
    |log(T(q)^-1 Ttarget)| quadratically wherever the Jacobian is well conditioned."""
    w = _soa_workload("ik_newton_step")

    def pose_err(q):
        T = w.kin.forwardSpace(q, layout=mb.SOA)
        X = mb.SE3_mul(mb.SE3_inv(T, layout=mb.SOA), w.Tt, layout=mb.SOA)
        return mb.SE3_log(X, layout=mb.SOA).norm(dim=0)

    e0, e1 = pose_err(w.q), pose_err(w.qo)
    assert bool(torch.isfinite(w.qo).all())
    good = e1 < 10 * e0 * e0 + 1e-12  # quadratic contraction
    assert float(good.double().mean()) > 0.97  # the rest sits near UR5 singularities
    assert float(e1.median()) < 0.05 * float(e0.median())


def test_full_size_inverse_space_and_newton_loop(mb):
    """SURVEY 8(f) rows 1 and 4 at bench size: kinematics<6>::inverseSpace on 2^24 targets and the Newton loop on 2^22,
    checked through what the solvers promise -- every instance reported solved reproduces its target pose, the answer
    stays inside the joint ranges, counts and status codes are in range -- all on the device."""
    w = _soa_workload("inverse_space_codo")
    q, st, fn = w.out
    assert bool(torch.isfinite(q).all()) and bool(((st == mb.CONVERGED) | (st == mb.DIVERGED)).all())
    solved = (st == mb.CONVERGED) & (fn <= 0.5 * 1.1920928955078125e-07)
    assert float(solved.double().mean()) > 0.99
    assert bool((q.abs() <= np.pi).all())  # the box of the demo, robotics.hpp:126-152 with joint ranges (-pi, pi)
    T = w.kin.forwardSpace(q)
    assert float((T - w.Tt)[solved].abs().max()) < 5e-7
    # instances that are not CONVERGED return the start (robotics.hpp:151)
    div = st == mb.DIVERGED
    if bool(div.any()):
        assert bool((q[div] == w.q0[div]).all())
    del w
    torch.cuda.empty_cache()
    n = 1 << 22
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    qi = torch.empty((n, 6), dtype=torch.float64, device="cuda")
    mb.fill_uniform(qi, synth.SEEDS["ik_newton_step"], 0, -np.pi, np.pi)
    d = torch.empty_like(qi)
    mb.fill_uniform(d, synth.SEEDS["ik_newton_step"] ^ 0xA5A5, 0, -0.05, 0.05)
    Tt = kin.forwardSpace(qi + d)
    qs, it, st = kin.inverseSpaceNewton(Tt, qi, 20)
    assert bool(((it >= 1) & (it <= 20)).all()) and bool(((st == mb.CONVERGED) | (st == mb.DIVERGED)).all())
    conv = st == mb.CONVERGED
    assert float(conv.double().mean()) > 0.99 and bool((it[~conv] == 20).all())
    # the loop stops on the residual BEFORE its last update (lie_test.cpp:103-110): the pose after it is off by the
    # square of that update, i.e. by ~(kappa * 1.2e-7)^2 -- inside the tolerance except next to a singularity
    err = (kin.forwardSpace(qs) - Tt).abs().amax(dim=1)[conv]
    assert float(err.kthvalue(int(0.999 * err.numel())).values) < 1.1920928955078125e-07 and float(err.max()) < 1e-3
    assert 3.5 < float(it.double().mean()) < 5.0


# --------------------------------------------------------------------------------------- out-of-bounds canaries
@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 1000])
def test_persistent_and_staged_kernels_stay_inside_their_buffers(mb, port, n):
    """The persistent kernels (ik_solve, ik_codo: instances claimed from a counter) and the cp.async-staged ones
    (inverse, linsolve<LU>) called through the raw C ABI on buffers embedded in canary-filled allocations, ragged n,
    both layouts: every output byte outside [0, n) keeps its canary, and the inside matches the oracle."""
    import ctypes as C
    from matrix_b200 import _lib
    lib, CAN = _lib.lib, -7.25
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    Sp = np.ascontiguousarray(S, dtype=np.float64)
    Mp = np.ascontiguousarray(M, dtype=np.float64)
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    q = synth.joint_angles(n, seed=0xD1) * 0.9
    Tt = port.poe_fk_jacobian(S, M, q + synth.ik_offsets(n), want_Js=False)[0]
    A = synth.dense(n, 6, 6, seed=0xD2)
    b = synth.dense(n, 6, 1, seed=0xD3)
    ref_codo = port.ik_codo(S, M, lo, hi, Tt, q)
    ref_newton = port.ik_solve(S, M, q, Tt, 20)
    ref_inv, (ref_x, _) = port.inverse(6, A), port.linsolve_lu(6, A, b)
    pad = 96
    for layout in (mb.AOS, mb.SOA):
        ld = n + 37 if layout == mb.SOA else n

        def put(a):  # device copy of the records in `layout`, leading dimension ld
            t = dev(a)
            if layout == mb.AOS:
                return t.contiguous()
            out = torch.zeros((a.shape[1], ld), dtype=torch.float64, device="cuda")
            out[:, :n] = t.t()
            return out

        def canary(width):  # (whole allocation, pointer to the embedded output)
            whole = torch.full((pad + width * ld + pad,), CAN, dtype=torch.float64, device="cuda")
            return whole, C.c_void_p(whole.data_ptr() + 8 * pad)

        def inside(whole, width):  # -> records [n, width]; asserts that everything else is still canary
            body = whole[pad:pad + width * ld]
            assert torch.all(whole[:pad] == CAN) and torch.all(whole[pad + width * ld:] == CAN)
            if layout == mb.AOS:
                return body.view(n, width).cpu().numpy()
            body = body.view(width, ld)
            assert torch.all(body[:, n:] == CAN)
            return body[:, :n].t().cpu().numpy()

        dq, dT, dA, db = put(q), put(Tt), put(A), put(b)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        ist = torch.full((n + 2 * pad,), -3, dtype=torch.int8, device="cuda")
        iit = torch.full((n + 2 * pad,), -3, dtype=torch.int32, device="cuda")
        ifn = torch.full((n + 2 * pad,), CAN, dtype=torch.float64, device="cuda")
        # inverseSpace (CoDo)
        wq, pq = canary(6)
        wx, px = canary(6)
        _lib.check(lib.ppx_cuda_batch_ik_codo(vp(Sp), vp(Mp), 6, vp(lo), vp(hi), p(dT), p(dq), pq, px,
                                              C.c_void_p(ifn.data_ptr() + 8 * pad), C.c_void_p(ist.data_ptr() + pad),
                                              n, layout, ld, None))
        torch.cuda.synchronize()
        qs, st = inside(wq, 6), ist[pad:pad + n].cpu().numpy()
        inside(wx, 6)
        assert torch.all(ist[:pad] == -3) and torch.all(ist[pad + n:] == -3)
        assert torch.all(ifn[:pad] == CAN) and torch.all(ifn[pad + n:] == CAN)
        fn = ifn[pad:pad + n].cpu().numpy()
        both = (st == mb.CONVERGED) & (ref_codo[3] == mb.CONVERGED) & (fn < 5e-8) & (ref_codo[2] < 5e-8)
        assert (st == ref_codo[3]).mean() >= 0.99 and both.sum() >= 0.9 * n - 1
        assert np.abs(port.poe_fk_jacobian(S, M, qs[both], want_Js=False)[0] - Tt[both]).max() < 5e-7
        # Newton loop
        ist.fill_(-3)
        wq, pq = canary(6)
        _lib.check(lib.ppx_cuda_batch_ik_solve(vp(Sp), vp(Mp), 6, p(dq), p(dT), 20, pq,
                                               C.c_void_p(iit.data_ptr() + 4 * pad), C.c_void_p(ist.data_ptr() + pad),
                                               n, layout, ld, None))
        torch.cuda.synchronize()
        qs = inside(wq, 6)
        assert torch.all(iit[:pad] == -3) and torch.all(iit[pad + n:] == -3)
        assert torch.all(ist[:pad] == -3) and torch.all(ist[pad + n:] == -3)
        it = iit[pad:pad + n].cpu().numpy()
        assert np.all((it >= 1) & (it <= 20)) and (it == ref_newton[1]).mean() >= 0.9
        conv = (ist[pad:pad + n].cpu().numpy() == mb.CONVERGED)
        assert conv.sum() >= 0.9 * n - 1
        assert np.abs(port.poe_fk_jacobian(S, M, qs[conv], want_Js=False)[0] - Tt[conv]).max() < 1.2e-7
        # inverse, linsolve<LU>
        wi, pi = canary(36)
        _lib.check(lib.ppx_cuda_batch_inverse(6, p(dA), pi, n, layout, ld, None))
        wl, pl = canary(6)
        ist.fill_(-3)
        _lib.check(lib.ppx_cuda_batch_linsolve_lu(6, p(dA), p(db), pl, C.c_void_p(ist.data_ptr() + pad), n, layout, ld,
                                                  None))
        torch.cuda.synchronize()
        kappa = np.linalg.cond(A.reshape(n, 6, 6).transpose(0, 2, 1))
        assert np.all(rel_err(inside(wi, 36), ref_inv) <= REL * np.maximum(1.0, kappa))
        assert np.all(rel_err(inside(wl, 6), ref_x) <= REL * np.maximum(1.0, kappa))
        assert torch.all(ist[:pad] == -3) and torch.all(ist[pad + n:] == -3) and torch.all(ist[pad:pad + n] == 0)
        # round 2: MatrixS::operator* (an unrolled shape and the generic kernel) and the IK Newton step in both solve
        # modes (for the UR5 these are the pattern kernels)
        for (mm, nn, ll) in ((6, 6, 6), (5, 4, 2)):
            Am, Bm = synth.dense(n, mm, nn, seed=0xD4), synth.dense(n, nn, ll, seed=0xD5)
            wc, pc = canary(mm * ll)
            dAm, dBm = put(Am), put(Bm)  # (kept alive past the launch)
            _lib.check(lib.ppx_cuda_batch_matmul(mm, nn, ll, p(dAm), p(dBm), pc, n, layout, ld, None))
            torch.cuda.synchronize()
            assert np.array_equal(inside(wc, mm * ll), port.matmul(mm, nn, ll, Am, Bm))
        ref_step = port.ik_newton_step(S, M, q, Tt)
        for lu in (1, 0):
            prev = lib.ppx_cuda_set_square_solve_lu(lu)
            try:
                wq, pq = canary(6)
                wv, pv = canary(6)
                _lib.check(lib.ppx_cuda_batch_ik_newton_step(vp(Sp), vp(Mp), 6, p(dq), p(dT), pq, pv, n, layout, ld, None))
                torch.cuda.synchronize()
            finally:
                lib.ppx_cuda_set_square_solve_lu(prev)
            assert rel_err(inside(wv, 6), ref_step[1]).max() < REL
            assert np.abs(inside(wq, 6) - ref_step[0]).max() < 1e-6


# --------------------------------------------------------------------------------------- square linsolve<SVD>: LU fast path
def test_square_pinv_solve_fast_path_and_fallback(mb, port, fixtures):
    """Square linsolve<SVD> systems and the IK Newton step take LU with partial pivoting unless a pivot is below 1e-8 of
    the largest; then the Jacobi SVD solve.  Both paths against the oracle's SVD::solve, the forced-SVD mode against
    the default one, and inputs that must take the fallback (near-singular Jacobians, the rank-deficient fixtures)."""
    from matrix_b200 import _lib
    n = 1 << 14
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    kin = mb.Kinematics(S, M)
    q, Tt, kappa = _ik_inputs(port, n, first=77)
    # near-singular configurations: wrist (q5 -> 0) and elbow (q3 -> 0) at distances 1e-3 .. 1e-11
    qs = q[:512].copy()
    qs[:256, 4] = 10.0 ** -np.linspace(3, 11, 256)
    qs[256:, 2] = 10.0 ** -np.linspace(3, 11, 256)
    Ts = port.poe_fk_jacobian(S, M, qs + synth.ik_offsets(512), want_Js=False)[0]
    Js = port.poe_fk_jacobian(S, M, qs, want_T=False)[1]
    sv = np.linalg.svd(Js.reshape(512, 6, 6).transpose(0, 2, 1), compute_uv=False)
    ks = sv[:, 0] / sv[:, -1]
    assert ks.max() > 1e9 and ks.min() < 1e5  # both sides of the pivot test
    A6, b6 = synth.dense(n, 6, 6, seed=0xE1), synth.dense(n, 6, 1, seed=0xE2)
    k6 = np.linalg.cond(A6.reshape(n, 6, 6).transpose(0, 2, 1))
    fxA, fxb = fixtures["dense6.A"], fixtures["dense6.b"]
    out = {}
    try:
        for mode in (True, False):
            _lib.set_square_solve_lu(mode)
            out[mode] = (
                run(mb, lambda q_, t_, layout: kin.ik_newton_step(q_, t_, layout=layout), "soa", q, Tt),
                run(mb, lambda q_, t_, layout: kin.ik_newton_step(q_, t_, layout=layout), "aos", qs, Ts),
                run(mb, lambda a_, b_, layout: mb.linsolve(mb.SVD, 6, 6, a_, b_, layout=layout), "soa", A6, b6)[0],
                run(mb, lambda a_, b_, layout: mb.linsolve(mb.SVD, 6, 6, a_, b_, layout=layout), "aos", fxA, fxb)[0],
                run(mb, lambda a_, layout: mb.pinv(6, 6, a_, layout=layout), "soa", A6),
                run(mb, lambda a_, layout: mb.pinv(6, 6, a_, layout=layout), "aos", fxA))
    finally:
        _lib.set_square_solve_lu(True)
    rq = port.ik_newton_step(S, M, q, Tt)[0]
    rqs = port.ik_newton_step(S, M, qs, Ts)[0]
    rx = port.linsolve_svd(6, 6, A6, b6)
    rP, rPf = port.pinv(6, 6, A6), port.pinv(6, 6, fxA)
    for mode in (True, False):
        dq, dqs, x, xf, P, Pf = out[mode]
        assert np.all(rel_err(P, rP) <= REL * np.maximum(1.0, k6))
        assert np.all(np.abs(dq - rq).max(axis=1) <= 1e-12 * kappa * np.maximum(1.0, np.abs(rq - q).max(axis=1)))
        step = np.maximum(1.0, np.abs(rqs - qs).max(axis=1))
        assert np.all(np.abs(dqs - rqs).max(axis=1) <= 1e-11 * ks * step)
        assert np.all(rel_err(x, rx) <= REL * np.maximum(1.0, k6))
        # the recorded reference outputs incl. zero / rank-deficient / Hilbert matrices (truncation active): the fallback
        ref = fixtures["dense6.x_svd"]
        ok = np.isfinite(ref).all(axis=1)
        kf = np.linalg.cond(fxA[ok].reshape(-1, 6, 6).transpose(0, 2, 1))
        good = kf < 1e12
        assert np.all(rel_err(xf[ok][good], ref[ok][good]) <= 1e-11 * np.maximum(1.0, kf[good]))
        okp = np.isfinite(rPf).all(axis=1)  # pinv of the same fixtures: zero, identity, rank 4, rank 1, Hilbert ...
        kp = np.linalg.cond(fxA[okp].reshape(-1, 6, 6).transpose(0, 2, 1))
        goodp = kp < 1e12
        assert np.all(rel_err(Pf[okp][goodp], rPf[okp][goodp]) <= 1e-11 * np.maximum(1.0, kp[goodp]))
        rank_def = okp.copy()
        rank_def[okp] = kp > 1e15  # truncation active: the SVD fallback must have been taken
        if rank_def.any():
            nz = np.abs(rPf[rank_def]).max(axis=1) > 0
            assert np.all(rel_err(Pf[rank_def][nz], rPf[rank_def][nz]) < 1e-6)
    # the two modes agree far inside the parity bar on well-conditioned systems
    well = kappa < 1e3
    assert np.abs(out[True][0][well] - out[False][0][well]).max() < 1e-11


def test_graded_triangular_systems_take_the_svd_fallback(mb, port):
    """I + t triu(ones,1): all partial-pivoting pivots are 1 while cond reaches 3.5e18, where the reference's
    SVD::solve / pinv drop the smallest singular value (linalg.hpp:889-901, 1211-1227).  The LU shortcut is gated by a
    rigorous condition bound (lu_cond_bound), so these lanes get the Jacobi SVD answer -- in the default mode, mixed
    into a batch of ordinary matrices, on both layouts."""
    from test_math_host import _graded_triangular
    G = _graded_triangular()
    A = np.ascontiguousarray(np.vstack([synth.dense(4000, 6, 6, seed=0xE7), G, synth.dense(77, 6, 6, seed=0xE8)]))
    n = A.shape[0]
    b = np.ascontiguousarray(synth.dense(n, 6, 1, seed=0xE9))
    rx, rP = port.linsolve_svd(6, 6, A, b), port.pinv(6, 6, A)
    kap = np.linalg.cond(A.reshape(n, 6, 6).transpose(0, 2, 1))
    for layout in LAYOUTS:
        x = run(mb, lambda a_, b_, layout: mb.linsolve(mb.SVD, 6, 6, a_, b_, layout=layout), layout, A, b)[0]
        P = run(mb, lambda a_, layout: mb.pinv(6, 6, a_, layout=layout), layout, A)
        low = kap < 1e13
        assert np.all(rel_err(x[low], rx[low]) <= 1e-11 * np.maximum(1.0, kap[low]))
        assert np.all(rel_err(P[low], rP[low]) <= 1e-11 * np.maximum(1.0, kap[low]))
        trunc = kap > 1e15  # truncation active in the reference: minimum-norm answers of size ~1/t, not ~t^5
        assert trunc.sum() >= 6
        assert np.abs(rx[trunc]).max() < 1.0 and np.abs(x[trunc]).max() < 1.0
        assert np.all(rel_err(x[trunc], rx[trunc]) < 1e-6) and np.all(rel_err(P[trunc], rP[trunc]) < 1e-6)


# --------------------------------------------------------------------------------------- several devices
def _fk_host(host, q, devices):
    host.set_devices(devices)
    try:
        return host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME).fk_jacobian(q)
    finally:
        host.set_devices(None)


def test_host_face_device_list(mb):
    """ppx_cuda_host_set_devices: explicit list = the worker-thread path (exercised with the one device every box has),
    bad lists are refused, and the result does not depend on the list."""
    from matrix_b200 import host
    from matrix_b200._lib import PpxCudaError
    n = 700_001  # > 5 chunks of 48 MiB / 416 B
    q = host.pinned_empty((n, 6))
    q[:] = synth.joint_angles(n, seed=0xD1)
    T0, J0 = _fk_host(host, q, None)
    T1, J1 = _fk_host(host, q, [0])
    assert np.array_equal(T0, T1) and np.array_equal(J0, J1)
    assert host.get_devices() == []
    host.set_devices([0])
    assert host.get_devices() == [0]
    host.set_devices(None)
    for bad in ([0, 0], [torch.cuda.device_count()], [-1]):
        with pytest.raises(PpxCudaError):
            host.set_devices(bad)
    assert host.get_devices() == []
    rates = host.copy_probe(q, "d2h", 0.2)
    assert len(rates) == 1 and rates[0] > 1.0  # GB/s
    devs, agg = host.tune_devices([0], q, "d2h", 0.05)  # one candidate: it is chosen, and becomes the device list
    assert devs == [0] and host.get_devices() == [0] and agg > 1.0
    host.set_devices(None)


def test_device_list_from_the_environment(mb):
    """PPX_CUDA_DEVICES is read once, at the first host-face call of a process: "all", a list, or garbage (ignored)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import numpy as np; from matrix_b200 import host; "
            "x = host.se3_exp_log(np.full((4, 6), 0.25)); print(host.get_devices(), float(x[0, 0]))")
    ndev = torch.cuda.device_count()
    for value, want in (("all", list(range(ndev))), ("0", [0]), ("0,0", []), ("7x", []), ("", [])):
        env = dict(os.environ, PPX_CUDA_DEVICES=value)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        got, val = r.stdout.strip().rsplit("] ", 1)
        assert eval(got + "]") == want, (value, r.stdout)
        assert abs(float(val) - 0.25) < 1e-12


def test_two_device_host_call_is_bitwise_the_one_device_result(mb, port):
    """one host call spread over two GPUs (one worker thread per device, chunks from a shared queue) gives
    bit for bit what one GPU gives -- and both match the oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from matrix_b200 import host
    n = 1_500_003
    q = host.pinned_empty((n, 6))
    q[:] = synth.joint_angles(n, seed=0xD2)
    T1, J1 = _fk_host(host, q, [0])
    T2, J2 = _fk_host(host, q, [0, 1])
    Tb, Jb = _fk_host(host, q, [1])
    Ta, Ja = _fk_host(host, q, "all")
    for T, J in ((T2, J2), (Tb, Jb), (Ta, Ja)):
        assert np.array_equal(T1, T) and np.array_equal(J1, J)
    m = 1 << 15
    rT, rJ = port.poe_fk_jacobian(synth.UR5_SCREWS, synth.UR5_HOME, q[-m:])
    assert rel_err(T2[-m:], rT).max() < REL and rel_err(J2[-m:], rJ).max() < REL
    # a solver with per-instance status and an optional output left out, over both devices
    A, b = synth.lstsq_4x3(2_000_001)
    host.set_devices([0, 1])
    try:
        x2, st2 = host.linsolve("QR", 4, 3, A, b)
    finally:
        host.set_devices(None)
    x1, st1 = host.linsolve("QR", 4, 3, A, b)
    assert np.array_equal(x1, x2) and np.array_equal(st1, st2)


def test_two_gpu_shards_concatenate_to_the_one_gpu_result(mb):
    """the torchrun form of the same property: block r of the batch computed on device r (device-resident, SoA),
    concatenated, is bitwise the whole batch computed on device 0"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from matrix_b200.shard import shard_ranges
    n = 1_000_003
    kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
    whole = torch.empty((6, n), dtype=torch.float64, device="cuda:0")
    mb.fill_uniform(whole, 0xD3, 0, -np.pi, np.pi)  # component-major fill: plane k holds counters k*n ..
    T, Js = kin.fk_jacobian(whole, layout=mb.SOA)
    parts_T, parts_J = [], []
    for r, (first, cnt) in enumerate(shard_ranges(n, 2)):
        with torch.cuda.device(r):
            qr = whole[:, first:first + cnt].to(f"cuda:{r}").contiguous()
            Tr, Jr = kin.fk_jacobian(qr, layout=mb.SOA)
            parts_T.append(Tr.cpu())
            parts_J.append(Jr.cpu())
    assert torch.equal(torch.cat(parts_T, dim=1), T.cpu()) and torch.equal(torch.cat(parts_J, dim=1), Js.cpu())


# --------------------------------------------------------------------------------------- the bench line itself
def test_bench_prints_the_contract_line(mb):
    """bench.py on a small batch: one JSON line with the contract's keys, a roofline that adds up, an end-to-end leg
    that moved the bytes it claims, and as many launches of this library's kernels as steps."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "4", "--warmup", "3", "--n", "1048576",
                        "--no-others", "--no-cpu", "--e2e-steps", "1"], capture_output=True, text=True, timeout=600,
                       cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "instances/sec" and d["n_gpus"] == 1 and d["steps"] == 4 and d["gpu_launches"] == 4
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert abs(rf["achieved"] - 1048576 * 464 / (rf["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]
    assert abs(d["value"] - 1048576 * 4 / (d["ms_per_step"] * 4e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 1048576 * 48 and e["d2h_bytes_per_step"] == 1048576 * 416 and 0 < e["value"] < d["value"]
    assert e["result_probe"] == 1.0  # T[-1][15] of the end-to-end call's output: the result really arrived
    assert d["configs"][0]["workload"] == "ur5_fk_jacobian" and d["config"]["instances_per_gpu"] == 1048576
    # the headline is the records kernel, the one the end-to-end call launches too
    assert "AOS" in rf["kernel"] and d["layout"].startswith("aos") and "records" in d["config"]["device_layout"]


def test_host_abi_is_thread_safe_and_keeps_its_scratch(mb, port):
    """Concurrent host-ABI calls from several threads (the streams and the scratch pool are shared, per device) give
    the results of serial calls; ppx_cuda_host_trim() hands the retained scratch back."""
    import threading
    from matrix_b200 import _lib, host
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    kin = host.Kinematics(S, M)
    qs = [synth.joint_angles(20000 + 977 * k, seed=0xF0 + k) for k in range(6)]
    ref = [port.poe_fk_jacobian(S, M, q) for q in qs]
    got, errs = [None] * len(qs), []

    def work(k):
        try:
            for _ in range(5):
                got[k] = kin.fk_jacobian(qs[k])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(qs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs
    for (T, Js), (rT, rJs) in zip(got, ref):
        assert rel_err(T, rT).max() < REL and rel_err(Js, rJs).max() < REL
    assert _lib.lib.ppx_cuda_host_trim() == 0
    T, Js = kin.fk_jacobian(qs[0])  # still works after the trim
    assert rel_err(T, ref[0][0]).max() < REL
