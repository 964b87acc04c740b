"""bench_cpu.py -- the CPU legs of bench.py (cpu_baseline and --impl reference).

This is the one place outside tests/ and __graft_entry__.smoke() that executes anything under oracle/: it times
the reference's own CPU routines (oracle/_ref/libppx_ref.so, the ppx headers behind ref_shim.cpp) or, when that
was not built, the C restatement, as the BASELINE the GPU numbers are reported next to.  It is never on the
product path: matrix_b200 does not import it.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from matrix_b200 import synth  # noqa: E402

# all the host threads this process may use -- stated explicitly on every call, because launchers such as torchrun
# export OMP_NUM_THREADS=1 and would silently turn the CPU baseline into a single-thread run
NTHREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def load_cpu_oracle():
    """-> (Oracle, kind): the compiled reference if oracle/_ref is there, else the C restatement."""
    from oracle_lib import Oracle, available
    if available("ref"):
        return Oracle("ref"), "reference"
    so = os.path.join(ROOT, "oracle", "libppx_oracle_fma.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    return Oracle("port_fma"), "port"


_dp = C.POINTER(C.c_double)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _ok(rc):
    if rc != 0:
        raise RuntimeError(f"CPU oracle call failed (rc={rc})")


def _touch(*arrays):
    for a in arrays:
        a.fill(0)  # fault the pages in before timing


def cpu_ur5_fk_jacobian(oracle, n, first=0):
    S = np.ascontiguousarray(synth.UR5_SCREWS)
    M = np.ascontiguousarray(synth.UR5_HOME)
    q = synth.joint_angles(n, first=first)
    T, Js = np.empty((n, 16)), np.empty((n, 36))
    _touch(T, Js)
    return lambda: _ok(oracle.lib.ppxo_poe_fk_jacobian(_d(S), _d(M), C.c_int(6), _d(q), _d(T), _d(Js), C.c_size_t(n),
                                                       C.c_int(NTHREADS)))


def cpu_se3_exp_log(oracle, n, first=0):
    xi = synth.se3_twists(n, first=first)
    out = np.empty_like(xi)
    _touch(out)
    return lambda: _ok(oracle.lib.ppxo_se3_exp_log(_d(xi), _d(out), C.c_size_t(n), C.c_int(NTHREADS)))


def cpu_linsolve_qr_4x3(oracle, n, first=0):
    A, b = synth.lstsq_4x3(n, first=first)
    x = np.empty((n, 3))
    st = np.empty(n, dtype=np.int8)
    _touch(x, st)
    return lambda: _ok(oracle.lib.ppxo_linsolve_qr(C.c_int(4), C.c_int(3), _d(A), _d(b), _d(x),
                                                   st.ctypes.data_as(C.POINTER(C.c_byte)), C.c_size_t(n), C.c_int(NTHREADS)))


def cpu_svd_6x6(oracle, n, first=0):
    A = synth.dense(n, 6, 6, first=first)
    U, w, V = np.empty((n, 36)), np.empty((n, 6)), np.empty((n, 36))
    _touch(U, w, V)
    return lambda: _ok(oracle.lib.ppxo_svd(C.c_int(6), C.c_int(6), _d(A), _d(U), _d(w), _d(V), C.c_size_t(n),
                                           C.c_int(NTHREADS)))


def cpu_eig_sym_4(oracle, n, first=0):
    S = synth.symmetric(n, 4, first=first)
    d, z = np.empty((n, 4)), np.empty((n, 16))
    _touch(d, z)
    return lambda: _ok(oracle.lib.ppxo_eig_sym(C.c_int(4), _d(S), _d(d), _d(z), C.c_int(1), C.c_size_t(n),
                                               C.c_int(NTHREADS)))


def cpu_ik_newton_step(oracle, n, first=0):
    S = np.ascontiguousarray(synth.UR5_SCREWS)
    M = np.ascontiguousarray(synth.UR5_HOME)
    q = synth.joint_angles(n, first=first, seed=synth.SEEDS["ik_newton_step"])
    Tt = oracle.poe_fk_jacobian(S, M, q + synth.ik_offsets(n, first=first), want_Js=False)[0]
    qo, Vs = np.empty((n, 6)), np.empty((n, 6))
    _touch(qo, Vs)
    return lambda: _ok(oracle.lib.ppxo_ik_newton_step(_d(S), _d(M), C.c_int(6), _d(q), _d(Tt), _d(qo), _d(Vs),
                                                      C.c_size_t(n), C.c_int(NTHREADS)))



def cpu_inverse_space_codo(oracle, n, first=0):
    S = np.ascontiguousarray(synth.UR5_SCREWS)
    M = np.ascontiguousarray(synth.UR5_HOME)
    q = synth.joint_angles(n, first=first, seed=synth.SEEDS["ik_newton_step"]) * 0.95
    Tt = oracle.poe_fk_jacobian(S, M, q, want_Js=False)[0]
    q0 = q - 0.05
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    qo = np.empty((n, 6))
    _touch(qo)
    return lambda: _ok(oracle.lib.ppxo_ik_codo(_d(S), _d(M), C.c_int(6), _d(lo), _d(hi), _d(Tt), _d(q0), _d(qo), None,
                                               None, None, C.c_size_t(n), C.c_int(NTHREADS)))


def _lie_inputs(oracle, n):
    xi = synth.se3_twists(n)
    T = oracle.se3_exp(xi)
    T2 = oracle.se3_exp(xi[::-1].copy())
    w = np.ascontiguousarray(xi[:, :3])
    R = oracle.so3_exp(w)
    return xi, T, T2, w, R


def _unary(oracle, fn, x, wout):
    out = np.empty((x.shape[0], wout))
    _touch(out)
    f = getattr(oracle.lib, fn)
    return lambda: _ok(f(_d(x), _d(out), C.c_size_t(x.shape[0]), C.c_int(NTHREADS)))


def cpu_op(name, oracle, n):
    """CPU closure for an entry of matrix_b200/ops_table.py (same distribution, n instances)"""
    S = np.ascontiguousarray(synth.UR5_SCREWS)
    M = np.ascontiguousarray(synth.UR5_HOME)
    if name in ("so3_exp", "SO3_log", "se3_exp", "SE3_log", "SE3_inv", "SE3_Adt", "se3_adt", "SE3_mul", "so3_ljacinv"):
        xi, T, T2, w, R = _lie_inputs(oracle, n)
        if name == "so3_exp":
            return _unary(oracle, "ppxo_so3_exp", w, 9)
        if name == "SO3_log":
            return _unary(oracle, "ppxo_SO3_log", R, 3)
        if name == "se3_exp":
            return _unary(oracle, "ppxo_se3_exp", xi, 16)
        if name == "SE3_log":
            return _unary(oracle, "ppxo_SE3_log", T, 6)
        if name == "SE3_inv":
            return _unary(oracle, "ppxo_SE3_inv", T, 16)
        if name == "SE3_Adt":
            return _unary(oracle, "ppxo_SE3_Adt", T, 36)
        if name == "se3_adt":
            return _unary(oracle, "ppxo_se3_adt", xi, 36)
        if name == "SE3_mul":
            out = np.empty_like(T)
            _touch(out)
            return lambda: _ok(oracle.lib.ppxo_SE3_mul(_d(T), _d(T2), _d(out), C.c_size_t(n), C.c_int(NTHREADS)))
        out = np.empty((n, 9))
        _touch(out)
        return lambda: _ok(oracle.lib.ppxo_so3_jac(C.c_int(1), _d(w), _d(out), C.c_size_t(n), C.c_int(NTHREADS)))
    if name.startswith("matmul_"):
        m, k, l = (int(x) for x in name[len("matmul_"):].split("x"))
        A, B = synth.dense(n, m, k, seed=0x301), synth.dense(n, k, l, seed=0x302)
        out = np.empty((n, m * l))
        _touch(out)
        return lambda: _ok(oracle.lib.ppxo_matmul(C.c_int(m), C.c_int(k), C.c_int(l), _d(A), _d(B), _d(out), C.c_size_t(n),
                                                  C.c_int(NTHREADS)))
    if name == "ik_solve_newton_loop":
        q = synth.joint_angles(n, seed=synth.SEEDS["ik_newton_step"])
        Tt = oracle.poe_fk_jacobian(S, M, q + synth.ik_offsets(n), want_Js=False)[0]
        qo = np.empty((n, 6))
        it = np.empty(n, dtype=np.int32)
        _touch(qo)
        return lambda: _ok(oracle.lib.ppxo_ik_solve(_d(S), _d(M), C.c_int(6), _d(q), _d(Tt), C.c_int(20), _d(qo),
                                                    it.ctypes.data_as(C.POINTER(C.c_int)), C.c_size_t(n), C.c_int(NTHREADS)))
    m, nn = {"linsolve_svd_6x6": (6, 6), "linsolve_lu_6x6": (6, 6), "inverse_6x6": (6, 6), "det_6x6": (6, 6),
             "pinv_4x3": (4, 3), "pinv_6x6": (6, 6), "norm2_6x6": (6, 6), "svd_4x3": (4, 3)}[name]
    A = synth.dense(n, m, nn, seed=0x301)
    b = synth.dense(n, m, 1, seed=0x302)
    st = np.empty(n, dtype=np.int8)
    bp = st.ctypes.data_as(C.POINTER(C.c_byte))
    if name == "linsolve_svd_6x6":
        x = np.empty((n, 6)); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_linsolve_svd(C.c_int(6), C.c_int(6), _d(A), _d(b), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "linsolve_lu_6x6":
        x = np.empty((n, 6)); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_linsolve_lu(C.c_int(6), _d(A), _d(b), _d(x), bp, C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "inverse_6x6":
        x = np.empty((n, 36)); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_inverse(C.c_int(6), _d(A), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "det_6x6":
        x = np.empty(n); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_det(C.c_int(6), _d(A), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "pinv_4x3":
        x = np.empty((n, 12)); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_pinv(C.c_int(4), C.c_int(3), _d(A), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "pinv_6x6":
        x = np.empty((n, 36)); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_pinv(C.c_int(6), C.c_int(6), _d(A), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    if name == "norm2_6x6":
        x = np.empty(n); _touch(x)
        return lambda: _ok(oracle.lib.ppxo_norm2(C.c_int(6), C.c_int(6), _d(A), _d(x), C.c_size_t(n), C.c_int(NTHREADS)))
    U, w, V = np.empty((n, 12)), np.empty((n, 3)), np.empty((n, 9))
    _touch(U, w, V)
    return lambda: _ok(oracle.lib.ppxo_svd(C.c_int(4), C.c_int(3), _d(A), _d(U), _d(w), _d(V), C.c_size_t(n), C.c_int(NTHREADS)))


CPU = {
    "ur5_fk_jacobian": cpu_ur5_fk_jacobian,
    "ur5_fk_jacobian_planes": cpu_ur5_fk_jacobian,
    "se3_exp_log": cpu_se3_exp_log,
    "se3_exp_log_1M": cpu_se3_exp_log,
    "linsolve_qr_4x3": cpu_linsolve_qr_4x3,
    "svd_6x6": cpu_svd_6x6,
    "eig_sym_4": cpu_eig_sym_4,
    "ik_newton_step": cpu_ik_newton_step,
    "ik_newton_step_svd": cpu_ik_newton_step,
    "inverse_space_codo": cpu_inverse_space_codo,
}
