"""Torch-free description of the BASELINE workloads: names, sizes and the CPU-side closure that runs the same
work through a CPU checker library (oracle/ppx_oracle.h interface).  Used by bench.py's cpu_baseline leg and
its --impl reference arm, neither of which needs a GPU."""
import ctypes as C

import numpy as np

from . import synth

HEADLINE = "ur5_fk_jacobian"

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
                                                       C.c_int(0)))


def cpu_se3_exp_log(oracle, n, first=0):
    xi = synth.se3_twists(n, first=first)
    out = np.empty_like(xi)
    _touch(out)
    return lambda: _ok(oracle.lib.ppxo_se3_exp_log(_d(xi), _d(out), C.c_size_t(n), C.c_int(0)))


def cpu_linsolve_qr_4x3(oracle, n, first=0):
    A, b = synth.lstsq_4x3(n, first=first)
    x = np.empty((n, 3))
    st = np.empty(n, dtype=np.int8)
    _touch(x, st)
    return lambda: _ok(oracle.lib.ppxo_linsolve_qr(C.c_int(4), C.c_int(3), _d(A), _d(b), _d(x),
                                                   st.ctypes.data_as(C.POINTER(C.c_byte)), C.c_size_t(n), C.c_int(0)))


def cpu_svd_6x6(oracle, n, first=0):
    A = synth.dense(n, 6, 6, first=first)
    U, w, V = np.empty((n, 36)), np.empty((n, 6)), np.empty((n, 36))
    _touch(U, w, V)
    return lambda: _ok(oracle.lib.ppxo_svd(C.c_int(6), C.c_int(6), _d(A), _d(U), _d(w), _d(V), C.c_size_t(n),
                                           C.c_int(0)))


def cpu_eig_sym_4(oracle, n, first=0):
    S = synth.symmetric(n, 4, first=first)
    d, z = np.empty((n, 4)), np.empty((n, 16))
    _touch(d, z)
    return lambda: _ok(oracle.lib.ppxo_eig_sym(C.c_int(4), _d(S), _d(d), _d(z), C.c_int(1), C.c_size_t(n),
                                               C.c_int(0)))


def cpu_ik_newton_step(oracle, n, first=0):
    S = np.ascontiguousarray(synth.UR5_SCREWS)
    M = np.ascontiguousarray(synth.UR5_HOME)
    q = synth.joint_angles(n, first=first, seed=synth.SEEDS["ik_newton_step"])
    Tt = oracle.poe_fk_jacobian(S, M, q + synth.ik_offsets(n, first=first), want_Js=False)[0]
    qo, Vs = np.empty((n, 6)), np.empty((n, 6))
    _touch(qo, Vs)
    return lambda: _ok(oracle.lib.ppxo_ik_newton_step(_d(S), _d(M), C.c_int(6), _d(q), _d(Tt), _d(qo), _d(Vs),
                                                      C.c_size_t(n), C.c_int(0)))


# cpu_sample: instances per CPU pass, sized for ~1-2 s per pass on a 16-core host
META = {
    "ur5_fk_jacobian": {"cpu": cpu_ur5_fk_jacobian, "cpu_sample": 1 << 22},
    "se3_exp_log": {"cpu": cpu_se3_exp_log, "cpu_sample": 1 << 24},
    "linsolve_qr_4x3": {"cpu": cpu_linsolve_qr_4x3, "cpu_sample": 1 << 25},
    "svd_6x6": {"cpu": cpu_svd_6x6, "cpu_sample": 1 << 22},
    "eig_sym_4": {"cpu": cpu_eig_sym_4, "cpu_sample": 1 << 23},
    "ik_newton_step": {"cpu": cpu_ik_newton_step, "cpu_sample": 1 << 21},
}
