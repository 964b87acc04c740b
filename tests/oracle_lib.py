"""ctypes front end for the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

Loads oracle/libppx_oracle.so (the C restatement, kind "port") or
oracle/_ref/libppx_ref*.so (the reference headers behind ref_shim.cpp, kind
"reference").  Both export oracle/ppx_oracle.h.  Arrays are numpy float64,
C-contiguous, one dense column-major reference record per row ("AoS").

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module; the product package matrix_b200 never does.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

PATHS = {
    "port": os.path.join(ORACLE_DIR, "libppx_oracle.so"),
    "port_fma": os.path.join(ORACLE_DIR, "libppx_oracle_fma.so"),
    "ref": os.path.join(ORACLE_DIR, "_ref", "libppx_ref.so"),
    "ref_strict": os.path.join(ORACLE_DIR, "_ref", "libppx_ref_strict.so"),
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_byte)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _in(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    a = a.reshape(a.shape[0], -1)
    assert a.shape[1] == cols, (a.shape, cols)
    return a


class Oracle:
    """One loaded checker library."""

    def __init__(self, which="port", nthreads=0):
        path = PATHS[which]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is not built (cd oracle && make)")
        self.which = which
        self.path = path
        self.nthreads = nthreads
        self.lib = C.CDLL(path)
        self.lib.ppxo_kind.restype = C.c_char_p
        self.kind = self.lib.ppxo_kind().decode()

    def max_threads(self):
        return int(self.lib.ppxo_max_threads())

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            raise ValueError(f"{name}: shape not available in oracle '{self.which}' (rc={rc})")

    # -- unary maps on fixed record sizes -------------------------------------------------
    def _map(self, name, x, cin, cout):
        x = _in(x, cin)
        out = np.empty((x.shape[0], cout))
        self._call(name, _d(x), _d(out), C.c_size_t(x.shape[0]), C.c_int(self.nthreads))
        return out

    def so3_exp(self, w):
        return self._map("ppxo_so3_exp", w, 3, 9)

    def SO3_log(self, R):
        return self._map("ppxo_SO3_log", R, 9, 3)

    def se3_exp(self, xi):
        return self._map("ppxo_se3_exp", xi, 6, 16)

    def SE3_log(self, T):
        return self._map("ppxo_SE3_log", T, 16, 6)

    def se3_exp_log(self, xi):
        return self._map("ppxo_se3_exp_log", xi, 6, 6)

    def SE3_inv(self, T):
        return self._map("ppxo_SE3_inv", T, 16, 16)

    def SE3_Adt(self, T):
        return self._map("ppxo_SE3_Adt", T, 16, 36)

    def se3_adt(self, xi):
        return self._map("ppxo_se3_adt", xi, 6, 36)

    def matmul(self, m, n, l, A, B):
        A = _in(A, m * n)
        B = _in(B, n * l)
        out = np.empty((A.shape[0], m * l))
        self._call("ppxo_matmul", C.c_int(m), C.c_int(n), C.c_int(l), _d(A), _d(B), _d(out), C.c_size_t(A.shape[0]),
                   C.c_int(self.nthreads))
        return out

    def SE3_mul(self, A, B):
        A = _in(A, 16)
        B = _in(B, 16)
        out = np.empty_like(A)
        self._call("ppxo_SE3_mul", _d(A), _d(B), _d(out), C.c_size_t(A.shape[0]), C.c_int(self.nthreads))
        return out

    def so3_jac(self, which, w):
        w = _in(w, 3)
        out = np.empty((w.shape[0], 9))
        self._call("ppxo_so3_jac", C.c_int(which), _d(w), _d(out), C.c_size_t(w.shape[0]), C.c_int(self.nthreads))
        return out

    # -- kinematics -----------------------------------------------------------------------
    def poe_fk_jacobian(self, screws, home, q, want_T=True, want_Js=True):
        screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        nj = screws.shape[0]
        home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        q = _in(q, nj)
        n = q.shape[0]
        T = np.empty((n, 16)) if want_T else None
        Js = np.empty((n, 6 * nj)) if want_Js else None
        self._call("ppxo_poe_fk_jacobian", _d(screws), _d(home), C.c_int(nj), _d(q), _d(T), _d(Js),
                   C.c_size_t(n), C.c_int(self.nthreads))
        return T, Js

    def ik_newton_step(self, screws, home, q, Tt):
        screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        nj = screws.shape[0]
        home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        q = _in(q, nj)
        Tt = _in(Tt, 16)
        n = q.shape[0]
        qo = np.empty((n, nj))
        Vs = np.empty((n, 6))
        self._call("ppxo_ik_newton_step", _d(screws), _d(home), C.c_int(nj), _d(q), _d(Tt), _d(qo), _d(Vs),
                   C.c_size_t(n), C.c_int(self.nthreads))
        return qo, Vs

    def ik_solve(self, screws, home, q0, Tt, max_iter=20):
        screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        nj = screws.shape[0]
        home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        q0 = _in(q0, nj)
        Tt = _in(Tt, 16)
        n = q0.shape[0]
        qo = np.empty((n, nj))
        it = np.empty(n, dtype=np.int32)
        self._call("ppxo_ik_solve", _d(screws), _d(home), C.c_int(nj), _d(q0), _d(Tt), C.c_int(max_iter), _d(qo),
                   it.ctypes.data_as(_ip), C.c_size_t(n), C.c_int(self.nthreads))
        return qo, it

    def ik_codo(self, screws, home, lo, hi, Tt, init):
        """kinematics<6>::inverseSpace (CoDo) -> (q_out, x_raw, fnorm, status)"""
        screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        nj = screws.shape[0]
        home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(nj)
        hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(nj)
        init = _in(init, nj)
        Tt = _in(Tt, 16)
        n = init.shape[0]
        qo, xr = np.empty((n, nj)), np.empty((n, nj))
        fn = np.empty(n)
        st = np.empty(n, dtype=np.int8)
        self._call("ppxo_ik_codo", _d(screws), _d(home), C.c_int(nj), _d(lo), _d(hi), _d(Tt), _d(init), _d(qo), _d(xr),
                   _d(fn), st.ctypes.data_as(_bp), C.c_size_t(n), C.c_int(self.nthreads))
        return qo, xr, fn, st

    # -- dense solvers --------------------------------------------------------------------
    def linsolve_qr(self, m, n, A, b):
        A = _in(A, m * n)
        b = _in(b, m)
        cnt = A.shape[0]
        x = np.empty((cnt, n))
        st = np.empty(cnt, dtype=np.int8)
        self._call("ppxo_linsolve_qr", C.c_int(m), C.c_int(n), _d(A), _d(b), _d(x), st.ctypes.data_as(_bp),
                   C.c_size_t(cnt), C.c_int(self.nthreads))
        return x, st

    def linsolve_lu(self, n, A, b):
        A = _in(A, n * n)
        b = _in(b, n)
        cnt = A.shape[0]
        x = np.empty((cnt, n))
        st = np.empty(cnt, dtype=np.int8)
        self._call("ppxo_linsolve_lu", C.c_int(n), _d(A), _d(b), _d(x), st.ctypes.data_as(_bp),
                   C.c_size_t(cnt), C.c_int(self.nthreads))
        return x, st

    def inverse(self, n, A):
        A = _in(A, n * n)
        out = np.empty_like(A)
        self._call("ppxo_inverse", C.c_int(n), _d(A), _d(out), C.c_size_t(A.shape[0]), C.c_int(self.nthreads))
        return out

    def det(self, n, A):
        A = _in(A, n * n)
        out = np.empty(A.shape[0])
        self._call("ppxo_det", C.c_int(n), _d(A), _d(out), C.c_size_t(A.shape[0]), C.c_int(self.nthreads))
        return out

    def svd(self, m, n, A):
        A = _in(A, m * n)
        cnt = A.shape[0]
        U = np.empty((cnt, m * n))
        w = np.empty((cnt, n))
        V = np.empty((cnt, n * n))
        self._call("ppxo_svd", C.c_int(m), C.c_int(n), _d(A), _d(U), _d(w), _d(V), C.c_size_t(cnt),
                   C.c_int(self.nthreads))
        return U, w, V

    def linsolve_svd(self, m, n, A, b):
        A = _in(A, m * n)
        b = _in(b, m)
        cnt = A.shape[0]
        x = np.empty((cnt, n))
        self._call("ppxo_linsolve_svd", C.c_int(m), C.c_int(n), _d(A), _d(b), _d(x), C.c_size_t(cnt),
                   C.c_int(self.nthreads))
        return x

    def pinv(self, m, n, A):
        A = _in(A, m * n)
        out = np.empty((A.shape[0], m * n))
        self._call("ppxo_pinv", C.c_int(m), C.c_int(n), _d(A), _d(out), C.c_size_t(A.shape[0]), C.c_int(self.nthreads))
        return out

    def norm2(self, m, n, A):
        A = _in(A, m * n)
        out = np.empty(A.shape[0])
        self._call("ppxo_norm2", C.c_int(m), C.c_int(n), _d(A), _d(out), C.c_size_t(A.shape[0]), C.c_int(self.nthreads))
        return out

    def eig_sym(self, n, S, sorted=True, vectors=True):
        S = _in(S, n * n)
        cnt = S.shape[0]
        d = np.empty((cnt, n))
        z = np.empty((cnt, n * n)) if vectors else None
        self._call("ppxo_eig_sym", C.c_int(n), _d(S), _d(d), _d(z), C.c_int(1 if sorted else 0), C.c_size_t(cnt),
                   C.c_int(self.nthreads))
        return d, z


def available(which):
    return os.path.exists(PATHS[which])
