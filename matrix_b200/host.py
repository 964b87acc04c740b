"""Host-buffer batch API: numpy arrays of dense reference records in, numpy arrays out.

This is the drop-in call for code that holds std::vector<ppx::MatrixS<M,N>>-style host data: each
function is one ppx_cuda_host_* call, which chunks the batch and overlaps H2D, kernel and D2H on three
streams.  Use pinned_empty() for the buffers to get asynchronous copies at full PCIe rate.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


class _PinnedBlock:
    """One ppx_cuda_malloc_host block; freed when the last numpy view of it is collected."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib.ppx_cuda_malloc_host(C.byref(p), max(int(nbytes), 1)))
        self.ptr = p.value

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                lib.ppx_cuda_free_host(C.c_void_p(ptr))
            except Exception:  # interpreter shutdown
                pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory from ppx_cuda_malloc_host (DMA-able by every device of the
    process).  The block is returned to the driver when the array and all its views are gone."""
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    block = _PinnedBlock(nbytes)
    buf = (C.c_char * max(nbytes, 1)).from_address(block.ptr)
    buf._ppx_block = block  # the ctypes buffer is the numpy array's base: it keeps the block alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def set_devices(devices):
    """Devices the host-buffer calls spread a batch over: a list of ordinals, "all", or None / [] for the calling
    thread's current device only (ppx_cuda_host_set_devices)."""
    if devices is None or (not isinstance(devices, str) and len(devices) == 0):
        check(lib.ppx_cuda_host_set_devices(None, 0))
    elif isinstance(devices, str):
        if devices != "all":
            raise ValueError('devices must be a list of ordinals, "all" or None')
        check(lib.ppx_cuda_host_set_devices(None, -1))
    else:
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        check(lib.ppx_cuda_host_set_devices(arr, len(devices)))


def get_devices():
    arr = (C.c_int * 64)()
    n = lib.ppx_cuda_host_get_devices(arr, 64)
    if n < 0:
        check(n)
    return [arr[i] for i in range(n)]


def copy_probe(host_array, direction, seconds=1.0):
    """GB/s each configured device sustains while all of them copy through `host_array` (pinned) at once;
    direction "h2d", "d2h" or "both" (ppx_cuda_host_copy_probe)."""
    n = max(len(get_devices()), 1)
    out = (C.c_double * n)()
    d = {"h2d": 0, "d2h": 1, "both": 2}[direction]
    check(lib.ppx_cuda_host_copy_probe(C.c_void_p(host_array.ctypes.data), host_array.nbytes, d, float(seconds), out))
    return [out[i] for i in range(n)]


def tune_devices(candidates, host_array, direction, seconds=0.15):
    """ppx_cuda_host_tune_devices: pick the subset of `candidates` (None: all visible devices) whose host links carry
    the most together and make it the device list -> (devices, aggregate GB/s)"""
    d = {"h2d": 0, "d2h": 1, "both": 2}[direction]
    agg = C.c_double(0.0)
    if candidates is None:
        arr, cnt = None, 0
    else:
        arr, cnt = (C.c_int * len(candidates))(*[int(x) for x in candidates]), len(candidates)
    check(lib.ppx_cuda_host_tune_devices(arr, cnt, C.c_void_p(host_array.ctypes.data), host_array.nbytes, d,
                                         float(seconds), C.byref(agg)))
    return get_devices(), agg.value


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, width):
    a = np.ascontiguousarray(a, dtype=np.float64)
    a = a.reshape(-1, width)
    return a


def _same_rows(n, **arrays):
    for name, a in arrays.items():
        if a.shape[0] != n:
            raise ValueError(f"{name} holds {a.shape[0]} records, expected {n}")


def _outbuf(out, n, width, dtype=np.float64):
    if out is None:
        return np.empty((n, width) if width else (n,), dtype=dtype)
    if not (isinstance(out, np.ndarray) and out.flags.c_contiguous and out.flags.writeable and out.dtype == dtype
            and out.size == n * max(width, 1)):
        raise ValueError(f"out must be a writable C-contiguous {np.dtype(dtype).name} array of {n} x {max(width, 1)} "
                         "elements")
    return out


def _unary(fn, x, win, wout, out):
    x = _arr(x, win)
    y = _outbuf(out, x.shape[0], wout)
    check(fn(_p(x), _p(y), x.shape[0]))
    return y


def matmul(m, n, l, A, B, out=None):
    """MatrixS<m,n>::operator* over host records (include/matrixs.hpp:596-640)"""
    A, B = _arr(A, m * n), _arr(B, n * l)
    _same_rows(A.shape[0], B=B)
    y = _outbuf(out, A.shape[0], m * l)
    check(lib.ppx_cuda_host_matmul(m, n, l, _p(A), _p(B), _p(y), A.shape[0]))
    return y


def so3_exp(w, out=None):
    return _unary(lib.ppx_cuda_host_so3_exp, w, 3, 9, out)


def SO3_log(R, out=None):
    return _unary(lib.ppx_cuda_host_SO3_log, R, 9, 3, out)


def se3_exp(xi, out=None):
    return _unary(lib.ppx_cuda_host_se3_exp, xi, 6, 16, out)


def SE3_log(T, out=None):
    return _unary(lib.ppx_cuda_host_SE3_log, T, 16, 6, out)


def se3_exp_log(xi, out=None):
    return _unary(lib.ppx_cuda_host_se3_exp_log, xi, 6, 6, out)


def SE3_inv(T, out=None):
    return _unary(lib.ppx_cuda_host_SE3_inv, T, 16, 16, out)


def SE3_Adt(T, out=None):
    return _unary(lib.ppx_cuda_host_SE3_Adt, T, 16, 36, out)


def se3_adt(xi, out=None):
    return _unary(lib.ppx_cuda_host_se3_adt, xi, 6, 36, out)


def SE3_mul(A, B, out=None):
    A, B = _arr(A, 16), _arr(B, 16)
    _same_rows(A.shape[0], B=B)
    y = _outbuf(out, A.shape[0], 16)
    check(lib.ppx_cuda_host_SE3_mul(_p(A), _p(B), _p(y), A.shape[0]))
    return y


def so3_jac(which, w, out=None):
    k = {"ljac": 0, "ljacinv": 1, "rjac": 2, "rjacinv": 3}[which] if isinstance(which, str) else int(which)
    w = _arr(w, 3)
    y = _outbuf(out, w.shape[0], 9)
    check(lib.ppx_cuda_host_so3_jac(k, _p(w), _p(y), w.shape[0]))
    return y


class Kinematics:
    """demo/robotics.hpp's kinematics<N> on host arrays."""

    def __init__(self, screws, home):
        self.screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        self.home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        self.njoints = self.screws.shape[0]

    def fk_jacobian(self, q, want_T=True, want_Js=True, out_T=None, out_Js=None):
        q = _arr(q, self.njoints)
        n = q.shape[0]
        T = _outbuf(out_T, n, 16) if want_T else None
        Js = _outbuf(out_Js, n, 6 * self.njoints) if want_Js else None
        check(lib.ppx_cuda_host_poe_fk_jacobian(_p(self.screws), _p(self.home), self.njoints, _p(q), _p(T), _p(Js), n))
        return T, Js

    def forwardSpace(self, q):
        return self.fk_jacobian(q, want_Js=False)[0]

    def jacobiSpace(self, q):
        return self.fk_jacobian(q, want_T=False)[1]

    def ik_newton_step(self, q, Ttarget, want_Vs=False, out=None):
        q, Tt = _arr(q, self.njoints), _arr(Ttarget, 16)
        n = q.shape[0]
        _same_rows(n, Ttarget=Tt)
        qo = _outbuf(out, n, self.njoints)
        Vs = np.empty((n, 6)) if want_Vs else None
        check(lib.ppx_cuda_host_ik_newton_step(_p(self.screws), _p(self.home), self.njoints, _p(q), _p(Tt), _p(qo),
                                               _p(Vs), n))
        return (qo, Vs) if want_Vs else qo

    def inverseSpace(self, Ttarget, init, lo=None, hi=None):
        """demo/robotics.hpp:126-152 (CoDo trust region) -> (q, status, fnorm)"""
        nj = self.njoints
        big = float(np.finfo(np.float32).max)
        lo = np.ascontiguousarray(np.full(nj, -big) if lo is None else lo, dtype=np.float64).reshape(nj)
        hi = np.ascontiguousarray(np.full(nj, big) if hi is None else hi, dtype=np.float64).reshape(nj)
        q, Tt = _arr(init, nj), _arr(Ttarget, 16)
        n = q.shape[0]
        _same_rows(n, Ttarget=Tt)
        qo = np.empty((n, nj))
        status = np.empty(n, dtype=np.int8)
        fnorm = np.empty(n)
        check(lib.ppx_cuda_host_ik_codo(_p(self.screws), _p(self.home), nj, _p(lo), _p(hi), _p(Tt), _p(q), _p(qo), None,
                                        _p(fnorm), _p(status), n))
        return qo, status, fnorm

    def inverseSpaceNewton(self, Ttarget, init, max_iter=20):
        q, Tt = _arr(init, self.njoints), _arr(Ttarget, 16)
        n = q.shape[0]
        _same_rows(n, Ttarget=Tt)
        qo = np.empty((n, self.njoints))
        iters = np.empty(n, dtype=np.int32)
        status = np.empty(n, dtype=np.int8)
        check(lib.ppx_cuda_host_ik_solve(_p(self.screws), _p(self.home), self.njoints, _p(q), _p(Tt), int(max_iter),
                                         _p(qo), _p(iters), _p(status), n))
        return qo, iters, status


def linsolve(factorization, m, n, A, b):
    A, b = _arr(A, m * n), _arr(b, m)
    cnt = A.shape[0]
    _same_rows(cnt, b=b)
    x = np.empty((cnt, n))
    if factorization == "LU":
        st = np.empty(cnt, dtype=np.int8)
        check(lib.ppx_cuda_host_linsolve_lu(n, _p(A), _p(b), _p(x), _p(st), cnt))
    elif factorization == "QR":
        st = np.empty(cnt, dtype=np.int8)
        check(lib.ppx_cuda_host_linsolve_qr(m, n, _p(A), _p(b), _p(x), _p(st), cnt))
    elif factorization == "SVD":
        st = np.zeros(cnt, dtype=np.int8)
        check(lib.ppx_cuda_host_linsolve_svd(m, n, _p(A), _p(b), _p(x), cnt))
    else:
        raise ValueError("factorization must be 'LU', 'QR' or 'SVD'")
    return x, st


def svdcmp(m, n, A, want_U=True, want_V=True):
    A = _arr(A, m * n)
    cnt = A.shape[0]
    U = np.empty((cnt, m * n)) if want_U else None
    w = np.empty((cnt, n))
    V = np.empty((cnt, n * n)) if want_V else None
    check(lib.ppx_cuda_host_svd(m, n, _p(A), _p(U), _p(w), _p(V), cnt))
    return U, w, V


def eig(n, S, sorted=True, vectors=True):
    S = _arr(S, n * n)
    cnt = S.shape[0]
    d = np.empty((cnt, n))
    z = np.empty((cnt, n * n)) if vectors else None
    check(lib.ppx_cuda_host_eig_sym(n, _p(S), _p(d), _p(z), 1 if sorted else 0, cnt))
    return d, z


def inverse(n, A):
    A = _arr(A, n * n)
    out = np.empty_like(A)
    check(lib.ppx_cuda_host_inverse(n, _p(A), _p(out), A.shape[0]))
    return out


def det(n, A):
    A = _arr(A, n * n)
    out = np.empty(A.shape[0])
    check(lib.ppx_cuda_host_det(n, _p(A), _p(out), A.shape[0]))
    return out


def pinv(m, n, A):
    A = _arr(A, m * n)
    out = np.empty((A.shape[0], m * n))
    check(lib.ppx_cuda_host_pinv(m, n, _p(A), _p(out), A.shape[0]))
    return out


def norm2(m, n, A):
    A = _arr(A, m * n)
    out = np.empty(A.shape[0])
    check(lib.ppx_cuda_host_norm2(m, n, _p(A), _p(out), A.shape[0]))
    return out
