"""Device-side batch API: the host-side mirror of the reference's call surface for the hot path.

Names follow the reference (include/liegroup.hpp, include/linalg.hpp, demo/robotics.hpp): so3_exp,
SO3_log, se3_exp, SE3_log, SE3_mul, SE3_inv (SE3::I), SE3_Adt, se3_adt, linsolve(factorization=...),
svdcmp, eig, Kinematics.forwardSpace / jacobiSpace / inverseSpace.  Every function evaluates the
reference routine for a whole batch with one kernel launch of libppx_cuda.so on torch's current
stream.  torch is used for device memory and streams only.

Batches are float64 CUDA tensors in one of two layouts:
  AOS  shape (n, W): one dense column-major reference record per row (std::vector<MatrixS<M,N>>)
  SOA  shape (W, ld): component-major, ld >= n instances per component (the fast layout)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

AOS, SOA = 0, 1
LU, QR, SVD = "LU", "QR", "SVD"
NORMAL, CONVERGED, DIVERGED, SINGULAR = 0, 1, 2, 3


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _in(t, width, layout, n=None):
    """-> (pointer, n, ld) after validating a batch tensor."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise TypeError("expected a contiguous float64 CUDA tensor")
    if t.dim() == 1 and layout == AOS:
        t = t.view(-1, width)
    if t.dim() != 2:
        raise ValueError("expected a 2-D batch tensor")
    if layout == AOS:
        if t.shape[1] != width:
            raise ValueError(f"AOS batch must have shape (n, {width}), got {tuple(t.shape)}")
        rows, ld = t.shape[0], 0
    elif layout == SOA:
        if t.shape[0] != width:
            raise ValueError(f"SOA batch must have shape ({width}, ld), got {tuple(t.shape)}")
        rows, ld = t.shape[1], t.shape[1]
    else:
        raise ValueError("layout must be AOS or SOA")
    n = rows if n is None else int(n)
    if not 0 <= n <= rows:  # every operand and every out= tensor of a call passes through here with the call's n
        raise ValueError(f"batch tensor of shape {tuple(t.shape)} holds {rows} instances, the call asks for {n}")
    return C.c_void_p(t.data_ptr()), n, ld


def _out(like, width, layout, n, ld, out=None):
    if out is not None:
        p, _, old = _in(out, width, layout, n)
        if layout == SOA and old != ld:
            raise ValueError("all SOA arrays of one call must share ld")
        return out, p
    shape = (n, width) if layout == AOS else (width, ld)
    t = torch.empty(shape, dtype=torch.float64, device=like.device)
    return t, C.c_void_p(t.data_ptr())


def _unary(fn, x, win, wout, layout, n, out):
    px, n, ld = _in(x, win, layout, n)
    y, py = _out(x, wout, layout, n, ld, out)
    with torch.cuda.device(x.device):
        check(fn(px, py, n, layout, ld, _stream()))
    return y


# ------------------------------------------------------------------------------ MatrixS::operator*
def matmul(m, n, l, A, B, layout=AOS, count=None, out=None):
    """MatrixS<m,n>::operator*(MatrixS<n,l>), include/matrixs.hpp:596-640: (., m*n), (., n*l) -> (., m*l);
    k-j-i accumulation without FMA contraction (bit-identical to the reference's strict build)"""
    pa, count, ld = _in(A, m * n, layout, count)
    pb, _, ldb = _in(B, n * l, layout, count)
    if ldb != ld:
        raise ValueError("all SOA arrays of one call must share ld")
    Cm, pc = _out(A, m * l, layout, count, ld, out)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_matmul(m, n, l, pa, pb, pc, count, layout, ld, _stream()))
    return Cm


# ------------------------------------------------------------------------------ Lie group / algebra
def so3_exp(w, layout=AOS, n=None, out=None):
    """so3::exp, include/liegroup.hpp:116-133: (.,3) -> (.,9)"""
    return _unary(lib.ppx_cuda_batch_so3_exp, w, 3, 9, layout, n, out)


def SO3_log(R, layout=AOS, n=None, out=None):
    """SO3::log, include/liegroup.hpp:60-92: (.,9) -> (.,3)"""
    return _unary(lib.ppx_cuda_batch_SO3_log, R, 9, 3, layout, n, out)


def se3_exp(xi, layout=AOS, n=None, out=None):
    """se3::exp, include/liegroup.hpp:253-269: (.,6) -> (.,16)"""
    return _unary(lib.ppx_cuda_batch_se3_exp, xi, 6, 16, layout, n, out)


def SE3_log(T, layout=AOS, n=None, out=None):
    """SE3::log, include/liegroup.hpp:231-250: (.,16) -> (.,6)"""
    return _unary(lib.ppx_cuda_batch_SE3_log, T, 16, 6, layout, n, out)


def se3_exp_log(xi, layout=AOS, n=None, out=None):
    """fused xi.exp().log() (BASELINE config 1): (.,6) -> (.,6)"""
    return _unary(lib.ppx_cuda_batch_se3_exp_log, xi, 6, 6, layout, n, out)


def SE3_inv(T, layout=AOS, n=None, out=None):
    """SE3::I, include/liegroup.hpp:227-230"""
    return _unary(lib.ppx_cuda_batch_SE3_inv, T, 16, 16, layout, n, out)


def SE3_Adt(T, layout=AOS, n=None, out=None):
    """SE3::Adt, include/liegroup.hpp:217-226: (.,16) -> (.,36)"""
    return _unary(lib.ppx_cuda_batch_SE3_Adt, T, 16, 36, layout, n, out)


def se3_adt(xi, layout=AOS, n=None, out=None):
    """se3::adt, include/liegroup.hpp:271-280: (.,6) -> (.,36)"""
    return _unary(lib.ppx_cuda_batch_se3_adt, xi, 6, 36, layout, n, out)


def SE3_mul(A, B, layout=AOS, n=None, out=None):
    """SE3::operator*, include/liegroup.hpp:204-207"""
    pa, n, ld = _in(A, 16, layout, n)
    pb, _, ldb = _in(B, 16, layout, n)
    if ldb != ld:
        raise ValueError("all SOA arrays of one call must share ld")
    y, py = _out(A, 16, layout, n, ld, out)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_SE3_mul(pa, pb, py, n, layout, ld, _stream()))
    return y


_JAC = {"ljac": 0, "ljacinv": 1, "rjac": 2, "rjacinv": 3}


def so3_jac(which, w, layout=AOS, n=None, out=None):
    """so3 ljac / ljacinv / rjac / rjacinv, include/liegroup.hpp:142-184: (.,3) -> (.,9)"""
    k = _JAC[which] if isinstance(which, str) else int(which)
    pw, n, ld = _in(w, 3, layout, n)
    y, py = _out(w, 9, layout, n, ld, out)
    with torch.cuda.device(w.device):
        check(lib.ppx_cuda_batch_so3_jac(k, pw, py, n, layout, ld, _stream()))
    return y


# ------------------------------------------------------------------------------ kinematics
class Kinematics:
    """Mirror of demo/robotics.hpp's kinematics<N>: a serial chain given by its space screw axes
    (JointList()[i].screw, rows (omega; v)) and the home pose of the last joint (JointList().back().pose)."""

    def __init__(self, screws, home):
        self.screws = np.ascontiguousarray(screws, dtype=np.float64).reshape(-1, 6)
        self.home = np.ascontiguousarray(home, dtype=np.float64).reshape(16)
        self.njoints = self.screws.shape[0]
        self._ps = self.screws.ctypes.data_as(C.c_void_p)
        self._ph = self.home.ctypes.data_as(C.c_void_p)

    def fk_jacobian(self, q, layout=AOS, n=None, want_T=True, want_Js=True, out_T=None, out_Js=None):
        """forwardSpace(q) and jacobiSpace(q) in one launch (demo/robotics.hpp:83-106)."""
        nj = self.njoints
        pq, n, ld = _in(q, nj, layout, n)
        T = pT = Js = pJ = None
        if want_T:
            T, pT = _out(q, 16, layout, n, ld, out_T)
        if want_Js:
            Js, pJ = _out(q, 6 * nj, layout, n, ld, out_Js)
        with torch.cuda.device(q.device):
            check(lib.ppx_cuda_batch_poe_fk_jacobian(self._ps, self._ph, nj, pq, pT, pJ, n, layout, ld, _stream()))
        return T, Js

    def forwardSpace(self, q, layout=AOS, n=None):
        return self.fk_jacobian(q, layout, n, want_Js=False)[0]

    def jacobiSpace(self, q, layout=AOS, n=None):
        return self.fk_jacobian(q, layout, n, want_T=False)[1]

    def ik_newton_step(self, q, Ttarget, layout=AOS, n=None, want_Vs=False, out=None):
        """one step of test/lie_test.cpp:104-110: q + linsolve<SVD>(Js(q), Ad(T) log(T^-1 Ttarget)).x"""
        nj = self.njoints
        pq, n, ld = _in(q, nj, layout, n)
        pt, _, ldt = _in(Ttarget, 16, layout, n)
        if ldt != ld:
            raise ValueError("all SOA arrays of one call must share ld")
        qo, pqo = _out(q, nj, layout, n, ld, out)
        Vs = pV = None
        if want_Vs:
            Vs, pV = _out(q, 6, layout, n, ld)
        with torch.cuda.device(q.device):
            check(lib.ppx_cuda_batch_ik_newton_step(self._ps, self._ph, nj, pq, pt, pqo, pV, n, layout, ld, _stream()))
        return (qo, Vs) if want_Vs else qo

    def inverseSpace(self, Ttarget, init, lo=None, hi=None, layout=AOS, n=None):
        """kinematics<6>::inverseSpace(pose, init), demo/robotics.hpp:126-152: bounded trust-region dogleg (CoDo) with
        the joint ranges [lo, hi] (default: unbounded, +-MAX_SP like joint::range) -> (q, status int8[n], fnorm[n])"""
        nj = self.njoints
        big = float(np.finfo(np.float32).max)
        lo = np.ascontiguousarray(np.full(nj, -big) if lo is None else lo, dtype=np.float64).reshape(nj)
        hi = np.ascontiguousarray(np.full(nj, big) if hi is None else hi, dtype=np.float64).reshape(nj)
        pq, n, ld = _in(init, nj, layout, n)
        pt, _, ldt = _in(Ttarget, 16, layout, n)
        if ldt != ld:
            raise ValueError("all SOA arrays of one call must share ld")
        qo, pqo = _out(init, nj, layout, n, ld)
        status = torch.empty(n, dtype=torch.int8, device=init.device)
        fnorm = torch.empty(n, dtype=torch.float64, device=init.device)
        with torch.cuda.device(init.device):
            check(lib.ppx_cuda_batch_ik_codo(self._ps, self._ph, nj, lo.ctypes.data_as(C.c_void_p),
                                             hi.ctypes.data_as(C.c_void_p), pt, pq, pqo, None,
                                             C.c_void_p(fnorm.data_ptr()), C.c_void_p(status.data_ptr()), n, layout, ld,
                                             _stream()))
        return qo, status, fnorm

    def inverseSpaceNewton(self, Ttarget, init, max_iter=20, layout=AOS, n=None):
        """the Newton-SVD loop of test/lie_test.cpp:96-113 -> (q, iterations int32[n], status int8[n])"""
        nj = self.njoints
        pq, n, ld = _in(init, nj, layout, n)
        pt, _, ldt = _in(Ttarget, 16, layout, n)
        if ldt != ld:
            raise ValueError("all SOA arrays of one call must share ld")
        qo, pqo = _out(init, nj, layout, n, ld)
        iters = torch.empty(n, dtype=torch.int32, device=init.device)
        status = torch.empty(n, dtype=torch.int8, device=init.device)
        with torch.cuda.device(init.device):
            check(lib.ppx_cuda_batch_ik_solve(self._ps, self._ph, nj, pq, pt, int(max_iter), pqo,
                                              C.c_void_p(iters.data_ptr()), C.c_void_p(status.data_ptr()), n, layout,
                                              ld, _stream()))
        return qo, iters, status


# ------------------------------------------------------------------------------ dense factorisations
def linsolve(factorization, m, n, A, b, layout=AOS, count=None):
    """linsolve<Factorization::QR|SVD>(A, b), include/linalg.hpp:1189-1209 -> (x, status int8[count])"""
    pa, count, ld = _in(A, m * n, layout, count)
    pb, _, ldb = _in(b, m, layout, count)
    if ldb != ld:
        raise ValueError("all SOA arrays of one call must share ld")
    x, px = _out(A, n, layout, count, ld)
    with torch.cuda.device(A.device):
        if factorization == LU:
            if m != n:
                raise ValueError("LU decomposation only supports square matrix.")
            status = torch.empty(count, dtype=torch.int8, device=A.device)
            check(lib.ppx_cuda_batch_linsolve_lu(n, pa, pb, px, C.c_void_p(status.data_ptr()), count, layout, ld,
                                                 _stream()))
        elif factorization == QR:
            status = torch.empty(count, dtype=torch.int8, device=A.device)
            check(lib.ppx_cuda_batch_linsolve_qr(m, n, pa, pb, px, C.c_void_p(status.data_ptr()), count, layout, ld,
                                                 _stream()))
        elif factorization == SVD:
            # linsolve<SVD> always reports NORMAL (include/linalg.hpp:1208)
            status = torch.zeros(count, dtype=torch.int8, device=A.device)
            check(lib.ppx_cuda_batch_linsolve_svd(m, n, pa, pb, px, count, layout, ld, _stream()))
        else:
            raise ValueError("factorization must be LU, QR or SVD")
    return x, status


def svdcmp(m, n, A, layout=AOS, count=None, want_U=True, want_V=True):
    """SVD<m,n> (documented svdcmp), include/linalg.hpp:605-884 -> (U, w, V), A = U diag(w) V^T"""
    pa, count, ld = _in(A, m * n, layout, count)
    U = pU = V = pV = None
    if want_U:
        U, pU = _out(A, m * n, layout, count, ld)
    w, pw = _out(A, n, layout, count, ld)
    if want_V:
        V, pV = _out(A, n * n, layout, count, ld)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_svd(m, n, pa, pU, pw, pV, count, layout, ld, _stream()))
    return U, w, V


def eig(n, S, sorted=True, vectors=True, layout=AOS, count=None):
    """EigenValue<n>(S, sorted) (documented eig<eigensystem::Sym*>), include/linalg.hpp:967-1167 -> (d, z)"""
    ps, count, ld = _in(S, n * n, layout, count)
    d, pd = _out(S, n, layout, count, ld)
    z = pz = None
    if vectors:
        z, pz = _out(S, n * n, layout, count, ld)
    with torch.cuda.device(S.device):
        check(lib.ppx_cuda_batch_eig_sym(n, ps, pd, pz, 1 if sorted else 0, count, layout, ld, _stream()))
    return d, z


def inverse(n, A, layout=AOS, count=None):
    """MatrixS::I(), include/linalg.hpp:933-948 (all-zero result for a singular matrix)"""
    pa, count, ld = _in(A, n * n, layout, count)
    Ai, pi = _out(A, n * n, layout, count, ld)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_inverse(n, pa, pi, count, layout, ld, _stream()))
    return Ai


def det(n, A, layout=AOS, count=None):
    """MatrixS::det(), include/linalg.hpp:950-965 -> (count, 1) / (1, ld)"""
    pa, count, ld = _in(A, n * n, layout, count)
    d, pd = _out(A, 1, layout, count, ld)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_det(n, pa, pd, count, layout, ld, _stream()))
    return d


def pinv(m, n, A, layout=AOS, count=None):
    """pinv(A), include/linalg.hpp:1211-1227: (., m*n) -> (., n*m)"""
    pa, count, ld = _in(A, m * n, layout, count)
    P, pp = _out(A, m * n, layout, count, ld)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_pinv(m, n, pa, pp, count, layout, ld, _stream()))
    return P


def norm2(m, n, A, layout=AOS, count=None):
    """matrix norm2 = largest singular value, include/linalg.hpp:926-931"""
    pa, count, ld = _in(A, m * n, layout, count)
    sv, ps = _out(A, 1, layout, count, ld)
    with torch.cuda.device(A.device):
        check(lib.ppx_cuda_batch_norm2(m, n, pa, ps, count, layout, ld, _stream()))
    return sv


# ------------------------------------------------------------------------------ utilities
def to_soa(aos, ld=None):
    """(n, W) AOS -> (W, ld) SOA on the device"""
    n, width = aos.shape
    ld = n if ld is None else ld
    soa = torch.empty((width, ld), dtype=torch.float64, device=aos.device)
    with torch.cuda.device(aos.device):
        check(lib.ppx_cuda_aos_to_soa(C.c_void_p(aos.data_ptr()), C.c_void_p(soa.data_ptr()), width, n, ld, _stream()))
    return soa


def to_aos(soa, n=None):
    """(W, ld) SOA -> (n, W) AOS on the device"""
    width, ld = soa.shape
    n = ld if n is None else n
    aos = torch.empty((n, width), dtype=torch.float64, device=soa.device)
    with torch.cuda.device(soa.device):
        check(lib.ppx_cuda_soa_to_aos(C.c_void_p(soa.data_ptr()), C.c_void_p(aos.data_ptr()), width, n, ld, _stream()))
    return aos


def fill_uniform(out, seed, first, lo, hi):
    """out.flatten()[j] = lo + (hi-lo)*u01(seed, first+j): the generator of matrix_b200.synth, on the device"""
    with torch.cuda.device(out.device):
        check(lib.ppx_cuda_fill_uniform(C.c_void_p(out.data_ptr()), out.numel(), seed, first, float(lo), float(hi),
                                        _stream()))
    return out


def fp64_peak_tflops(device=None):
    t = C.c_double(0.0)
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        check(lib.ppx_cuda_fp64_peak_probe(C.byref(t), _stream()))
    return t.value
