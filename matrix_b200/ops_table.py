"""Every batched entry point of the C ABI as a timeable micro-workload (bench.py --all-ops).

Each entry: algorithmic bytes per instance (dense reference records in + out), a setup() that builds seeded SoA inputs
on the device, and a step() that issues exactly one kernel launch.  The headline / BASELINE workloads live in
workloads.py; this table exists so that every SURVEY section-8 row has a measured line next to its roofline.
"""
import ctypes as C

import numpy as np
import torch

from . import batch as mb
from . import synth
from ._lib import check, lib


def _rec(n, width, seed, lo, hi, device):
    t = torch.empty((n, width), dtype=torch.float64, device=device)
    mb.fill_uniform(t, seed, 0, lo, hi)
    return t


def _twists(n, device):
    r = _rec(n, 6, synth.SEEDS["se3_exp_log"], 0.0, 1.0, device)
    z = 2.0 * r[:, 0] - 1.0
    ph = 2.0 * np.pi * r[:, 1]
    s = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    th = 0.05 + (np.pi - 0.1) * r[:, 2]
    xi = torch.empty((6, n), dtype=torch.float64, device=device)
    xi[0], xi[1], xi[2] = th * s * torch.cos(ph), th * s * torch.sin(ph), th * z
    xi[3:] = (2.0 * r[:, 3:] - 1.0).t()
    return xi


def _p(t):
    return C.c_void_p(t.data_ptr())


class Op:
    def __init__(self, name, n, bytes_per_instance, setup, ref="", touched=None):
        self.name, self.n, self.bytes, self._setup, self.ref = name, n, bytes_per_instance, setup, ref
        # bytes the kernel really moves: in SoA the four constant planes (0,0,0,1) of an SE3 INPUT are never read
        self.touched = bytes_per_instance if touched is None else touched

    def setup(self, device):
        self.step = self._setup(self.n, device)


def _lie(n, dev):
    xi = _twists(n, dev)
    T = mb.se3_exp(xi, layout=mb.SOA)
    T2 = mb.se3_exp(_twists(n, dev).flip(1).contiguous(), layout=mb.SOA)
    w = xi[:3].contiguous()
    R = mb.so3_exp(w, layout=mb.SOA)
    return xi, T, T2, w, R


def make_ops(layout=mb.SOA):
    """layout = mb.SOA (component planes, the fast layout) or mb.AOS (arrays of the reference's dense records)"""
    S, M = synth.UR5_SCREWS, synth.UR5_HOME
    ops = []
    aos = layout == mb.AOS

    def lay(t):  # an SoA tensor [width, n] in the requested layout
        return mb.to_aos(t) if aos else t

    def out(w, cnt, dev):
        return torch.empty((cnt, w) if aos else (w, cnt), dtype=torch.float64, device=dev)

    def lie_op(name, pick, fn, win, wout, ref):
        def setup(n, dev):
            xi, T, T2, w, R = _lie(n, dev)
            x = lay({"xi": xi, "T": T, "w": w, "R": R}[pick])
            o = out(wout, n, dev)
            return lambda: fn(x, layout=layout, out=o)
        ops.append(Op(name, 1 << 24, 8 * (win + wout), setup, ref,
                      8 * ((12 if pick == "T" and not aos else win) + wout)))

    lie_op("so3_exp", "w", mb.so3_exp, 3, 9, "liegroup.hpp:116-133")
    lie_op("SO3_log", "R", mb.SO3_log, 9, 3, "liegroup.hpp:60-92")
    lie_op("se3_exp", "xi", mb.se3_exp, 6, 16, "liegroup.hpp:253-269")
    lie_op("SE3_log", "T", mb.SE3_log, 16, 6, "liegroup.hpp:231-250")
    lie_op("SE3_inv", "T", mb.SE3_inv, 16, 16, "liegroup.hpp:227-230")
    lie_op("SE3_Adt", "T", mb.SE3_Adt, 16, 36, "liegroup.hpp:217-226")
    lie_op("se3_adt", "xi", mb.se3_adt, 6, 36, "liegroup.hpp:271-280")

    def setup_mul(n, dev):
        xi, T, T2, w, R = _lie(n, dev)
        T, T2 = lay(T), lay(T2)
        o = torch.empty_like(T)
        return lambda: mb.SE3_mul(T, T2, layout=layout, out=o)
    ops.append(Op("SE3_mul", 1 << 24, 8 * 48, setup_mul, "liegroup.hpp:204-207", 8 * (48 if aos else 40)))

    def matmul_op(m, n_, l):  # MatrixS::operator*
        def setup(cnt, dev):
            A = _rec(cnt, m * n_, 0x301, -1.0, 1.0, dev)
            B = _rec(cnt, n_ * l, 0x302, -1.0, 1.0, dev)
            if not aos:
                A, B = mb.to_soa(A), mb.to_soa(B)
            o = out(m * l, cnt, dev)
            return lambda: mb.matmul(m, n_, l, A, B, layout=layout, out=o)
        ops.append(Op(f"matmul_{m}x{n_}x{l}", 1 << 24, 8 * (m * n_ + n_ * l + m * l), setup, "matrixs.hpp:596-640"))

    matmul_op(3, 3, 3)
    matmul_op(4, 4, 4)
    matmul_op(6, 6, 6)
    matmul_op(6, 6, 1)

    def setup_jac(n, dev):
        w = lay(_lie(n, dev)[3])
        o = out(9, n, dev)
        return lambda: mb.so3_jac("ljacinv", w, layout=layout, out=o)
    ops.append(Op("so3_ljacinv", 1 << 24, 8 * 12, setup_jac, "liegroup.hpp:157-170"))

    def dense_op(name, m, n_, count, seed, bytes_, ref, launch, out_widths, with_status=False):
        """launch(A, b, outs, status, count) issues the C-ABI call; outs are SoA tensors of the given widths"""
        def setup(cnt, dev):
            A = _rec(cnt, m * n_, seed, -1.0, 1.0, dev)
            b = _rec(cnt, m, seed + 1, -1.0, 1.0, dev)
            if not aos:
                A, b = mb.to_soa(A), mb.to_soa(b)
            outs = [out(w, cnt, dev) for w in out_widths]
            status = torch.empty(cnt, dtype=torch.int8, device=dev) if with_status else None
            return lambda: check(launch(A, b, outs, status, cnt))
        ops.append(Op(name, count, bytes_, setup, ref))

    def tail(c):
        return (c, layout, c, mb._stream())

    dense_op("linsolve_svd_6x6", 6, 6, 1 << 23, 0x201, 8 * (36 + 6 + 6), "linalg.hpp:886-913, 1202-1209",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_linsolve_svd(6, 6, _p(A), _p(b), _p(o[0]), *tail(c)), [6])
    dense_op("linsolve_lu_6x6", 6, 6, 1 << 24, 0x203, 8 * (36 + 6 + 6) + 1, "linalg.hpp:448-533, 1176-1187",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_linsolve_lu(6, _p(A), _p(b), _p(o[0]), _p(s), *tail(c)), [6], True)
    dense_op("inverse_6x6", 6, 6, 1 << 24, 0x205, 8 * 72, "linalg.hpp:933-948",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_inverse(6, _p(A), _p(o[0]), *tail(c)), [36])
    dense_op("det_6x6", 6, 6, 1 << 24, 0x207, 8 * 37, "linalg.hpp:950-965",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_det(6, _p(A), _p(o[0]), *tail(c)), [1])
    dense_op("pinv_4x3", 4, 3, 1 << 24, 0x209, 8 * 24, "linalg.hpp:1211-1227",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_pinv(4, 3, _p(A), _p(o[0]), *tail(c)), [12])
    dense_op("pinv_6x6", 6, 6, 1 << 24, 0x20F, 8 * 72, "linalg.hpp:1211-1227",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_pinv(6, 6, _p(A), _p(o[0]), *tail(c)), [36])
    dense_op("norm2_6x6", 6, 6, 1 << 23, 0x20B, 8 * 37, "linalg.hpp:926-931",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_norm2(6, 6, _p(A), _p(o[0]), *tail(c)), [1])
    dense_op("svd_4x3", 4, 3, 1 << 24, 0x20D, 8 * (12 + 12 + 3 + 9), "linalg.hpp:605-884",
             lambda A, b, o, s, c: lib.ppx_cuda_batch_svd(4, 3, _p(A), _p(o[0]), _p(o[1]), _p(o[2]), *tail(c)),
             [12, 3, 9])

    def setup_iksolve(n, dev):
        kin = mb.Kinematics(S, M)
        q = _rec(n, 6, synth.SEEDS["ik_newton_step"], -np.pi, np.pi, dev)
        d = _rec(n, 6, synth.SEEDS["ik_newton_step"] ^ 0xA5A5, -0.05, 0.05, dev)
        Tt = lay(kin.forwardSpace(mb.to_soa(q + d), layout=mb.SOA))
        q = q if aos else mb.to_soa(q)
        return lambda: kin.inverseSpaceNewton(Tt, q, 20, layout=layout)
    ops.append(Op("ik_solve_newton_loop", 1 << 22, 8 * (6 + 16 + 6) + 5, setup_iksolve, "test/lie_test.cpp:96-113"))
    return ops
