"""The BASELINE.json configurations as runnable workloads (bench.py and the full-size tests use these).

Each workload knows how to build its synthetic inputs on the device (from the counter-based generator of
matrix_b200.synth, so any shard can be regenerated anywhere), how to run one hot-path step on device-resident
SoA data, how to run the same step through the host ABI on host AoS buffers, and how many bytes an instance
moves at minimum (the dense reference record sizes).
"""
import numpy as np
import torch

from . import batch as mb
from . import synth
from .workloads_meta import HEADLINE, META  # noqa: F401


def _device_records(n, width, seed, first, lo, hi, device):
    """(n, width) AoS tensor, row i = instance first+i: same values as synth.uniform_records"""
    t = torch.empty((n, width), dtype=torch.float64, device=device)
    mb.fill_uniform(t, seed, first * width, lo, hi)
    return t


class Workload:
    name = ""
    kernel = ""  # the kernel one step() launches
    layout = "soa"  # device layout of the batch: component planes, or "aos" = arrays of the reference's records
    n_default = 0
    bytes_in = 0      # algorithmic bytes per instance
    bytes_out = 0
    bound = "hbm"

    def __init__(self, n, device, first=0):
        self.n, self.device, self.first = n, device, first

    @classmethod
    def bytes_per_instance(cls):
        return cls.bytes_in + cls.bytes_out


class Ur5FkJacobian(Workload):
    """config 2: UR5 product-of-exponentials forward kinematics + space Jacobian"""
    name = "ur5_fk_jacobian"
    kernel = "batch_kernel_pf<SOA, OpFkJacobian<6, true, true>>"
    n_default = 1 << 24
    bytes_in, bytes_out = 48, 128 + 288

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        q = _device_records(self.n, 6, synth.SEEDS[self.name], self.first, -np.pi, np.pi, self.device)
        self.q = mb.to_soa(q)
        del q
        self.T = torch.empty((16, self.n), dtype=torch.float64, device=self.device)
        self.Js = torch.empty((36, self.n), dtype=torch.float64, device=self.device)

    def step(self):
        self.kin.fk_jacobian(self.q, layout=mb.SOA, out_T=self.T, out_Js=self.Js)

    # host ABI
    def host_setup(self, n):
        from . import host
        self.hkin = host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        self.hq = host.pinned_empty((n, 6))
        self.hq[:] = synth.joint_angles(n, first=self.first)
        self.hT = host.pinned_empty((n, 16))
        self.hJ = host.pinned_empty((n, 36))
        return self.hq.nbytes, self.hT.nbytes + self.hJ.nbytes

    def host_step(self):
        self.hkin.fk_jacobian(self.hq, out_T=self.hT, out_Js=self.hJ)
        return float(self.hT[-1, 12])



class Ur5FkJacobianAoS(Ur5FkJacobian):
    """config 2 on the drop-in layout: device-resident arrays of dense reference records (std::vector<MatrixS>)"""
    name = "ur5_fk_jacobian_aos"
    kernel = "batch_kernel_pf<AOS, OpFkJacobian<6, true, true>>"
    layout = "aos"

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        self.q = _device_records(self.n, 6, synth.SEEDS["ur5_fk_jacobian"], self.first, -np.pi, np.pi, self.device)
        self.T = torch.empty((self.n, 16), dtype=torch.float64, device=self.device)
        self.Js = torch.empty((self.n, 36), dtype=torch.float64, device=self.device)

    def step(self):
        self.kin.fk_jacobian(self.q, layout=mb.AOS, out_T=self.T, out_Js=self.Js)


class Se3ExpLog(Workload):
    """config 1: se3::exp -> SE3::log round trip, fused"""
    name = "se3_exp_log"
    kernel = "batch_kernel_pf<SOA, OpSe3ExpLog>"
    n_default = 1 << 26  # the 1M-instance case of BASELINE fits in L2; 64M is the HBM-resident size
    bytes_in, bytes_out = 48, 48

    def setup(self):
        r = _device_records(self.n, 6, synth.SEEDS[self.name], self.first, 0.0, 1.0, self.device)
        z = 2.0 * r[:, 0] - 1.0
        ph = 2.0 * np.pi * r[:, 1]
        s = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
        th = 0.05 + (np.pi - 0.1) * r[:, 2]
        xi = torch.empty((6, self.n), dtype=torch.float64, device=self.device)
        xi[0], xi[1], xi[2] = th * s * torch.cos(ph), th * s * torch.sin(ph), th * z
        xi[3:] = (2.0 * r[:, 3:] - 1.0).t()
        self.xi = xi
        del r
        self.out = torch.empty_like(xi)

    def step(self):
        mb.se3_exp_log(self.xi, layout=mb.SOA, out=self.out)



class LinsolveQR43(Workload):
    """config 3: least-squares linsolve<QR> on 4x3"""
    name = "linsolve_qr_4x3"
    kernel = "batch_kernel<SOA, OpLinsolveQR<4, 3>>"
    n_default = 1 << 26
    bytes_in, bytes_out = 96 + 32, 24 + 1

    def setup(self):
        r = _device_records(self.n, 16, synth.SEEDS[self.name], self.first, -1.0, 1.0, self.device)
        soa = mb.to_soa(r)
        del r
        self.A, self.b = soa[:12], soa[12:]
        self.x = torch.empty((3, self.n), dtype=torch.float64, device=self.device)
        self.status = torch.empty(self.n, dtype=torch.int8, device=self.device)
        self._keep = soa

    def step(self):
        import ctypes as C
        from ._lib import check, lib
        check(lib.ppx_cuda_batch_linsolve_qr(4, 3, C.c_void_p(self.A.data_ptr()), C.c_void_p(self.b.data_ptr()),
                                             C.c_void_p(self.x.data_ptr()), C.c_void_p(self.status.data_ptr()),
                                             self.n, mb.SOA, self.n, mb._stream()))



class Svd66(Workload):
    """config 4a: SVD<6,6> (U, w, V)"""
    name = "svd_6x6"
    kernel = "batch_kernel<SOA, OpSVD<6, 6>>"
    n_default = 1 << 24
    bytes_in, bytes_out = 288, 288 + 48 + 288
    bound = "fp64"

    def setup(self):
        r = _device_records(self.n, 36, synth.SEEDS[self.name], self.first, -1.0, 1.0, self.device)
        self.A = mb.to_soa(r)
        del r
        self.U = torch.empty_like(self.A)
        self.V = torch.empty_like(self.A)
        self.w = torch.empty((6, self.n), dtype=torch.float64, device=self.device)

    def step(self):
        import ctypes as C
        from ._lib import check, lib
        check(lib.ppx_cuda_batch_svd(6, 6, C.c_void_p(self.A.data_ptr()), C.c_void_p(self.U.data_ptr()),
                                     C.c_void_p(self.w.data_ptr()), C.c_void_p(self.V.data_ptr()), self.n, mb.SOA,
                                     self.n, mb._stream()))



class EigSym4(Workload):
    """config 4b: EigenValue<4>(S, sorted=true) with eigenvectors"""
    name = "eig_sym_4"
    kernel = "batch_kernel<SOA, OpEigSym<4, true>>"
    n_default = 1 << 24
    bytes_in, bytes_out = 128, 32 + 128
    bound = "fp64"

    def setup(self):
        r = _device_records(self.n, 16, synth.SEEDS[self.name], self.first, -1.0, 1.0, self.device)
        a = r.view(self.n, 4, 4)
        iu = torch.triu_indices(4, 4, 1)
        a[:, iu[0], iu[1]] = a[:, iu[1], iu[0]]
        self.S = mb.to_soa(r)
        del r, a
        self.d = torch.empty((4, self.n), dtype=torch.float64, device=self.device)
        self.z = torch.empty_like(self.S)

    def step(self):
        import ctypes as C
        from ._lib import check, lib
        check(lib.ppx_cuda_batch_eig_sym(4, C.c_void_p(self.S.data_ptr()), C.c_void_p(self.d.data_ptr()),
                                         C.c_void_p(self.z.data_ptr()), 1, self.n, mb.SOA, self.n, mb._stream()))



class IkNewtonStep(Workload):
    """config 5: UR5 IK Newton step (FK + Jacobian + log + adjoint + linsolve<SVD>: LU fast path, Jacobi SVD fallback)"""
    name = "ik_newton_step"
    kernel = "batch_kernel_pf<SOA, OpIkStep<6>>"
    n_default = 1 << 25  # 256M over 8 GPUs
    bytes_in, bytes_out = 48 + 128, 48
    bound = "fp64"

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        seed = synth.SEEDS[getattr(self, "seed_name", self.name)]
        q = _device_records(self.n, 6, seed, self.first, -np.pi, np.pi, self.device)
        d = _device_records(self.n, 6, seed ^ 0xA5A5, self.first, -0.05, 0.05, self.device)
        self.q = mb.to_soa(q)
        qd = mb.to_soa(q + d)
        del q, d
        self.Tt = self.kin.forwardSpace(qd, layout=mb.SOA)  # T_target = FK(q + delta), demo/main.cpp:374
        del qd
        self.qo = torch.empty_like(self.q)

    def step(self):
        self.kin.ik_newton_step(self.q, self.Tt, layout=mb.SOA, out=self.qo)


class IkNewtonStepSvd(IkNewtonStep):
    """config 5 with the pseudo-inverse solve forced through the Jacobi SVD on every instance (the fp64-bound variant)"""
    name = "ik_newton_step_svd"
    kernel = "batch_kernel_pf<SOA, OpIkStep<6>> (allow_lu = false)"
    seed_name = "ik_newton_step"

    def step(self):
        from . import _lib
        prev = _lib.set_square_solve_lu(False)
        try:
            self.kin.ik_newton_step(self.q, self.Tt, layout=mb.SOA, out=self.qo)
        finally:
            _lib.set_square_solve_lu(prev)



class IkCodo(Workload):
    """kinematics<6>::inverseSpace, the trust-region IK the reference's demo and benchmark time
    (demo/main.cpp:371-377, benchmark/benchmark.cpp:167-189): start 0.05 rad off the solution on every joint"""
    name = "inverse_space_codo"
    kernel = "ik_codo_kernel<AOS>"
    layout = "aos"
    n_default = 1 << 24  # as config 2; the ~0.3 % of targets on which CoDo stalls for ITMAX = 300 trips leave a ~10 ms tail per launch
    bytes_in, bytes_out = 48 + 128, 48 + 8 + 1
    bound = "fp64 / latency (data-dependent iteration counts)"

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        q = _device_records(self.n, 6, synth.SEEDS["ik_newton_step"], self.first, -np.pi, np.pi, self.device) * 0.95
        self.Tt = self.kin.forwardSpace(q.contiguous())
        self.q0 = (q - 0.05).contiguous()
        self.lo, self.hi = -np.pi * np.ones(6), np.pi * np.ones(6)

    def step(self):
        self.out = self.kin.inverseSpace(self.Tt, self.q0, self.lo, self.hi)


WORKLOADS = {w.name: w for w in (Ur5FkJacobian, Ur5FkJacobianAoS, Se3ExpLog, LinsolveQR43, Svd66, EigSym4, IkNewtonStep, IkNewtonStepSvd, IkCodo)}
