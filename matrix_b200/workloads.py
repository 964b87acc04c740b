"""The BASELINE.json configurations as runnable workloads (bench.py and the full-size tests use these).

Each workload knows how to build its synthetic inputs on the device (from the counter-based generator of
matrix_b200.synth, so any shard can be regenerated anywhere), how to run one hot-path step on device-resident
data (one kernel launch), how to run the same step through the host ABI on pinned host arrays of reference
records (ppx_cuda_host_*: H2D + kernel + D2H per call), and how many bytes an instance moves at minimum (the
dense reference record sizes).
"""
import ctypes as C

import numpy as np
import torch

from . import batch as mb
from . import synth
from ._lib import check, lib
from .workloads_meta import HEADLINE, META  # noqa: F401


def _device_records(n, width, seed, first, lo, hi, device):
    """(n, width) AoS tensor, row i = instance first+i: same values as synth.uniform_records"""
    t = torch.empty((n, width), dtype=torch.float64, device=device)
    mb.fill_uniform(t, seed, first * width, lo, hi)
    return t


def _vp(t):
    return C.c_void_p(t.data_ptr())


def _hp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Arena:
    """One pinned host block (ppx_cuda_malloc_host) that the end-to-end legs of all workloads carve their arrays
    from, so that the bench page-locks host memory once."""

    def __init__(self, nbytes):
        from . import host
        self.buf = host.pinned_empty((int(nbytes),), np.uint8)
        self.reset()

    def reset(self):
        self.off = 0

    def take(self, shape, dtype=np.float64):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        start = (self.off + 4095) // 4096 * 4096
        if start + nbytes > self.buf.size:
            raise MemoryError("pinned arena too small")
        self.off = start + nbytes
        return self.buf[start:start + nbytes].view(dtype).reshape(shape)


class Workload:
    name = ""
    kernel = ""  # the kernel one step() launches
    layout = "soa"  # device layout of the batch: component planes, or "aos" = arrays of the reference's records
    n_default = 0
    bytes_in = 0      # algorithmic bytes per instance
    bytes_out = 0
    bound = "hbm"
    host_api = ""     # the ppx_cuda_host_* entry point of the end-to-end leg
    host_out = ()     # ((width, numpy dtype), ...) of the host-leg outputs

    def __init__(self, n, device, first=0):
        self.n, self.device, self.first = n, device, first

    @classmethod
    def bytes_per_instance(cls):
        return cls.bytes_in + cls.bytes_out

    # ---- host ABI leg: pinned host arrays of dense records in and out ---------------------------------------
    def host_inputs(self):
        """-> list of (n, width) float64 DEVICE tensors in record layout (the leg copies them to the host once)"""
        raise NotImplementedError

    def host_call(self, n, ins, outs):
        raise NotImplementedError

    def host_setup(self, n, arena):
        """inputs of the first n instances to pinned host arrays, pinned outputs; -> (h2d bytes, d2h bytes) per step"""
        srcs = self.host_inputs()
        self.h_in = []
        for t in srcs:
            a = arena.take((n, t.shape[1]))
            torch.from_numpy(a).copy_(t[:n])
            self.h_in.append(a)
        torch.cuda.synchronize()
        self.h_out = [arena.take((n, w) if w > 1 else (n,), dt) for w, dt in self.host_out]
        self.h_n = n
        return sum(a.nbytes for a in self.h_in), sum(a.nbytes for a in self.h_out)

    def host_step(self):
        """one blocking call of the host ABI; -> a value read from the result (proof the outputs arrived)"""
        self.host_call(self.h_n, self.h_in, self.h_out)
        return float(self.h_out[0].reshape(self.h_n, -1)[-1, -1])

    def host_release(self):
        self.h_in = self.h_out = None


class Ur5FkJacobian(Workload):
    """config 2, the headline: UR5 product-of-exponentials forward kinematics + space Jacobian on device-resident arrays
    of the reference's dense records (what a std::vector<MatrixS> holds) -- the layout the host face hands the kernel, so
    `value`, `roofline` and `e2e` of the bench line describe one and the same kernel"""
    name = "ur5_fk_jacobian"
    kernel = "batch_kernel_pf<AOS, OpFkJacobian<6, true, true>>"
    layout = "aos"
    n_default = 1 << 24
    bytes_in, bytes_out = 48, 128 + 288
    host_api = "ppx_cuda_host_poe_fk_jacobian"
    host_out = ((16, np.float64), (36, np.float64))

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        self.q = _device_records(self.n, 6, synth.SEEDS["ur5_fk_jacobian"], self.first, -np.pi, np.pi, self.device)
        self.T = torch.empty((self.n, 16), dtype=torch.float64, device=self.device)
        self.Js = torch.empty((self.n, 36), dtype=torch.float64, device=self.device)

    def step(self):
        self.kin.fk_jacobian(self.q, layout=mb.AOS, out_T=self.T, out_Js=self.Js)

    def free_outputs(self):
        self.T = self.Js = None

    def host_inputs(self):
        return [mb.to_aos(self.q) if self.layout == "soa" else self.q]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_poe_fk_jacobian(_hp(self.kin.screws), _hp(self.kin.home), 6, _hp(ins[0]), _hp(outs[0]),
                                                _hp(outs[1]), n))


class Ur5FkJacobianPlanes(Ur5FkJacobian):
    """config 2 on component planes (one plane per matrix element, the layout every other config is benched on)"""
    name = "ur5_fk_jacobian_planes"
    kernel = "batch_kernel_pf<SOA, OpFkJacobian<6, true, true>>"
    layout = "soa"

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        q = _device_records(self.n, 6, synth.SEEDS["ur5_fk_jacobian"], self.first, -np.pi, np.pi, self.device)
        self.q = mb.to_soa(q)
        del q
        self.T = torch.empty((16, self.n), dtype=torch.float64, device=self.device)
        self.Js = torch.empty((36, self.n), dtype=torch.float64, device=self.device)

    def step(self):
        self.kin.fk_jacobian(self.q, layout=mb.SOA, out_T=self.T, out_Js=self.Js)


def _twists_soa(n, first, device):
    r = _device_records(n, 6, synth.SEEDS["se3_exp_log"], first, 0.0, 1.0, device)
    z = 2.0 * r[:, 0] - 1.0
    ph = 2.0 * np.pi * r[:, 1]
    s = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    th = 0.05 + (np.pi - 0.1) * r[:, 2]
    xi = torch.empty((6, n), dtype=torch.float64, device=device)
    xi[0], xi[1], xi[2] = th * s * torch.cos(ph), th * s * torch.sin(ph), th * z
    xi[3:] = (2.0 * r[:, 3:] - 1.0).t()
    return xi


class Se3ExpLog(Workload):
    """config 1 at an HBM-resident size (64M twists): se3::exp -> SE3::log round trip, fused"""
    name = "se3_exp_log"
    kernel = "batch_kernel_pf<SOA, OpSe3ExpLog>"
    n_default = 1 << 26
    bytes_in, bytes_out = 48, 48
    host_api = "ppx_cuda_host_se3_exp_log"
    host_out = ((6, np.float64),)

    def setup(self):
        self.xi = _twists_soa(self.n, self.first, self.device)
        self.out = torch.empty_like(self.xi)

    def step(self):
        mb.se3_exp_log(self.xi, layout=mb.SOA, out=self.out)

    def free_outputs(self):
        self.out = None

    def host_inputs(self):
        return [mb.to_aos(self.xi)]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_se3_exp_log(_hp(ins[0]), _hp(outs[0]), n))


class Se3ExpLog1M(Se3ExpLog):
    """config 1 at its stated size, 1M twists (96 MB per launch: it would sit in the 126 MB L2), over NBUF = 8 rotating
    buffer sets (768 MB touched per rotation) so that every launch streams from and to HBM.  A launch lasts ~17 us, the
    scale of a launch from Python and of a kernel's ramp-up and tail, so the rotation is captured once in a CUDA graph
    -- the 8 independent launches forked over LANES = 4 streams, so that one launch's tail overlaps the next one's
    start -- and step() replays it: one step = NBUF launches over NBUF x 1M instances."""
    name = "se3_exp_log_1M"
    kernel = "batch_kernel_pf<SOA, OpSe3ExpLog> x8 (CUDA graph: 8 rotating 1M-twist buffer sets, 4 launches in flight)"
    n_default = 1 << 20
    NBUF = 8
    LANES = 4

    def setup(self):
        self.bufs = []
        for b in range(self.NBUF):
            xi = _twists_soa(self.n, self.first + b * self.n, self.device)
            self.bufs.append((xi, torch.empty_like(xi)))
        self.xi = self.bufs[0][0]
        for xi, out in self.bufs:  # warm-up outside the capture: occupancy queries, attribute calls
            mb.se3_exp_log(xi, layout=mb.SOA, out=out)
        torch.cuda.synchronize()
        side = [torch.cuda.Stream(device=self.device) for _ in range(self.LANES - 1)]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            cur = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(cur)
            for st in side:
                st.wait_event(fork)
            for b, (xi, out) in enumerate(self.bufs):
                lane = b % self.LANES
                with torch.cuda.stream(cur if lane == 0 else side[lane - 1]):
                    mb.se3_exp_log(xi, layout=mb.SOA, out=out)  # launched on torch's current stream
            for st in side:
                join = torch.cuda.Event()
                join.record(st)
                cur.wait_event(join)
        self.launches_per_step = self.NBUF
        self.instances_per_step = self.NBUF * self.n

    def step(self):
        self.graph.replay()

    def free_outputs(self):
        self.graph = None
        self.bufs = [(xi, None) for xi, _ in self.bufs]


class LinsolveQR43(Workload):
    """config 3: least-squares linsolve<QR> on 4x3"""
    name = "linsolve_qr_4x3"
    kernel = "batch_kernel<SOA, OpLinsolveQR<4, 3>>"
    n_default = 1 << 26
    bytes_in, bytes_out = 96 + 32, 24 + 1
    host_api = "ppx_cuda_host_linsolve_qr"
    host_out = ((3, np.float64), (1, np.int8))

    def setup(self):
        r = _device_records(self.n, 16, synth.SEEDS[self.name], self.first, -1.0, 1.0, self.device)
        soa = mb.to_soa(r)
        del r
        self.A, self.b = soa[:12], soa[12:]
        self.x = torch.empty((3, self.n), dtype=torch.float64, device=self.device)
        self.status = torch.empty(self.n, dtype=torch.int8, device=self.device)
        self._keep = soa

    def step(self):
        check(lib.ppx_cuda_batch_linsolve_qr(4, 3, _vp(self.A), _vp(self.b), _vp(self.x), _vp(self.status), self.n,
                                             mb.SOA, self.n, mb._stream()))

    def free_outputs(self):
        self.x = self.status = None

    def host_inputs(self):
        r = mb.to_aos(self._keep)
        return [r[:, :12].contiguous(), r[:, 12:].contiguous()]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_linsolve_qr(4, 3, _hp(ins[0]), _hp(ins[1]), _hp(outs[0]), _hp(outs[1]), n))


class Svd66(Workload):
    """config 4a: SVD<6,6> (U, w, V)"""
    name = "svd_6x6"
    kernel = "batch_kernel<SOA, OpSVD<6, 6>>"
    n_default = 1 << 24
    bytes_in, bytes_out = 288, 288 + 48 + 288
    bound = "fp64"
    host_api = "ppx_cuda_host_svd"
    host_out = ((36, np.float64), (6, np.float64), (36, np.float64))

    def setup(self):
        r = _device_records(self.n, 36, synth.SEEDS[self.name], self.first, -1.0, 1.0, self.device)
        self.A = mb.to_soa(r)
        del r
        self.U = torch.empty_like(self.A)
        self.V = torch.empty_like(self.A)
        self.w = torch.empty((6, self.n), dtype=torch.float64, device=self.device)

    def step(self):
        check(lib.ppx_cuda_batch_svd(6, 6, _vp(self.A), _vp(self.U), _vp(self.w), _vp(self.V), self.n, mb.SOA, self.n,
                                     mb._stream()))

    def free_outputs(self):
        self.U = self.V = self.w = None

    def host_inputs(self):
        return [mb.to_aos(self.A)]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_svd(6, 6, _hp(ins[0]), _hp(outs[0]), _hp(outs[1]), _hp(outs[2]), n))


class EigSym4(Workload):
    """config 4b: EigenValue<4>(S, sorted=true) with eigenvectors"""
    name = "eig_sym_4"
    kernel = "batch_kernel<SOA, OpEigSym<4, true>>"
    n_default = 1 << 24
    bytes_in, bytes_out = 128, 32 + 128
    bound = "fp64"
    host_api = "ppx_cuda_host_eig_sym"
    host_out = ((4, np.float64), (16, np.float64))

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
        check(lib.ppx_cuda_batch_eig_sym(4, _vp(self.S), _vp(self.d), _vp(self.z), 1, self.n, mb.SOA, self.n,
                                         mb._stream()))

    def free_outputs(self):
        self.d = self.z = None

    def host_inputs(self):
        return [mb.to_aos(self.S)]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_eig_sym(4, _hp(ins[0]), _hp(outs[0]), _hp(outs[1]), 1, n))


class IkNewtonStep(Workload):
    """config 5: UR5 IK Newton step (FK + Jacobian + log + adjoint + linsolve<SVD>: LU fast path, Jacobi SVD fallback)"""
    name = "ik_newton_step"
    kernel = "batch_kernel_pf<SOA, OpIkStep<6>>"
    n_default = 1 << 25  # 256M over 8 GPUs
    bytes_in, bytes_out = 48 + 128, 48
    bound = "fp64"
    host_api = "ppx_cuda_host_ik_newton_step"
    host_out = ((6, np.float64),)
    allow_lu = True

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        seed = synth.SEEDS["ik_newton_step"]
        q = _device_records(self.n, 6, seed, self.first, -np.pi, np.pi, self.device)
        d = _device_records(self.n, 6, seed ^ 0xA5A5, self.first, -0.05, 0.05, self.device)
        self.q = mb.to_soa(q)
        qd = mb.to_soa(q + d)
        del q, d
        self.Tt = self.kin.forwardSpace(qd, layout=mb.SOA)  # T_target = FK(q + delta), demo/main.cpp:374
        del qd
        self.qo = torch.empty_like(self.q)

    def _mode(self):
        from . import _lib
        return _lib.set_square_solve_lu(self.allow_lu)

    def step(self):
        from . import _lib
        prev = self._mode()
        try:
            self.kin.ik_newton_step(self.q, self.Tt, layout=mb.SOA, out=self.qo)
        finally:
            _lib.set_square_solve_lu(prev)

    def free_outputs(self):
        self.qo = None

    def host_inputs(self):
        return [mb.to_aos(self.q), mb.to_aos(self.Tt)]

    def host_call(self, n, ins, outs):
        from . import _lib
        prev = self._mode()
        try:
            check(lib.ppx_cuda_host_ik_newton_step(_hp(self.kin.screws), _hp(self.kin.home), 6, _hp(ins[0]), _hp(ins[1]),
                                                   _hp(outs[0]), None, n))
        finally:
            _lib.set_square_solve_lu(prev)


class IkNewtonStepSvd(IkNewtonStep):
    """config 5 with the pseudo-inverse solve forced through the Jacobi SVD on every instance (the fp64-bound variant)"""
    name = "ik_newton_step_svd"
    kernel = "batch_kernel_pf<SOA, OpIkStep<6>> (allow_lu = false)"
    allow_lu = False


class IkCodo(Workload):
    """kinematics<6>::inverseSpace, the trust-region IK the reference's demo and benchmark time
    (demo/main.cpp:371-377, benchmark/benchmark.cpp:167-189): start 0.05 rad off the solution on every joint"""
    name = "inverse_space_codo"
    kernel = "ik_codo_kernel<AOS>"
    layout = "aos"
    n_default = 1 << 24  # as config 2; the ~0.3 % of targets on which CoDo stalls for ITMAX = 300 trips leave a tail per launch
    bytes_in, bytes_out = 48 + 128, 48 + 8 + 1
    bound = "fp64 / latency (data-dependent iteration counts)"
    host_api = "ppx_cuda_host_ik_codo"
    host_out = ((6, np.float64), (1, np.float64), (1, np.int8))

    def setup(self):
        self.kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
        q = _device_records(self.n, 6, synth.SEEDS["ik_newton_step"], self.first, -np.pi, np.pi, self.device) * 0.95
        self.Tt = self.kin.forwardSpace(q.contiguous())
        self.q0 = (q - 0.05).contiguous()
        self.lo, self.hi = -np.pi * np.ones(6), np.pi * np.ones(6)

    def step(self):
        self.out = self.kin.inverseSpace(self.Tt, self.q0, self.lo, self.hi)

    def free_outputs(self):
        self.out = None

    def host_inputs(self):
        return [self.Tt, self.q0]

    def host_call(self, n, ins, outs):
        check(lib.ppx_cuda_host_ik_codo(_hp(self.kin.screws), _hp(self.kin.home), 6, _hp(self.lo), _hp(self.hi),
                                        _hp(ins[0]), _hp(ins[1]), _hp(outs[0]), None, _hp(outs[1]), _hp(outs[2]), n))


WORKLOADS = {w.name: w for w in (Ur5FkJacobian, Ur5FkJacobianPlanes, Se3ExpLog1M, Se3ExpLog, LinsolveQR43, Svd66, EigSym4,
                                 IkNewtonStep, IkNewtonStepSvd, IkCodo)}
