"""ctypes binding of libppx_cuda.so (include/ppx_cuda.h).

The library is the product: there is no Python or CPU implementation behind these calls.  If the
shared object is missing (not built, or not shipped to the GPU box) importing this module raises --
build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C matrix_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libppx_cuda.so")


class PpxCudaError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"libppx_cuda: {what} (code {code})")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is not built: matrix_b200 has no fallback path. Run `make -C matrix_b200/csrc` "
        "(nvcc, sm_100a) or __graft_entry__.build() first.")

lib = C.CDLL(LIB_PATH)

_vp, _sz, _i, _u64, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_double

# name -> argtypes; every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    "ppx_cuda_device_count": [],
    "ppx_cuda_set_device": [_i],
    "ppx_cuda_malloc": [C.POINTER(_vp), _sz],
    "ppx_cuda_free": [_vp],
    "ppx_cuda_malloc_host": [C.POINTER(_vp), _sz],
    "ppx_cuda_free_host": [_vp],
    "ppx_cuda_memcpy_h2d": [_vp, _vp, _sz, _vp],
    "ppx_cuda_memcpy_d2h": [_vp, _vp, _sz, _vp],
    "ppx_cuda_stream_synchronize": [_vp],
    "ppx_cuda_aos_to_soa": [_vp, _vp, _i, _sz, _sz, _vp],
    "ppx_cuda_soa_to_aos": [_vp, _vp, _i, _sz, _sz, _vp],
    "ppx_cuda_fill_uniform": [_vp, _sz, _u64, _u64, _d, _d, _vp],
    "ppx_cuda_fp64_peak_probe": [C.POINTER(_d), _vp],
    "ppx_cuda_probe_write": [_vp, _i, _sz, _sz, _i, _vp],
    "ppx_cuda_batch_matmul": [_i, _i, _i, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_so3_exp": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_SO3_log": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_se3_exp": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_SE3_log": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_se3_exp_log": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_SE3_mul": [_vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_SE3_inv": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_SE3_Adt": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_se3_adt": [_vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_so3_jac": [_i, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_poe_fk_jacobian": [_vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_ik_newton_step": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_ik_solve": [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_ik_codo": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_linsolve_qr": [_i, _i, _vp, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_svd": [_i, _i, _vp, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_linsolve_svd": [_i, _i, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_eig_sym": [_i, _vp, _vp, _vp, _i, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_linsolve_lu": [_i, _vp, _vp, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_inverse": [_i, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_det": [_i, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_pinv": [_i, _i, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_batch_norm2": [_i, _i, _vp, _vp, _sz, _i, _sz, _vp],
    "ppx_cuda_host_matmul": [_i, _i, _i, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_so3_exp": [_vp, _vp, _sz],
    "ppx_cuda_host_SO3_log": [_vp, _vp, _sz],
    "ppx_cuda_host_se3_exp": [_vp, _vp, _sz],
    "ppx_cuda_host_SE3_log": [_vp, _vp, _sz],
    "ppx_cuda_host_se3_exp_log": [_vp, _vp, _sz],
    "ppx_cuda_host_SE3_mul": [_vp, _vp, _vp, _sz],
    "ppx_cuda_host_SE3_inv": [_vp, _vp, _sz],
    "ppx_cuda_host_SE3_Adt": [_vp, _vp, _sz],
    "ppx_cuda_host_se3_adt": [_vp, _vp, _sz],
    "ppx_cuda_host_so3_jac": [_i, _vp, _vp, _sz],
    "ppx_cuda_host_poe_fk_jacobian": [_vp, _vp, _i, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_ik_newton_step": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_ik_solve": [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_ik_codo": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_linsolve_qr": [_i, _i, _vp, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_svd": [_i, _i, _vp, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_linsolve_svd": [_i, _i, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_eig_sym": [_i, _vp, _vp, _vp, _i, _sz],
    "ppx_cuda_host_linsolve_lu": [_i, _vp, _vp, _vp, _vp, _sz],
    "ppx_cuda_host_inverse": [_i, _vp, _vp, _sz],
    "ppx_cuda_host_det": [_i, _vp, _vp, _sz],
    "ppx_cuda_host_pinv": [_i, _i, _vp, _vp, _sz],
    "ppx_cuda_host_norm2": [_i, _i, _vp, _vp, _sz],
    "ppx_cuda_host_trim": [],
    "ppx_cuda_host_set_devices": [C.POINTER(C.c_int), _i],
    "ppx_cuda_host_get_devices": [C.POINTER(C.c_int), _i],
    "ppx_cuda_host_copy_probe": [_vp, _sz, _i, _d, C.POINTER(_d)],
    "ppx_cuda_host_tune_devices": [C.POINTER(C.c_int), _i, _vp, _sz, _i, _d, C.POINTER(_d)],
}
_RESTYPES = {
    "ppx_cuda_version": (C.c_char_p, []),
    "ppx_cuda_error_string": (C.c_char_p, [_i]),
    "ppx_cuda_launch_count": (C.c_uint64, []),
    "ppx_cuda_set_square_solve_lu": (C.c_int, [_i]),
}

EXPORTS = sorted(list(_SIGNATURES) + list(_RESTYPES))

for _name, _args in _SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = header and library out of sync: fail at import
    _f.argtypes = _args
    _f.restype = C.c_int
for _name, (_res, _args) in _RESTYPES.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = _res


def check(code):
    if code != 0:
        raise PpxCudaError(code, lib.ppx_cuda_error_string(code).decode())


def version():
    return lib.ppx_cuda_version().decode()


def launch_count():
    return int(lib.ppx_cuda_launch_count())


def set_square_solve_lu(enable):
    """LU fast path of square linsolve<SVD> / the IK Newton step on (default) or off; -> previous setting"""
    return bool(lib.ppx_cuda_set_square_solve_lu(1 if enable else 0))
