"""Torch-free description of the BASELINE workloads (names, sizes, bytes) shared by bench.py's GPU and CPU legs."""

HEADLINE = "ur5_fk_jacobian"

# n: instances per GPU per step (the BASELINE size); bytes: algorithmic bytes per instance;
# cpu_sample: instances per CPU pass of the cpu_baseline leg, sized for ~1 s per pass on a 16-core host (the CPU
# throughput does not depend on the batch size); counters_as: the ncu counter record that describes the same kernel;
# layout: how the GPU arm keeps the batch in HBM -- "records" = arrays of the reference's dense MatrixS records (what the
# host face hands the kernels and what the CPU reference computes on), default component planes
META = {
    "ur5_fk_jacobian": {"n": 1 << 24, "bytes": 464, "chain": True, "cpu_sample": 1 << 22, "layout": "records"},
    "ur5_fk_jacobian_planes": {"n": 1 << 24, "bytes": 464, "chain": True, "cpu_sample": 1 << 22},
    "se3_exp_log_1M": {"n": 1 << 20, "bytes": 96, "cpu_sample": 1 << 24, "counters_as": "se3_exp_log_1M"},
    "se3_exp_log": {"n": 1 << 26, "bytes": 96, "cpu_sample": 1 << 24},
    "linsolve_qr_4x3": {"n": 1 << 26, "bytes": 153, "cpu_sample": 1 << 25},
    "svd_6x6": {"n": 1 << 24, "bytes": 912, "cpu_sample": 1 << 22},
    "eig_sym_4": {"n": 1 << 24, "bytes": 288, "cpu_sample": 1 << 23},
    "ik_newton_step": {"n": 1 << 25, "bytes": 224, "chain": True, "cpu_sample": 1 << 21},
    "ik_newton_step_svd": {"n": 1 << 25, "bytes": 224, "chain": True, "cpu_sample": 1 << 21},
    "inverse_space_codo": {"n": 1 << 24, "bytes": 233, "chain": True, "cpu_sample": 1 << 19, "layout": "records"},
}
