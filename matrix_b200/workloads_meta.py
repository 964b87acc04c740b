"""Torch-free description of the BASELINE workloads (names and sizes) shared by bench.py's GPU and CPU legs."""

HEADLINE = "ur5_fk_jacobian"

# cpu_sample: instances per CPU pass of the cpu_baseline / --impl reference legs, sized for ~1 s per pass on a
# 16-core host (the CPU throughput does not depend on the batch size)
META = {
    "ur5_fk_jacobian": {"cpu_sample": 1 << 22},
    "ur5_fk_jacobian_aos": {"cpu_sample": 1 << 22},
    "se3_exp_log": {"cpu_sample": 1 << 24},
    "linsolve_qr_4x3": {"cpu_sample": 1 << 25},
    "svd_6x6": {"cpu_sample": 1 << 22},
    "eig_sym_4": {"cpu_sample": 1 << 23},
    "ik_newton_step": {"cpu_sample": 1 << 21},
    "ik_newton_step_svd": {"cpu_sample": 1 << 21},
    "inverse_space_codo": {"cpu_sample": 1 << 19},
}
