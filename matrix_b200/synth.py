"""Counter-based synthetic inputs.

Element j of a stream is  u01(j) = (splitmix64(seed + (j+1)*GOLDEN) >> 11) * 2^-53,
so any host thread, any GPU shard and the CPU baseline can regenerate any slice
[j0, j1) of a workload independently and bit-identically.  The same mixer is
implemented on the device by ppx_cuda_fill_uniform (csrc/util.cu); tests pin the
two against each other.

Workloads are returned as dense reference records ("AoS": one column-major
record per row, exactly std::vector<ppx::MatrixS<M,N>>::data()).
"""
import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB

SEEDS = {
    "se3_exp_log": 0x5EED0001,
    "ur5_fk_jacobian": 0x5EED0002,
    "linsolve_qr_4x3": 0x5EED0003,
    "svd_6x6": 0x5EED0004,
    "eig_sym_4": 0x5EED0014,
    "ik_newton_step": 0x5EED0005,
}

# UR5 fixture of the reference's tests (test/lie_test.cpp:187-196); (omega; v) per joint.
UR5_SCREWS = np.array([
    [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],
    [0.0, 1.0, 0.0, -0.089, 0.0, 0.0],
    [0.0, 1.0, 0.0, -0.089, 0.0, 0.425],
    [0.0, 1.0, 0.0, -0.089, 0.0, 0.817],
    [0.0, 0.0, -1.0, -0.109, 0.817, 0.0],
    [0.0, 1.0, 0.0, 0.006, 0.0, 0.817],
])
# column-major SE3 record
UR5_HOME = np.array([-1.0, 0.0, 0.0, 0.0,
                     0.0, 0.0, 1.0, 0.0,
                     0.0, 1.0, 0.0, 0.0,
                     0.817, 0.191, -0.006, 1.0])


def u01(seed, start, count):
    """count uniform doubles in [0,1) for stream elements start .. start+count-1."""
    with np.errstate(over="ignore"):
        j = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + j * np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed, start, count, lo, hi):
    """Same arithmetic as the device kernel: lo + (hi - lo) * u01 (one mul, one add, no FMA)."""
    return lo + (hi - lo) * u01(seed, start, count)


def uniform_records(seed, first, n, width, lo, hi):
    """n records of `width` doubles for instances first .. first+n-1."""
    return uniform(seed, first * width, n * width, lo, hi).reshape(n, width)


def se3_twists(n, first=0, seed=SEEDS["se3_exp_log"], theta_lo=0.05, theta_hi=np.pi - 0.05):
    """cfg 1: xi = (theta * axis, v), axis uniform on S^2, theta ~ U(theta_lo, theta_hi), v ~ U(-1,1)^3."""
    r = uniform_records(seed, first, n, 6, 0.0, 1.0)
    z = 2.0 * r[:, 0] - 1.0
    ph = 2.0 * np.pi * r[:, 1]
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    theta = theta_lo + (theta_hi - theta_lo) * r[:, 2]
    xi = np.empty((n, 6))
    xi[:, 0] = theta * s * np.cos(ph)
    xi[:, 1] = theta * s * np.sin(ph)
    xi[:, 2] = theta * z
    xi[:, 3:] = 2.0 * r[:, 3:] - 1.0
    return xi


def joint_angles(n, first=0, seed=SEEDS["ur5_fk_jacobian"], njoints=6):
    """cfg 2 / 5: q ~ U(-pi, pi)^njoints."""
    return uniform_records(seed, first, n, njoints, -np.pi, np.pi)


def lstsq_4x3(n, first=0, seed=SEEDS["linsolve_qr_4x3"]):
    """cfg 3: A ~ U(-1,1)^{4x3} (12 doubles, column-major), b ~ U(-1,1)^4."""
    r = uniform_records(seed, first, n, 16, -1.0, 1.0)
    return np.ascontiguousarray(r[:, :12]), np.ascontiguousarray(r[:, 12:])


def dense(n, rows, cols, first=0, seed=SEEDS["svd_6x6"]):
    """cfg 4a: A ~ U(-1,1)^{rows x cols}."""
    return uniform_records(seed, first, n, rows * cols, -1.0, 1.0)


def symmetric(n, dim, first=0, seed=SEEDS["eig_sym_4"]):
    """cfg 4b: symmetric, upper triangle ~ U(-1,1) mirrored into the lower one."""
    a = uniform_records(seed, first, n, dim * dim, -1.0, 1.0).reshape(n, dim, dim)  # [i, col, row]
    iu = np.triu_indices(dim, 1)
    a[:, iu[0], iu[1]] = a[:, iu[1], iu[0]]
    return a.reshape(n, dim * dim)


def ik_offsets(n, first=0, seed=SEEDS["ik_newton_step"] ^ 0xA5A5, njoints=6, amp=0.05):
    """cfg 5: delta ~ U(-amp, amp)^njoints; T_target = FK(q + delta) (mirrors demo/main.cpp:374)."""
    return uniform_records(seed, first, n, njoints, -amp, amp)
