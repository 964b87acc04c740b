"""Latency of small host-ABI calls (what the C++ header's single-value routines cost): tools/host_latency.py"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from matrix_b200 import host, synth  # noqa: E402

for n in (1, 64, 4096, 1 << 18):
    xi = synth.se3_twists(n)
    out = host.se3_exp(xi)  # warm-up
    reps = 2000 if n <= 4096 else 50
    t0 = time.perf_counter()
    for _ in range(reps):
        out = host.se3_exp(xi)
    dt = (time.perf_counter() - t0) / reps
    print(f"se3_exp host call, n = {n:7d}: {dt * 1e6:9.1f} us per call  ({n / dt / 1e6:8.2f} M inst/s)")
kin = host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
q = synth.joint_angles(1)
kin.fk_jacobian(q)
t0 = time.perf_counter()
for _ in range(2000):
    kin.fk_jacobian(q)
print(f"fk_jacobian host call, n = 1: {(time.perf_counter() - t0) / 2000 * 1e6:9.1f} us per call")
