"""Latency of small device-ABI calls of the persistent kernels (launch + stream sync): tools/device_latency.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from matrix_b200 import batch as mb, synth  # noqa: E402

kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
for n in (32, 4096):
    q = torch.from_numpy(synth.joint_angles(n)).cuda()
    Tt = kin.forwardSpace(q + 0.03)
    lo, hi = -np.pi * np.ones(6), np.pi * np.ones(6)
    for name, fn in (("ik_solve", lambda: kin.inverseSpaceNewton(Tt, q, 20)),
                     ("ik_codo", lambda: kin.inverseSpace(Tt, q, lo, hi)),
                     ("ik_newton_step", lambda: kin.ik_newton_step(q, Tt))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            fn()
            torch.cuda.synchronize()
        print(f"{name:16s} n = {n:5d}: {(time.perf_counter() - t0) / 300 * 1e6:8.1f} us per call (launch + sync)")
