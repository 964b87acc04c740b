"""Prints a checksum of the raw bits of svd_6x6 / pinv / IK-step outputs on seeded inputs (to compare library builds)."""
import hashlib
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from matrix_b200 import batch as mb, synth  # noqa: E402

n = 1 << 18
A = torch.from_numpy(synth.dense(n, 6, 6)).cuda()
U, w, V = mb.svdcmp(6, 6, A)
h = hashlib.sha256()
for t in (U, w, V):
    h.update(t.cpu().numpy().tobytes())
A43 = torch.from_numpy(synth.dense(n, 4, 3, seed=5)).cuda()
h.update(mb.pinv(4, 3, A43).cpu().numpy().tobytes())
kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
q = torch.from_numpy(synth.joint_angles(n)).cuda()
Tt = kin.forwardSpace(q + 0.03)
out = kin.ik_newton_step(q, Tt)
h.update((out[0] if isinstance(out, tuple) else out).cpu().numpy().tobytes())
print("bits", h.hexdigest()[:24])
