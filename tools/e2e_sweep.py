"""tools/e2e_sweep.py -- host-ABI (pinned AoS in/out) throughput of cfg 2 for pipeline settings given by env vars."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matrix_b200 import host, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
kin = host.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
q = host.pinned_empty((n, 6)); q[:] = synth.joint_angles(n)
T = host.pinned_empty((n, 16)); J = host.pinned_empty((n, 36))
kin.fk_jacobian(q, out_T=T, out_Js=J)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); kin.fk_jacobian(q, out_T=T, out_Js=J); ts.append(time.perf_counter() - t0)
b = n * 464
print(f"slots={os.environ.get('PPX_HOST_SLOTS','3')} chunk={os.environ.get('PPX_HOST_CHUNK_MIB','48')}MiB: best {min(ts)*1e3:.1f} ms  {n/min(ts)/1e6:.1f} M inst/s  {b/min(ts)/1e9:.1f} GB/s total, d2h {n*416/min(ts)/1e9:.1f} GB/s")
