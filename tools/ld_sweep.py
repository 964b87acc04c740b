"""tools/ld_sweep.py -- kernel time of the SoA kinematics kernel vs the component stride ld (DRAM bank/row locality)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matrix_b200 import batch as mb, synth
n = 1 << 24
kin = mb.Kinematics(synth.UR5_SCREWS, synth.UR5_HOME)
for pad in (0, 32, 96, 1056, 4128, 16416, 65568, 1 << 20):
    ld = n + pad
    q = torch.empty((6, ld), dtype=torch.float64, device="cuda")
    mb.fill_uniform(q, 1, 0, -np.pi, np.pi)
    T = torch.empty((16, ld), dtype=torch.float64, device="cuda")
    J = torch.empty((36, ld), dtype=torch.float64, device="cuda")
    for _ in range(3):
        kin.fk_jacobian(q, layout=mb.SOA, n=n, out_T=T, out_Js=J)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    torch.cuda.synchronize(); ev[0].record()
    for i in range(10):
        kin.fk_jacobian(q, layout=mb.SOA, n=n, out_T=T, out_Js=J); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
    print(f"ld = n + {pad:8d}: {ms:.3f} ms  {n*464/ms/1e6:.0f} GB/s")
    del q, T, J
