"""Kernel-only timing of named rows of matrix_b200/ops_table.py (the --all-ops workloads) without the CPU legs:
    python tools/quick_op.py ik_solve_newton_loop inverse_6x6 ...
"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from matrix_b200 import ops_table  # noqa: E402

from matrix_b200 import batch as mb  # noqa: E402

dev = torch.device("cuda", 0)
argv = sys.argv[1:]
layout = mb.SOA
if argv and argv[0] in ("--aos", "--soa"):
    layout = mb.AOS if argv[0] == "--aos" else mb.SOA
    argv = argv[1:]
want = set(argv)
for op in ops_table.make_ops(layout):
    if want and op.name not in want:
        continue
    op.setup(dev)
    for _ in range(3):
        op.step()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(5):
        op.step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    med = statistics.median(ms)
    print(f"{op.name:24s} n={op.n:9d}  {med:8.3f} ms (min {min(ms):.3f} max {max(ms):.3f})  "
          f"{op.n / med / 1e6:9.3f} G inst/s  {op.n * op.bytes / med / 1e6:7.0f} GB/s")
    op.step = None
    torch.cuda.empty_cache()
