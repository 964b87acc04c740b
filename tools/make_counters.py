#!/usr/bin/env python
"""tools/make_counters.py prof_TAG_<workload>.ncu-rep ... -> profiles/counters.json

Per-workload counters taken from one `ncu --set full` capture of the bench kernel on the seeded BASELINE workload
(deterministic inputs => deterministic counts): DRAM bytes per launch and fp64 flops per instance
(2*DFMA + DADD + DMUL, thread-level, predicated-on)."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import summarise  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matrix_b200.workloads_meta import META  # noqa: E402

SIZES = {"ur5_fk_jacobian": 1 << 24, "ur5_fk_jacobian_planes": 1 << 24, "se3_exp_log": 1 << 26, "linsolve_qr_4x3": 1 << 26,
         "svd_6x6": 1 << 24, "eig_sym_4": 1 << 24, "ik_newton_step": 1 << 25,
         "ik_newton_step_svd": 1 << 25}
out = {}
for path in sys.argv[1:]:
    m = re.search(r"prof_[^_]+_(.+?)\.(?:ncu-rep|raw\.csv)$", os.path.basename(path))
    wl = m.group(1)
    s = summarise(path)[0]
    n = SIZES[wl]
    out[wl] = {"instances": n, "kernel": s["kernel"], "report": os.path.basename(path),
               "duration_ms_under_ncu": s.get("duration_ms"),
               "dram_bytes_per_launch": s.get("dram_bytes"), "dram_gbs": s.get("dram_gbs"),
               "fp64_flops_per_instance": s.get("fp64_flops", 0) / n if s.get("fp64_flops") else None,
               "fp64_instr_per_instance": (s.get("dfma", 0) + s.get("dadd", 0) + s.get("dmul", 0)) / n,
               "warp_instr_per_instance": s.get("inst_executed", 0) * 32 / n,  # one instance per thread
               "fp64_pipe_pct": s.get("fp64_pipe_pct"), "issue_active_pct": s.get("issue_active_pct"),
               "pipe_fma_pct": s.get("pipe_fma_pct"), "pipe_alu_pct": s.get("pipe_alu_pct"),
               "pipe_xu_pct": s.get("pipe_xu_pct"), "eligible_warps_per_cycle": s.get("eligible_warps"),
               "registers": s.get("regs"), "warps_active_pct": s.get("warps_active_pct")}
json.dump(out, open(os.path.join(ROOT, "profiles", "counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
