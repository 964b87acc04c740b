#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "duration_ms": ("gpu__time_duration.sum", 1.0),
    "dram_read": ("dram__bytes_read.sum", 1.0),
    "dram_write": ("dram__bytes_write.sum", 1.0),
    "dram_pct": ("dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "fp64_pipe_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "regs": ("launch__registers_per_thread", 1.0),
    "inst_executed": ("smsp__inst_executed.sum", 1.0),
    "dfma_per_cycle": ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", 1.0),
    "dadd_per_cycle": ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", 1.0),
    "dmul_per_cycle": ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", 1.0),
    "smsp_cycles": ("smsp__cycles_elapsed.avg", 1.0),
    "local_loads": ("sass__inst_executed_local_loads", 1.0),
    "local_stores": ("sass__inst_executed_local_stores", 1.0),
    "sm_cycles": ("sm__cycles_elapsed.avg", 1.0),
    "sm_mhz": ("sm__cycles_elapsed.avg.per_second", 1.0),
    "l1tex_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "lts_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "grid": ("launch__grid_size", 1.0),
    "block": ("launch__block_size", 1.0),
    "eligible_warps": ("smsp__warps_eligible.avg.per_cycle_active", 1.0),
    "pipe_fma_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    "pipe_alu_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    "pipe_xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3,
        "Ghz": 1e3, "Mhz": 1.0, "cycle/nsecond": 1e3, "cycle/usecond": 1.0}


def raw_csv(path):
    """raw counter page of a report: either an .ncu-rep (read through ncu -i) or an already exported .raw.csv"""
    if path.endswith(".csv"):
        return open(path).read()
    return subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def summarise(path):
    out = raw_csv(path)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        s = {"kernel": rec.get("Kernel Name", "")}
        for k, (name, _) in KEYS.items():
            if name in rec and rec[name] != "":
                try:
                    v = float(rec[name].replace(",", ""))
                except ValueError:
                    continue
                s[k] = v * UNIT.get(u[name], 1.0)
        if "dfma_per_cycle" in s and "smsp_cycles" in s:
            # thread-level counts = (sum over SMSPs of inst/cycle) x elapsed cycles
            for op in ("dfma", "dadd", "dmul"):
                s[op] = s.get(op + "_per_cycle", 0.0) * s["smsp_cycles"]
            s["fp64_flops"] = 2 * s["dfma"] + s["dadd"] + s["dmul"]
            if "duration_ms" in s:
                s["fp64_tflops"] = s["fp64_flops"] / (s["duration_ms"] * 1e-3) / 1e12
        if "dram_read" in s:
            s["dram_bytes"] = s["dram_read"] + s.get("dram_write", 0)
            if "duration_ms" in s:
                s["dram_gbs"] = s["dram_bytes"] / (s["duration_ms"] * 1e-3) / 1e9
        res.append(s)
    return res


def stall_table(path, top=8):
    """per-warp-state sampling percentages for the (single) kernel in the report"""
    out = raw_csv(path)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, vals = rows[0], rows[2]
    st = {}
    for h, v in zip(hdr, vals):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
            except ValueError:
                pass
    return sorted(st.items(), key=lambda kv: -kv[1])[:top]


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    allres = {}
    for p in args:
        for s in summarise(p):
            allres[p] = s
            print(f"== {p}\n   {s['kernel'][:110]}")
            for k, v in s.items():
                if k != "kernel":
                    print(f"   {k:18s} {v:,.4g}")
            print("   stalls (warps per issue):", ", ".join(f"{k}={v:.2f}" for k, v in stall_table(p)))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(allres, f, indent=1)
