"""tools/op_flops_merge.py all_ops.json gpurun_out/op_flops_*.csv -> the same all-ops JSON with fp64 flops per instance
(2 DFMA + DADD + DMUL of the op's last launch, ncu) and the fp64 rate each row sustains; written next to the input."""
import csv
import json
import os
import re
import sys


def last_launch(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    if not rows:
        return None
    last = max(int(r[0]) for r in rows)
    vals = {r[-3]: float(r[-1].replace(",", "")) for r in rows if int(r[0]) == last}
    get = lambda k: next((v for n, v in vals.items() if k in n), 0.0)
    return {"dfma": get("dfma"), "dadd": get("dadd"), "dmul": get("dmul"),
            "kernel": next(r[4] for r in rows if int(r[0]) == last)}


def main():
    path = sys.argv[1]
    d = json.loads([ln for ln in open(path) if ln.startswith("{")][-1])
    counts = {}
    for p in sys.argv[2:]:
        m = re.search(r"op_flops_(.+)\.csv$", os.path.basename(p))
        c = last_launch(p)
        if m and c:
            counts[m.group(1)] = c
    peak = float(sys.argv[sys.argv.index("--fp64-peak") + 1]) if "--fp64-peak" in sys.argv else 34.8
    for row in d["all_ops"]:
        c = counts.get(row["op"])
        if not c:
            continue
        flops = (2 * c["dfma"] + c["dadd"] + c["dmul"]) / row["instances"]
        row["fp64_instr_per_instance_ncu"] = (c["dfma"] + c["dadd"] + c["dmul"]) / row["instances"]
        row["fp64_flops_per_instance_ncu"] = flops
        row["fp64_tflops"] = flops * row["value"] / 1e12
        row["fp64_frac_of_measured_peak"] = row["fp64_tflops"] / peak
        row["fp64_counted_kernel"] = c["kernel"]
    d["fp64_peak_tflops"] = peak
    open(path, "w").write(json.dumps(d) + "\n")
    for row in d["all_ops"]:
        if "fp64_tflops" in row:
            print(f"{row['op']:24s} {row['fp64_instr_per_instance_ncu']:8.0f} instr {row['fp64_flops_per_instance_ncu']:8.0f} flop  "
                  f"{row['fp64_tflops']:6.2f} TF/s ({row['fp64_frac_of_measured_peak']:.2f})  hbm {row['hbm_frac_touched']:.2f}")


if __name__ == "__main__":
    args = [a for a in sys.argv]
    main()
