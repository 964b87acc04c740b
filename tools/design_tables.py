"""tools/design_tables.py bench_n1.json [bench_n2.json bench_n8.json ...] -> the markdown tables of DESIGN.md section 3 / 5 / 6
from bench.py's JSON lines (so that the document quotes the measured numbers, not retyped ones)."""
import json
import sys


def load(path):
    return json.loads([ln for ln in open(path) if ln.startswith("{")][-1])


def g(x, digits=2):
    return f"{x / 1e9:.{digits}f} G" if x >= 1e9 else f"{x / 1e6:.1f} M"


def main():
    runs = [load(p) for p in sys.argv[1:]]
    d = runs[0]
    print("| workload (BASELINE cfg) | instances / launch | B/inst | kernel | inst/s | HBM GB/s (frac of peak) | fp64 TFLOP/s (frac) | bound | e2e inst/s (PCIe GB/s) | CPU ref inst/s (cores) | kernel / CPU | e2e / CPU |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for t in d["configs"]:
        if "error" in t:
            print(f"| `{t['workload']}` | error: {t['error']} |")
            continue
        r, e, c = t["roofline"], t.get("e2e") or {}, t.get("cpu_baseline") or {}
        fp = f"{r['fp64_tflops']:.1f} ({r['fp64_frac']:.2f})" if "fp64_frac" in r else "—"
        print(f"| `{t['workload']}` | {r['instances_per_launch']} | {r['algorithmic_bytes_per_instance']} | {r['kernel_ms']:.3f} ms | "
              f"{g(t['value'])} | {r['hbm_gbs']:.0f} ({r['hbm_frac']:.2f}) | {fp} | {r['bound']} | "
              f"{g(e['value']) if e else '—'} ({e.get('pcie_gbs', 0):.0f}) | {g(c['value']) if c else '—'} ({c.get('cores', '')}) | "
              f"{t.get('kernel_vs_cpu', 0):.0f}x | {t.get('e2e_vs_cpu', 0):.1f}x |")
    if len(runs) > 1:
        print()
        names = [t["workload"] for t in d["configs"] if "error" not in t]
        print("| workload | " + " | ".join(f"N={r['n_gpus']} kernel inst/s (eff.) | N={r['n_gpus']} e2e inst/s" for r in runs) + " |")
        print("|---|" + "---|---|" * len(runs))
        base = {t["workload"]: t for t in d["configs"] if "error" not in t}
        for nm in names:
            row = [f"`{nm}`"]
            for r in runs:
                t = {x["workload"]: x for x in r["configs"] if "error" not in x}.get(nm)
                if not t:
                    row += ["—", "—"]
                    continue
                eff = t["value"] / (r["n_gpus"] * base[nm]["value"])
                e = t.get("e2e") or {}
                row += [f"{g(t['value'])} ({eff:.3f})", g(e["value"]) if e else "—"]
            print("| " + " | ".join(row) + " |")


if __name__ == "__main__":
    main()
