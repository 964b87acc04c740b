#!/bin/bash
# quick kernel-only timing of a list of workloads: tools/quick_bench.sh svd_6x6 eig_sym_4 ...
for wl in "$@"; do
  python bench.py --workload $wl --no-others --no-cpu --no-e2e --steps 5 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(f\"{d['config']['workload']:18s} {d['value']/1e9:8.3f} G inst/s  kernel {r['kernel_ms']:8.3f} ms  hbm {r['hbm_gbs']:7.0f} GB/s ({r['hbm_frac']*100:5.1f}%)  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}\")
    else: print(line)
"
done
