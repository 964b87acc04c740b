#!/bin/bash
# tools/ab_variants.sh WORKLOAD N variant.so...  -- swaps each prebuilt library variant (build/variants/*.so) over
# matrix_b200/libppx_cuda.so in the gpurun scratch copy and prints the kernel-only line of bench.py for it
WL=$1; N=$2; shift 2
cp matrix_b200/libppx_cuda.so /tmp/libppx_cuda.orig.so
for v in "$@"; do
  cp build/variants/$v.so matrix_b200/libppx_cuda.so
  echo "== $v"
  python bench.py --workload $WL $( [ "$N" != "0" ] && echo --n $N ) --no-others --no-cpu --no-e2e --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(f\"{d['config']['workload']:18s} {d['value']/1e6:10.2f} M inst/s  kernel {r['kernel_ms']:8.3f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']} {d['clocks'].get('power_w_max')}\")
    else: print(line)
"
done
cp /tmp/libppx_cuda.orig.so matrix_b200/libppx_cuda.so
