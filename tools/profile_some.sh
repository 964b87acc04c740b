#!/bin/bash
# tools/profile_some.sh TAG workload...   -> gpurun_out/prof_TAG_<workload>.ncu-rep (one launch, --set full)
TAG=$1; shift
B="python bench.py --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
for wl in "$@"; do
  $B --workload $wl > gpurun_out/plain_${TAG}_$wl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_$wl $B --workload $wl > gpurun_out/ncuf_${TAG}_$wl.log 2>&1
  echo $wl rc=$?
done
