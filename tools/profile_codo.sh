#!/bin/bash
# tools/profile_codo.sh TAG [N]  -> ncu --set full of one ik_codo_kernel launch, exported as text/csv under gpurun_out/
TAG=$1; N=${2:-1048576}
B="python bench.py --workload inverse_space_codo --n $N --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
$B > gpurun_out/plain_${TAG}_codo.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ik_codo -s 2 -c 1 -o gpurun_out/prof_${TAG}_codo $B > gpurun_out/ncuf_${TAG}_codo.log 2>&1
echo codo rc=$?
ncu -i gpurun_out/prof_${TAG}_codo.ncu-rep --page details > gpurun_out/prof_${TAG}_codo.details.txt 2>&1
ncu -i gpurun_out/prof_${TAG}_codo.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_codo.raw.csv 2>&1
