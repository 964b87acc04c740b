# ncu --set full of the two cfg-2 kernels under their present workload names (headline = records)
mkdir -p gpurun_out
B="python bench.py --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
for wl in ur5_fk_jacobian ur5_fk_jacobian_planes; do
  $B --workload $wl > gpurun_out/plain_r2ab_$wl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 3 -c 1 -o gpurun_out/prof_r2ab_$wl $B --workload $wl > gpurun_out/ncuf_r2ab_$wl.log 2>&1
  echo $wl rc=$?
  ncu -i gpurun_out/prof_r2ab_$wl.ncu-rep --page raw --csv > gpurun_out/prof_r2ab_$wl.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r2ab_$wl.ncu-rep --page details > gpurun_out/prof_r2ab_$wl.details.txt 2>/dev/null
  ncu -i gpurun_out/prof_r2ab_$wl.ncu-rep --page source --csv > gpurun_out/prof_r2ab_$wl.source.csv 2>/dev/null
  [ $wl = ur5_fk_jacobian ] || rm -f gpurun_out/prof_r2ab_$wl.ncu-rep
done
ls -la gpurun_out | grep r2ab | cut -c25-100
