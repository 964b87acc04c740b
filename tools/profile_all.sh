B="python bench.py --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
$B > gpurun_out/plain_fk.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fk.csv $B > gpurun_out/ncu_fk.log 2>&1
echo launches rc=$?
for wl in ur5_fk_jacobian se3_exp_log linsolve_qr_4x3 svd_6x6 eig_sym_4 ik_newton_step; do
  $B --workload $wl > gpurun_out/plain_$wl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 3 -c 1 -o gpurun_out/prof_$wl $B --workload $wl > gpurun_out/ncuf_$wl.log 2>&1
  echo $wl rc=$?
done
ls -la gpurun_out/
