#!/bin/bash
# tools/profile_all.sh TAG : (1) launch list of the default bench command, (2) one `ncu --set full` capture of the
# bench kernel of every workload.  Each ncu command runs only after the same command line exited 0 without ncu
# (B200_PROFILING.md).  Outputs land in gpurun_out/; summaries are produced on the build box with
# tools/ncu_summary.py / tools/make_counters.py and committed under profiles/.
TAG=${1:-r1}
B="python bench.py --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
$B > gpurun_out/plain_${TAG}_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_${TAG}_launches.log 2>&1
echo launches rc=$?
for wl in ur5_fk_jacobian ur5_fk_jacobian_planes se3_exp_log linsolve_qr_4x3 svd_6x6 eig_sym_4 ik_newton_step ik_newton_step_svd; do
  $B --workload $wl > gpurun_out/plain_${TAG}_$wl.log 2>&1 && ncu --set full --clock-control none -k regex:batch_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_$wl $B --workload $wl > gpurun_out/ncuf_${TAG}_$wl.log 2>&1
  echo $wl rc=$?
  # the raw counter page and the details page are what profiles/ quotes; keep the (10 MB) report itself only for
  # the headline kernel and the SVD so that gpurun_out stays under the 64 MiB return limit
  ncu -i gpurun_out/prof_${TAG}_$wl.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$wl.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$wl.ncu-rep --page details > gpurun_out/prof_${TAG}_$wl.details.txt 2>/dev/null
  case $wl in ur5_fk_jacobian|svd_6x6) ;; *) rm -f gpurun_out/prof_${TAG}_$wl.ncu-rep ;; esac
done
