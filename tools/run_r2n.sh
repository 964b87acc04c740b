mkdir -p gpurun_out
bash tools/ab_variants.sh ik_newton_step 0 base ikqr ikqr4 > gpurun_out/r2n_ab.log 2>&1
grep -v "^+" gpurun_out/r2n_ab.log | tail -8
