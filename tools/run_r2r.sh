mkdir -p gpurun_out
{
bash tools/ab_variants.sh ik_newton_step 0 base fixedkinds generalonly
bash tools/ab_variants.sh ur5_fk_jacobian 0 base fixedkinds generalonly
bash tools/ab_variants.sh ik_newton_step_svd 0 base fixedkinds
for v in base fixedkinds; do cp build/variants/$v.so matrix_b200/libppx_cuda.so; echo "== $v"; python tools/quick_op.py ik_solve_newton_loop; done
} > gpurun_out/r2r_ab.log 2>&1
grep -v "^+" gpurun_out/r2r_ab.log | tail -24
