mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
{
bash tools/ab_variants.sh svd_6x6 0 base v_b128 v_nop v_sw2
bash tools/ab_variants.sh ik_newton_step 0 base k_nop
bash tools/ab_variants.sh ik_newton_step_svd 0 base k_nop
bash tools/ab_variants.sh ur5_fk_jacobian_planes 0 base
bash tools/ab_variants.sh ur5_fk_jacobian 0 base
} > gpurun_out/r2e_ab.log 2>&1
grep -v "^+" gpurun_out/r2e_ab.log | tail -40
timeout 600 python bench.py --all-ops > gpurun_out/r2e_all_ops.json 2> gpurun_out/r2e_all_ops.err; echo "allops rc=$?"
cp build/variants/v_nop.so matrix_b200/libppx_cuda.so
timeout 600 python bench.py --all-ops > gpurun_out/r2e_all_ops_noprecond.json 2>> gpurun_out/r2e_all_ops.err; echo "allops nop rc=$?"
