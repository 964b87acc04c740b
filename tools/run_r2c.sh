set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
{
for wl in svd_6x6 ik_newton_step_svd eig_sym_4; do bash tools/ab_variants.sh $wl 0 base noprecond svd_minb3 ik_noprecond; done
for wl in ik_newton_step; do bash tools/ab_variants.sh $wl 0 base ik_nocold ik_minb2 ik_minb4; done
for wl in ur5_fk_jacobian_planes ur5_fk_jacobian; do bash tools/ab_variants.sh $wl 0 base fk_minb3 fk_minb5; done
} > gpurun_out/r2c_ab.log 2>&1
cat gpurun_out/r2c_ab.log | grep -v "^+" | tail -60
