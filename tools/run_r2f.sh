mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ik or pinv or mixed or bench or smoke or persistent" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
{
bash tools/ab_variants.sh ik_newton_step 0 base ik4 ik2 iknop
bash tools/ab_variants.sh ik_newton_step_svd 0 base iknop
} > gpurun_out/r2f_ab.log 2>&1
grep -v "^+" gpurun_out/r2f_ab.log | tail -20
