mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log; tail -12 gpurun_out/r2s_pytest.log
bash tools/quick_bench.sh ik_newton_step ik_newton_step_svd ur5_fk_jacobian 2>&1 | grep -v "^+" | tail -4
python tools/quick_op.py ik_solve_newton_loop | tail -1
PPX_IK_PATTERNS=0 bash tools/quick_bench.sh ik_newton_step 2>&1 | grep -v "^+" | tail -1
