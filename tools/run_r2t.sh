mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "codo or inverse_space or persistent" > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
bash tools/quick_bench.sh inverse_space_codo 2>&1 | grep -v "^+" | tail -1
PPX_IK_PATTERNS=0 bash tools/quick_bench.sh inverse_space_codo 2>&1 | grep -v "^+" | tail -1
PPX_IK_PATTERNS=0 timeout 600 python -m pytest tests -m gpu -x -q -k "codo or inverse_space" 2>&1 | tail -2
