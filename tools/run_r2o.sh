mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "svd or pinv or norm2 or scale" > gpurun_out/r2o_pytest.log 2>&1; tail -2 gpurun_out/r2o_pytest.log
bash tools/quick_bench.sh svd_6x6 eig_sym_4 2>&1 | grep -v "^+" | tail -3
python tools/quick_op.py svd_4x3 pinv_4x3 norm2_6x6 2>&1 | tail -3
