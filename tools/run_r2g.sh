mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench rc=$?"
python tools/design_tables.py gpurun_out/r2g_bench_n1.json
timeout 600 python bench.py --all-ops > gpurun_out/r2g_all_ops.json 2> gpurun_out/r2g_all_ops.err; echo "allops rc=$?"
timeout 600 python bench.py --all-ops --layout aos > gpurun_out/r2g_all_ops_aos.json 2>> gpurun_out/r2g_all_ops.err; echo "allops aos rc=$?"
tail -c 800 gpurun_out/r2g_bench_n1.err gpurun_out/r2g_all_ops.err
