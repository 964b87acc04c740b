mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2k_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"
python tools/design_tables.py gpurun_out/r2k_bench_n1.json | cut -c1-260
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_reference_arm.json 2>&1; echo "ref rc=$?"
bash tools/op_flops.sh 2>&1 | tail -30
