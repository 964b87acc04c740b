set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err; echo "bench2 rc=$?"
timeout 300 python tools/pcie_diag.py --devices all > gpurun_out/r2a_diag_n2.json 2> gpurun_out/r2a_diag_n2.err; echo "diag rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2>&1; echo "ref rc=$?"
tail -c 600 gpurun_out/r2a_bench_n1.err gpurun_out/r2a_bench_n2.err gpurun_out/r2a_diag_n2.err
