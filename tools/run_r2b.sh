set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.txt
timeout 300 python tools/pcie_diag.py --devices all --gb 1.5 --e2e-n 4194304 > gpurun_out/r2b_diag_n8.json 2> gpurun_out/r2b_diag_n8.err; echo "diag rc=$?"
for cfg in "4 16" "2 96" "6 24" "3 128"; do set -- $cfg
  PPX_HOST_SLOTS=$1 PPX_HOST_CHUNK_MIB=$2 timeout 200 python tools/pcie_diag.py --devices all --gb 0.5 --e2e-n 4194304 --skip-box --skip-solo > gpurun_out/r2b_diag_n8_s$1_c$2.json 2>> gpurun_out/r2b_diag_n8.err; echo "sweep $cfg rc=$?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err; echo "bench8 rc=$?"
timeout 200 python bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2b_ref_n8.json 2>&1; echo "ref rc=$?"
tail -c 1500 gpurun_out/r2b_bench_n8.err gpurun_out/r2b_diag_n8.err
