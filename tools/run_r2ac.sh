# 2 GPUs: the bench line with the records headline (short form) and the multi-device tests
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
  --steps 5 --warmup 3 --only ur5_fk_jacobian_planes,ik_newton_step --no-cpu > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_bench_n2.err
echo "bench rc $?"
python bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --n 1048576 | cut -c1-200
timeout 300 python -m pytest tests -m gpu -x -q -k "device or shard or two_gpu or 2gpu" 2>&1 | tail -2
