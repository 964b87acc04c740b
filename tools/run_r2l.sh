mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
timeout 200 python tools/host_latency.py > gpurun_out/r2l_host_latency.txt 2>&1; cat gpurun_out/r2l_host_latency.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2l_bench_n$N.json') if l.startswith('{')][-1])
e=d['e2e']; print('value %.3e e2e %.1f M/s devices %s all %.1f per_rank %.1f probe %s'%(d['value'],e['value']/1e6,e.get('devices_used'),e.get('all_devices',{}).get('value',0)/1e6,e.get('per_rank',{}).get('value',0)/1e6,e.get('result_probe')))
print(json.dumps(e.get('ceiling'))[:900])
PY
timeout 300 python bench.py --impl reference --gpus $N --steps 5 --warmup 1 | head -c 300
