mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench_n$N.json') if l.startswith('{')][-1])
e=d['e2e']; print('value %.3e e2e %.1f M/s devices %s all %.1f per_rank %.1f probe %s'%(d['value'],e['value']/1e6,e.get('devices_used'),e.get('all_devices',{}).get('value',0)/1e6,e.get('per_rank',{}).get('value',0)/1e6,e.get('result_probe')))
print(json.dumps(e.get('ceiling'))[:900])
PY
nvidia-smi topo -m | head -8
