mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench_n$N.json') if l.startswith('{')][-1])
e=d['e2e']; print('value %.3e e2e %.1f M/s devices %s all %.1f per_rank %.1f probe %s'%(d['value'],e['value']/1e6,e.get('devices_used'),e.get('all_devices',{}).get('value',0)/1e6,e.get('per_rank',{}).get('value',0)/1e6,e.get('result_probe')))
for t in d['configs']: print('%-22s %8.2f G/s  e2e %8.1f M/s'%(t['workload'],t['value']/1e9,(t.get('e2e') or {}).get('value',0)/1e6))
PY
