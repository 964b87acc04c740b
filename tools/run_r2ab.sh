# round 2, last pass: the headline moved to the records kernel -- bench line, launch list of the same command, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; tail -1 gpurun_out/r2ab_smoke.log
python bench.py > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ab_bench_reference_arm.json 2>/dev/null; echo "ref rc $?"
timeout 300 python -m pytest tests -m gpu -x -q -k "bench or smoke or fk" 2>&1 | tail -2
B="python bench.py --no-others --no-cpu --no-e2e --steps 3 --warmup 3"
$B > gpurun_out/r2ab_plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2ab_launches_bench.csv $B > gpurun_out/r2ab_ncu_launches.log 2>&1; echo "launches rc $?"
