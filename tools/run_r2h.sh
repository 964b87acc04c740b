mkdir -p gpurun_out
bash tools/profile_all.sh r2 2>&1 | tail -12
timeout 600 python tests/parity_report.py > gpurun_out/r2_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2_parity.log
ls -la gpurun_out | head -40
