mkdir -p gpurun_out
timeout 600 python tests/parity_report.py > gpurun_out/r2w_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r2w_parity.log
