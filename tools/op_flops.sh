#!/bin/bash
# tools/op_flops.sh -> gpurun_out/op_flops_<op>.csv: fp64 instruction counts (ncu, thread level, predicated on) of the last
# kernel launch of every row of matrix_b200/ops_table.py (the --all-ops workloads), one ncu run per row.
# tools/op_flops_merge.py turns them into fp64 flops per instance next to the all-ops timings.
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
mkdir -p gpurun_out
for op in $(python - <<'PY'
import sys
sys.path.insert(0, ".")
from matrix_b200 import ops_table, batch as mb
print(" ".join(o.name for o in ops_table.make_ops(mb.SOA)))
PY
); do
  python tools/quick_op.py $op > /dev/null 2>&1 || { echo "$op failed without ncu"; continue; }
  ncu --metrics $M --clock-control none --csv -k 'regex:batch_kernel|matmul_dyn|ik_solve_kernel' --log-file gpurun_out/op_flops_$op.csv python tools/quick_op.py $op > /dev/null 2>&1
  echo "$op rc=$?"
done
