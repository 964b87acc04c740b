#!/bin/bash
# tools/ab_ops.sh "op op ..." variant.so...  -- like ab_variants.sh for rows of matrix_b200/ops_table.py (tools/quick_op.py)
OPS=$1; shift
cp matrix_b200/libppx_cuda.so /tmp/libppx_cuda.orig.so
for v in "$@"; do
  cp build/variants/$v.so matrix_b200/libppx_cuda.so
  echo "== $v"
  python tools/quick_op.py $OPS
done
cp /tmp/libppx_cuda.orig.so matrix_b200/libppx_cuda.so
