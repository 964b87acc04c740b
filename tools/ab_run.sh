#!/bin/bash
# tools/ab_run.sh "command" variant.so...  -- runs a command once per prebuilt library variant (build/variants/*.so)
CMD=$1; shift
cp matrix_b200/libppx_cuda.so /tmp/libppx_cuda.orig.so
for v in "$@"; do
  cp build/variants/$v.so matrix_b200/libppx_cuda.so
  echo "== $v"
  bash -c "$CMD"
done
cp /tmp/libppx_cuda.orig.so matrix_b200/libppx_cuda.so
