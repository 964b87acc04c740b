#!/bin/bash
# tools/build_variants.sh name:file.cu:"-DFLAG=.." ...  -- builds build/variants/<name>.so = the base library with one
# translation unit recompiled with extra flags (A/B experiments: tools/ab_variants.sh runs them on the GPU box)
set -e
cd "$(dirname "$0")/../matrix_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../build/variants build/var
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC -diag-suppress 177"
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; file=${rest%%:*}; flags=${rest#*:}
  (
    $NV $flags -c $file -o build/var/$name.o
    objs=""
    for f in lie matmul kinematics solvers util host; do
      if [ "$f.cu" == "$file" ]; then objs="$objs build/var/$name.o"; else objs="$objs build/$f.o"; fi
    done
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/$name.so $objs -cudart static
    echo "built $name ($file $flags)"
  ) &
done
wait
