mkdir -p gpurun_out
{
bash tools/ab_variants.sh ik_newton_step 0 base ik2 ik4
for v in base ik2 ik4; do cp build/variants/$v.so matrix_b200/libppx_cuda.so; echo "== $v"; python tools/quick_op.py ik_solve_newton_loop | tail -1; done
} > gpurun_out/r2y_ab.log 2>&1
grep -v "^+" gpurun_out/r2y_ab.log | tail -14
