mkdir -p gpurun_out
{
bash tools/ab_variants.sh svd_6x6 0 s0_noprecond s1_base s2_b256 s3_b256_d4 s4_b256_d8 s5_b256_noprecond s6_b256_d4_sw2
bash tools/ab_variants.sh ik_newton_step 0 k_fkminb3_nocold
bash tools/ab_variants.sh ik_newton_step_svd 0 k_fkminb3_nocold
bash tools/ab_variants.sh ur5_fk_jacobian 0 k_fkminb3_nocold
} > gpurun_out/r2d_ab.log 2>&1
grep -v "^+" gpurun_out/r2d_ab.log | tail -40
