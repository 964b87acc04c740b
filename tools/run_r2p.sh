mkdir -p gpurun_out
{
bash tools/ab_variants.sh ur5_fk_jacobian_planes 0 base fk_b256_m2 fk_b512_m1 fk_b64_m6 fk_b256_m1
bash tools/ab_variants.sh ur5_fk_jacobian 0 base fk_b256_m2 fk_b512_m1 fk_b64_m6 fk_b256_m1
} > gpurun_out/r2p_ab.log 2>&1
grep -v "^+" gpurun_out/r2p_ab.log | tail -24
