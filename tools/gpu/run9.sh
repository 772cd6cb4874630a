mkdir -p gpurun_out
L=gpurun_out/r2_poly_sweep.log
: > $L
for round in 1 2; do
for v in "" _p8 _p6 _p5 _p4 _p3; do
  K5_VARIANT_NOCHECK=$([ $round = 2 ] && echo 1 || echo 0) K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 200 python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1" 2>&1 | grep -E "attn S|parity" >> $L
done
done
cat $L
timeout 1500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_shard.py tests/test_gpu_pipeline.py tests/test_gpu_vae.py -m gpu -x -q 2>&1 | grep -v "DeprecationWarning\|warnings.warn" | tail -15 > gpurun_out/r2_pytest4.log
cat gpurun_out/r2_pytest4.log
