mkdir -p gpurun_out
L=gpurun_out/r2_poly_nc.log
: > $L
for round in 1 2; do
for v in "" _q8 _q6 _q5 _q4a _q4b; do
  K5_VARIANT_NOCHECK=$([ $round = 2 ] && echo 1 || echo 0) K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 200 python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1" 2>&1 | grep -E "attn S|parity" >> $L
done
done
cat $L
echo "== nabla selection" | tee gpurun_out/r2_nabla_sel.log
for P in 0.9 0.5 0.0; do K5_NABLA_P=$P timeout 200 python tests/gpu_bench_nabla.py 2>&1 | tail -2 | tee -a gpurun_out/r2_nabla_sel.log; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "DeprecationWarning\|warnings.warn" | tail -15 > gpurun_out/r2_pytest5.log
cat gpurun_out/r2_pytest5.log
