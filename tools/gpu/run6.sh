# A/B of issuer wait forms (isolated dense attention at S = 47 616), two rounds to see the box noise; then the GPU tests
mkdir -p gpurun_out
L=gpurun_out/r2_iss_wait.log
: > $L
for round in 1 2; do
for v in "" _h300 _h1500 _s100 _s400; do
  K5_VARIANT_NOCHECK=1 K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 100 python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1" 2>&1 | grep "attn S" >> $L
done
done
cat $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest2.log
cat gpurun_out/r2_pytest2.log
