mkdir -p gpurun_out
L=gpurun_out/r2_ab_w8_w16.log
: > $L
for v in "" _w8 "" _w8; do
  K5_VARIANT_NOCHECK=1 K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 100 python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1" 2>&1 | grep "attn S" >> $L
done
for v in "" _w8; do
  echo "== sparse lib$v" >> $L
  K5_VARIANT_BOUND=1 K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 200 python tests/gpu_bench_sparse.py >> $L 2>&1
done
echo "== sparse general kernel" >> $L
K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5.so timeout 200 python tests/gpu_bench_sparse.py >> $L 2>&1
cat $L
