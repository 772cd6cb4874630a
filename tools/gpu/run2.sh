mkdir -p gpurun_out
L=gpurun_out/${OUT:-r2_variants.log}
: > $L
for v in "$@"; do
  [ "$v" = "-" ] && v=""
  echo "=== lib$v" >> $L
  K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1,K5_VARIANT_SAMPLE:1" >> $L 2>&1
done
K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5.so python tests/gpu_attn_variants.py "general=" >> $L 2>&1
cat $L
