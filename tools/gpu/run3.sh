mkdir -p gpurun_out
L=gpurun_out/${OUT:-r2_variants.log}
: > $L
run() { K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$1.so timeout 300 python tests/gpu_attn_variants.py "$2" >> $L 2>&1; }
run "" "bounded=K5_VARIANT_BOUND:1"
run "" "bounded_stagger1100=K5_VARIANT_BOUND:1,K5_ATTN_STAGGER:1100"
run "" "bounded_stagger1600=K5_VARIANT_BOUND:1,K5_ATTN_STAGGER:1600"
run _pp "bounded_pp=K5_VARIANT_BOUND:1,K5_VARIANT_SAMPLE:1"
run "" "general="
cat $L
