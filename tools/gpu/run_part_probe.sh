# A/B of the s_full probe position in the PART attention kernels (isolated timing, one GPU): libk5.so (pair 56) and variants.
mkdir -p gpurun_out
L=gpurun_out/r2_attention_part_probe.log
: > $L
for v in "" _p57 _p58 _p60; do
  lib=""
  [ -n "$v" ] && lib=$PWD/kandinsky-5_b200/libk5$v.so
  K5_LIB_PATH=$lib timeout 60 python tests/gpu_attn_part_probe.py 2>&1 | tail -1 >> $L
done
cat $L
