# usage: ncu_attn.sh <lib suffix> <report name>   (full-size bounded dense attention, one launch)
mkdir -p gpurun_out
export K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$1.so K5_VARIANT_NAME=ncu K5_VARIANT_BOUND=1 K5_VARIANT_NOCHECK=1
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 9 -c 1 -f -o gpurun_out/$2 python tests/gpu_attn_variants.py > gpurun_out/$2.log 2>&1
tail -3 gpurun_out/$2.log
